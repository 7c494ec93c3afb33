// k_blur_tc.cu - K1 on the tensor cores (OPT-IN: VBS_BLUR_TC=1 or vbs_set_blur_tc, the default
// stays the integer-dot-product kernel of k_blur.cu).  Same contract as k_blur.cu: gray frames -> two fixed-point
// Gaussian blurs (cv2.GaussianBlur on CV_8U, MD:117-126) -> wrapping DoG -> inRange -> bit-packed area mask +
// per-frame popcount, bit-exact.
//
// Both separable passes are banded-Toeplitz int8 GEMMs on tcgen05.mma kind::i8 (u8/s8 x u8/s8 -> s32, exact):
//
//   pass 1 (horizontal)  D1[m, y] = sum_x' A1[m, x'] * img[y, x']      M = 128 = 64 output columns x 2 blurs
//        A1 = the two tap rows of every output column (REFLECT_101 at the left/right image edge FOLDED INTO the
//        matrix: one variant per 64-px strip, built on the host), B = 64 image rows as they lie in memory (K-major),
//        staged by TMA with 128-byte swizzle; out-of-image columns arrive as zeros, which is what the folded matrix wants.
//   epilogue 1           D1 (<= 65280) leaves TMEM, is split into high / low bytes, both flipped to signed (^0x80),
//        and stored K-major (y contiguous) as the B operand of pass 2 - no transposition: a thread owns one
//        (blur, column) TMEM lane and writes its own 64-byte row.
//   pass 2 (vertical)    D2_b[y, (x,hi|lo)] = sum_y' A2_b[y, y'] * Hb[(x,hi|lo), y']   per blur b, M = 128 rows
//        A2 = plain Toeplitz.  REFLECT_101 at the top / bottom edge is handled by giving pass 1 MIRRORED image rows
//        (single-row TMA loads) for the virtual rows above row 0 and below row H-1.
//   epilogue 2           u_b = 256 * D2_b[hi] + D2_b[lo];  blur_b = (u_b >> 16) + const  (the rounding constant
//        32768 and the sign flips cancel: sum(taps) = 256, so sum tap*(v-128) = sum tap*v - 32768, and the two
//        constants are equal for both blurs)  ->  dog = uint8(b_large - b_small + 15), lo <= dog <= hi, 64 bits per row.
//
// A work item = one 64-column strip of one frame, marched down in 64-row chunks; persistent CTAs (one per SM) run a
// contiguous range of items without draining the pipeline in between.  Roles: TMA warp, one MMA-issuing warp per pass,
// 4 warps epilogue 1, 8 warps epilogue 2; accumulators in TMEM (D1 x3 stages, D2), mbarrier hand-offs.
// What bounds it (per-CTA timeline, VBS_TC_TIMELINE=1): a kind::i8 MMA reads 32 K-bytes of both operands from shared
// memory, (4 KB + N x 32 B) per instruction, and at N <= 128 that - not the multipliers - sets its duration (N = 64:
// 48.9 cycles = 6 KB at 128 B/clk, N = 128: 64.6 cycles); both passes together read 92 KB per chunk, epilogue 1 and TMA
// write another 28 KB, so the shared-memory ports allow ~0.94 k cycles per chunk (~0.5 ms per 256 1080p frames) and the
// kernel runs at ~1.3 k.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "vbs_ctx.h"

namespace {

constexpr int SW = 64;             // output columns per CTA
constexpr int CHR = 64;            // image rows per chunk (pass 1 step) = K bytes per ring slab
constexpr int BLK = 128;           // output rows per pass-2 block
constexpr int KEXT = 256;          // K extent of both passes in bytes (two 128-byte swizzle slabs)
constexpr int NSLAB = 6;           // ring of pass-1 results (4 are read by a block, 2 are being produced: epilogue 1 never waits for the block in flight)
constexpr int NSTAGE = 3;          // image-row stages in flight and D1 accumulators (a TMA load takes ~3000 cycles from issue to arrival; a 4th stage
                                   // and an L2 prefetch 8 chunks ahead measured no gain: epilogue 1 and the shared-memory ports set the pace)
constexpr int NTHREADS = 512;      // warp 0: TMA, 1: MMA pass 1, 2: TMEM allocator, 3: MMA pass 2, 4-7: epilogue 1, 8-15: epilogue 2

// Pass 1 only has taps in K bytes 0..191 (image columns x0 - 64 .. x0 + 127), so its operands are a 128-byte slab
// (SWIZZLE_128B) plus a 64-byte half slab (SWIZZLE_64B): three image stages fit where two full ones did.
constexpr uint32_t OFF_A1 = 0;                         // [128 rows][128 B] + [128 rows][64 B]    24 KB
constexpr uint32_t OFF_A1H = 16384;
constexpr uint32_t OFF_A2 = 24576;                     // large blur: [2 K-slabs][128 rows][128 B]  32 KB
constexpr uint32_t OFF_A2S = OFF_A2 + 32768;           // small blur, K bytes 32..223 only: [128 rows][128 B] + [128 rows][64 B]  24 KB
constexpr uint32_t OFF_A2SH = OFF_A2S + 16384;
constexpr uint32_t OFF_B1 = OFF_A2SH + 8192;           // [NSTAGE]{[64 rows][128 B] + [64 rows][64 B]}   48 KB
constexpr uint32_t B1_STAGE = 12288, B1_HALF = 8192;
constexpr uint32_t OFF_H = OFF_B1 + NSTAGE * B1_STAGE; // [NSLAB][2 blurs][128 rows][64 B]        96 KB
constexpr uint32_t OFF_BAR = OFF_H + NSLAB * 16384;    // mbarriers + TMEM base + abort flag
constexpr uint32_t SMEM_BYTES = OFF_BAR + 256 + 1024;  // + slack to align the base to 1024 B

enum { BAR_A1 = 0, BAR_A2 = 1, BAR_D2_FULL = 2, BAR_D2_EMPTY = 3, BAR_B1_FULL = 4, BAR_B1_EMPTY = BAR_B1_FULL + NSTAGE, BAR_D1_FULL = BAR_B1_EMPTY + NSTAGE,
       BAR_D1_EMPTY = BAR_D1_FULL + NSTAGE, BAR_H_FULL = BAR_D1_EMPTY + NSTAGE, BAR_BLK = BAR_H_FULL + NSLAB, NBARS = BAR_BLK + 4 };
static_assert(8 * NBARS + 8 <= 256, "barrier block");

__host__ __device__ constexpr uint32_t tm_d1(int st) { return st < 2 ? 64u * (uint32_t)st : 384u + 64u * (uint32_t)(st - 2); }      // TMEM columns of D1 stage st (D2 sits at 128..383)
constexpr uint32_t TM_D2 = 128;    // D2 of blur b at 128 + 128 b: columns [0,64) high bytes, [64,128) low bytes
constexpr uint32_t TM_COLS = 512;

struct TcParams {
    int H, W, WW, lo, hi;
    int nreal, nchunks, nblocks;   // real chunks (image rows), chunk slots incl. the mirrored ones, blocks of 128 output rows
    int radius;                    // rows mirrored above row 0 and below row H-1 (radius of the large blur)
    int batch, nstrips, total;     // work items = (strip, frame) pairs; CTA c runs items [total c / grid, total (c + 1) / grid)
    int strip_major;               // 1: item e = strip * batch + frame (persistent CTAs keep their operator matrix), 0: frame * nstrips + strip
    int dbg_item;                  // which of the CTA's items the timeline records
    uint32_t *area_bits, *area_count, *status;
    long long *dbg;                // optional timeline of one CTA (VBS_TC_TIMELINE=1): clock64 at pipeline events
};
#define TC_DBG(id) do { if (dbg_on) P.dbg[id] = clock64(); } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {         // non-blocking
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
// bounded wait (a mis-programmed pipeline must not hang the GPU): false on time-out or when another role gave up
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, volatile uint32_t *abort_flag) {
    for (int spin = 0; spin < (1 << 20); ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return true;
        if ((spin & 63) == 63 && *abort_flag) return false;
    }
    *abort_flag = 1;
    return false;
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int x, int y, int z, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int x, int y, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], 8-bit integer operands, int32 accumulators
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
// one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, swizzled (cute::UMMA::SmemDescriptor): start address and stride between
// 8-row groups in 16-byte units, version 1, layout 2 = SWIZZLE_128B, 4 = SWIZZLE_64B.  Moving the start by n bytes
// (a K step inside the swizzle row, another slab) is adding n / 16 to the descriptor.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)layout << 61);
}
// instruction descriptor for kind::i8 (cute::UMMA::InstrDescriptor): D = s32, A/B format 0 = u8, 1 = s8, both K-major
__host__ __device__ constexpr uint32_t idesc_i8(int M, int N, uint32_t afmt, uint32_t bfmt) {
    return (2u << 4) | (afmt << 7) | (bfmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// K steps (32 bytes) that hold non-zero taps: pass 1 reads image columns x0 - 64 + k, pass 2 virtual rows 128 b - 64 + k
__host__ __device__ constexpr int k1_lo(int R) { return (64 - R) / 32; }
__host__ __device__ constexpr int k1_hi(int R) { return (64 + 63 + R) / 32 + 1; }
__host__ __device__ constexpr int k2_lo(int R) { return (64 - R) / 32; }
__host__ __device__ constexpr int k2_hi(int R) { return (64 + 127 + R) / 32 + 1; }

// byte offset of (row, k) inside a [128 rows][64 B] K-major SWIZZLE_64B tile (tile base 1024-byte aligned)
__device__ __forceinline__ uint32_t swz64(uint32_t row, uint32_t k) {
    const uint32_t off = row * 64u + k;
    return off ^ (((off >> 7) & 3u) << 4);
}

// RL / RS: radii of the large / small blur (50 / 19 above 480 rows, 17 / 10 below: MD:117-126)
template <int RL, int RS>
__global__ void __launch_bounds__(NTHREADS, 1)
blur_area_tc_kernel(const __grid_constant__ CUtensorMap map_img, const __grid_constant__ CUtensorMap map_imgh,
                    const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_a1h,
                    const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_a2h, const TcParams P) {
    static_assert(k2_lo(RS) == 1 && k2_hi(RS) == 7, "the small blur's pass-2 matrix is stored for K steps 1..6");
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *gen = smem_raw + (base - smem_u32(smem_raw));          // generic pointer to the aligned base
    const uint32_t bars = base + OFF_BAR;
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + OFF_BAR + 8 * NBARS);
    volatile uint32_t *abort_flag = tmem_slot + 1;
    auto bar = [&](int i) -> uint32_t { return bars + 8u * (uint32_t)i; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int CR = P.nreal, C = P.nchunks, NB = P.nblocks;    // chunk slot cc covers virtual rows [64 (cc - 1), 64 cc); real: 1 .. CR
    // Persistent CTA: a contiguous range of work items.  Every role walks the same item sequence and keeps RUNNING
    // counters (g1: real chunks, gc0: chunk slots, gb0: blocks) from which all stage indices and barrier parities
    // follow, so the pipeline never drains between items: the loads and pass 1 of item n+1 run under pass 2 and the
    // epilogues of item n.
    const int e0 = (int)((long long)P.total * blockIdx.x / gridDim.x), e1 = (int)((long long)P.total * (blockIdx.x + 1) / gridDim.x);
    auto item_strip = [&](int e) -> int { return P.strip_major ? e / P.batch : e % P.nstrips; };
    auto item_frame = [&](int e) -> int { return P.strip_major ? e % P.batch : e / P.nstrips; };
    const bool dbg_cta = P.dbg != nullptr && blockIdx.x == 5;
    bool dbg_on = dbg_cta;
    if (tid == 0) TC_DBG(0);

    if (tid == 0) {
        mbar_init(bar(BAR_A1), 1); mbar_init(bar(BAR_A2), 1);
        for (int i = 0; i < NSTAGE; ++i) {
            mbar_init(bar(BAR_B1_FULL + i), 1); mbar_init(bar(BAR_B1_EMPTY + i), 1);
            mbar_init(bar(BAR_D1_FULL + i), 1); mbar_init(bar(BAR_D1_EMPTY + i), 128);
        }
        for (int i = 0; i < NSLAB; ++i) mbar_init(bar(BAR_H_FULL + i), 128);
        for (int i = 0; i < 4; ++i) mbar_init(bar(BAR_BLK + i), 1);
        mbar_init(bar(BAR_D2_FULL), 1); mbar_init(bar(BAR_D2_EMPTY), 256);
        *abort_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_img) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_imgh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a1h) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a2) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a2h) : "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void *)tmem_slot)), "r"(TM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<const uint32_t *>(gen + OFF_BAR + 8 * NBARS);     // plain load at a uniform address: a uniform value
    if (tid == 0) TC_DBG(1);

    if (warp == 0) {
        // ================= TMA producer (warp-uniform control flow, one elected lane issues): operator matrices, then
        // 64 image rows per chunk =================
        {
            bool ok = true;
            int st = 0; uint32_t ph = 0;
            int g1 = 0, cur_strip = -1;
            for (int e = e0; e < e1 && ok; ++e) {
            const int strip = item_strip(e), f = item_frame(e), x0 = strip * SW;
            dbg_on = dbg_cta && e - e0 == P.dbg_item;
            if (strip != cur_strip) {                                           // a new operator matrix: every pass-1 MMA that reads the old one must be done
                if (g1 > 0) {
                    const int gl = g1 - 1;                                      // the last chunk issued commits to its stage's EMPTY barrier
                    ok = __all_sync(0xffffffffu, mbar_wait(bar(BAR_B1_EMPTY + gl % NSTAGE), (uint32_t)(gl / NSTAGE) & 1u, abort_flag));
                    if (!ok) break;
                }
                if (elect_one()) {
                    mbar_expect_tx(bar(BAR_A1), 16384u + 8192u);
                    tma_load_2d(base + OFF_A1, &map_a1, 0, 128 * strip, bar(BAR_A1));
                    tma_load_2d(base + OFF_A1H, &map_a1h, 128, 128 * strip, bar(BAR_A1));
                }
                __syncwarp();
                cur_strip = strip;
            }
            for (int i = 0; i < CR && ok; ++i, ++g1) {                          // real chunk i + 1 = image rows [64 i, 64 i + 64)
                if (g1 >= NSTAGE) ok = __all_sync(0xffffffffu, mbar_wait(bar(BAR_B1_EMPTY + st), ph ^ 1u, abort_flag));
                if (!ok) break;
                if (elect_one()) {
                    const uint32_t dst = base + OFF_B1 + B1_STAGE * st;
                    mbar_expect_tx(bar(BAR_B1_FULL + st), B1_STAGE);            // rows / columns outside the image arrive as zeros
                    tma_load_3d(dst, &map_img, x0 - 64, CHR * i, f, bar(BAR_B1_FULL + st));
                    tma_load_3d(dst + B1_HALF, &map_imgh, x0 + 64, CHR * i, f, bar(BAR_B1_FULL + st));
                    TC_DBG(120 + i);
                    if (e == e0 && (i == 1 || CR == 1)) {                       // pass 2 starts four chunks in: its matrices load behind the first rows
                        mbar_expect_tx(bar(BAR_A2), 32768u + 16384u + 8192u);
                        for (int s = 0; s < 2; ++s) tma_load_2d(base + OFF_A2 + 16384u * s, &map_a2, 128 * s, 0, bar(BAR_A2));
                        tma_load_2d(base + OFF_A2S, &map_a2, 32, 128, bar(BAR_A2));             // small blur: K bytes 32..159
                        tma_load_2d(base + OFF_A2SH, &map_a2h, 160, 128, bar(BAR_A2));          //             K bytes 160..223
                    }
                }
                __syncwarp();
                if (++st == NSTAGE) { st = 0; ph ^= 1u; }
            }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer, pass 1 (warp-uniform control flow, one elected lane issues).  Everything between two
        // MMAs is scalar code of one thread, so the descriptors are built once and moved by constants, the K loop is unrolled over
        // a compile-time range, and the waits are the blocking kind (a poll is a ~150-cycle round trip: one warp polling for both
        // passes paced the whole pipeline at ~900 cycles per issue - hence one issuing warp per pass). =================
        constexpr uint32_t ID1 = idesc_i8(128, 64, 0, 0);                       // taps (u8, <= 26) x pixels (u8)
        const uint64_t a1d = smem_desc(base + OFF_A1, 1024, 2), b1d = smem_desc(base + OFF_B1, 1024, 2);
        const uint64_t a1hd = smem_desc(base + OFF_A1H, 512, 4), b1hd = smem_desc(base + OFF_B1 + B1_HALF, 512, 4);      // K bytes 128..191
        bool ok = true;
        int st = 0; uint32_t ph = 0;
        int g1 = 0, cur_strip = -1; uint32_t a1_loads = 0;
        for (int e = e0; e < e1 && ok; ++e) {
        const int strip = item_strip(e);
        dbg_on = dbg_cta && e - e0 == P.dbg_item;
        if (strip != cur_strip) {
            ok = __all_sync(0xffffffffu, mbar_wait(bar(BAR_A1), a1_loads & 1u, abort_flag));
            ++a1_loads; cur_strip = strip;
            if (!ok) break;
        }
        for (int i = 0; i < CR && ok; ++i, ++g1) {                              // real chunk i + 1
            ok = mbar_wait(bar(BAR_B1_FULL + st), ph, abort_flag);
            if (ok && g1 >= NSTAGE) ok = mbar_wait(bar(BAR_D1_EMPTY + st), ph ^ 1u, abort_flag);
            ok = __all_sync(0xffffffffu, ok);
            if (!ok) break;
            tc_fence_after();
            const uint64_t so = (uint64_t)((B1_STAGE / 16) * st);
            const uint32_t dd = tmem + tm_d1(st);
            if (elect_one()) {
#pragma unroll
                for (int j = k1_lo(RL); j < k1_hi(RL); ++j) {
                    if (j < 4) tc_mma_i8(dd, a1d + (uint64_t)(2 * j), b1d + so + (uint64_t)(2 * j), ID1, j > k1_lo(RL));
                    else tc_mma_i8(dd, a1hd + (uint64_t)(2 * (j - 4)), b1hd + so + (uint64_t)(2 * (j - 4)), ID1, j > k1_lo(RL));
                }
                tc_commit(bar(BAR_B1_EMPTY + st));
                tc_commit(bar(BAR_D1_FULL + st));
                TC_DBG(10 + i);
            }
            __syncwarp();
            if (++st == NSTAGE) { st = 0; ph ^= 1u; }
        }
        }
    } else if (warp == 3) {
        // ================= MMA issuer, pass 2: output rows [128 b, 128 b + 128) read chunk slots 2 b .. 2 b + 3 of the ring =================
        constexpr uint32_t ID2 = idesc_i8(128, 128, 1, 1);                      // taps (s8, <= 13) x sign-flipped bytes (s8)
        const uint64_t a2d = smem_desc(base + OFF_A2, 1024, 2), hd = smem_desc(base + OFF_H, 512, 4);
        const uint64_t a2sd = smem_desc(base + OFF_A2S, 1024, 2), a2shd = smem_desc(base + OFF_A2SH, 512, 4);
        bool ok = __all_sync(0xffffffffu, mbar_wait(bar(BAR_A2), 0, abort_flag));
        for (int e = e0; e < e1 && ok; ++e) {
        const int gc0 = (e - e0) * C, gb0 = (e - e0) * NB;                      // running chunk-slot / block numbers of this item's first
        dbg_on = dbg_cta && e - e0 == P.dbg_item;
        int slab0 = gc0 % NSLAB;                                                // ring slot of chunk 2 b
        for (int b2 = 0; b2 < NB && ok; ++b2) {
            const int gb = gb0 + b2;
            const int c_need = min(2 * b2 + 3, C - 1);                          // newest chunk slot the block reads (epilogue 1 finishes chunks in order)
            ok = mbar_wait(bar(BAR_H_FULL + (gc0 + c_need) % NSLAB), (uint32_t)((gc0 + c_need) / NSLAB) & 1u, abort_flag);
            if (ok && gb >= 1) ok = mbar_wait(bar(BAR_D2_EMPTY), (uint32_t)(gb - 1) & 1u, abort_flag);
            ok = __all_sync(0xffffffffu, ok);
            if (!ok) break;
            uint64_t hs[4];
            int sl = slab0, cc = 2 * b2;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (cc > C - 1) { hs[q] = hs[q - (q > 0)]; continue; }          // slots past the last one: taps there only feed rows that are never stored
                hs[q] = hd + (uint64_t)(1024u * sl);
                ++cc; if (++sl == NSLAB) sl = 0;
            }
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int j = k2_lo(RL); j < k2_hi(RL); ++j)
                    tc_mma_i8(tmem + TM_D2, a2d + (uint64_t)(1024 * (j >> 2) + 2 * (j & 3)), hs[j >> 1] + (uint64_t)(2 * (j & 1)), ID2, j > k2_lo(RL));
#pragma unroll
                for (int j = k2_lo(RS); j < k2_hi(RS); ++j)
                    tc_mma_i8(tmem + TM_D2 + 128u, j < 5 ? a2sd + (uint64_t)(2 * (j - 1)) : a2shd + (uint64_t)(2 * (j - 5)), hs[j >> 1] + (uint64_t)(512 + 2 * (j & 1)), ID2,
                              j > k2_lo(RS));
                tc_commit(bar(BAR_BLK + (gb & 3)));
                tc_commit(bar(BAR_D2_FULL));
                TC_DBG(40 + b2);
            }
            __syncwarp();
            slab0 += 2; if (slab0 >= NSLAB) slab0 -= NSLAB;
        }
        }
    } else if (warp >= 4 && warp < 8) {
        // ================= epilogue 1: D1 -> signed high / low bytes, K-major, into the ring =================
        const int q = warp & 3, row = 32 * q + lane;          // TMEM lane = (blur, column): blur = row / 64
        const int bl = row >> 6, n = row & 63;
        int gc0 = 0, gb0 = 0;                                  // running chunk-slot / block numbers of the current item's first
        auto tile_of = [&](int cc) -> unsigned char * { return gen + OFF_H + 16384u * (uint32_t)((gc0 + cc) % NSLAB) + 8192u * (uint32_t)bl; };
        // the four bytes of virtual rows t .. t + 3 (t % 4 == 0) of one of this thread's two ring rows
        auto vword = [&](int prow, int t) -> uint32_t * {
            return reinterpret_cast<uint32_t *>(tile_of((t >> 6) + 1) + swz64((uint32_t)prow, (uint32_t)(t & 63)));
        };
        // the slot of chunk cc held the chunk NSLAB slots earlier (of this item or the previous one), last read by block
        // min(NB - 1, chunk / 2) of that item (a block reads chunks 2 b .. 2 b + 3)
        int blk_known = -1;                                    // newest block known complete (blocks complete in order; a wait is a ~150-cycle round trip)
        auto slot_free = [&](int cc) -> bool {
            const int g = gc0 + cc - NSLAB;
            if (g < 0) return true;
            int pc = cc - NSLAB, pb0 = gb0;
            while (pc < 0) { pc += C; pb0 -= NB; }              // (frames of fewer than NSLAB chunk slots: further back)
            const int bdone = pb0 + min(NB - 1, pc >> 1);
            if (bdone <= blk_known) return true;
            if (!mbar_wait(bar(BAR_BLK + (bdone & 3)), (uint32_t)(bdone >> 2) & 1u, abort_flag)) return false;
            blk_known = bdone;
            return true;
        };
        bool ok = true;
        int g1 = 0;
        for (int e = e0; e < e1 && ok; ++e, gc0 += C, gb0 += NB) {
        dbg_on = dbg_cta && e - e0 == P.dbg_item;
        for (int i = 0; i < CR && ok; ++i, ++g1) {
            const int st = g1 % NSTAGE, cc = i + 1;
            ok = mbar_wait(bar(BAR_D1_FULL + st), (uint32_t)(g1 / NSTAGE) & 1u, abort_flag);
            if (!ok) break;
            tc_fence_after();
            uint32_t v[4][16];
#pragma unroll
            for (int g = 0; g < 4; ++g) tmem_ld16(tmem + ((uint32_t)(32 * q) << 16) + tm_d1(st) + 16u * g, v[g]);
            tmem_ld_wait();
            if (tid == 128) TC_DBG(150 + i);
            tc_fence_before();
            mbar_arrive(bar(BAR_D1_EMPTY + st));                // the accumulator stage may be overwritten
            if (cc == 1 && !(ok = slot_free(0))) break;         // the top mirror goes into the slot before this chunk's
            if (!(ok = slot_free(cc))) break;
            if (tid == 128) TC_DBG(170 + i);
            unsigned char *tile = tile_of(cc);                  // [128 rows][64 B]: rows 0..63 high bytes, 64..127 low bytes
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint32_t p[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) p[t] = __byte_perm(v[g][2 * t], v[g][2 * t + 1], 0x5140);   // {a.b0, b.b0, a.b1, b.b1}
                uint4 lo4, hi4;
                lo4.x = __byte_perm(p[0], p[1], 0x5410) ^ 0x80808080u; hi4.x = __byte_perm(p[0], p[1], 0x7632) ^ 0x80808080u;
                lo4.y = __byte_perm(p[2], p[3], 0x5410) ^ 0x80808080u; hi4.y = __byte_perm(p[2], p[3], 0x7632) ^ 0x80808080u;
                lo4.z = __byte_perm(p[4], p[5], 0x5410) ^ 0x80808080u; hi4.z = __byte_perm(p[4], p[5], 0x7632) ^ 0x80808080u;
                lo4.w = __byte_perm(p[6], p[7], 0x5410) ^ 0x80808080u; hi4.w = __byte_perm(p[6], p[7], 0x7632) ^ 0x80808080u;
                const uint32_t ch = (uint32_t)g ^ ((uint32_t)(n >> 1) & 3u);            // SWIZZLE_64B: 16-byte chunk index ^= (row / 2) & 3
                *reinterpret_cast<uint4 *>(tile + 64u * n + 16u * ch) = hi4;
                *reinterpret_cast<uint4 *>(tile + 64u * (64 + n) + 16u * ch) = lo4;      // ((64 + n) / 2) & 3 == (n / 2) & 3
            }
            // REFLECT_101 in y: the horizontal pass commutes with it, so the virtual rows above row 0 and below row H-1 are
            // COPIES of rows this thread has just produced (its own two ring rows: no cross-thread hazard)
            if (cc == 1) {                                      // rows -j <- rows j: chunk slot 0, k = 64 - j, straight from the registers
                unsigned char *t0 = tile_of(0);
#pragma unroll
                for (int c = 0; c < 4; ++c) {                   // 16-byte chunk c of slot 0: k = 16 c .. 16 c + 15  <-  rows 64 - k
                    uint32_t lw[4], hw[4];
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        constexpr int Z = 63;                   // (k = 0 would need row 64: never read, the taps end at 63 rows)
                        const int y0 = 64 - 16 * c - 4 * w;
                        const uint32_t va = v[(y0 > Z ? Z : y0) >> 4][(y0 > Z ? Z : y0) & 15], vb = v[(y0 - 1) >> 4][(y0 - 1) & 15];
                        const uint32_t vc = v[(y0 - 2) >> 4][(y0 - 2) & 15], vd = v[(y0 - 3) >> 4][(y0 - 3) & 15];
                        const uint32_t pa = __byte_perm(va, vb, 0x5140), pb = __byte_perm(vc, vd, 0x5140);
                        lw[w] = __byte_perm(pa, pb, 0x5410) ^ 0x80808080u;
                        hw[w] = __byte_perm(pa, pb, 0x7632) ^ 0x80808080u;
                    }
                    const uint32_t ch = (uint32_t)c ^ ((uint32_t)(n >> 1) & 3u);
                    *reinterpret_cast<uint4 *>(t0 + 64u * n + 16u * ch) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                    *reinterpret_cast<uint4 *>(t0 + 64u * (64 + n) + 16u * ch) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                }
            }
            if (cc == CR) {                                     // rows H + i <- rows H - 2 - i, i < radius (sources may sit in the previous chunk)
                for (int ct = CR + 1; ct < C && ok; ++ct) ok = slot_free(ct);
                if (!ok) break;
                // word-wise: target bytes t0 .. t0 + 3 are source bytes M - t0 - 3 .. M - t0 reversed (M = 2 H - 2); the source span
                // starts (M + 1) & 3 bytes into an aligned word, the same for every word
                const int M2 = 2 * P.H - 2, sh = (M2 + 1) & 3;
#pragma unroll 2
                for (int t0 = P.H & ~3; t0 < P.H + P.radius; t0 += 4) {         // both planes per iteration: two independent chains
                    const int a = (M2 - t0 - 3) & ~3;
                    const uint32_t h0 = *vword(n, a), l0 = *vword(64 + n, a);
                    const uint32_t h1 = sh ? *vword(n, a + 4) : 0u, l1 = sh ? *vword(64 + n, a + 4) : 0u;
                    const uint32_t hrev = __byte_perm(__funnelshift_r(h0, h1, 8 * sh), 0u, 0x0123);
                    const uint32_t lrev = __byte_perm(__funnelshift_r(l0, l1, 8 * sh), 0u, 0x0123);
                    uint32_t *hd_ = vword(n, t0), *ld_ = vword(64 + n, t0);
                    if (t0 >= P.H) { *hd_ = hrev; *ld_ = lrev; }
                    else {                                      // the word that straddles row H: keep the real rows below H
                        const uint32_t keep = (1u << (8 * (P.H - t0))) - 1u;
                        *hd_ = (*hd_ & keep) | (hrev & ~keep);
                        *ld_ = (*ld_ & keep) | (lrev & ~keep);
                    }
                }
            }
            if (tid == 128) TC_DBG(190 + i);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                  // generic writes -> visible to the tensor core's reads
            if (cc == 1) mbar_arrive(bar(BAR_H_FULL + gc0 % NSLAB));
            mbar_arrive(bar(BAR_H_FULL + (gc0 + cc) % NSLAB));
            if (cc == CR)
                for (int ct = CR + 1; ct < C; ++ct) mbar_arrive(bar(BAR_H_FULL + (gc0 + ct) % NSLAB));
            if (tid == 128) TC_DBG(60 + i);
        }
        }
    } else if (warp >= 8) {
        // ================= epilogue 2: D2 -> rounding, wrapping DoG, inRange, 32 bits per thread and row =================
        const int q = warp & 3, half = (warp - 8) >> 2, r = 32 * q + lane;     // TMEM lane = output row within the block; columns 32 half ..
        const uint32_t lane_addr = tmem + ((uint32_t)(32 * q) << 16) + TM_D2 + 32u * half;
        const uint32_t bias16 = (uint32_t)(15 - P.lo) << 16, lim16 = (uint32_t)(P.hi - P.lo + 1) << 16;
        bool ok = true;
        int gb = 0;
        for (int e = e0; e < e1 && ok; ++e) {
        const int f = item_frame(e), x0 = item_strip(e) * SW;
        dbg_on = dbg_cta && e - e0 == P.dbg_item;
        const int nvalid = min(32, P.W - x0 - 32 * half);       // columns of this half strip inside the image
        const uint32_t colmask = nvalid >= 32 ? ~0u : nvalid > 0 ? ((1u << nvalid) - 1u) : 0u;
        const int wx = (x0 >> 5) + half;
        uint32_t count = 0;
        for (int b = 0; b < NB && ok; ++b, ++gb) {
            ok = mbar_wait(bar(BAR_D2_FULL), (uint32_t)gb & 1u, abort_flag);
            if (!ok) break;
            tc_fence_after();
            uint32_t bits = 0;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                uint32_t hL[16], lL[16], hS[16], lS[16];
                tmem_ld16(lane_addr + 16u * g, hL);
                tmem_ld16(lane_addr + 64u + 16u * g, lL);
                tmem_ld16(lane_addr + 128u + 16u * g, hS);
                tmem_ld16(lane_addr + 192u + 16u * g, lS);
                tmem_ld_wait();
                if (g == 1) { tc_fence_before(); mbar_arrive(bar(BAR_D2_EMPTY)); }      // this thread's part of D2 is in registers
                uint32_t bitsg = 0;
                // uint8(b_large - b_small + 15) (MD:128) only needs the blurs mod 256 = byte 2 of u = 256 hi + lo; subtracting the
                // small blur with its low half cleared leaves the difference in byte 2 (no borrow from below), the bias is added
                // in place, and `in range` (MD:129) becomes the sign of (byte 2) - (span + 1), shifted into the row word
#pragma unroll
                for (int i = 15; i >= 0; --i) {
                    const uint32_t uL = (uint32_t)((int)hL[i] * 256 + (int)lL[i]);
                    const uint32_t uS = (uint32_t)((int)hS[i] * 256 + (int)lS[i]);
                    const uint32_t x = (uL - (uS & 0xffff0000u) + bias16) & 0x00ff0000u;         // d << 16
                    bitsg = __funnelshift_l(x - lim16, bitsg, 1);                                 // bit 31 of (d - span - 1) << 16: d <= span
                }
                bits |= (bitsg & 0xffffu) << (16 * g);
            }
            if (tid == 256) TC_DBG(90 + b);
            const int y = BLK * b + r;
            if (y < P.H && wx < P.WW) {
                bits &= colmask;
                P.area_bits[((size_t)f * P.H + y) * P.WW + wx] = bits;
                count += __popc(bits);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
        if (lane == 0 && count) atomicAdd(P.area_count + f, count);              // mean of area_mask for the NCC (MD:153)
        }
    }

    tc_fence_before();
    __syncthreads();
    dbg_on = dbg_cta;
    if (tid == 0) TC_DBG(2);
    if (tid == 0 && *abort_flag) atomicOr(P.status, VBS_DEV_TMA_TIMEOUT);
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TM_COLS) : "memory");
    }
}

// ---- host side -------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

bool encode_u8(CUtensorMap *map, const void *ptr, int rank, const cuuint64_t *dims, const cuuint64_t *strides, const cuuint32_t *box, bool sw64 = false) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, (cuuint32_t)rank, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               sw64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int host_reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * (n - 1) - i;
    return i;
}

}  // namespace

// Operator matrices of both passes for the context's geometry (built once, kept on the device):
//   A1 [nstrips][128][256]  rows 0..63 large blur, 64..127 small blur of output column x0 + (row & 63); K byte k = image column
//                           x0 - 64 + k; taps that REFLECT_101 sends back into the image are added to the mirrored column
//   A2 [2][128][256]        output row r of a block, K byte k = virtual row 128 b - 64 + k: tap index k - 64 - r + radius
cudaError_t vbs_blur_tc_setup(vbs_ctx *ctx) {
    if (ctx->tc_a1) return cudaSuccess;
    const int W = ctx->W, KS = ctx->br.ks, KL = ctx->br.kl;
    const int nstrips = (W + SW - 1) / SW;
    int tapsL[128], tapsS[128];
    vbs_host_taps(KL, ctx->big ? 20.0 : 11.4, tapsL);
    vbs_host_taps(KS, ctx->big ? 8.0 : 4.56, tapsS);
    const int RL = KL / 2, RS = KS / 2;
    if (!((RL == 50 && RS == 19) || (RL == 17 && RS == 10))) return cudaErrorInvalidValue;     // the kernel's two instantiations
    std::vector<uint8_t> a1((size_t)nstrips * 128 * KEXT, 0), a2((size_t)2 * 128 * KEXT, 0);
    for (int s = 0; s < nstrips; ++s)
        for (int row = 0; row < 128; ++row) {
            const int bl = row >> 6, xo = s * SW + (row & 63);
            if (xo >= W) continue;
            const int R = bl ? RS : RL, *taps = bl ? tapsS : tapsL;
            for (int t = 0; t <= 2 * R; ++t) {
                const int k = host_reflect101(xo + t - R, W) - (s * SW - 64);
                if (k < 0 || k >= KEXT) return cudaErrorInvalidValue;           // cannot happen: |reflected offset| <= radius <= 63
                a1[((size_t)s * 128 + row) * KEXT + k] += (uint8_t)taps[t];
            }
        }
    for (int bl = 0; bl < 2; ++bl) {
        const int R = bl ? RS : RL, *taps = bl ? tapsS : tapsL;
        for (int r = 0; r < 128; ++r)
            for (int t = 0; t <= 2 * R; ++t) a2[((size_t)bl * 128 + r) * KEXT + (64 + r + t - R)] = (uint8_t)taps[t];
    }
    cudaError_t e;
    if ((e = cudaMalloc((void **)&ctx->tc_a1, a1.size())) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&ctx->tc_a2, a2.size())) != cudaSuccess) return e;
    if ((e = cudaMemcpy(ctx->tc_a1, a1.data(), a1.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
    return cudaMemcpy(ctx->tc_a2, a2.data(), a2.size(), cudaMemcpyHostToDevice);
}

// returns cudaErrorNotSupported when this batch cannot take the tensor-core path (BGR input, frames the TMA unit cannot
// describe): the caller then runs the integer-dot-product kernel
cudaError_t vbs_launch_blur_tc(vbs_ctx *ctx, const uint8_t *frames, int batch, int64_t frame_stride, int64_t row_pitch) {
    if (ctx->C != 1 || ctx->W < 128 || ctx->H < CHR || !encode_fn()) return cudaErrorNotSupported;
    if ((reinterpret_cast<uintptr_t>(frames) & 15) || (row_pitch & 15) || (frame_stride & 15)) return cudaErrorNotSupported;
    cudaError_t e = vbs_blur_tc_setup(ctx);
    if (e != cudaSuccess) return e;
    const int H = ctx->H, W = ctx->W, nstrips = (W + SW - 1) / SW;
    CUtensorMap m_img, m_imgh, m_a1, m_a1h, m_a2, m_a2h;
    {
        const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)batch};
        const cuuint64_t strides[2] = {(cuuint64_t)row_pitch, (cuuint64_t)frame_stride};
        const cuuint32_t box[3] = {128, CHR, 1}, boxh[3] = {64, CHR, 1};
        if (!encode_u8(&m_img, frames, 3, dims, strides, box) || !encode_u8(&m_imgh, frames, 3, dims, strides, boxh, true)) return cudaErrorNotSupported;
    }
    {
        const cuuint64_t d1[2] = {KEXT, (cuuint64_t)128 * nstrips}, d2[2] = {KEXT, 256}, st[1] = {KEXT};
        const cuuint32_t box[2] = {128, 128}, boxh[2] = {64, 128};
        if (!encode_u8(&m_a1, ctx->tc_a1, 2, d1, st, box) || !encode_u8(&m_a1h, ctx->tc_a1, 2, d1, st, boxh, true) ||
            !encode_u8(&m_a2, ctx->tc_a2, 2, d2, st, box) || !encode_u8(&m_a2h, ctx->tc_a2, 2, d2, st, boxh, true))
            return cudaErrorNotSupported;
    }
    const int R = ctx->br.kl / 2;
    TcParams P;
    P.H = H; P.W = W; P.WW = ctx->WW; P.lo = ctx->br.lo; P.hi = ctx->br.hi;
    P.nreal = (H + CHR - 1) / CHR;
    P.nchunks = (H - 1 + R) / CHR + 2;                           // slots for virtual rows [-64, H + radius)
    P.nblocks = (H + BLK - 1) / BLK;
    P.radius = R;
    // persistent CTAs, one per SM, over strip-major items (a CTA changes its operator matrix at most a few times); neighbouring
    // strips of a frame are still in flight at about the same time, so their halo columns meet in L2.  VBS_TC_PERSIST=0: one
    // CTA per item, frame-major.
    const char *pe = getenv("VBS_TC_PERSIST");
    const bool persist = !(pe && pe[0] == '0');
    P.batch = batch; P.nstrips = nstrips; P.total = nstrips * batch;
    P.strip_major = persist ? 1 : 0;
    const int grid = persist ? (P.total < ctx->sm_count ? P.total : ctx->sm_count) : P.total;
    P.dbg_item = (P.total / grid) > 1 ? 1 : 0;
    P.area_bits = ctx->area_bits; P.area_count = ctx->area_count; P.status = ctx->d_status;
    P.dbg = nullptr;
    static long long *dbg_buf = nullptr;
    if (const char *env = getenv("VBS_TC_TIMELINE"); env && env[0] == '1') {
        if (!dbg_buf) { cudaMalloc((void **)&dbg_buf, 256 * sizeof(long long)); }
        cudaMemsetAsync(dbg_buf, 0, 256 * sizeof(long long), ctx->stream);
        P.dbg = dbg_buf;
    }
    auto kern = ctx->big ? blur_area_tc_kernel<50, 19> : blur_area_tc_kernel<17, 10>;
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES)) != cudaSuccess) return e;
    kern<<<dim3(grid), NTHREADS, SMEM_BYTES, ctx->stream>>>(m_img, m_imgh, m_a1, m_a1h, m_a2, m_a2h, P);
    ctx->launches += 1;
    ctx->tc_launches += 1;
    if (P.dbg) {                                                 // developer aid: print the timeline of one item of CTA 5 relative to the CTA's start
        long long h[256];
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "tc timeline (cycles after CTA start): setup %lld end %lld\n", h[1] - h[0], h[2] - h[0]);
        const char *names[6] = {"tma", "p1", "e1ld", "slot", "stored", "e1"};
        const int base[6] = {120, 10, 150, 170, 190, 60};
        for (int i = 0; i < P.nreal; ++i) {
            fprintf(stderr, "  chunk %2d:", i);
            for (int k = 0; k < 6; ++k) fprintf(stderr, " %s %6lld", names[k], h[base[k] + i] ? h[base[k] + i] - h[0] : -1);
            fprintf(stderr, "\n");
        }
        for (int b = 0; b < P.nblocks; ++b) fprintf(stderr, "  block %2d: p2 %6lld e2 %6lld\n", b, h[40 + b] - h[0], h[90 + b] - h[0]);
    }
    return cudaGetLastError();
}
