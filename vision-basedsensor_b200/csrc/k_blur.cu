// k_blur.cu - K1: gray -> two fixed-point Gaussian blurs -> wrapping DoG -> inRange -> bit-packed
// area mask + per-frame popcount.  Replaces MD:114-129 (cvtColor, GaussianBlur x2, uint8
// subtraction, inRange) bit-exactly.
//
// One CTA marches down a 128-pixel-wide column strip.  Per step of 8 rows it
//   1. stages 8 input rows (+ halo) in shared memory (REFLECT_101 resolved while loading),
//   2. runs both horizontal passes with IDP.4A (u8 x u8 -> u32) on aligned words, the tap
//      alignment of each of the 4 pixels a thread owns being folded into compile-time weights,
//   3. packs vertically adjacent rows as u16 pairs into a ring of the last ~112 rows,
//   4. runs both vertical passes with IDP.2A (u16 x u8 -> u32), 8 output rows per thread so
//      every ring word is loaded once per 8 outputs,
//   5. rounds, forms uint8(b_large - b_small + 15), range-tests and ballots 32 pixels per word.
// Nothing but the 1 bit/pixel result leaves the SM.  No tensor cores: integer dot products only.
#include <cuda.h>
#include <cstring>
#include <type_traits>
#include <utility>
#include "vbs_ctx.h"

namespace {

// ---- baked 8.8 fixed-point taps (cv2.GaussianBlur on CV_8U); verified against the host recipe
// in vbs_check_taps() at context creation and against cv2 in tests/test_oracle_exact.py -------
template <int K> struct Taps;
template <> struct Taps<39> {
    static __host__ __device__ constexpr int at(int j) {
        constexpr int h[20] = {1, 1, 1, 2, 2, 3, 3, 5, 5, 6, 6, 8, 9, 10, 11, 11, 12, 13, 13, 12};
        return (j < 0 || j >= 39) ? 0 : h[j < 20 ? j : 38 - j];
    }
};
template <> struct Taps<101> {
    static __host__ __device__ constexpr int at(int j) {
        constexpr int h[51] = {0, 0, 1, 0, 0, 1, 0, 1, 0, 1, 1, 1, 0, 1, 1, 1, 2, 1, 1, 2, 2, 1, 2, 2, 3, 2,
                               3, 2, 3, 3, 3, 3, 4, 4, 3, 4, 4, 4, 5, 4, 5, 4, 5, 5, 5, 5, 5, 5, 5, 5, 6};
        return (j < 0 || j >= 101) ? 0 : h[j < 51 ? j : 100 - j];
    }
};
template <> struct Taps<21> {
    static __host__ __device__ constexpr int at(int j) {
        constexpr int h[11] = {2, 3, 5, 7, 10, 12, 16, 18, 21, 23, 22};
        return (j < 0 || j >= 21) ? 0 : h[j < 11 ? j : 20 - j];
    }
};
template <> struct Taps<35> {
    static __host__ __device__ constexpr int at(int j) {
        constexpr int h[18] = {3, 4, 4, 5, 6, 6, 6, 7, 7, 8, 9, 9, 9, 10, 10, 10, 10, 10};
        return (j < 0 || j >= 35) ? 0 : h[j < 18 ? j : 34 - j];
    }
};

template <int K> __host__ __device__ constexpr uint32_t pack4(int j0) {
    return (uint32_t)Taps<K>::at(j0) | ((uint32_t)Taps<K>::at(j0 + 1) << 8) |
           ((uint32_t)Taps<K>::at(j0 + 2) << 16) | ((uint32_t)Taps<K>::at(j0 + 3) << 24);
}

template <int B, int E, class F> __device__ __forceinline__ void static_for(F &&f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

constexpr int cdiv(int a, int b) { return (a + b - 1) / b; }
constexpr int rup(int a, int b) { return cdiv(a, b) * b; }

constexpr int TW = 128;       // strip width (pixels) = threads per CTA
constexpr int RB = 8;         // rows per step

template <int KS, int KL> struct Geo {
    static constexpr int RL = KL / 2, RS = KS / 2;
    static constexpr int HL = rup(RL, 4);                 // left/right halo, word aligned
    static constexpr int TWORDS = (TW + 2 * HL) / 4;      // words per staged input row
    static constexpr int NW = (3 + HL + RL) / 4 + 1;      // words a thread reads per row
    static constexpr int PL = rup(RL, 2), PS = rup(RS, 2);
    static constexpr int LEAD = cdiv(RL + PL, RB);        // horizontal steps ahead of the vertical pass
    static constexpr int NGL = LEAD + 1;                  // ring groups (8 rows each), large kernel
    static constexpr int NPL = (PL + RB + RL + 1) / 2;    // row pairs the large vertical pass reads
    static constexpr int OFFS = (PL - PS) / 2;            // first pair of the small pass, relative
    static constexpr int NPS = (PS + RB + RS + 1) / 2;    // row pairs the small vertical pass reads
    static constexpr int GS0 = OFFS / 4;                  // first group the small pass touches
    static constexpr int GS1 = (OFFS + NPS - 1) / 4;      // last group
    static constexpr int NGS = NGL - GS0;                 // ring groups kept for the small kernel
    static constexpr int PRE = cdiv(RB * TWORDS, TW);     // prefetch registers per thread
    static constexpr int HLA = rup(HL, 16);               // TMA tiles must START on a 16-byte boundary (measured: an unaligned
                                                          // inner coordinate raises 'illegal instruction' on sm_100a)
    static constexpr int OFFW = (HLA - HL) / 4;           // first word of a staged row the horizontal pass uses
    static constexpr int BOXW = rup(HLA - HL + TWORDS * 4, 16);   // bytes per staged row (<= 256)
    static constexpr int TSTRIDE = BOXW / 4;              // words per staged row
    static constexpr int TILE_BYTES = RB * BOXW;          // one staged tile (multiple of 128 B)
    static constexpr size_t SMEM = 2 * (size_t)TILE_BYTES + 128 + (size_t)(NGL + NGS) * TW * 16;
};

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * (n - 1) - i;
    return i;
}

template <bool BGR>
__device__ __forceinline__ uint32_t load_px(const uint8_t *row, int x) {
    if constexpr (BGR) {
        const uint8_t *p = row + 3 * (size_t)x;
        return (3735u * p[0] + 19235u * p[1] + 9798u * p[2] + 16384u) >> 15;   // cvtColor BGR2GRAY, MD:114
    } else {
        return row[x];
    }
}

// one staged word = 4 horizontally adjacent gray pixels starting at image column col0
template <bool BGR>
__device__ __forceinline__ uint32_t load_word(const uint8_t *row, int col0, int W, bool aligned4) {
    if (!BGR && aligned4 && col0 >= 0 && col0 + 3 < W)
        return __ldg(reinterpret_cast<const uint32_t *>(row + col0));
    uint32_t v = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) v |= load_px<BGR>(row, reflect101(col0 + b, W)) << (8 * b);
    return v;
}

// ---- TMA (cp.async.bulk.tensor) + mbarrier helpers ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: returns false if the phase never completed (a mis-programmed copy must not hang the GPU)
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity) {
    for (int spin = 0; spin < (1 << 22); ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}
// 8 rows x BOXW bytes of frame z starting at column x, row y -> shared memory, completion on the mbarrier
__device__ __forceinline__ void tma_load_tile(void *dst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

template <int KS, int KL, bool BGR>
__global__ void __launch_bounds__(TW, 4)
blur_area_kernel(const __grid_constant__ CUtensorMap tmap, int use_tma, const uint8_t *__restrict__ frames, int64_t frame_stride,
                 int64_t row_pitch, int H, int W, int WW, VbsSegPlan plan, int strips, int lo, int hi, uint32_t *__restrict__ area_bits,
                 uint32_t *__restrict__ area_count, uint32_t *__restrict__ status) {
    using G = Geo<KS, KL>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t *tiles = reinterpret_cast<uint32_t *>(smem_raw);                 // [2][RB][TSTRIDE] staged input rows (double buffer)
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + 2 * G::TILE_BYTES);   // [2] TMA completion barriers
    uint4 *ringL = reinterpret_cast<uint4 *>(smem_raw + 2 * G::TILE_BYTES + 128);  // [NGL][TW]
    uint4 *ringS = ringL + G::NGL * TW;                                       // [NGS][TW]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // CTA -> (frame, strip, row segment): whole-height items first, the tail of the grid in row segments (VbsSegPlan)
    int item = blockIdx.x, ys = 0, ye = H;
    if (item >= plan.n_full) {
        const int j = item - plan.n_full, q = j / plan.vsegs;
        item = plan.n_full + q;
        ys = (j - q * plan.vsegs) * plan.seg_rows;
        ye = min(H, ys + plan.seg_rows);
    }
    const int f = item / strips;
    const int x0 = (item - f * strips) * TW;
    const uint8_t *fbase = frames + (size_t)f * frame_stride;
    const bool aligned4 = ((reinterpret_cast<uintptr_t>(fbase) | (uintptr_t)row_pitch) & 3) == 0;
    const int nk = (ye - ys + RB - 1) / RB;
    const int nsteps = nk + G::LEAD;
    uint32_t count = 0;

    // loader: the (row-in-step, word) slots a thread fills never change, so their column and the
    // "plain 32-bit load" test are hoisted; per step only the source row moves (REFLECT_101 outside [0,H))
    uint32_t pre[G::PRE];
    int lrow[G::PRE], lcol[G::PRE], lslot[G::PRE];
    bool lfast[G::PRE];
#pragma unroll
    for (int i = 0; i < G::PRE; ++i) {
        const int wi = tid + i * TW;
        lrow[i] = wi / G::TWORDS;
        lcol[i] = x0 - G::HL + 4 * (wi - lrow[i] * G::TWORDS);
        lslot[i] = lrow[i] * G::TSTRIDE + G::OFFW + (wi - lrow[i] * G::TWORDS);
        lfast[i] = !BGR && aligned4 && lcol[i] >= 0 && lcol[i] + 3 < W;
    }
    auto fetch = [&](int m) {
        const int p0 = ys - G::PL + RB * m;
#pragma unroll
        for (int i = 0; i < G::PRE; ++i) {
            if (tid + i * TW < RB * G::TWORDS) {
                const int p = p0 + lrow[i];
                const int pr = (p >= 0 && p < H) ? p : reflect101(p, H);
                const uint8_t *row = fbase + (size_t)pr * row_pitch;
                if (lfast[i]) pre[i] = __ldg(reinterpret_cast<const uint32_t *>(row + lcol[i]));
                else pre[i] = load_word<BGR>(row, lcol[i], W, false);
            }
        }
    };
    auto stash = [&](uint32_t *tile) {
#pragma unroll
        for (int i = 0; i < G::PRE; ++i)
            if (tid + i * TW < RB * G::TWORDS) tile[lslot[i]] = pre[i];
    };
    // A step's tile comes through TMA when it lies wholly inside the image (the tensor map fills
    // out-of-bounds bytes with zeros, the blur needs REFLECT_101), otherwise through the generic loader.
    const bool x_inside = use_tma && !BGR && x0 - G::HL >= 0 && x0 - G::HL + G::TWORDS * 4 <= W;
    auto by_tma = [&](int m) -> bool {
        const int p0 = ys - G::PL + RB * m;
        return x_inside && p0 >= 0 && p0 + RB <= H;
    };
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t phase = 0;                              // bit b: parity of the next completion of barrier b
    auto issue = [&](int m) {                        // start loading the tile of step m into buffer m & 1
        if (by_tma(m)) {
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of this buffer are done
                mbar_expect_tx(&mbar[m & 1], G::TILE_BYTES);
                tma_load_tile(tiles + (m & 1) * (G::TILE_BYTES / 4), &tmap, x0 - G::HLA, ys - G::PL + RB * m, f, &mbar[m & 1]);
            }
        } else {
            fetch(m);
        }
    };

    issue(0);
    int gl = 0, gs = 0;          // ring group written by horizontal step m: m % NGL, m % NGS
    for (int m = 0; m < nsteps; ++m) {
        uint32_t *tile = tiles + (m & 1) * (G::TILE_BYTES / 4);
        const bool tma_now = by_tma(m);
        if (!tma_now) stash(tile);
        __syncthreads();                                // tile visible; the vertical pass of the previous step is done
        if (tma_now) {
            if (!mbar_wait(&mbar[m & 1], (phase >> (m & 1)) & 1u) && tid == 0) atomicOr(status, VBS_DEV_TMA_TIMEOUT);
            phase ^= 1u << (m & 1);
        }
        if (m + 1 < nsteps) issue(m + 1);               // the next tile flies during the math (its buffer was last read in step m-1)

        // ---- horizontal passes: warp = row pair, lane = pixel quad ------------------------------
        {
            uint32_t outL[4], outS[4];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const uint32_t *trow = tile + (2 * warp + half) * G::TSTRIDE + G::OFFW + lane;
                uint32_t x[G::NW];
#pragma unroll
                for (int w = 0; w < G::NW; ++w) x[w] = trow[w];
                uint32_t aL[4] = {0, 0, 0, 0}, aS[4] = {0, 0, 0, 0};
                static_for<0, 4>([&](auto S_) {
                    constexpr int s = decltype(S_)::value;
                    static_for<0, G::NW>([&](auto W_) {
                        constexpr int w = decltype(W_)::value;
                        constexpr uint32_t wl = pack4<KL>(4 * w - (G::HL - G::RL) - s);
                        constexpr uint32_t ws = pack4<KS>(4 * w - (G::HL - G::RS) - s);
                        if constexpr (wl != 0) aL[s] = __dp4a(x[w], wl, aL[s]);
                        if constexpr (ws != 0) aS[s] = __dp4a(x[w], ws, aS[s]);
                    });
                });
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    if (half == 0) { outL[s] = aL[s]; outS[s] = aS[s]; }
                    else { outL[s] |= aL[s] << 16; outS[s] |= aS[s] << 16; }
                }
            }
            // ring element (group, column ^ swz).e[warp]: the column index is XOR-swizzled with two bits of
            // the quad number, so the 32 lanes of one store spread over 8 banks (4-way instead of 16-way
            // conflict) while the vertical pass still reads one conflict-free LDS.128 per group
            uint32_t *dstL = reinterpret_cast<uint32_t *>(ringL + gl * TW) + warp;
            uint32_t *dstS = reinterpret_cast<uint32_t *>(ringS + gs * TW) + warp;
            const int swz = (lane >> 1) & 3;
#pragma unroll
            for (int sft = 0; sft < 4; ++sft) {
                const int c = (4 * lane + sft) ^ swz;
                dstL[c * 4] = outL[sft];
                dstS[c * 4] = outS[sft];
            }
        }
        __syncthreads();

        // ---- vertical passes: thread = column, 8 output rows -----------------------------------
        if (m >= G::LEAD) {
            const int k = m - G::LEAD;
            const int yb = ys + RB * k;
            uint32_t accL[RB], accS[RB];
#pragma unroll
            for (int r = 0; r < RB; ++r) { accL[r] = 32768u; accS[r] = 32768u; }
            // oldest live group of the large ring is (gl + 1) % NGL (written at step m - LEAD = k)
            int g0 = gl + 1; if (g0 >= G::NGL) g0 -= G::NGL;
            const int vcol = tid ^ ((tid >> 3) & 3);                  // same swizzle as the stores
            const uint4 *bL0 = ringL + g0 * TW + vcol, *bL1 = bL0 - G::NGL * TW;
            const int wrapL = G::NGL - g0;                            // first group index that wraps
            static_for<0, G::NGL>([&](auto G_) {
                constexpr int g = decltype(G_)::value;
                if constexpr (4 * g < G::NPL) {
                    const uint4 v = (g < wrapL ? bL0 : bL1)[g * TW];
                    static_for<0, RB>([&](auto R_) {
                        constexpr int r = decltype(R_)::value;
                        constexpr uint32_t w01 = pack4<KL>(2 * (4 * g) - r - G::PL + G::RL);
                        constexpr uint32_t w23 = pack4<KL>(2 * (4 * g + 2) - r - G::PL + G::RL);
                        if constexpr ((w01 & 0xffffu) != 0) accL[r] = __dp2a_lo(v.x, w01, accL[r]);
                        if constexpr ((w01 >> 16) != 0) accL[r] = __dp2a_hi(v.y, w01, accL[r]);
                        if constexpr ((w23 & 0xffffu) != 0) accL[r] = __dp2a_lo(v.z, w23, accL[r]);
                        if constexpr ((w23 >> 16) != 0) accL[r] = __dp2a_hi(v.w, w23, accL[r]);
                    });
                }
            });
            // small ring holds groups k+GS0 .. m; group written at step j sits in slot j % NGS
            int s0 = gs + 1 + 0; if (s0 >= G::NGS) s0 -= G::NGS;       // slot of step m - NGS + 1 = k + GS0
            const uint4 *bS0 = ringS + s0 * TW + vcol, *bS1 = bS0 - G::NGS * TW;
            const int wrapS = G::NGS - s0;
            static_for<G::GS0, G::GS1 + 1>([&](auto G_) {
                constexpr int g = decltype(G_)::value;                 // group index relative to step k
                const uint4 v = ((g - G::GS0) < wrapS ? bS0 : bS1)[(g - G::GS0) * TW];
                static_for<0, RB>([&](auto R_) {
                    constexpr int r = decltype(R_)::value;
                    // pair index relative to the small pass start: i = 4g + e - OFFS; tap = 2i - r - PS + RS
                    constexpr uint32_t w01 = pack4<KS>(2 * (4 * g - G::OFFS) - r - G::PS + G::RS);
                    constexpr uint32_t w23 = pack4<KS>(2 * (4 * g + 2 - G::OFFS) - r - G::PS + G::RS);
                    if constexpr ((w01 & 0xffffu) != 0) accS[r] = __dp2a_lo(v.x, w01, accS[r]);
                    if constexpr ((w01 >> 16) != 0) accS[r] = __dp2a_hi(v.y, w01, accS[r]);
                    if constexpr ((w23 & 0xffffu) != 0) accS[r] = __dp2a_lo(v.z, w23, accS[r]);
                    if constexpr ((w23 >> 16) != 0) accS[r] = __dp2a_hi(v.w, w23, accS[r]);
                });
            });
            // round, wrapping DoG, inRange, 32 pixels per ballot word
            const bool col_ok = (x0 + tid) < W;
            uint32_t myword = 0;
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const uint32_t bl = accL[r] >> 16, bs = accS[r] >> 16;
                const uint32_t dog = (bl - bs + 15u) & 255u;             // uint8 wrap, MD:128
                const bool in = col_ok && dog >= (uint32_t)lo && dog <= (uint32_t)hi;
                const uint32_t word = __ballot_sync(0xffffffffu, in);
                if (lane == r) myword = word;
            }
            const int wx = (x0 >> 5) + warp;
            if (lane < RB && yb + lane < ye && wx < WW) {
                area_bits[((size_t)f * H + (yb + lane)) * WW + wx] = myword;
                count += __popc(myword);
            }
        }
        if (++gl == G::NGL) gl = 0;
        if (++gs == G::NGS) gs = 0;
    }
    // per-frame popcount -> mean of area_mask for the NCC (MD:153)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
    if (lane == 0 && count) atomicAdd(area_count + f, count);
}

// ---- K1, column-sum form of the large vertical pass (default for gray frames of the > 480 branch) --------------
// The 8.8 fixed-point taps of the 101-tap kernel are small integers (0..6) that change by at most one from tap to tap,
// so with the running column sum P[q] = sum_{q' <= q} h[q'] of the horizontal results
//     sum_j k[j] h[q + j] = sum_t (k[t] - k[t+1]) P[q + t]
// has 56 terms of weight +-1 instead of 101 multiply-adds: per 8 output rows of a column 233 integer adds (two thirds of
// them three-input, after common sub-sums) instead of 404 IDP.2A (fma pipe, half rate); the compiler spreads the adds over
// the ALU pipe (IADD3) and, for a third of them, the fma pipe (IMAD.IADD).  All sums are taken mod 2^32 and the true value is below
// 2^24: exact.  The column sums are private to the thread that owns the column.  The ring holds 32-bit sums (56 KB
// instead of 28 KB of u16 pairs), so two CTAs are resident per SM instead of four, and each pipe gets its own warps:
//   warps 0-3  (warp = row pair, lane = pixel quad) load the tiles and run both horizontal passes of step i (IDP.4A),
//   warps 4-7  (thread = column) take step i - 1: its 8 rows join the column sums, then the large vertical pass as adds;
//              the eight sums of each column cross to the last role through shared memory,
//   warps 8-11 (thread = column) take step i - 2: small vertical pass (IDP.2A), rounding, DoG, inRange, ballot.
// One CTA barrier per step; the large horizontal results cross through a double-buffered stage, the small ring has two
// slots more than the last role reads so the producers' next group never lands on a live one.
// Measured (256 1080p frames): 3.83 -> 3.43 ms.  The kernel executes about as many instructions as blur_area_kernel
// (5.5 vs 5.7 warp instructions per pixel: the adds, the wider ring loads and the hand-over replace the dot products
// one for one) but spreads them over both pipes; it issues on 73 % of the cycles (ncu: profiles/r02_ncu_full_blur_cs_batch64.txt),
// blur_area_kernel on 65 % with the fma pipe at 85 % of its IDP rate.  Tried on the way: the same arithmetic in the
// four-warp CTA of blur_area_kernel with the adds written between the dot products (ptxas schedules them after the dot
// products anyway: 4.08 ms), two roles instead of three (3.53 ms), and the small horizontal pass moved from the first to
// the third role to even out the roles (3.44 against 3.46 ms: the sum of the work, not its split, sets the pace).
template <int KL> struct DTaps {                                    // D[t] = k[t] - k[t+1], t = -1 .. KL-1
    static __host__ __device__ constexpr int at(int t) { return Taps<KL>::at(t) - Taps<KL>::at(t + 1); }
    static __host__ __device__ constexpr bool unit() {
        for (int t = -1; t < KL; ++t)
            if (at(t) > 1 || at(t) < -1) return false;
        return at(-1) == 0;
    }
    // e-index (0..7) of the n-th non-zero term of ring group g for output row r, -1 when there is none
    static __host__ __device__ constexpr int term(int g, int r, int n) {
        int c = 0;
        for (int e = 0; e < RB; ++e)
            if (at(RB * g + e - r) != 0) {
                if (c == n) return e;
                ++c;
            }
        return -1;
    }
};

// acc[r] += sum_e D[8g + e - r] * p[e] for the 8 output rows, two terms per add
template <int KL, int g>
__device__ __forceinline__ void psum_group(uint32_t (&acc)[RB], const uint32_t (&p)[RB]) {
    static_for<0, RB>([&](auto R_) {
        constexpr int r = decltype(R_)::value;
        static_for<0, RB / 2>([&](auto N_) {
            constexpr int n = decltype(N_)::value;
            constexpr int e1 = DTaps<KL>::term(g, r, 2 * n), e2 = DTaps<KL>::term(g, r, 2 * n + 1);
            if constexpr (e1 >= 0 && e2 >= 0) {
                constexpr int s1 = DTaps<KL>::at(RB * g + e1 - r), s2 = DTaps<KL>::at(RB * g + e2 - r);
                if constexpr (s1 > 0 && s2 > 0) acc[r] = acc[r] + p[e1] + p[e2];
                else if constexpr (s1 > 0 && s2 < 0) acc[r] = acc[r] + p[e1] - p[e2];
                else if constexpr (s1 < 0 && s2 > 0) acc[r] = acc[r] - p[e1] + p[e2];
                else acc[r] = acc[r] - p[e1] - p[e2];
            } else if constexpr (e1 >= 0) {
                constexpr int s1 = DTaps<KL>::at(RB * g + e1 - r);
                if constexpr (s1 > 0) acc[r] += p[e1];
                else acc[r] -= p[e1];
            }
        });
    });
}

template <int KS, int KL> struct GeoCS : Geo<KS, KL> {
    using G = Geo<KS, KL>;
    static constexpr int NT = 2;                                                 // tile buffers
    static constexpr int NGSW = G::NGS + 2;
    static constexpr int OFFSLOT = (NGSW - (G::LEAD - G::GS0) % NGSW) % NGSW;    // slot of step m - LEAD + GS0, relative to m % NGSW
    static constexpr size_t SMEM = NT * (size_t)G::TILE_BYTES + 128 + 2 * (size_t)TW * 16 + 4 * (size_t)TW * 16 + (size_t)G::NGL * 2 * TW * 16 + (size_t)NGSW * TW * 16;
};

__device__ __forceinline__ void role_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the 4 producer warps only

template <int KS, int KL>
__global__ void __launch_bounds__(3 * TW, 2)
blur_area_cs_kernel(const __grid_constant__ CUtensorMap tmap, int use_tma, const uint8_t *__restrict__ frames, int64_t frame_stride,
                    int64_t row_pitch, int H, int W, int WW, VbsSegPlan plan, int strips, int lo, int hi, uint32_t *__restrict__ area_bits,
                    uint32_t *__restrict__ area_count, uint32_t *__restrict__ status) {
    using G = GeoCS<KS, KL>;
    static_assert(DTaps<KL>::unit(), "the column-sum vertical pass needs taps that change by at most one and start at zero");
    static_assert(G::PL == G::RL, "ring row q of output row r and tap j is q = r + j");
    static_assert(TW == 128, "bar.sync 1, 128 names the producer warps");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t *tiles = reinterpret_cast<uint32_t *>(smem_raw);                 // [NT][RB][TSTRIDE] staged input rows
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + G::NT * G::TILE_BYTES);   // [NT] TMA completion barriers
    uint4 *stage = reinterpret_cast<uint4 *>(smem_raw + G::NT * G::TILE_BYTES + 128);  // [2][TW] large horizontal results of a step
    uint4 *sumsL = stage + 2 * TW;                                            // [2][2][TW] large vertical sums of a step, rows 0-3 / 4-7
    uint4 *ringP = sumsL + 4 * TW;                                            // [NGL][2][TW] column sums, rows 0-3 / 4-7 of a group
    uint4 *ringS = ringP + G::NGL * 2 * TW;                                   // [NGSW][TW]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool producer = warp < 4, summer = warp < 8;
    const int tid = threadIdx.x & (TW - 1);                                   // index within the role
    int item = blockIdx.x, ys = 0, ye = H;
    if (item >= plan.n_full) {
        const int j = item - plan.n_full, q = j / plan.vsegs;
        item = plan.n_full + q;
        ys = (j - q * plan.vsegs) * plan.seg_rows;
        ye = min(H, ys + plan.seg_rows);
    }
    const int f = item / strips;
    const int x0 = (item - f * strips) * TW;
    const int nk = (ye - ys + RB - 1) / RB;
    const int nsteps = nk + G::LEAD;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < G::NT; ++b) mbar_init(&mbar[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (producer) {
        // =========================== producers: tiles + horizontal passes ===========================
        const uint8_t *fbase = frames + (size_t)f * frame_stride;
        const bool aligned4 = ((reinterpret_cast<uintptr_t>(fbase) | (uintptr_t)row_pitch) & 3) == 0;
        uint32_t pre[G::PRE];
        int lrow[G::PRE], lcol[G::PRE], lslot[G::PRE];
        bool lfast[G::PRE];
#pragma unroll
        for (int i = 0; i < G::PRE; ++i) {
            const int wi = tid + i * TW;
            lrow[i] = wi / G::TWORDS;
            lcol[i] = x0 - G::HL + 4 * (wi - lrow[i] * G::TWORDS);
            lslot[i] = lrow[i] * G::TSTRIDE + G::OFFW + (wi - lrow[i] * G::TWORDS);
            lfast[i] = aligned4 && lcol[i] >= 0 && lcol[i] + 3 < W;
        }
        auto fetch = [&](int m) {
            const int p0 = ys - G::PL + RB * m;
#pragma unroll
            for (int i = 0; i < G::PRE; ++i) {
                if (tid + i * TW < RB * G::TWORDS) {
                    const int p = p0 + lrow[i];
                    const int pr = (p >= 0 && p < H) ? p : reflect101(p, H);
                    const uint8_t *row = fbase + (size_t)pr * row_pitch;
                    if (lfast[i]) pre[i] = __ldg(reinterpret_cast<const uint32_t *>(row + lcol[i]));
                    else pre[i] = load_word<false>(row, lcol[i], W, false);
                }
            }
        };
        auto stash = [&](uint32_t *tile) {
#pragma unroll
            for (int i = 0; i < G::PRE; ++i)
                if (tid + i * TW < RB * G::TWORDS) tile[lslot[i]] = pre[i];
        };
        const bool x_inside = use_tma && x0 - G::HL >= 0 && x0 - G::HL + G::TWORDS * 4 <= W;
        auto by_tma = [&](int m) -> bool {
            const int p0 = ys - G::PL + RB * m;
            return x_inside && p0 >= 0 && p0 + RB <= H;
        };
        uint32_t phase = 0;
        auto issue = [&](int m) {
            if (by_tma(m)) {
                if (tid == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx(&mbar[m % G::NT], G::TILE_BYTES);
                    tma_load_tile(tiles + (m % G::NT) * (G::TILE_BYTES / 4), &tmap, x0 - G::HLA, ys - G::PL + RB * m, f, &mbar[m % G::NT]);
                }
            } else {
                fetch(m);
            }
        };
        issue(0);
        int gs = 0;                                          // small-ring slot of step i: i % NGSW
        for (int i = 0; i <= nsteps + 1; ++i) {
            if (i < nsteps) {
                const int tb = i % G::NT;
                uint32_t *tile = tiles + tb * (G::TILE_BYTES / 4);
                const bool tma_now = by_tma(i);
                if (!tma_now) stash(tile);
                role_barrier();                              // tile visible to the four producer warps
                if (tma_now) {
                    if (!mbar_wait(&mbar[tb], (phase >> tb) & 1u) && tid == 0) atomicOr(status, VBS_DEV_TMA_TIMEOUT);
                    phase ^= 1u << tb;
                }
                if (i + 1 < nsteps) issue(i + 1);            // its buffer was last read an iteration ago, before the CTA barrier
                uint32_t outL[4], outS[4];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const uint32_t *trow = tile + (2 * warp + half) * G::TSTRIDE + G::OFFW + lane;
                    uint32_t x[G::NW];
#pragma unroll
                    for (int w = 0; w < G::NW; ++w) x[w] = trow[w];
                    uint32_t aL[4] = {0, 0, 0, 0}, aS[4] = {0, 0, 0, 0};
                    static_for<0, 4>([&](auto S_) {
                        constexpr int s = decltype(S_)::value;
                        static_for<0, G::NW>([&](auto W_) {
                            constexpr int w = decltype(W_)::value;
                            constexpr uint32_t wl = pack4<KL>(4 * w - (G::HL - G::RL) - s);
                            constexpr uint32_t ws = pack4<KS>(4 * w - (G::HL - G::RS) - s);
                            if constexpr (wl != 0) aL[s] = __dp4a(x[w], wl, aL[s]);
                            if constexpr (ws != 0) aS[s] = __dp4a(x[w], ws, aS[s]);
                        });
                    });
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        if (half == 0) { outL[s] = aL[s]; outS[s] = aS[s]; }
                        else { outL[s] |= aL[s] << 16; outS[s] |= aS[s] << 16; }
                    }
                }
                uint32_t *dstL = reinterpret_cast<uint32_t *>(stage + (i & 1) * TW) + warp;
                uint32_t *dstS = reinterpret_cast<uint32_t *>(ringS + gs * TW) + warp;
                const int swz = (lane >> 1) & 3;
#pragma unroll
                for (int sft = 0; sft < 4; ++sft) {
                    const int c = (4 * lane + sft) ^ swz;
                    dstL[c * 4] = outL[sft];
                    dstS[c * 4] = outS[sft];
                }
                if (++gs == G::NGSW) gs = 0;
            }
            __syncthreads();
        }
    } else if (summer) {
        // =========================== column sums + large vertical pass (step i - 1) ===========================
        uint32_t prun = 0;                                   // column sum up to the last row of the previous step
        int gl = 0;                                          // slot of step m: m % NGL
        const int vcol = tid ^ ((tid >> 3) & 3);             // swizzle of the producers' stores
        __syncthreads();                                     // iteration 0: nothing to consume yet
        for (int m = 0; m < nsteps; ++m) {
            uint32_t pn[RB];
            {
                const uint4 v = stage[(m & 1) * TW + vcol];
                pn[0] = prun + (v.x & 0xffffu);  pn[1] = pn[0] + (v.x >> 16);
                pn[2] = pn[1] + (v.y & 0xffffu); pn[3] = pn[2] + (v.y >> 16);
                pn[4] = pn[3] + (v.z & 0xffffu); pn[5] = pn[4] + (v.z >> 16);
                pn[6] = pn[5] + (v.w & 0xffffu); pn[7] = pn[6] + (v.w >> 16);
                prun = pn[7];
                uint4 *dst = ringP + gl * 2 * TW + tid;
                dst[0] = make_uint4(pn[0], pn[1], pn[2], pn[3]);
                dst[TW] = make_uint4(pn[4], pn[5], pn[6], pn[7]);
            }
            if (m >= G::LEAD) {
                uint32_t accL[RB];
#pragma unroll
                for (int r = 0; r < RB; ++r) accL[r] = 32768u;
                // oldest live group of the sum ring is (gl + 1) % NGL (step m - LEAD); the newest is still in registers
                int g0 = gl + 1; if (g0 >= G::NGL) g0 -= G::NGL;
                const uint4 *bP0 = ringP + g0 * 2 * TW + tid, *bP1 = bP0 - G::NGL * 2 * TW;
                const int wrapL = G::NGL - g0;
                static_for<0, G::NGL - 1>([&](auto G_) {
                    constexpr int g = decltype(G_)::value;
                    const uint4 *src = (g < wrapL ? bP0 : bP1) + g * 2 * TW;
                    const uint4 a = src[0], b = src[TW];
                    const uint32_t p[RB] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                    psum_group<KL, g>(accL, p);
                });
                psum_group<KL, G::NGL - 1>(accL, pn);
                uint4 *dst = sumsL + (m & 1) * 2 * TW + tid;
                dst[0] = make_uint4(accL[0], accL[1], accL[2], accL[3]);
                dst[TW] = make_uint4(accL[4], accL[5], accL[6], accL[7]);
            }
            if (++gl == G::NGL) gl = 0;
            __syncthreads();
        }
        __syncthreads();                                     // iteration nsteps + 1
    } else {
        // =========================== small vertical pass + decision (step i - 2) ===========================
        uint32_t count = 0;
        int gs = 0;                                          // slot of step m: m % NGSW
        const int vcol = tid ^ ((tid >> 3) & 3);
        const bool col_ok = (x0 + tid) < W;
        const int wx = (x0 >> 5) + (warp - 8);
        __syncthreads();                                     // iterations 0 and 1: nothing to consume yet
        __syncthreads();
        for (int m = 0; m < nsteps; ++m) {                   // iteration m + 2
            if (m >= G::LEAD) {
                const int k = m - G::LEAD;
                const int yb = ys + RB * k;
                uint32_t accS[RB];
#pragma unroll
                for (int r = 0; r < RB; ++r) accS[r] = 32768u;
                int s0 = gs + G::OFFSLOT; if (s0 >= G::NGSW) s0 -= G::NGSW;      // slot of step k + GS0
                const uint4 *bS0 = ringS + s0 * TW + vcol, *bS1 = bS0 - G::NGSW * TW;
                const int wrapS = G::NGSW - s0;
                static_for<G::GS0, G::GS1 + 1>([&](auto G_) {
                    constexpr int g = decltype(G_)::value;                 // group index relative to step k
                    const uint4 v = ((g - G::GS0) < wrapS ? bS0 : bS1)[(g - G::GS0) * TW];
                    static_for<0, RB>([&](auto R_) {
                        constexpr int r = decltype(R_)::value;
                        constexpr uint32_t w01 = pack4<KS>(2 * (4 * g - G::OFFS) - r - G::PS + G::RS);
                        constexpr uint32_t w23 = pack4<KS>(2 * (4 * g + 2 - G::OFFS) - r - G::PS + G::RS);
                        if constexpr ((w01 & 0xffffu) != 0) accS[r] = __dp2a_lo(v.x, w01, accS[r]);
                        if constexpr ((w01 >> 16) != 0) accS[r] = __dp2a_hi(v.y, w01, accS[r]);
                        if constexpr ((w23 & 0xffffu) != 0) accS[r] = __dp2a_lo(v.z, w23, accS[r]);
                        if constexpr ((w23 >> 16) != 0) accS[r] = __dp2a_hi(v.w, w23, accS[r]);
                    });
                });
                const uint4 *src = sumsL + (m & 1) * 2 * TW + tid;
                const uint4 la = src[0], lb = src[TW];
                const uint32_t accL[RB] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
                uint32_t myword = 0;
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const uint32_t bl = accL[r] >> 16, bs = accS[r] >> 16;
                    const uint32_t dog = (bl - bs + 15u) & 255u;             // uint8 wrap, MD:128
                    const bool in = col_ok && dog >= (uint32_t)lo && dog <= (uint32_t)hi;
                    const uint32_t word = __ballot_sync(0xffffffffu, in);
                    if (lane == r) myword = word;
                }
                if (lane < RB && yb + lane < ye && wx < WW) {
                    area_bits[((size_t)f * H + (yb + lane)) * WW + wx] = myword;
                    count += __popc(myword);
                }
            }
            if (++gs == G::NGSW) gs = 0;
            __syncthreads();
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
        if (lane == 0 && count) atomicAdd(area_count + f, count);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
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

// 3-D uint8 tensor map {W, H, batch} with an {BOXW, 8, 1} box; returns false when the frames cannot be
// described to the TMA unit (unaligned base / pitches: e.g. a crop view) - the kernel then loads generically
bool make_frame_map(CUtensorMap *map, const uint8_t *frames, int W, int H, int batch, int64_t row_pitch, int64_t frame_stride, int boxw) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    if ((reinterpret_cast<uintptr_t>(frames) & 15) || (row_pitch & 15) || (frame_stride & 15) || W < boxw) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)batch};
    const cuuint64_t strides[2] = {(cuuint64_t)row_pitch, (cuuint64_t)frame_stride};
    const cuuint32_t box[3] = {(cuuint32_t)boxw, (cuuint32_t)RB, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t *>(frames), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int KS, int KL>
cudaError_t launch(vbs_ctx *ctx, const uint8_t *frames, int batch, int64_t frame_stride, int64_t row_pitch) {
    using G = Geo<KS, KL>;
    const int strips = (ctx->W + TW - 1) / TW;
    // 4 CTAs per SM are resident; a lead step runs the horizontal passes only (about a third of a full step)
    const VbsSegPlan plan = vbs_seg_plan(ctx->H, (long long)strips * batch, 4 * ctx->sm_count, G::LEAD, RB, 0.35, ctx->seg_plan != 0);
    dim3 grid(plan.ctas), block(TW);
    cudaError_t e;
    CUtensorMap map;
    std::memset(&map, 0, sizeof(map));
    const int use_tma = (ctx->C == 1 && !ctx->no_tma && make_frame_map(&map, frames, ctx->W, ctx->H, batch, row_pitch, frame_stride, G::BOXW)) ? 1 : 0;
    if (ctx->C == 3) {
        auto kern = blur_area_kernel<KS, KL, true>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM)) != cudaSuccess) return e;
        kern<<<grid, block, G::SMEM, ctx->stream>>>(map, 0, frames, frame_stride, row_pitch, ctx->H, ctx->W, ctx->WW, plan, strips,
                                                    ctx->br.lo, ctx->br.hi, ctx->area_bits, ctx->area_count, ctx->d_status);
    } else {
        auto kern = blur_area_kernel<KS, KL, false>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM)) != cudaSuccess) return e;
        kern<<<grid, block, G::SMEM, ctx->stream>>>(map, use_tma, frames, frame_stride, row_pitch, ctx->H, ctx->W, ctx->WW, plan, strips,
                                                    ctx->br.lo, ctx->br.hi, ctx->area_bits, ctx->area_count, ctx->d_status);
    }
    ctx->launches += 1;
    ctx->tma_launches += use_tma;
    return cudaGetLastError();
}

// column-sum kernel: producer / summer / decider warps, 2 CTAs per SM; during a lead step only the producers work
cudaError_t launch_cs(vbs_ctx *ctx, const uint8_t *frames, int batch, int64_t frame_stride, int64_t row_pitch) {
    using G = GeoCS<39, 101>;
    const int strips = (ctx->W + TW - 1) / TW;
    const VbsSegPlan plan = vbs_seg_plan(ctx->H, (long long)strips * batch, 2 * ctx->sm_count, G::LEAD, RB, 0.75, ctx->seg_plan != 0);
    dim3 grid(plan.ctas), block(3 * TW);
    cudaError_t e;
    CUtensorMap map;
    std::memset(&map, 0, sizeof(map));
    const int use_tma = (!ctx->no_tma && make_frame_map(&map, frames, ctx->W, ctx->H, batch, row_pitch, frame_stride, G::BOXW)) ? 1 : 0;
    auto kern = blur_area_cs_kernel<39, 101>;
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM)) != cudaSuccess) return e;
    kern<<<grid, block, G::SMEM, ctx->stream>>>(map, use_tma, frames, frame_stride, row_pitch, ctx->H, ctx->W, ctx->WW, plan, strips,
                                                ctx->br.lo, ctx->br.hi, ctx->area_bits, ctx->area_count, ctx->d_status);
    ctx->launches += 1;
    ctx->tma_launches += use_tma;
    return cudaGetLastError();
}

template <int K> bool taps_match(const int *ref) {
    for (int j = 0; j < K; ++j)
        if (Taps<K>::at(j) != ref[j]) return false;
    return true;
}

// host recipe of OpenCV's fixed-point Gaussian kernel: float64 taps, error diffusion from the
// outside in (round half to even), the centre takes the remainder so the sum is exactly 256
void host_taps(int ksize, double sigma, int *out) {
    double k[128], sum = 0;
    for (int i = 0; i < ksize; ++i) {
        const double x = i - (ksize - 1) / 2.0;
        k[i] = exp(-(x * x) / (2.0 * sigma * sigma));
        sum += k[i];
    }
    double err = 0;
    int acc = 0;
    for (int i = 0; i < ksize / 2; ++i) {
        const double adj = k[i] / sum * 256.0 + err;
        const int v = (int)nearbyint(adj);
        err = adj - v;
        out[i] = out[ksize - 1 - i] = v;
        acc += v;
    }
    out[ksize / 2] = 256 - 2 * acc;
}

}  // namespace

void vbs_host_taps(int ksize, double sigma, int *out) { host_taps(ksize, sigma, out); }

int vbs_check_taps(std::string &err) {
    int t[128];
    host_taps(39, 8.0, t);    if (!taps_match<39>(t))  { err = "baked 39-tap kernel differs from the fixed-point recipe"; return -1; }
    host_taps(101, 20.0, t);  if (!taps_match<101>(t)) { err = "baked 101-tap kernel differs from the fixed-point recipe"; return -1; }
    host_taps(21, 4.56, t);   if (!taps_match<21>(t))  { err = "baked 21-tap kernel differs from the fixed-point recipe"; return -1; }
    host_taps(35, 11.4, t);   if (!taps_match<35>(t))  { err = "baked 35-tap kernel differs from the fixed-point recipe"; return -1; }
    return 0;
}

cudaError_t vbs_launch_blur(vbs_ctx *ctx, const uint8_t *frames, int batch, int64_t frame_stride, int64_t row_pitch) {
    VbsRange range("vbs:blur");
    cudaError_t e = cudaMemsetAsync(ctx->area_count, 0, sizeof(uint32_t) * batch, ctx->stream);
    if (e != cudaSuccess) return e;
    if (ctx->undist_on) {                       // MD:88-89: lens correction of the (cropped) frame comes first
        if ((e = vbs_launch_remap(ctx, frames, batch, frame_stride, row_pitch, ctx->d_undist)) != cudaSuccess) return e;
        frames = ctx->d_undist;
        row_pitch = (int64_t)ctx->W * ctx->C;
        frame_stride = row_pitch * ctx->H;
    }
    if (ctx->blur_tc) {                         // opt-in: both blurs as int8 GEMMs on the tensor cores (k_blur_tc.cu)
        e = vbs_launch_blur_tc(ctx, frames, batch, frame_stride, row_pitch);
        if (e != cudaErrorNotSupported) return e;
    }                                           // BGR input / frames the TMA unit cannot describe: integer-dot-product kernel
    if (ctx->big && ctx->C == 1 && ctx->blur_variant != 0) return launch_cs(ctx, frames, batch, frame_stride, row_pitch);
    if (ctx->big) return launch<39, 101>(ctx, frames, batch, frame_stride, row_pitch);
    return launch<21, 35>(ctx, frames, batch, frame_stride, row_pitch);
}
