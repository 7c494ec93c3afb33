// k_ncc.cu - K2: bit-packed area mask -> mask = (normxcorr2(gkern, area_mask) > 0.1), bit-packed.
// Replaces MD:132-133 + MD:146-164 (three float64 FFT convolutions in the reference).
//
// Closed form used (pinned in tests/test_oracle_exact.py).  With b = area_mask/255,
// n the 1-D template factor, L its length, window [i-OFF, i-OFF+L-1] per axis, m = mean(b):
//     G  = sum n[a] n[c] b      (separable Gaussian of a BINARY image)
//     S  = sum b                (integer box sum)          A  = window pixels inside the image
//     G1 = sum n[a] n[c] inside                            st2 = (sum n^2)^2 - 1/L^2
//     ncc > 0.1  <=>  G > thr := m G1 + (S - m A)/L^2 + 0.1 sqrt(st2 q)/L,
//                     q = S(L^2-S) - 2 m S (L^2-A) + m^2 A (L^2-A)  (q <= 0 -> mask 0)
// Inside the image (A = L^2, G1 = 1) m cancels: thr is a function of S alone -> one table lookup.
//
// The horizontal pass works on bits: per row, h(x) = sum over run ends of Cn[e-x+OFF] minus the
// same over run starts (Cn = prefix sums of n, float64) - a handful of additions per pixel
// instead of L multiply-adds.  The vertical pass is L float32 FMAs per pixel on a shared-memory
// ring (8 output rows per thread).  float32 is only a filter: any pixel with |G - thr| below a
// rigorous rounding bound is queued and re-decided in float64 with the reference's literal
// formula (a few dozen pixels per frame).
#include <type_traits>
#include <utility>
#include "vbs_ctx.h"

namespace {

template <int B, int E, class F> __device__ __forceinline__ void static_for(F &&f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}
constexpr int cdiv(int a, int b) { return (a + b - 1) / b; }

constexpr int TW = 128;
constexpr int RB = 8;
constexpr int TWP = TW + TW / 8;          // padded column index: c + c/8
constexpr float BAND = 6e-6f;             // > 80 * 2^-24 (FMA chain) + input / table roundings

__constant__ float c_n32[2][96];          // template factor n, float32 (vertical pass weights); [0]: L=33, [1]: L=80

// two independent float32 FMAs in one instruction (sm_100 FFMA2), both IEEE rn:
// acc.x += a.x * w, acc.y += a.y * w: ptxas folds the {w, w} pack into a broadcast operand (SASS `UR.F32`)
__device__ __forceinline__ void ffma2(float2 &acc, const float2 a, const float w) {
    asm("{ .reg .b64 t; mov.b64 t, {%2, %2}; fma.rn.f32x2 %0, %1, t, %0; }"
        : "+l"(reinterpret_cast<unsigned long long &>(acc))
        : "l"(reinterpret_cast<const unsigned long long &>(a)), "f"(w));
}

template <int TL> struct Geo {
    static constexpr int OFF = (TL - 1) - (TL - 1) / 2;   // 40 / 16
    static constexpr int HI = TL - 1 - OFF;               // 39 / 16
    static constexpr int LEAD = cdiv(TL - 1, RB);         // 10 / 4
    static constexpr int NG = LEAD + 1;                   // ring groups of 8 rows
    static constexpr int UL = RB + TL - 1;                // widest union window the prefix table must cover
    static constexpr int CNX = 8 + UL + 9;                // guarded prefix table entries, index d+8
    static constexpr size_t SMEM = (size_t)NG * 2 * TWP * 16 + (size_t)NG * TWP * 8 + (size_t)(CNX + (CNX & 1)) * 8 + 4 * 112 * 4 +
                                   2 * 4 * TW * 4 + 4 * TW * 4;
};

struct NccParams {
    int H, W, WW, seg_rows;
    VbsSegPlan plan; int strips;                   // thread-per-column kernel: CTA -> (frame, strip, row segment)
    double st2, hw;
    const uint32_t *area_bits; const uint32_t *area_count;
    uint32_t *mask_bits;
    const float *thr_lut; const double *cn64;      // cn64[d+8] = Cn[clamp(d,0,TL)]
    const int *cnfix;                              // [4][112] fixed-point copies: cnfix[s][k] = round(2^30 cn64[k+s])
    int2 *recheck; uint32_t *recheck_n; int recheck_cap;
    uint32_t *status;
};

__device__ __forceinline__ uint32_t ld_bits(const uint32_t *row, int wi, int WW) {
    return (wi >= 0 && wi < WW) ? __ldg(row + wi) : 0u;
}

// Decision for a pixel whose window is clipped by the image border (A < L^2 window pixels inside the image).
// Two window states make EVERY term of G - thr proportional to m or to 1 - m (m = mean(b)), so that an absolute
// float32 band would queue the whole border of a nearly empty / nearly full frame; both have closed answers:
//   S == 0 (no area pixel in the window): G = 0 and thr = m (g1 - A/L^2 + 0.1 sqrt(st2 A (L^2-A)) / L) > 0
//           (a window that holds its own centre has g1 >= A/L^2)                       -> mask 0, for any m
//   S == A (every in-image pixel set):    ncc = (g1 - A/L^2) / sqrt(st2 A (1 - A/L^2)), m cancels  -> float64, inline
// Otherwise float32 is enough for the filter: q is evaluated as a sum of NON-NEGATIVE terms
//   q = (L^2-A) (S (1-m)^2 + m^2 (A-S)) + S (A-S)        (= S(L^2-S) - 2 m S (L^2-A) + m^2 A (L^2-A), no cancellation)
// so thr carries ~1e-6 relative error; such pixels use the wider BAND_BORDER before the float64 re-decision.
constexpr float BAND_BORDER = 2e-5f;
struct BorderGeo { int A; double g1; };
template <int TL>
__device__ __forceinline__ BorderGeo border_geo(int y, int x, int H, int W, const double *cn) {
    using G = Geo<TL>;
    const int ylo = max(0, G::OFF - y), yhi = min(TL - 1, H - 1 - y + G::OFF);
    const int xlo = max(0, G::OFF - x), xhi = min(TL - 1, W - 1 - x + G::OFF);
    BorderGeo g;
    g.A = (yhi - ylo + 1) * (xhi - xlo + 1);
    g.g1 = (cn[yhi + 1 + 8] - cn[ylo + 8]) * (cn[xhi + 1 + 8] - cn[xlo + 8]);
    return g;
}
template <int TL>
__device__ __forceinline__ bool border_full_window_on(const BorderGeo &g, double st2, float m) {
    if (!(m < 1.0f)) return false;                         // the whole frame is set: the reference divides 0 by 0 -> 0
    const double L2 = (double)(TL * TL), A = (double)g.A;
    const double num = g.g1 - A / L2, den = sqrt((A - A * A / L2) * st2);
    const double ncc = num / den;
    return isfinite(ncc) && ncc > 0.1;
}
template <int TL>
__device__ __forceinline__ float border_threshold(const BorderGeo &g, float S, float m, float mc, float st2) {
    const float A = (float)g.A, L2 = (float)(TL * TL);
    const float q = (L2 - A) * (S * mc * mc + m * m * (A - S)) + S * (A - S);
    if (!(q > 0.0f)) return INFINITY;
    return m * (float)g.g1 + (S - m * A) * (1.0f / L2) + (0.1f / (float)TL) * sqrtf(st2 * q);
}

// 256 threads per 128-pixel strip, three phases per 8-row step:
//  H  (threads 0..127): one (row, 8-pixel octet) item each.  Run transitions of the 87-bit union
//     window are turned into h(x) with a FIXED-POINT prefix table (2^30 scale, int32 adds are exact
//     and order-free); any 8 consecutive table entries come from two aligned LDS.128 (four shifted
//     copies of the table).  Results go to the float ring (scale folded into the vertical weights).
//  V  (all 256 threads): thread = (column, half); half h accumulates taps [h L/2, (h+1) L/2) of all
//     8 output rows, so it reads only half of the ring; partial sums of the 4 rows the other half
//     decides are swapped through shared memory together with the box sums.
//  D  decision: LUT threshold on the box sum; a CTA-uniform fast path when the whole step is interior.
constexpr int NT = 256;
constexpr int VR = 4;                    // output rows each half decides
constexpr int FIX_SHIFT = 30;            // fixed-point scale of the prefix table
constexpr int FXN = 112;                 // entries per shifted copy of the fixed-point table (>= CNX + 3, multiple of 4)

template <int TL> struct GeoR : Geo<TL> {
    static constexpr int NR = Geo<TL>::NG;                // ring groups (a spare slot would save a barrier but costs the 3rd CTA/SM)
    static constexpr size_t SMEM = (size_t)NR * 2 * TWP * 16 + (size_t)NR * TWP * 8 + (size_t)(Geo<TL>::CNX + (Geo<TL>::CNX & 1)) * 8 +
                                   4 * 112 * 4 + 2 * 4 * TW * 4 + 4 * TW * 4;
};

template <int TL>
__global__ void __launch_bounds__(NT, 3) ncc_mask_kernel(NccParams P) {
    using G = GeoR<TL>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *ringH = reinterpret_cast<float4 *>(smem_raw);                        // [NR*2][TWP]
    uint2 *ringB = reinterpret_cast<uint2 *>(ringH + G::NR * 2 * TWP);           // [NR][TWP]
    double *cn = reinterpret_cast<double *>(ringB + G::NR * TWP);                // [CNX] float64 prefix sums (border formula)
    int4 *fx = reinterpret_cast<int4 *>(cn + G::CNX + (G::CNX & 1));             // [4][FXN/4] shifted fixed-point copies
    float *xbuf = reinterpret_cast<float *>(fx + FXN);                           // [2][VR][TW] partial sums in flight

    const int tid = threadIdx.x, lane = tid & 31;
    const int x0 = blockIdx.x * TW;
    const int ys = blockIdx.y * P.seg_rows;
    const int ye = min(P.H, ys + P.seg_rows);
    const int f = blockIdx.z;
    const int H = P.H, W = P.W, WW = P.WW;
    const uint32_t *abits = P.area_bits + (size_t)f * H * WW;
    const double mfrac64 = (double)P.area_count[f] / P.hw;
    const float mfrac = (float)mfrac64, mcomp = (float)(1.0 - mfrac64);   // mean(area_mask)/255 and its complement
    for (int i = tid; i < G::CNX; i += NT) cn[i] = P.cn64[i];
    for (int i = tid; i < 4 * FXN; i += NT) reinterpret_cast<int *>(fx)[i] = P.cnfix[i];

    const int nk = (ye - ys + RB - 1) / RB;
    const int nsteps = nk + G::LEAD;
    // horizontal role (threads 0..127): row hr of the step, pixel octet ho
    const bool hrole = tid < TW;
    const int hr = (tid >> 4) & 7, ho = tid & 15;
    const int sb = x0 + 8 * ho - G::OFF;                            // first bit of the union window
    const int wi0 = sb >> 5, bo = sb & 31;                          // arithmetic shift: floor
    const int vcol = tid & (TW - 1), vhalf = tid >> 7;              // vertical role
    const int cp = vcol + (vcol >> 3);
    const bool strip_interior = x0 >= G::OFF && x0 + TW - 1 + G::HI < W;

    uint32_t pw[4] = {0, 0, 0, 0};
    auto fetch = [&](int m) {
        const int p = ys - G::OFF + RB * m + hr;
        if (p >= 0 && p < H) {
            const uint32_t *row = abits + (size_t)p * WW;
#pragma unroll
            for (int i = 0; i < 4; ++i) pw[i] = ld_bits(row, wi0 + i, WW);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) pw[i] = 0u;
        }
    };
    int s_prev = 0; uint32_t hb_m1 = 0;         // S of the previous output row and the box row that left the window

    if (hrole) fetch(0);
    __syncthreads();
    int gw = 0;                                  // ring slot written by step m (m % NR)
    for (int m = 0; m < nsteps; ++m) {
        // ---- H: horizontal pass on bits ------------------------------------------------------------
        if (hrole) {
            uint32_t U0 = __funnelshift_r(pw[0], pw[1], bo);
            uint32_t U1 = __funnelshift_r(pw[1], pw[2], bo);
            uint32_t U2 = __funnelshift_r(pw[2], pw[3], bo);
            if (m + 1 < nsteps) fetch(m + 1);
            if constexpr (G::UL <= 64) { U2 = 0; U1 &= (G::UL == 64) ? 0xffffffffu : ((1u << (G::UL - 32)) - 1u); }
            else { U2 &= (1u << (G::UL - 64)) - 1u; }
            // transitions: bit t set when bit(t) != bit(t-1)  (bit(-1) = 0, bit(UL) = 0)
            const uint32_t T0 = U0 ^ (U0 << 1);
            const uint32_t T1 = U1 ^ __funnelshift_l(U0, U1, 1);
            const uint32_t T2 = U2 ^ __funnelshift_l(U1, U2, 1);
            int acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0;
            auto consume = [&](uint32_t T, uint32_t Uw, int base) {
                while (T) {
                    const int b = __ffs(T) - 1;
                    T &= T - 1;
                    const int sgn = ((Uw >> b) & 1u) ? -1 : 1;       // run start: -Cn, run end: +Cn
                    const int i0 = base + b + 1;                     // table index of d = t - 7 (pixel j = 7)
                    const int4 *src = fx + (i0 & 3) * (FXN / 4) + (i0 >> 2);
                    const int4 lo = src[0], hi = src[1];
                    acc[7] += sgn * lo.x; acc[6] += sgn * lo.y; acc[5] += sgn * lo.z; acc[4] += sgn * lo.w;
                    acc[3] += sgn * hi.x; acc[2] += sgn * hi.y; acc[1] += sgn * hi.z; acc[0] += sgn * hi.w;
                }
            };
            consume(T0, U0, 0);
            consume(T1, U1, 32);
            if constexpr (G::UL > 64) consume(T2, U2, 64);
            auto bit = [&](int t) -> uint32_t {
                return t < 32 ? (U0 >> t) & 1u : t < 64 ? (U1 >> (t - 32)) & 1u : (U2 >> (t - 64)) & 1u;
            };
            uint32_t hb[8];
            if constexpr (TL >= 64) hb[0] = __popc(U0) + __popc(U1) + __popc(U2 & ((1u << (TL - 64)) - 1u));
            else hb[0] = __popc(U0) + __popc(U1 & ((1u << (TL - 32)) - 1u));
#pragma unroll
            for (int j = 1; j < 8; ++j) hb[j] = hb[j - 1] + bit(j - 1 + TL) - bit(j - 1);
            // ring stores: H as float4 units [group*2 + hr/4][col + col/8].f[hr%4]; box rows as bytes
            float *dstH = reinterpret_cast<float *>(ringH + (gw * 2 + (hr >> 2)) * TWP) + (hr & 3);
            unsigned char *dstB = reinterpret_cast<unsigned char *>(ringB + gw * TWP) + hr;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = 9 * ho + j;                            // (8 ho + j) + (8 ho + j) / 8
                dstH[c * 4] = (float)acc[j];                         // 2^30 h; the scale lives in the vertical weights
                dstB[c * 8] = (unsigned char)hb[j];
            }
        }
        __syncthreads();

        // ---- V: vertical pass ------------------------------------------------------------------------
        if (m >= G::LEAD) {
            const int k = m - G::LEAD;
            const int yb = ys + RB * k;
            int g0 = gw + 1; if (g0 >= G::NR) g0 -= G::NR;           // slot of step k = m - LEAD (NR = LEAD + 1)
            // box sums of the 8 rows (both halves run the same integer chain: each needs them to decide
            // whether the window of its warp is empty, and it saves an exchange)
            int S[VR];
            bool empty;
            {
                auto box_at = [&](auto I_) -> int {
                    constexpr int idx = decltype(I_)::value;
                    int gi = g0 + idx / 8; if (gi >= G::NR) gi -= G::NR;
                    const uint2 b = ringB[gi * TWP + cp];
                    const uint32_t w = (idx % 8) < 4 ? b.x : b.y;
                    return (int)((w >> (8 * (idx % 4))) & 255u);
                };
                int S8[RB];
                if (k == 0) {                   // first step of the segment: sum the whole window once (rolled: code size)
                    int s = 0;
#pragma unroll 1
                    for (int t = 0; t < TL; ++t) {
                        int gi = g0 + (t >> 3); if (gi >= G::NR) gi -= G::NR;
                        s += reinterpret_cast<const unsigned char *>(ringB + gi * TWP + cp)[t & 7];
                    }
                    S8[0] = s;
                } else {
                    S8[0] = s_prev + box_at(std::integral_constant<int, TL - 1>{}) - (int)hb_m1;
                }
                static_for<1, RB>([&](auto R_) {
                    constexpr int r = decltype(R_)::value;
                    S8[r] = S8[r - 1] + box_at(std::integral_constant<int, r + TL - 1>{}) - box_at(std::integral_constant<int, r - 1>{});
                });
                s_prev = S8[RB - 1]; hb_m1 = (uint32_t)box_at(std::integral_constant<int, RB - 1>{});
                int any = 0;
#pragma unroll
                for (int r = 0; r < RB; ++r) any |= S8[r];
#pragma unroll
                for (int r = 0; r < VR; ++r) S[r] = vhalf ? S8[VR + r] : S8[r];
                // S == 0 <=> no area pixel in the whole L x L window <=> G == 0 and mask == 0: when that holds
                // for all 8 rows of all 32 columns of the warp, the L-tap column sums are skipped (exact)
                empty = !__any_sync(0xffffffffu, any != 0);
            }
            // partial Gaussian column sums: taps [A0, A1) of all 8 rows; ring rows A0 .. A1+6
            // Packed form: ring rows come as float4 = two aligned pairs (h_t, h_t+1), t even.  With the accumulators
            // paired as (row r, row r+1) one FFMA2 does  part[r] += n[t-r] h_t  and  part[r+1] += n[t-r] h_t+1  (the same
            // tap).  Pairs with r even (E) cover the even taps of even rows / odd rows; pairs with r odd (O) plus the two
            // singles o0, o7 cover the rest.  9 instructions per pair of ring rows instead of 16 FFMAs.
            float part[RB];
            float2 E[4], O[3];
            float o0 = 0.f, o7 = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) E[q] = make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 3; ++q) O[q] = make_float2(0.f, 0.f);
            auto vpart = [&](auto H_) {
                constexpr int hh = decltype(H_)::value;
                constexpr int A0 = hh == 0 ? 0 : (TL + 1) / 2 / 4 * 4;          // tap split on a 4-row unit boundary
                constexpr int A1 = hh == 0 ? (TL + 1) / 2 / 4 * 4 : TL;
                constexpr int U0_ = A0 / 4, U1_ = (A1 - 1 + RB - 1) / 4;        // ring units touched
                // units before the ring wraps use base 0, the others base 1 (= base 0 - ring size)
                const int ub = 2 * g0;
                const float4 *b0 = ringH + ub * TWP + cp;
                const float4 *b1 = b0 - 2 * G::NR * TWP;
                const int wrap_at = 2 * G::NR - ub;                            // first unit index that wraps
                static_for<U0_, U1_ + 1>([&](auto U_) {
                    constexpr int u = decltype(U_)::value;
                    const float4 v = (u < wrap_at ? b0 : b1)[u * TWP];
                    static_for<0, 2>([&](auto P_) {
                        constexpr int t = 4 * u + 2 * decltype(P_)::value;      // even ring row of the pair
                        const float2 hp = decltype(P_)::value ? make_float2(v.z, v.w) : make_float2(v.x, v.y);
                        static_for<0, 4>([&](auto Q_) {
                            constexpr int a = t - 2 * decltype(Q_)::value;
                            if constexpr (a >= A0 && a < A1) ffma2(E[decltype(Q_)::value], hp, c_n32[TL == 80][a]);
                        });
                        static_for<0, 3>([&](auto Q_) {
                            constexpr int a = t - 2 * decltype(Q_)::value - 1;
                            if constexpr (a >= A0 && a < A1) ffma2(O[decltype(Q_)::value], hp, c_n32[TL == 80][a]);
                        });
                        if constexpr (t + 1 >= A0 && t + 1 < A1) o0 = fmaf(c_n32[TL == 80][t + 1], hp.y, o0);
                        if constexpr (t - 7 >= A0 && t - 7 < A1) o7 = fmaf(c_n32[TL == 80][t - 7], hp.x, o7);
                    });
                });
            };
            if (!empty) { if (vhalf == 0) vpart(std::integral_constant<int, 0>{}); else vpart(std::integral_constant<int, 1>{}); }
            part[0] = E[0].x + o0;     part[1] = E[0].y + O[0].x; part[2] = E[1].x + O[0].y; part[3] = E[1].y + O[1].x;
            part[4] = E[2].x + O[1].y; part[5] = E[2].y + O[2].x; part[6] = E[3].x + O[2].y; part[7] = E[3].y + o7;
            // swap: each half sends the partials of the 4 rows the other half decides
            {
                float *dst = xbuf + (vhalf * VR) * TW + vcol;
#pragma unroll
                for (int r = 0; r < VR; ++r) dst[r * TW] = vhalf ? part[r] : part[VR + r];
            }
            __syncthreads();
            float acc[VR];
            {
                const float *src = xbuf + ((1 - vhalf) * VR) * TW + vcol;
#pragma unroll
                for (int r = 0; r < VR; ++r) {
                    const float mine = vhalf ? part[VR + r] : part[r];
                    const float other = src[r * TW];
                    acc[r] = vhalf ? other + mine : mine + other;     // low taps first on both sides
                }
            }
            // ---- D: decision --------------------------------------------------------------------------
            const int x = x0 + vcol;
            const int y0 = yb + VR * vhalf;
            uint32_t myword = 0;
            auto queue = [&](int y) {                 // float32 cannot decide: queue for the float64 pass
                const uint32_t slot = atomicAdd(P.recheck_n + f, 1u);
                if (slot < (uint32_t)P.recheck_cap) P.recheck[(size_t)f * P.recheck_cap + slot] = make_int2(x, y);
                else atomicOr(P.status, VBS_DEV_RECHECK_OVERFLOW);
            };
            if (strip_interior && yb >= G::OFF && yb + RB - 1 + G::HI < H && yb + RB <= ye) {
                // whole step inside the image: threshold is the table entry of the box sum
                float thr[VR];
#pragma unroll
                for (int r = 0; r < VR; ++r) thr[r] = __ldg(P.thr_lut + S[r]);
#pragma unroll
                for (int r = 0; r < VR; ++r) {
                    const float d = acc[r] - thr[r];
                    bool on = d > 0.f;
                    if (fabsf(d) <= BAND) { on = false; queue(y0 + r); }
                    const uint32_t word = __ballot_sync(0xffffffffu, on);
                    if (lane == r) myword = word;
                }
            } else {
                const bool xin = x >= G::OFF && x + G::HI < W;
#pragma unroll
                for (int r = 0; r < VR; ++r) {
                    const int y = y0 + r;
                    bool on = false;
                    if (x < W && y < ye && S[r] != 0) {                     // S == 0: G = 0 and thr > 0 (or infinite) -> 0
                        float thr = INFINITY;
                        float band = BAND;
                        bool decided = false;
                        if (xin && y >= G::OFF && y + G::HI < H) thr = __ldg(P.thr_lut + S[r]);
                        else {
                            const BorderGeo bg = border_geo<TL>(y, x, H, W, cn);
                            if (S[r] == bg.A) { on = border_full_window_on<TL>(bg, P.st2, mfrac); decided = true; }
                            else { thr = border_threshold<TL>(bg, (float)S[r], mfrac, mcomp, (float)P.st2); band = BAND_BORDER; }
                        }
                        if (!decided) {
                            const float d = acc[r] - thr;
                            on = d > 0.f;
                            if (fabsf(d) <= band) { on = false; queue(y); }
                        }
                    }
                    const uint32_t word = __ballot_sync(0xffffffffu, on);
                    if (lane == r) myword = word;
                }
            }
            const int wx = (x0 >> 5) + (vcol >> 5);
            const int yw = y0 + lane;
            if (lane < VR && yw < ye && wx < WW) P.mask_bits[((size_t)f * H + yw) * WW + wx] = myword;
        }
        // No barrier here: the next horizontal step overwrites the oldest ring group, whose last readers (the
        // vertical pass of this step) all finished before the barrier of the partial-sum swap; the swap buffer is
        // rewritten only after the next step's post-H barrier.
        if (++gw == G::NR) gw = 0;
    }
}

// One pixel of a step that touches the image border (or of a strip that does): 0 = off, 1 = on, 2 = float32 cannot
// decide.  Kept out of line: it runs in a small minority of the steps and would otherwise be unrolled 8 times into
// the hot loop's instruction footprint.
template <int TL>
__device__ __noinline__ int border_decide(int y, int x, int H, int W, int S, float acc, float mfrac, float mcomp, double st2, const float *thr_lut,
                                          const double *cn) {
    using G = Geo<TL>;
    if (S == 0) return 0;                                   // G = 0 and thr > 0 (or infinite)
    float thr, band = BAND;
    if (x >= G::OFF && x + G::HI < W && y >= G::OFF && y + G::HI < H) thr = __ldg(thr_lut + S);
    else {
        const BorderGeo bg = border_geo<TL>(y, x, H, W, cn);
        if (S == bg.A) return border_full_window_on<TL>(bg, st2, mfrac) ? 1 : 0;
        thr = border_threshold<TL>(bg, (float)S, mfrac, mcomp, (float)st2);
        band = BAND_BORDER;
    }
    const float d = acc - thr;
    if (fabsf(d) <= band) return 2;
    return d > 0.f ? 1 : 0;
}

// ---- column-per-thread variant (default) -------------------------------------------------------------------------
// The kernel above keeps half of its 256 threads idle in the horizontal pass (8 rows x 16 octets = 128 items) and pays
// two block barriers per 8-row step (ring hand-over, partial-sum swap).  Here a CTA has as many threads as its strip
// has columns (TW2 = 128 or 192):
//   H  8 rows x TW2/8 octets = TW2 items: every thread works
//   V  thread = column, all L taps of its 8 rows: nothing to swap, the box-sum chain is computed once per column
//   one barrier per step: the ring has a spare group, so the horizontal pass of step m+1 never writes a group the
//   vertical pass of step m may still be reading.
template <int TL, int TW2> struct GeoW : Geo<TL> {
    static constexpr int TWP2 = TW2 + TW2 / 8;
    static constexpr int NR = Geo<TL>::NG + 1;
    static constexpr size_t SMEM = (size_t)NR * 2 * TWP2 * 16 + (size_t)NR * TWP2 * 8 + (size_t)(Geo<TL>::CNX + (Geo<TL>::CNX & 1)) * 8 + 4 * 112 * 4;
};

template <int TL, int TW2>
__global__ void __launch_bounds__(TW2, TW2 == 128 ? 3 : 2) ncc_mask_wide_kernel(NccParams P) {
    using G = GeoW<TL, TW2>;
    constexpr int NT = TW2, TWP2 = G::TWP2, OCT = TW2 / 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *ringH = reinterpret_cast<float4 *>(smem_raw);                        // [NR*2][TWP2]
    uint2 *ringB = reinterpret_cast<uint2 *>(ringH + G::NR * 2 * TWP2);          // [NR][TWP2]
    double *cn = reinterpret_cast<double *>(ringB + G::NR * TWP2);               // [CNX] float64 prefix sums (border formula)
    int4 *fx = reinterpret_cast<int4 *>(cn + G::CNX + (G::CNX & 1));             // [4][FXN/4] shifted fixed-point copies

    const int tid = threadIdx.x, lane = tid & 31;
    // CTA -> (frame, strip, row segment): whole-height items first, the tail of the grid in row segments (VbsSegPlan)
    int item = blockIdx.x, ys = 0, ye = P.H;
    if (item >= P.plan.n_full) {
        const int j = item - P.plan.n_full, q = j / P.plan.vsegs;
        item = P.plan.n_full + q;
        ys = (j - q * P.plan.vsegs) * P.plan.seg_rows;
        ye = min(P.H, ys + P.plan.seg_rows);
    }
    const int f = item / P.strips;
    const int x0 = (item - f * P.strips) * TW2;
    const int H = P.H, W = P.W, WW = P.WW;
    const uint32_t *abits = P.area_bits + (size_t)f * H * WW;
    const double mfrac64 = (double)P.area_count[f] / P.hw;
    const float mfrac = (float)mfrac64, mcomp = (float)(1.0 - mfrac64);   // mean(area_mask)/255 and its complement
    for (int i = tid; i < G::CNX; i += NT) cn[i] = P.cn64[i];
    for (int i = tid; i < 4 * FXN; i += NT) reinterpret_cast<int *>(fx)[i] = P.cnfix[i];

    const int nk = (ye - ys + RB - 1) / RB;
    const int nsteps = nk + G::LEAD;
    const int hr = tid / OCT, ho = tid - hr * OCT;                  // horizontal role: row hr of the step, pixel octet ho
    const int sb = x0 + 8 * ho - G::OFF;                            // first bit of the union window
    const int wi0 = sb >> 5, bo = sb & 31;                          // arithmetic shift: floor
    const int vcol = tid;                                           // vertical role: one column
    const int cp = vcol + (vcol >> 3);
    const bool strip_interior = x0 >= G::OFF && x0 + TW2 - 1 + G::HI < W;
    const bool warp_outside = x0 + (tid & ~31) >= W;                // this warp's 32 columns lie right of the image

    uint32_t pw[4] = {0, 0, 0, 0};
    int woff[4]; uint32_t wmsk[4];                                  // the four words of the union window: step-invariant
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool ok = wi0 + i >= 0 && wi0 + i < WW;
        woff[i] = ok ? wi0 + i : 0; wmsk[i] = ok ? 0xffffffffu : 0u;
    }
    auto fetch = [&](int m) {
        const int p = ys - G::OFF + RB * m + hr;
        if (p >= 0 && p < H) {
            const uint32_t *row = abits + (size_t)p * WW;
#pragma unroll
            for (int i = 0; i < 4; ++i) pw[i] = __ldg(row + woff[i]);     // masked where used: the load stays in flight for a whole step
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) pw[i] = 0u;
        }
    };
    int s_prev = 0; uint32_t hb_m1 = 0;         // S of the previous output row and the box row that left the window
    uint32_t s_lead = 0;                        // box sum of the lead groups = rows 0 .. 8 (TL / 8) - 1 of the first window
    static_assert(G::LEAD == TL / 8 && TL % 8 <= 4, "first-window box sum is accumulated group by group during the lead steps");

    fetch(0);
    __syncthreads();
    int gw = 0;                                  // ring slot written by step m (m % NR)
    for (int m = 0; m < nsteps; ++m) {
        // ---- H: horizontal pass on bits ------------------------------------------------------------
        {
#pragma unroll
            for (int i = 0; i < 4; ++i) pw[i] &= wmsk[i];
            uint32_t U0 = __funnelshift_r(pw[0], pw[1], bo);
            uint32_t U1 = __funnelshift_r(pw[1], pw[2], bo);
            uint32_t U2 = __funnelshift_r(pw[2], pw[3], bo);
            if (m + 1 < nsteps) fetch(m + 1);
            if constexpr (G::UL <= 64) { U2 = 0; U1 &= (G::UL == 64) ? 0xffffffffu : ((1u << (G::UL - 32)) - 1u); }
            else { U2 &= (1u << (G::UL - 64)) - 1u; }
            const uint32_t T0 = U0 ^ (U0 << 1);
            const uint32_t T1 = U1 ^ __funnelshift_l(U0, U1, 1);
            const uint32_t T2 = U2 ^ __funnelshift_l(U1, U2, 1);
            int acc[8];
            uint32_t hb[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { acc[j] = 0; hb[j] = 0; }
            auto consume = [&](uint32_t T, uint32_t Uw, int base) {
                while (T) {
                    const int b = __ffs(T) - 1;
                    T &= T - 1;
                    const int sgn = ((Uw >> b) & 1u) ? -1 : 1;       // run start: -Cn, run end: +Cn
                    const int i0 = base + b + 1;                     // table index of d = t - 7 (pixel j = 7)
                    const int4 *src = fx + (i0 & 3) * (FXN / 4) + (i0 >> 2);
                    const int4 lo = src[0], hi = src[1];
                    acc[7] += sgn * lo.x; acc[6] += sgn * lo.y; acc[5] += sgn * lo.z; acc[4] += sgn * lo.w;
                    acc[3] += sgn * hi.x; acc[2] += sgn * hi.y; acc[1] += sgn * hi.z; acc[0] += sgn * hi.w;
                }
            };
            if (U0 | U1 | U2) {                                      // a window without area pixels leaves h = 0, box = 0
                consume(T0, U0, 0);
                consume(T1, U1, 32);
                if constexpr (G::UL > 64) consume(T2, U2, 64);
                auto bit = [&](int t) -> uint32_t {
                    return t < 32 ? (U0 >> t) & 1u : t < 64 ? (U1 >> (t - 32)) & 1u : (U2 >> (t - 64)) & 1u;
                };
                if constexpr (TL >= 64) hb[0] = __popc(U0) + __popc(U1) + __popc(U2 & ((1u << (TL - 64)) - 1u));
                else hb[0] = __popc(U0) + __popc(U1 & ((1u << (TL - 32)) - 1u));
#pragma unroll
                for (int j = 1; j < 8; ++j) hb[j] = hb[j - 1] + bit(j - 1 + TL) - bit(j - 1);
            }
            float *dstH = reinterpret_cast<float *>(ringH + (gw * 2 + (hr >> 2)) * TWP2) + (hr & 3);
            unsigned char *dstB = reinterpret_cast<unsigned char *>(ringB + gw * TWP2) + hr;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = 9 * ho + j;                            // (8 ho + j) + (8 ho + j) / 8
                dstH[c * 4] = (float)acc[j];                         // 2^30 h; the scale lives in the vertical weights
                dstB[c * 8] = (unsigned char)hb[j];
            }
        }
        __syncthreads();                                             // the only barrier of the step

        // ---- V + D: vertical pass and decision, thread = column ------------------------------------
        if (m < G::LEAD) {                                           // lead: only the box sum of the first window grows
            const uint2 b = ringB[gw * TWP2 + cp];
            s_lead = __dp4a(b.x, 0x01010101u, s_lead);
            s_lead = __dp4a(b.y, 0x01010101u, s_lead);
        } else {
            const int k = m - G::LEAD;
            const int yb = ys + RB * k;
            int g0 = gw - G::LEAD; if (g0 < 0) g0 += G::NR;          // slot of step k
            int S8[RB];
            bool empty;
            {
                auto box_at = [&](auto I_) -> int {
                    constexpr int idx = decltype(I_)::value;
                    int gi = g0 + idx / 8; if (gi >= G::NR) gi -= G::NR;
                    const uint2 b = ringB[gi * TWP2 + cp];
                    const uint32_t w = (idx % 8) < 4 ? b.x : b.y;
                    return (int)((w >> (8 * (idx % 4))) & 255u);
                };
                if (k == 0) {                   // first step of the segment: the lead groups + the TL % 8 rows of this step's group
                    int s = (int)s_lead;
                    if constexpr (TL % 8 != 0) s += (int)__dp4a(ringB[gw * TWP2 + cp].x & ((1u << (8 * (TL % 8))) - 1u), 0x01010101u, 0u);
                    S8[0] = s;
                } else {
                    S8[0] = s_prev + box_at(std::integral_constant<int, TL - 1>{}) - (int)hb_m1;
                }
                static_for<1, RB>([&](auto R_) {
                    constexpr int r = decltype(R_)::value;
                    S8[r] = S8[r - 1] + box_at(std::integral_constant<int, r + TL - 1>{}) - box_at(std::integral_constant<int, r - 1>{});
                });
                s_prev = S8[RB - 1]; hb_m1 = (uint32_t)box_at(std::integral_constant<int, RB - 1>{});
                int any = 0;
#pragma unroll
                for (int r = 0; r < RB; ++r) any |= S8[r];
                // S == 0 <=> no area pixel in the whole L x L window <=> G == 0 and mask == 0: when that holds
                // for all 8 rows of all 32 columns of the warp, the L-tap column sums are skipped (exact)
                empty = warp_outside || !__any_sync(0xffffffffu, any != 0);
            }
            // Gaussian column sums of the 8 rows, packed: ring rows come as float4 = two aligned pairs (h_t, h_t+1),
            // accumulators are paired as (row r, row r+1), one FFMA2 applies one tap to two rows
            float2 E[4], O[3];
            float o0 = 0.f, o7 = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) E[q] = make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 3; ++q) O[q] = make_float2(0.f, 0.f);
            if (!empty) {
                constexpr int U1_ = (TL - 1 + RB - 1) / 4;                       // last ring unit touched
                const int ub = 2 * g0;
                const float4 *b0 = ringH + ub * TWP2 + cp;
                const float4 *b1 = b0 - 2 * G::NR * TWP2;
                const int wrap_at = 2 * G::NR - ub;                             // first unit index that wraps
                auto unit = [&](auto U_, const float4 v) {
                    constexpr int u = decltype(U_)::value;
                    static_for<0, 2>([&](auto P_) {
                        constexpr int t = 4 * u + 2 * decltype(P_)::value;      // even ring row of the pair
                        const float2 hp = decltype(P_)::value ? make_float2(v.z, v.w) : make_float2(v.x, v.y);
                        static_for<0, 4>([&](auto Q_) {
                            constexpr int a = t - 2 * decltype(Q_)::value;
                            if constexpr (a >= 0 && a < TL) ffma2(E[decltype(Q_)::value], hp, c_n32[TL == 80][a]);
                        });
                        static_for<0, 3>([&](auto Q_) {
                            constexpr int a = t - 2 * decltype(Q_)::value - 1;
                            if constexpr (a >= 0 && a < TL) ffma2(O[decltype(Q_)::value], hp, c_n32[TL == 80][a]);
                        });
                        if constexpr (t + 1 >= 0 && t + 1 < TL) o0 = fmaf(c_n32[TL == 80][t + 1], hp.y, o0);
                        if constexpr (t - 7 >= 0 && t - 7 < TL) o7 = fmaf(c_n32[TL == 80][t - 7], hp.x, o7);
                    });
                };
                static_for<0, U1_ + 1>([&](auto U_) {
                    constexpr int u = decltype(U_)::value;
                    unit(U_, (u < wrap_at ? b0 : b1)[u * TWP2]);
                });
            }
            float acc[RB];
            acc[0] = E[0].x + o0;     acc[1] = E[0].y + O[0].x; acc[2] = E[1].x + O[0].y; acc[3] = E[1].y + O[1].x;
            acc[4] = E[2].x + O[1].y; acc[5] = E[2].y + O[2].x; acc[6] = E[3].x + O[2].y; acc[7] = E[3].y + o7;
            // ---- D: decision --------------------------------------------------------------------------
            const int x = x0 + vcol;
            uint32_t onmask = 0, needmask = 0;        // bit r: row r of this column is on / needs the float64 pass
            if (empty) {
                // S == 0 in the whole warp: off on every path (interior table entry 0 is +inf, border_decide returns 0)
            } else if (strip_interior && yb >= G::OFF && yb + RB - 1 + G::HI < H && yb + RB <= ye) {
                float thr[RB];                        // whole step inside the image: threshold is the table entry of the box sum
#pragma unroll
                for (int r = 0; r < RB; ++r) thr[r] = __ldg(P.thr_lut + S8[r]);
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const float d = acc[r] - thr[r];
                    const bool need = fabsf(d) <= BAND;
                    onmask |= (uint32_t)(d > 0.f && !need) << r;
                    needmask |= (uint32_t)need << r;
                }
            } else {
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const int y = yb + r;
                    if (x < W && y < ye) {
                        const int dec = border_decide<TL>(y, x, H, W, S8[r], acc[r], mfrac, mcomp, P.st2, P.thr_lut, cn);
                        onmask |= (uint32_t)(dec == 1) << r;
                        needmask |= (uint32_t)(dec == 2) << r;
                    }
                }
            }
            uint32_t myword = 0;
            if (!empty) {
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const uint32_t word = __ballot_sync(0xffffffffu, (onmask >> r) & 1u);
                    if (lane == r) myword = word;
                }
            }
            while (needmask) {                        // float32 cannot decide: queue for the float64 pass (rare)
                const int r = __ffs(needmask) - 1;
                needmask &= needmask - 1;
                const uint32_t slot = atomicAdd(P.recheck_n + f, 1u);
                if (slot < (uint32_t)P.recheck_cap) P.recheck[(size_t)f * P.recheck_cap + slot] = make_int2(x, yb + r);
                else atomicOr(P.status, VBS_DEV_RECHECK_OVERFLOW);
            }
            const int wx = (x0 >> 5) + (vcol >> 5);
            const int yw = yb + lane;
            if (lane < RB && yw < ye && wx < WW) P.mask_bits[((size_t)f * H + yw) * WW + wx] = myword;
        }
        if (++gw == G::NR) gw = 0;
    }
}

// ---- column-per-thread variant with warp roles -----------------------------------------------------------------
// Same arithmetic and ring as ncc_mask_wide_kernel, twice the threads: warps 0-3 run the horizontal pass (integer adds,
// table loads) of step i while warps 4-7 run the vertical pass (FFMA2) and the decision of step i - 1; the ring's spare
// group already keeps the two apart, so there is still one CTA barrier per step.  24 warps per SM instead of 12.
template <int TL, int TW2>
__global__ void __launch_bounds__(2 * TW2, 3) ncc_mask_roles_kernel(NccParams P) {
    using G = GeoW<TL, TW2>;
    constexpr int NT = TW2, TWP2 = G::TWP2, OCT = TW2 / 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *ringH = reinterpret_cast<float4 *>(smem_raw);                        // [NR*2][TWP2]
    uint2 *ringB = reinterpret_cast<uint2 *>(ringH + G::NR * 2 * TWP2);          // [NR][TWP2]
    double *cn = reinterpret_cast<double *>(ringB + G::NR * TWP2);               // [CNX] float64 prefix sums (border formula)
    int4 *fx = reinterpret_cast<int4 *>(cn + G::CNX + (G::CNX & 1));             // [4][FXN/4] shifted fixed-point copies

    const int tid = threadIdx.x % TW2, lane = tid & 31;              // index within the role
    const bool hrole = threadIdx.x < TW2;                            // warps 0..: horizontal pass of step i; the others: vertical pass of step i - 1
    // CTA -> (frame, strip, row segment): whole-height items first, the tail of the grid in row segments (VbsSegPlan)
    int item = blockIdx.x, ys = 0, ye = P.H;
    if (item >= P.plan.n_full) {
        const int j = item - P.plan.n_full, q = j / P.plan.vsegs;
        item = P.plan.n_full + q;
        ys = (j - q * P.plan.vsegs) * P.plan.seg_rows;
        ye = min(P.H, ys + P.plan.seg_rows);
    }
    const int f = item / P.strips;
    const int x0 = (item - f * P.strips) * TW2;
    const int H = P.H, W = P.W, WW = P.WW;
    const uint32_t *abits = P.area_bits + (size_t)f * H * WW;
    const double mfrac64 = (double)P.area_count[f] / P.hw;
    const float mfrac = (float)mfrac64, mcomp = (float)(1.0 - mfrac64);   // mean(area_mask)/255 and its complement
    for (int i = threadIdx.x; i < G::CNX; i += 2 * NT) cn[i] = P.cn64[i];
    for (int i = threadIdx.x; i < 4 * FXN; i += 2 * NT) reinterpret_cast<int *>(fx)[i] = P.cnfix[i];

    const int nk = (ye - ys + RB - 1) / RB;
    const int nsteps = nk + G::LEAD;
    const int hr = tid / OCT, ho = tid - hr * OCT;                  // horizontal role: row hr of the step, pixel octet ho
    const int sb = x0 + 8 * ho - G::OFF;                            // first bit of the union window
    const int wi0 = sb >> 5, bo = sb & 31;                          // arithmetic shift: floor
    const int vcol = tid;                                           // vertical role: one column
    const int cp = vcol + (vcol >> 3);
    const bool strip_interior = x0 >= G::OFF && x0 + TW2 - 1 + G::HI < W;
    const bool warp_outside = x0 + (tid & ~31) >= W;                // this warp's 32 columns lie right of the image

    uint32_t pw[4] = {0, 0, 0, 0};
    int woff[4]; uint32_t wmsk[4];                                  // the four words of the union window: step-invariant
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool ok = wi0 + i >= 0 && wi0 + i < WW;
        woff[i] = ok ? wi0 + i : 0; wmsk[i] = ok ? 0xffffffffu : 0u;
    }
    auto fetch = [&](int m) {
        const int p = ys - G::OFF + RB * m + hr;
        if (p >= 0 && p < H) {
            const uint32_t *row = abits + (size_t)p * WW;
#pragma unroll
            for (int i = 0; i < 4; ++i) pw[i] = __ldg(row + woff[i]);     // masked where used: the load stays in flight for a whole step
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) pw[i] = 0u;
        }
    };
    int s_prev = 0; uint32_t hb_m1 = 0;         // S of the previous output row and the box row that left the window
    uint32_t s_lead = 0;                        // box sum of the lead groups = rows 0 .. 8 (TL / 8) - 1 of the first window
    static_assert(G::LEAD == TL / 8 && TL % 8 <= 4, "first-window box sum is accumulated group by group during the lead steps");

    __syncthreads();                             // tables loaded
    int gw = 0;                                  // ring slot written by step m (m % NR)
    if (hrole) {
      fetch(0);
      for (int m = 0; m <= nsteps; ++m) {        // iteration m: H of step m beside V of step m - 1; the ring's spare group keeps them apart
        // ---- H: horizontal pass on bits ------------------------------------------------------------
        if (m < nsteps) {
#pragma unroll
            for (int i = 0; i < 4; ++i) pw[i] &= wmsk[i];
            uint32_t U0 = __funnelshift_r(pw[0], pw[1], bo);
            uint32_t U1 = __funnelshift_r(pw[1], pw[2], bo);
            uint32_t U2 = __funnelshift_r(pw[2], pw[3], bo);
            if (m + 1 < nsteps) fetch(m + 1);
            if constexpr (G::UL <= 64) { U2 = 0; U1 &= (G::UL == 64) ? 0xffffffffu : ((1u << (G::UL - 32)) - 1u); }
            else { U2 &= (1u << (G::UL - 64)) - 1u; }
            const uint32_t T0 = U0 ^ (U0 << 1);
            const uint32_t T1 = U1 ^ __funnelshift_l(U0, U1, 1);
            const uint32_t T2 = U2 ^ __funnelshift_l(U1, U2, 1);
            int acc[8];
            uint32_t hb[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { acc[j] = 0; hb[j] = 0; }
            auto consume = [&](uint32_t T, uint32_t Uw, int base) {
                while (T) {
                    const int b = __ffs(T) - 1;
                    T &= T - 1;
                    const int sgn = ((Uw >> b) & 1u) ? -1 : 1;       // run start: -Cn, run end: +Cn
                    const int i0 = base + b + 1;                     // table index of d = t - 7 (pixel j = 7)
                    const int4 *src = fx + (i0 & 3) * (FXN / 4) + (i0 >> 2);
                    const int4 lo = src[0], hi = src[1];
                    acc[7] += sgn * lo.x; acc[6] += sgn * lo.y; acc[5] += sgn * lo.z; acc[4] += sgn * lo.w;
                    acc[3] += sgn * hi.x; acc[2] += sgn * hi.y; acc[1] += sgn * hi.z; acc[0] += sgn * hi.w;
                }
            };
            if (U0 | U1 | U2) {                                      // a window without area pixels leaves h = 0, box = 0
                consume(T0, U0, 0);
                consume(T1, U1, 32);
                if constexpr (G::UL > 64) consume(T2, U2, 64);
                auto bit = [&](int t) -> uint32_t {
                    return t < 32 ? (U0 >> t) & 1u : t < 64 ? (U1 >> (t - 32)) & 1u : (U2 >> (t - 64)) & 1u;
                };
                if constexpr (TL >= 64) hb[0] = __popc(U0) + __popc(U1) + __popc(U2 & ((1u << (TL - 64)) - 1u));
                else hb[0] = __popc(U0) + __popc(U1 & ((1u << (TL - 32)) - 1u));
#pragma unroll
                for (int j = 1; j < 8; ++j) hb[j] = hb[j - 1] + bit(j - 1 + TL) - bit(j - 1);
            }
            float *dstH = reinterpret_cast<float *>(ringH + (gw * 2 + (hr >> 2)) * TWP2) + (hr & 3);
            unsigned char *dstB = reinterpret_cast<unsigned char *>(ringB + gw * TWP2) + hr;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = 9 * ho + j;                            // (8 ho + j) + (8 ho + j) / 8
                dstH[c * 4] = (float)acc[j];                         // 2^30 h; the scale lives in the vertical weights
                dstB[c * 8] = (unsigned char)hb[j];
            }
            if (++gw == G::NR) gw = 0;
        }
        __syncthreads();                                             // the only barrier of the step
      }
    } else {
      __syncthreads();                                               // iteration 0: nothing to consume yet
      for (int m = 0; m < nsteps; ++m) {
        // ---- V + D: vertical pass and decision, thread = column ------------------------------------
        if (m < G::LEAD) {                                           // lead: only the box sum of the first window grows
            const uint2 b = ringB[gw * TWP2 + cp];
            s_lead = __dp4a(b.x, 0x01010101u, s_lead);
            s_lead = __dp4a(b.y, 0x01010101u, s_lead);
        } else {
            const int k = m - G::LEAD;
            const int yb = ys + RB * k;
            int g0 = gw - G::LEAD; if (g0 < 0) g0 += G::NR;          // slot of step k
            int S8[RB];
            bool empty;
            {
                auto box_at = [&](auto I_) -> int {
                    constexpr int idx = decltype(I_)::value;
                    int gi = g0 + idx / 8; if (gi >= G::NR) gi -= G::NR;
                    const uint2 b = ringB[gi * TWP2 + cp];
                    const uint32_t w = (idx % 8) < 4 ? b.x : b.y;
                    return (int)((w >> (8 * (idx % 4))) & 255u);
                };
                if (k == 0) {                   // first step of the segment: the lead groups + the TL % 8 rows of this step's group
                    int s = (int)s_lead;
                    if constexpr (TL % 8 != 0) s += (int)__dp4a(ringB[gw * TWP2 + cp].x & ((1u << (8 * (TL % 8))) - 1u), 0x01010101u, 0u);
                    S8[0] = s;
                } else {
                    S8[0] = s_prev + box_at(std::integral_constant<int, TL - 1>{}) - (int)hb_m1;
                }
                static_for<1, RB>([&](auto R_) {
                    constexpr int r = decltype(R_)::value;
                    S8[r] = S8[r - 1] + box_at(std::integral_constant<int, r + TL - 1>{}) - box_at(std::integral_constant<int, r - 1>{});
                });
                s_prev = S8[RB - 1]; hb_m1 = (uint32_t)box_at(std::integral_constant<int, RB - 1>{});
                int any = 0;
#pragma unroll
                for (int r = 0; r < RB; ++r) any |= S8[r];
                // S == 0 <=> no area pixel in the whole L x L window <=> G == 0 and mask == 0: when that holds
                // for all 8 rows of all 32 columns of the warp, the L-tap column sums are skipped (exact)
                empty = warp_outside || !__any_sync(0xffffffffu, any != 0);
            }
            // Gaussian column sums of the 8 rows, packed: ring rows come as float4 = two aligned pairs (h_t, h_t+1),
            // accumulators are paired as (row r, row r+1), one FFMA2 applies one tap to two rows
            float2 E[4], O[3];
            float o0 = 0.f, o7 = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) E[q] = make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 3; ++q) O[q] = make_float2(0.f, 0.f);
            if (!empty) {
                constexpr int U1_ = (TL - 1 + RB - 1) / 4;                       // last ring unit touched
                const int ub = 2 * g0;
                const float4 *b0 = ringH + ub * TWP2 + cp;
                const float4 *b1 = b0 - 2 * G::NR * TWP2;
                const int wrap_at = 2 * G::NR - ub;                             // first unit index that wraps
                auto unit = [&](auto U_, const float4 v) {
                    constexpr int u = decltype(U_)::value;
                    static_for<0, 2>([&](auto P_) {
                        constexpr int t = 4 * u + 2 * decltype(P_)::value;      // even ring row of the pair
                        const float2 hp = decltype(P_)::value ? make_float2(v.z, v.w) : make_float2(v.x, v.y);
                        static_for<0, 4>([&](auto Q_) {
                            constexpr int a = t - 2 * decltype(Q_)::value;
                            if constexpr (a >= 0 && a < TL) ffma2(E[decltype(Q_)::value], hp, c_n32[TL == 80][a]);
                        });
                        static_for<0, 3>([&](auto Q_) {
                            constexpr int a = t - 2 * decltype(Q_)::value - 1;
                            if constexpr (a >= 0 && a < TL) ffma2(O[decltype(Q_)::value], hp, c_n32[TL == 80][a]);
                        });
                        if constexpr (t + 1 >= 0 && t + 1 < TL) o0 = fmaf(c_n32[TL == 80][t + 1], hp.y, o0);
                        if constexpr (t - 7 >= 0 && t - 7 < TL) o7 = fmaf(c_n32[TL == 80][t - 7], hp.x, o7);
                    });
                };
                static_for<0, U1_ + 1>([&](auto U_) {
                    constexpr int u = decltype(U_)::value;
                    unit(U_, (u < wrap_at ? b0 : b1)[u * TWP2]);
                });
            }
            float acc[RB];
            acc[0] = E[0].x + o0;     acc[1] = E[0].y + O[0].x; acc[2] = E[1].x + O[0].y; acc[3] = E[1].y + O[1].x;
            acc[4] = E[2].x + O[1].y; acc[5] = E[2].y + O[2].x; acc[6] = E[3].x + O[2].y; acc[7] = E[3].y + o7;
            // ---- D: decision --------------------------------------------------------------------------
            const int x = x0 + vcol;
            uint32_t onmask = 0, needmask = 0;        // bit r: row r of this column is on / needs the float64 pass
            if (empty) {
                // S == 0 in the whole warp: off on every path (interior table entry 0 is +inf, border_decide returns 0)
            } else if (strip_interior && yb >= G::OFF && yb + RB - 1 + G::HI < H && yb + RB <= ye) {
                float thr[RB];                        // whole step inside the image: threshold is the table entry of the box sum
#pragma unroll
                for (int r = 0; r < RB; ++r) thr[r] = __ldg(P.thr_lut + S8[r]);
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const float d = acc[r] - thr[r];
                    const bool need = fabsf(d) <= BAND;
                    onmask |= (uint32_t)(d > 0.f && !need) << r;
                    needmask |= (uint32_t)need << r;
                }
            } else {
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const int y = yb + r;
                    if (x < W && y < ye) {
                        const int dec = border_decide<TL>(y, x, H, W, S8[r], acc[r], mfrac, mcomp, P.st2, P.thr_lut, cn);
                        onmask |= (uint32_t)(dec == 1) << r;
                        needmask |= (uint32_t)(dec == 2) << r;
                    }
                }
            }
            uint32_t myword = 0;
            if (!empty) {
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const uint32_t word = __ballot_sync(0xffffffffu, (onmask >> r) & 1u);
                    if (lane == r) myword = word;
                }
            }
            while (needmask) {                        // float32 cannot decide: queue for the float64 pass (rare)
                const int r = __ffs(needmask) - 1;
                needmask &= needmask - 1;
                const uint32_t slot = atomicAdd(P.recheck_n + f, 1u);
                if (slot < (uint32_t)P.recheck_cap) P.recheck[(size_t)f * P.recheck_cap + slot] = make_int2(x, yb + r);
                else atomicOr(P.status, VBS_DEV_RECHECK_OVERFLOW);
            }
            const int wx = (x0 >> 5) + (vcol >> 5);
            const int yw = yb + lane;
            if (lane < RB && yw < ye && wx < WW) P.mask_bits[((size_t)f * H + yw) * WW + wx] = myword;
        }
        if (++gw == G::NR) gw = 0;
        __syncthreads();
      }
    }
}

// float64 re-decision of queued pixels with the reference's literal formula (MD:152-163).
// One warp per pixel: lanes split the window rows; fixed-order butterfly keeps it deterministic.
template <int TL>
__global__ void ncc_recheck_kernel(NccParams P, const double *__restrict__ n64, int batch) {
    using G = Geo<TL>;
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int f = blockIdx.y;
    const uint32_t n = min(P.recheck_n[f], (uint32_t)P.recheck_cap);
    const int H = P.H, W = P.W, WW = P.WW;
    const uint32_t *abits = P.area_bits + (size_t)f * H * WW;
    const double mu = (255.0 * (double)P.area_count[f]) / P.hw;       // np.mean(area_mask): exact sum, one division
    for (uint32_t e = blockIdx.x * wpb + (threadIdx.x >> 5); e < n; e += gridDim.x * wpb) {
        const int2 px = P.recheck[(size_t)f * P.recheck_cap + e];
        const int x = px.x, y = px.y;
        double gsum = 0.0; int ssum = 0;
        for (int a = lane; a < TL; a += 32) {
            const int p = y - G::OFF + a;
            if (p < 0 || p >= H) continue;
            const uint32_t *row = abits + (size_t)p * WW;
            double h = 0.0; int cnt = 0;
            for (int c = 0; c < TL; ++c) {
                const int xx = x - G::OFF + c;
                if (xx >= 0 && xx < W && ((row[xx >> 5] >> (xx & 31)) & 1u)) { h += n64[c]; ++cnt; }
            }
            gsum += n64[a] * h; ssum += cnt;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            gsum += __shfl_xor_sync(0xffffffffu, gsum, o);
            ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
        }
        if (lane == 0) {
            const int ylo = max(0, G::OFF - y), yhi = min(TL - 1, H - 1 - y + G::OFF);
            const int xlo = max(0, G::OFF - x), xhi = min(TL - 1, W - 1 - x + G::OFF);
            const double A = (double)(yhi - ylo + 1) * (double)(xhi - xlo + 1);
            const double g1 = (P.cn64[yhi + 1 + 8] - P.cn64[ylo + 8]) * (P.cn64[xhi + 1 + 8] - P.cn64[xlo + 8]);
            const double L2 = (double)(TL * TL), S = (double)ssum;
            const double sep = 255.0 * gsum - mu * g1;
            const double box1 = 255.0 * S - mu * A;
            const double box2 = 65025.0 * S - 510.0 * mu * S + mu * mu * A;
            const double num = sep - box1 / L2;
            double isq = box2 - box1 * box1 / L2;
            if (isq < 0) isq = 0;
            double ncc = num / sqrt(isq * P.st2);
            if (!isfinite(ncc)) ncc = 0.0;
            if (ncc > 0.1) atomicOr(P.mask_bits + ((size_t)f * H + y) * WW + (x >> 5), 1u << (x & 31));
        }
    }
}

template <int TL> cudaError_t launch(vbs_ctx *ctx, int batch, double st2) {
    NccParams P;
    P.H = ctx->H; P.W = ctx->W; P.WW = ctx->WW;
    const int variant = ctx->ncc_variant;          // 2 (default): thread = column, warp roles; 1: thread = column; 0: 256 threads, two tap halves per column (round 1)
    const int tw = TW;
    const int strips = (ctx->W + tw - 1) / tw;
    // vertical segments cost LEAD halo steps each: split only while the grid is short of ~8 waves (296 or 444 CTA slots)
    const long long want = 4096;
    int vsegs = 1;
    while ((long long)strips * batch * vsegs < want && (ctx->H + vsegs) / (vsegs + 1) >= 64) ++vsegs;
    P.seg_rows = ((ctx->H + vsegs - 1) / vsegs + RB - 1) / RB * RB;
    vsegs = (ctx->H + P.seg_rows - 1) / P.seg_rows;
    P.st2 = st2; P.hw = (double)ctx->H * (double)ctx->W;
    P.area_bits = ctx->area_bits; P.area_count = ctx->area_count; P.mask_bits = ctx->mask_bits;
    P.thr_lut = ctx->thr_lut; P.cn64 = ctx->d_cn64; P.cnfix = ctx->d_cnfix;
    P.recheck = ctx->recheck; P.recheck_n = ctx->recheck_n; P.recheck_cap = ctx->recheck_cap;
    P.status = ctx->d_status;
    cudaError_t e = cudaMemsetAsync(ctx->recheck_n, 0, sizeof(uint32_t) * batch, ctx->stream);
    if (e != cudaSuccess) return e;
    if (variant == 2) {
        P.strips = strips;
        P.plan = vbs_seg_plan(ctx->H, (long long)strips * batch, 3 * ctx->sm_count, Geo<TL>::LEAD, RB, 0.4, ctx->seg_plan != 0);
        auto kern = ncc_mask_roles_kernel<TL, 128>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GeoW<TL, 128>::SMEM)) != cudaSuccess) return e;
        kern<<<dim3(P.plan.ctas), 256, GeoW<TL, 128>::SMEM, ctx->stream>>>(P);
    } else if (variant == 1) {
        // 3 CTAs per SM are resident; a lead step runs the horizontal pass only
        P.strips = strips;
        P.plan = vbs_seg_plan(ctx->H, (long long)strips * batch, 3 * ctx->sm_count, Geo<TL>::LEAD, RB, 0.4, ctx->seg_plan != 0);
        auto kern = ncc_mask_wide_kernel<TL, 128>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GeoW<TL, 128>::SMEM)) != cudaSuccess) return e;
        kern<<<dim3(P.plan.ctas), 128, GeoW<TL, 128>::SMEM, ctx->stream>>>(P);
    } else {
        auto kern = ncc_mask_kernel<TL>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GeoR<TL>::SMEM)) != cudaSuccess) return e;
        kern<<<dim3(strips, vsegs, batch), NT, GeoR<TL>::SMEM, ctx->stream>>>(P);
    }
    ncc_recheck_kernel<TL><<<dim3(16, batch), 256, 0, ctx->stream>>>(P, ctx->d_n64, batch);
    ctx->launches += 2;
    return cudaGetLastError();
}

}  // namespace

// host-side tables: n (float64 + float32), guarded prefix sums, interior threshold LUT, st2
cudaError_t vbs_ncc_setup(vbs_ctx *ctx) {
    const int TL = ctx->br.tl;
    double n[96], e[96], sum = 0;
    for (int i = 0; i < TL; ++i) {           // np.linspace(-(l-1)/2, (l-1)/2, l); exp(-0.5 ax^2 / sig^2); / sum
        const double lo = -(TL - 1) / 2.0, step = (double)(TL - 1) / (double)(TL - 1);
        const double ax = (i == TL - 1) ? (TL - 1) / 2.0 : lo + step * i;
        e[i] = exp(-0.5 * (ax * ax) / (ctx->br.tsigma * ctx->br.tsigma));
        sum += e[i];
    }
    double ss = 0;
    float n32[96] = {0};
    // the horizontal pass hands over 2^30 h (fixed point): fold the exact power of two into the weights
    for (int i = 0; i < TL; ++i) { n[i] = e[i] / sum; ss += n[i] * n[i]; n32[i] = (float)ldexp(n[i], -30); }
    const double L2 = (double)TL * TL;
    const double st2 = ss * ss - 1.0 / L2;
    ctx->st2 = st2;
    // guarded prefix sums: index d+8, d in [-8, UL+8]
    const int UL = 8 + TL - 1, CNX = 8 + UL + 9;
    double cn[160], pre[97];
    pre[0] = 0; for (int i = 0; i < TL; ++i) pre[i + 1] = pre[i] + n[i];
    for (int i = 0; i < CNX; ++i) { int d = i - 8; d = d < 0 ? 0 : (d > TL ? TL : d); cn[i] = pre[d]; }
    // interior threshold on G as a function of the box sum S
    const int NL = TL * TL + 1;
    float *lut = new float[NL];
    for (int S = 0; S < NL; ++S) {
        const double q = (double)S * (L2 - S);
        lut[S] = (q > 0) ? (float)(S / L2 + 0.1 * sqrt(st2 * q) / (double)TL) : INFINITY;
    }
    cudaError_t err;
    int fixt[4 * 112];
    for (int sft = 0; sft < 4; ++sft)
        for (int k2 = 0; k2 < 112; ++k2) {
            const int idx = k2 + sft;
            fixt[sft * 112 + k2] = idx < CNX ? (int)llrint(ldexp(cn[idx], 30)) : (int)llrint(ldexp(pre[TL], 30));
        }
    if ((err = cudaMemcpy(ctx->d_cnfix, fixt, sizeof(fixt), cudaMemcpyHostToDevice)) != cudaSuccess) { delete[] lut; return err; }
    if ((err = cudaMemcpy(ctx->d_n64, n, sizeof(double) * TL, cudaMemcpyHostToDevice)) != cudaSuccess) { delete[] lut; return err; }
    if ((err = cudaMemcpy(ctx->d_cn64, cn, sizeof(double) * CNX, cudaMemcpyHostToDevice)) != cudaSuccess) { delete[] lut; return err; }
    err = cudaMemcpy(ctx->thr_lut, lut, sizeof(float) * NL, cudaMemcpyHostToDevice);
    delete[] lut;
    if (err != cudaSuccess) return err;
    return cudaMemcpyToSymbol(c_n32, n32, sizeof(float) * 96, sizeof(float) * 96 * (TL == 80 ? 1 : 0));
}

cudaError_t vbs_launch_ncc(vbs_ctx *ctx, int batch) {
    VbsRange range("vbs:ncc");
    if (ctx->big) return launch<80>(ctx, batch, ctx->st2);
    return launch<33>(ctx, batch, ctx->st2);
}
