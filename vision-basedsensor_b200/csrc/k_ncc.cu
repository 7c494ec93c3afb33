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

template <int TL> struct Geo {
    static constexpr int OFF = (TL - 1) - (TL - 1) / 2;   // 40 / 16
    static constexpr int HI = TL - 1 - OFF;               // 39 / 16
    static constexpr int LEAD = cdiv(TL - 1, RB);         // 10 / 4
    static constexpr int NG = LEAD + 1;                   // ring groups of 8 rows
    static constexpr int UL = RB + TL - 1;                // widest union window the prefix table must cover
    static constexpr int CNX = 8 + UL + 9;                // guarded prefix table entries, index d+8
    static constexpr size_t SMEM = (size_t)NG * 2 * TWP * 16 + (size_t)NG * TWP * 8 + (size_t)CNX * 8;
};

struct NccParams {
    int H, W, WW, seg_rows;
    double st2, hw;
    const uint32_t *area_bits; const uint32_t *area_count;
    uint32_t *mask_bits;
    const float *thr_lut; const double *cn64;      // cn64[d+8] = Cn[clamp(d,0,TL)]
    int2 *recheck; uint32_t *recheck_n; int recheck_cap;
    uint32_t *status;
};

__device__ __forceinline__ uint32_t ld_bits(const uint32_t *row, int wi, int WW) {
    return (wi >= 0 && wi < WW) ? __ldg(row + wi) : 0u;
}

// threshold on G for a pixel whose window is clipped by the image border (float64)
template <int TL>
__device__ double border_threshold(int y, int x, int H, int W, double S, double m, double st2, const double *cn) {
    using G = Geo<TL>;
    const int ylo = max(0, G::OFF - y), yhi = min(TL - 1, H - 1 - y + G::OFF);
    const int xlo = max(0, G::OFF - x), xhi = min(TL - 1, W - 1 - x + G::OFF);
    const double A = (double)(yhi - ylo + 1) * (double)(xhi - xlo + 1);
    const double g1 = (cn[yhi + 1 + 8] - cn[ylo + 8]) * (cn[xhi + 1 + 8] - cn[xlo + 8]);
    const double L2 = (double)(TL * TL);
    const double q = S * (L2 - S) - 2.0 * m * S * (L2 - A) + m * m * A * (L2 - A);
    if (!(q > 0.0)) return INFINITY;
    return m * g1 + (S - m * A) / L2 + 0.1 * sqrt(st2 * q) / (double)TL;
}

// 256 threads per 128-pixel strip.  Horizontal role: warp = row of the step, lane = pixel quad.
// Vertical role: thread = (column, half); half h owns output rows 4h..4h+3 of the 8-row step and
// reads the ring one 4-row unit later, so both halves run the same unrolled code.
constexpr int NT = 256;
constexpr int HPX = 4;      // pixels per thread in the horizontal pass
constexpr int VR = 4;       // output rows per thread in the vertical pass

template <int TL>
__global__ void __launch_bounds__(NT, 3) ncc_mask_kernel(NccParams P) {
    using G = Geo<TL>;
    constexpr int UL = HPX + TL - 1;                       // union window of 4 adjacent pixels (bits)
    constexpr int NU = (VR + TL - 1 + 3) / 4;              // 4-row ring units one half reads
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *ringH = reinterpret_cast<float4 *>(smem_raw);                        // [NG*2][TWP]
    uint2 *ringB = reinterpret_cast<uint2 *>(ringH + G::NG * 2 * TWP);           // [NG][TWP]
    double *cn = reinterpret_cast<double *>(ringB + G::NG * TWP);                // [CNX]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * TW;
    const int ys = blockIdx.y * P.seg_rows;
    const int ye = min(P.H, ys + P.seg_rows);
    const int f = blockIdx.z;
    const int H = P.H, W = P.W, WW = P.WW;
    const uint32_t *abits = P.area_bits + (size_t)f * H * WW;
    const double mfrac = (double)P.area_count[f] / P.hw;          // mean(area_mask)/255
    for (int i = tid; i < G::CNX; i += NT) cn[i] = P.cn64[i];

    const int nk = (ye - ys + RB - 1) / RB;
    const int nsteps = nk + G::LEAD;
    const int hr = warp;                                            // horizontal role: row of the step
    const int hcol = HPX * lane;                                    // strip-relative first column
    const int sb = x0 + hcol - G::OFF;                              // first bit of the union window
    const int wi0 = sb >> 5, bo = sb & 31;                          // arithmetic shift: floor
    const int vcol = tid & (TW - 1), vhalf = tid >> 7;              // vertical role

    uint32_t pw[4];
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

    fetch(0);
    __syncthreads();
    int gw = 0;                                  // ring group written by step m (m % NG)
    for (int m = 0; m < nsteps; ++m) {
        // ---- horizontal pass on bits -------------------------------------------------------------
        {
            uint32_t U0 = __funnelshift_r(pw[0], pw[1], bo);
            uint32_t U1 = __funnelshift_r(pw[1], pw[2], bo);
            uint32_t U2 = __funnelshift_r(pw[2], pw[3], bo);
            if (m + 1 < nsteps) fetch(m + 1);
            if constexpr (UL <= 64) { U2 = 0; U1 &= (UL == 64) ? 0xffffffffu : ((1u << (UL - 32)) - 1u); }
            else { U2 &= (1u << (UL - 64)) - 1u; }
            // transitions: bit t set when bit(t) != bit(t-1)  (bit(-1) = 0, bit(UL) = 0)
            uint32_t T0 = U0 ^ (U0 << 1);
            uint32_t T1 = U1 ^ __funnelshift_l(U0, U1, 1);
            uint32_t T2 = U2 ^ __funnelshift_l(U1, U2, 1);
            double acc[HPX];
#pragma unroll
            for (int j = 0; j < HPX; ++j) acc[j] = 0.0;
            auto consume = [&](uint32_t T, uint32_t U, int base) {
                while (T) {
                    const int b = __ffs(T) - 1;
                    T &= T - 1;
                    const bool is_start = (U >> b) & 1u;
                    const double *c = cn + (base + b + 8);           // cn[d + 8], d = t - j
#pragma unroll
                    for (int j = 0; j < HPX; ++j) {
                        const double v = c[-j];
                        acc[j] += is_start ? -v : v;
                    }
                }
            };
            consume(T0, U0, 0);
            consume(T1, U1, 32);
            if constexpr (UL > 64) consume(T2, U2, 64);
            auto bit = [&](int t) -> uint32_t {
                return t < 32 ? (U0 >> t) & 1u : t < 64 ? (U1 >> (t - 32)) & 1u : (U2 >> (t - 64)) & 1u;
            };
            uint32_t hb[HPX];
            if constexpr (TL >= 64) hb[0] = __popc(U0) + __popc(U1) + __popc(U2 & ((1u << (TL - 64)) - 1u));
            else hb[0] = __popc(U0) + __popc(U1 & ((1u << (TL - 32)) - 1u));
#pragma unroll
            for (int j = 1; j < HPX; ++j) hb[j] = hb[j - 1] + bit(j - 1 + TL) - bit(j - 1);
            // ring stores: H as float4 units [group*2 + hr/4][col + col/8].f[hr%4]; box rows as bytes
            float *dstH = reinterpret_cast<float *>(ringH + (gw * 2 + (hr >> 2)) * TWP) + (hr & 3);
            unsigned char *dstB = reinterpret_cast<unsigned char *>(ringB + gw * TWP) + hr;
#pragma unroll
            for (int j = 0; j < HPX; ++j) {
                const int c = hcol + j, cp = c + (c >> 3);
                dstH[cp * 4] = (float)acc[j];
                dstB[cp * 8] = (unsigned char)hb[j];
            }
        }
        __syncthreads();

        // ---- vertical pass ---------------------------------------------------------------------------
        if (m >= G::LEAD) {
            const int k = m - G::LEAD;
            const int yb = ys + RB * k;
            const int cp = vcol + (vcol >> 3);
            int g0 = gw + 1; if (g0 >= G::NG) g0 -= G::NG;           // group of step k
            // box sums of all 8 rows of the step (cheap; both halves run the same chain, no exchange)
            int S[RB];
            auto box_at = [&](auto I_) -> int {
                constexpr int idx = decltype(I_)::value;
                int gi = g0 + idx / 8; if (gi >= G::NG) gi -= G::NG;
                const uint2 b = ringB[gi * TWP + cp];
                const uint32_t w = (idx % 8) < 4 ? b.x : b.y;
                return (int)((w >> (8 * (idx % 4))) & 255u);
            };
            if (k == 0) {
                int s = 0;
                static_for<0, TL>([&](auto I_) { s += box_at(I_); });
                S[0] = s;
            } else {
                S[0] = s_prev + box_at(std::integral_constant<int, TL - 1>{}) - (int)hb_m1;
            }
            static_for<1, RB>([&](auto R_) {
                constexpr int r = decltype(R_)::value;
                S[r] = S[r - 1] + box_at(std::integral_constant<int, r + TL - 1>{}) - box_at(std::integral_constant<int, r - 1>{});
            });
            s_prev = S[RB - 1]; hb_m1 = (uint32_t)box_at(std::integral_constant<int, RB - 1>{});
            // Gaussian column sums of this half: G(yb + 4h + r) = sum_a n[a] h[4h + r + a]
            float acc[VR];
#pragma unroll
            for (int r = 0; r < VR; ++r) acc[r] = 0.f;
            int u0 = 2 * g0 + vhalf; if (u0 >= 2 * G::NG) u0 -= 2 * G::NG;
            static_for<0, NU>([&](auto U_) {
                constexpr int u = decltype(U_)::value;
                int ui = u0 + u; if (ui >= 2 * G::NG) ui -= 2 * G::NG;
                const float4 v = ringH[ui * TWP + cp];
                const float hv[4] = {v.x, v.y, v.z, v.w};
                static_for<0, 4>([&](auto E_) {
                    constexpr int t = 4 * u + decltype(E_)::value;
                    static_for<0, VR>([&](auto R_) {
                        constexpr int r = decltype(R_)::value;
                        constexpr int a = t - r;
                        if constexpr (a >= 0 && a < TL) acc[r] = fmaf(c_n32[TL == 80][a], hv[t - 4 * u], acc[r]);
                    });
                });
            });
            // decision
            const int x = x0 + vcol;
            const bool xin = x >= G::OFF && x + G::HI < W;
            uint32_t myword = 0;
#pragma unroll
            for (int r = 0; r < VR; ++r) {
                const int y = yb + VR * vhalf + r;
                const int Sr = vhalf ? S[VR + r] : S[r];
                bool on = false;
                if (x < W && y < ye) {
                    float thr;
                    if (xin && y >= G::OFF && y + G::HI < H) thr = __ldg(P.thr_lut + Sr);
                    else thr = (float)border_threshold<TL>(y, x, H, W, (double)Sr, mfrac, P.st2, cn);
                    const float d = acc[r] - thr;
                    on = d > 0.f;
                    if (fabsf(d) <= BAND) {          // float32 cannot decide: queue for the float64 pass
                        on = false;
                        const uint32_t slot = atomicAdd(P.recheck_n + f, 1u);
                        if (slot < (uint32_t)P.recheck_cap) P.recheck[(size_t)f * P.recheck_cap + slot] = make_int2(x, y);
                        else atomicOr(P.status, VBS_DEV_RECHECK_OVERFLOW);
                    }
                }
                const uint32_t word = __ballot_sync(0xffffffffu, on);
                if (lane == r) myword = word;
            }
            const int wx = (x0 >> 5) + (vcol >> 5);
            const int yw = yb + VR * vhalf + lane;
            if (lane < VR && yw < ye && wx < WW) P.mask_bits[((size_t)f * H + yw) * WW + wx] = myword;
        }
        __syncthreads();
        if (++gw == G::NG) gw = 0;
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
    using G = Geo<TL>;
    NccParams P;
    P.H = ctx->H; P.W = ctx->W; P.WW = ctx->WW;
    const int strips = (ctx->W + TW - 1) / TW;
    int vsegs = 1;
    while ((long long)strips * batch * vsegs < 4096 && (ctx->H + vsegs) / (vsegs + 1) >= 64) ++vsegs;
    P.seg_rows = ((ctx->H + vsegs - 1) / vsegs + RB - 1) / RB * RB;
    vsegs = (ctx->H + P.seg_rows - 1) / P.seg_rows;
    P.st2 = st2; P.hw = (double)ctx->H * (double)ctx->W;
    P.area_bits = ctx->area_bits; P.area_count = ctx->area_count; P.mask_bits = ctx->mask_bits;
    P.thr_lut = ctx->thr_lut; P.cn64 = ctx->d_cn64;
    P.recheck = ctx->recheck; P.recheck_n = ctx->recheck_n; P.recheck_cap = ctx->recheck_cap;
    P.status = ctx->d_status;
    cudaError_t e = cudaMemsetAsync(ctx->recheck_n, 0, sizeof(uint32_t) * batch, ctx->stream);
    if (e != cudaSuccess) return e;
    auto kern = ncc_mask_kernel<TL>;
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM)) != cudaSuccess) return e;
    kern<<<dim3(strips, vsegs, batch), NT, G::SMEM, ctx->stream>>>(P);
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
    for (int i = 0; i < TL; ++i) { n[i] = e[i] / sum; ss += n[i] * n[i]; n32[i] = (float)n[i]; }
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
    if ((err = cudaMemcpy(ctx->d_n64, n, sizeof(double) * TL, cudaMemcpyHostToDevice)) != cudaSuccess) { delete[] lut; return err; }
    if ((err = cudaMemcpy(ctx->d_cn64, cn, sizeof(double) * CNX, cudaMemcpyHostToDevice)) != cudaSuccess) { delete[] lut; return err; }
    err = cudaMemcpy(ctx->thr_lut, lut, sizeof(float) * NL, cudaMemcpyHostToDevice);
    delete[] lut;
    if (err != cudaSuccess) return err;
    return cudaMemcpyToSymbol(c_n32, n32, sizeof(float) * 96, sizeof(float) * 96 * (TL == 80 ? 1 : 0));
}

cudaError_t vbs_launch_ncc(vbs_ctx *ctx, int batch) {
    if (ctx->big) return launch<80>(ctx, batch, ctx->st2);
    return launch<33>(ctx, batch, ctx->st2);
}
