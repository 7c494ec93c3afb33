// k_ccl.cu - K3b/K4b: connected components on bit images by union-find over WORD SEGMENTS.
//
// A segment is a maximal run of set bits inside one 32-bit word; its id is the pixel index of
// its first bit, so ids live in a dense [H*W] int32 array that is only ever touched at segment
// starts (sparse traffic, no clearing pass).  Links always point to the smaller index, hence a
// component's root is its first pixel in raster order - exactly what the reference needs:
//   * scipy.ndimage.label numbers 4-connected components by first pixel (MD:176)  -> rank roots
//   * cv2.findContours(RETR_EXTERNAL) starts each outer border at the blob's topmost-leftmost
//     pixel and lists contours in descending start order (MD:196)     -> slot = n-1-rank
// The opened image is labelled twice in the same array: foreground 8-connected and background
// 4-connected (disjoint index sets); background touching the frame is united with a virtual
// "outside" (-1), so a blob is external iff the background left of its start pixel reaches -1.
#include "vbs_ctx.h"

namespace {

constexpr int OUTSIDE = -1;

__device__ __forceinline__ uint32_t valid_mask(int wx, int W) {
    const int rem = W - 32 * wx;
    return rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
}
template <bool INV>
__device__ __forceinline__ uint32_t get_bits(const uint32_t *img, int y, int wx, int W, int WW) {
    const uint32_t w = __ldg(img + (size_t)y * WW + wx);
    return INV ? (~w & valid_mask(wx, W)) : w;
}
// first bit of the run of set bits that contains set bit b
__device__ __forceinline__ int seg_start(uint32_t w, int b) {
    const uint32_t t = ~w & ((2u << b) - 1u);
    return t ? 32 - __clz(t) : 0;
}
__device__ __forceinline__ uint32_t run_mask(uint32_t w, int s) {      // run of set bits starting at s
    const uint32_t t = ~w >> s;
    const int len = t ? __ffs(t) - 1 : 32 - s;
    return (len >= 32 ? 0xffffffffu : ((1u << len) - 1u)) << s;
}
__device__ __forceinline__ int find_root(const int32_t *par, int i) {
    while (i >= 0) {
        const int p = par[i];
        if (p == i) break;
        i = p;
    }
    return i;          // root index, or a negative terminal (OUTSIDE / encoded label)
}
// find with path halving (every visited node is re-pointed at its grandparent; links only ever
// move towards smaller indices, so concurrent atomicMin links stay consistent)
__device__ __forceinline__ int find_root_halving(int32_t *par, int i) {
    if (i < 0) return i;
    int p = par[i];
    while (p != i) {
        if (p < 0) return p;
        const int g = par[p];
        if (g == p) return p;
        if (g < 0) return g;
        par[i] = g;
        i = g;
        p = par[i];
    }
    return i;
}
__device__ void unite(int32_t *par, int a, int b) {
    for (;;) {
        a = find_root_halving(par, a);
        b = find_root_halving(par, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(par + a, b);      // a >= 0 here (a > b >= -1)
        if (old == a) return;
        a = old;
    }
}

// word kernels: block (64, 4) = 64 words x 4 rows, grid (ceil(WW/64), ceil(H/4), batch)
#define VBS_WORD_COORDS                                          \
    const int wx = blockIdx.x * 64 + threadIdx.x;               \
    const int y = blockIdx.y * 4 + threadIdx.y;                 \
    const int f = blockIdx.z;                                   \
    if (wx >= WW || y >= H) return;

// ---- 1. every segment start becomes its own root ------------------------------------------------
// FG: set bits.  BG (conditional on holes[f] != 0): cleared bits inside the image.
template <bool FG, bool BG>
__global__ void __launch_bounds__(256) ccl_init_kernel(const uint32_t *__restrict__ bits, int32_t *__restrict__ parent,
                                                        const int32_t *__restrict__ holes, int H, int W, int WW) {
    VBS_WORD_COORDS
    const uint32_t raw = __ldg(bits + ((size_t)f * H + y) * WW + wx);
    int32_t *par = parent + (size_t)f * H * W;
    const int base = y * W + 32 * wx;
    if (FG) {
        uint32_t starts = raw & ~(raw << 1);
        while (starts) { const int s = __ffs(starts) - 1; starts &= starts - 1; par[base + s] = base + s; }
    }
    if (BG && holes[f] != 0) {
        const uint32_t w = ~raw & valid_mask(wx, W);
        uint32_t starts = w & ~(w << 1);
        while (starts) { const int s = __ffs(starts) - 1; starts &= starts - 1; par[base + s] = base + s; }
    }
}

// ---- 2. link segments that touch -----------------------------------------------------------------
template <bool CONN8, bool INV, bool BORDER>
__global__ void __launch_bounds__(256) ccl_merge_kernel(const uint32_t *__restrict__ bits, int32_t *__restrict__ parent,
                                                         const int32_t *__restrict__ holes, int H, int W, int WW) {
    VBS_WORD_COORDS
    if (INV && holes[f] == 0) return;          // no hole anywhere in this frame: background labels are not needed
    const uint32_t *img = bits + (size_t)f * H * WW;
    const uint32_t w = get_bits<INV>(img, y, wx, W, WW);
    if (!w) return;
    int32_t *par = parent + (size_t)f * H * W;
    const int base = y * W + 32 * wx;
    if ((w & 1u) && wx > 0) {
        const uint32_t pw = get_bits<INV>(img, y, wx - 1, W, WW);
        if (pw >> 31) unite(par, base, base - 32 + seg_start(pw, 31));
    }
    const uint32_t up = y > 0 ? get_bits<INV>(img, y - 1, wx, W, WW) : 0u;
    const uint32_t upl = (CONN8 && y > 0 && wx > 0) ? get_bits<INV>(img, y - 1, wx - 1, W, WW) : 0u;
    const uint32_t upr = (CONN8 && y > 0 && wx + 1 < WW) ? get_bits<INV>(img, y - 1, wx + 1, W, WW) : 0u;
    const int last_bit = (W - 1) - 32 * wx;         // position of pixel W-1 in this word (may be >= 32)
    uint32_t rem = w;
    while (rem) {
        const int s = __ffs(rem) - 1;
        const uint32_t seg = run_mask(w, s);
        rem &= ~seg;
        const int id = base + s;
        uint32_t nb = seg;
        if (CONN8) nb |= (seg << 1) | (seg >> 1);
        uint32_t ov = up & nb;
        while (ov) {
            const int b = __ffs(ov) - 1;
            const int us = seg_start(up, b);
            ov &= ~run_mask(up, us);
            unite(par, id, base - W + us);
        }
        if (CONN8) {
            if ((seg & 1u) && (upl >> 31)) unite(par, id, base - W - 32 + seg_start(upl, 31));
            if ((seg >> 31) && (upr & 1u)) unite(par, id, base - W + 32);
        }
        if (BORDER) {
            const bool touches = y == 0 || y == H - 1 || (wx == 0 && (seg & 1u)) ||
                                 (last_bit >= 0 && last_bit < 32 && ((seg >> last_bit) & 1u));
            if (touches) unite(par, id, OUTSIDE);
        }
    }
}

// ---- 3. flatten: every segment points at its root; mark roots (root bits + per-row counts) --------
__global__ void __launch_bounds__(256) ccl_roots_kernel(const uint32_t *__restrict__ bits, int32_t *__restrict__ parent,
                                                         uint32_t *__restrict__ root_bits, int32_t *__restrict__ rowcnt,
                                                         int H, int W, int WW) {
    VBS_WORD_COORDS
    const size_t i = ((size_t)f * H + y) * WW + wx;
    const uint32_t w = __ldg(bits + i);
    int32_t *par = parent + (size_t)f * H * W;
    const int base = y * W + 32 * wx;
    uint32_t roots = 0;
    uint32_t starts = w & ~(w << 1);
    while (starts) {
        const int s = __ffs(starts) - 1;
        starts &= starts - 1;
        const int r = find_root(par, base + s);
        if (r == base + s) roots |= 1u << s;
        else par[base + s] = r;
    }
    root_bits[i] = roots;
    if (roots) atomicAdd(rowcnt + (size_t)f * H + y, __popc(roots));
}
// background segments of frames with holes: point straight at the root (or at OUTSIDE)
__global__ void __launch_bounds__(256) ccl_flatten_bg_kernel(const uint32_t *__restrict__ bits, int32_t *__restrict__ parent,
                                                              const int32_t *__restrict__ holes, int H, int W, int WW) {
    VBS_WORD_COORDS
    if (holes[f] == 0) return;
    const uint32_t bg = ~__ldg(bits + ((size_t)f * H + y) * WW + wx) & valid_mask(wx, W);
    int32_t *par = parent + (size_t)f * H * W;
    const int base = y * W + 32 * wx;
    uint32_t bs = bg & ~(bg << 1);
    while (bs) {
        const int s = __ffs(bs) - 1;
        bs &= bs - 1;
        const int r = find_root(par, base + s);
        if (r != base + s) par[base + s] = r;
    }
}

// ---- 3b. Euler number of the 8-connected foreground by bit quads ----------------------------------
// 4 E = #Q1 - #Q3 - 2 #QD over all 2x2 windows of the zero-padded image (Gray's formula).  The
// number of holes of the whole image is (#blobs - E); when it is 0 no blob can lie inside
// another one, so the background labelling that RETR_EXTERNAL would need is skipped for the frame.
__global__ void __launch_bounds__(256) euler_kernel(const uint32_t *__restrict__ bits, int32_t *__restrict__ euler4, int H, int W, int WW) {
    const int wx = blockIdx.x * 64 + threadIdx.x;            // 0 .. WW (one extra all-zero column)
    const int yy = (int)(blockIdx.y * 4 + threadIdx.y) - 1;  // top row of the quad: -1 .. H-1
    const int f = blockIdx.z;
    int v = 0;
    if (wx <= WW && yy < H) {
        const uint32_t *img = bits + (size_t)f * H * WW;
        auto ld = [&](int y, int w) -> uint32_t { return (y >= 0 && y < H && w >= 0 && w < WW) ? __ldg(img + (size_t)y * WW + w) : 0u; };
        const uint32_t b = ld(yy, wx), d = ld(yy + 1, wx);
        const uint32_t a = (b << 1) | (ld(yy, wx - 1) >> 31), c = (d << 1) | (ld(yy + 1, wx - 1) >> 31);
        const uint32_t x1 = a ^ b, x2 = c ^ d, n1 = a & b, n2 = c & d;
        const uint32_t q1 = (x1 & ~x2 & ~n2) | (x2 & ~x1 & ~n1);
        const uint32_t q3 = (x1 & n2) | (x2 & n1);
        const uint32_t qd = x1 & x2 & ~(a ^ d);
        v = __popc(q1) - __popc(q3) - 2 * __popc(qd);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (((threadIdx.y * 64 + threadIdx.x) & 31) == 0 && v) atomicAdd(euler4 + f, v);
}

__global__ void holes_kernel(const int32_t *__restrict__ euler4, const int32_t *__restrict__ ncont, int32_t *__restrict__ holes, int batch) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < batch) holes[f] = ncont[f] - euler4[f] / 4;
}

// ---- 4. exclusive scan of the per-row root counts (one CTA per frame) ----------------------------
__global__ void __launch_bounds__(1024) row_scan_kernel(const int32_t *__restrict__ rowcnt, int32_t *__restrict__ rowoff,
                                                         int32_t *__restrict__ total, int H) {
    __shared__ int32_t warp_sum[32];
    __shared__ int32_t carry_s;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t *cnt = rowcnt + (size_t)f * H;
    int32_t *off = rowoff + (size_t)f * H;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int y0 = 0; y0 < H; y0 += 1024) {
        const int y = y0 + tid;
        const int v = y < H ? cnt[y] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int ws = warp_sum[lane], wi = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
            warp_sum[lane] = wi - ws;          // exclusive prefix of warp totals
        }
        __syncthreads();
        const int base = carry_s + warp_sum[warp];
        if (y < H) off[y] = base + incl - v;
        __syncthreads();
        if (tid == 1023) carry_s = base + incl;
        __syncthreads();
    }
    if (tid == 0) total[f] = carry_s;
}

// ---- 5. rank roots in raster order (one warp per row) --------------------------------------------
// MODE 0 (ring maxima): parent[root] = -2 - label.       MODE 1 (contours): croot[n-1-rank] = root.
template <int MODE>
__global__ void rank_roots_kernel(const uint32_t *__restrict__ root_bits, const int32_t *__restrict__ rowoff,
                                  const int32_t *__restrict__ total, int32_t *__restrict__ parent, int32_t *__restrict__ croot,
                                  int H, int W, int WW, int M, size_t nrows, uint32_t *status) {
    const size_t row = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const int lane = threadIdx.x & 31;
    const int y = (int)(row % H);
    const size_t f = row / H;
    int carry = rowoff[row];
    const int n = total[f];
    if (n > M && lane == 0 && y == 0) atomicOr(status, MODE == 0 ? VBS_DEV_LABEL_OVERFLOW : VBS_DEV_CONTOUR_OVERFLOW);
    for (int w0 = 0; w0 < WW; w0 += 32) {
        const int wx = w0 + lane;
        uint32_t r = wx < WW ? root_bits[row * WW + wx] : 0u;
        const int c = __popc(r);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        int rank = carry + incl - c;
        while (r) {
            const int b = __ffs(r) - 1;
            r &= r - 1;
            const int idx = y * W + 32 * wx + b;
            if (MODE == 0) { if (rank < M) parent[f * (size_t)H * W + idx] = -2 - rank; }
            else { const int slot = n - 1 - rank; if (slot < M) croot[f * (size_t)M + slot] = idx; }
            ++rank;
        }
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// ---- 6. ring components: integer moments per label (MD:181 center_of_mass on a 0/1 mask) ---------
__global__ void __launch_bounds__(256) moments_kernel(const uint32_t *__restrict__ bits, const int32_t *__restrict__ parent,
                                                       uint32_t *__restrict__ cnt, unsigned long long *__restrict__ sx,
                                                       unsigned long long *__restrict__ sy, int H, int W, int WW, int M) {
    VBS_WORD_COORDS
    const uint32_t w = __ldg(bits + ((size_t)f * H + y) * WW + wx);
    if (!w) return;
    const int32_t *par = parent + (size_t)f * H * W;
    const int base = y * W + 32 * wx;
    uint32_t rem = w;
    while (rem) {
        const int s = __ffs(rem) - 1;
        const uint32_t seg = run_mask(w, s);
        rem &= ~seg;
        int p = par[base + s];
        if (p >= 0) p = par[p];                 // non-root segments point straight at their root
        if (p >= 0) continue;                   // root beyond capacity (flagged elsewhere)
        const int label = -2 - p;
        if (label < 0 || label >= M) continue;
        const unsigned len = __popc(seg);
        const unsigned long long xs = (unsigned long long)len * (unsigned)(32 * wx + s) + (unsigned long long)len * (len - 1) / 2;
        atomicAdd(cnt + (size_t)f * M + label, len);
        atomicAdd(sx + (size_t)f * M + label, xs);
        atomicAdd(sy + (size_t)f * M + label, (unsigned long long)len * (unsigned)y);
    }
}

__global__ void centres_kernel(const uint32_t *__restrict__ cnt, const unsigned long long *__restrict__ sx,
                               const unsigned long long *__restrict__ sy, const int32_t *__restrict__ total,
                               double *__restrict__ centres, int M, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t f = i / M;
    const int l = (int)(i % M);
    if (l >= total[f]) return;
    const double c = (double)cnt[i];
    centres[2 * i + 0] = (double)sy[i] / c;     // row: exact integer sum, one float64 division
    centres[2 * i + 1] = (double)sx[i] / c;     // col
}

}  // namespace

cudaError_t vbs_launch_components(vbs_ctx *ctx, int batch) {
    const int H = ctx->H, W = ctx->W, WW = ctx->WW, M = ctx->M;
    const size_t nrows = (size_t)batch * H;
    const dim3 wb(64, 4);
    const dim3 wg((WW + 63) / 64, (H + 3) / 4, batch);
    const dim3 eg((WW + 1 + 63) / 64, (H + 1 + 3) / 4, batch);
    const unsigned gr = (unsigned)((nrows + 7) / 8);
    cudaStream_t st = ctx->stream;
    cudaError_t e;
    // ---- ring maxima: 4-connected, labels in raster order, centroids -------------------------------
    if ((e = cudaMemsetAsync(ctx->rowcnt, 0, sizeof(int32_t) * nrows, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(ctx->lab_cnt, 0, sizeof(uint32_t) * (size_t)batch * M, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(ctx->lab_sx, 0, sizeof(unsigned long long) * (size_t)batch * M, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(ctx->lab_sy, 0, sizeof(unsigned long long) * (size_t)batch * M, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(ctx->euler4, 0, sizeof(int32_t) * batch, st)) != cudaSuccess) return e;
    ccl_init_kernel<true, false><<<wg, wb, 0, st>>>(ctx->max_bits, ctx->parent, nullptr, H, W, WW);
    ccl_merge_kernel<false, false, false><<<wg, wb, 0, st>>>(ctx->max_bits, ctx->parent, nullptr, H, W, WW);
    ccl_roots_kernel<<<wg, wb, 0, st>>>(ctx->max_bits, ctx->parent, ctx->root_bits, ctx->rowcnt, H, W, WW);
    row_scan_kernel<<<batch, 1024, 0, st>>>(ctx->rowcnt, ctx->rowoff, ctx->d_nlabels, H);
    rank_roots_kernel<0><<<gr, 256, 0, st>>>(ctx->root_bits, ctx->rowoff, ctx->d_nlabels, ctx->parent, nullptr, H, W, WW, M, nrows, ctx->d_status);
    moments_kernel<<<wg, wb, 0, st>>>(ctx->max_bits, ctx->parent, ctx->lab_cnt, ctx->lab_sx, ctx->lab_sy, H, W, WW, M);
    centres_kernel<<<(unsigned)(((size_t)batch * M + 255) / 256), 256, 0, st>>>(ctx->lab_cnt, ctx->lab_sx, ctx->lab_sy, ctx->d_nlabels,
                                                                                ctx->centres, M, (size_t)batch * M);
    // ---- opened area mask: foreground 8-connected; background 4-connected only in frames with holes --
    if ((e = cudaMemsetAsync(ctx->rowcnt, 0, sizeof(int32_t) * nrows, st)) != cudaSuccess) return e;
    ccl_init_kernel<true, false><<<wg, wb, 0, st>>>(ctx->open_bits, ctx->parent2, nullptr, H, W, WW);
    ccl_merge_kernel<true, false, false><<<wg, wb, 0, st>>>(ctx->open_bits, ctx->parent2, nullptr, H, W, WW);
    euler_kernel<<<eg, wb, 0, st>>>(ctx->open_bits, ctx->euler4, H, W, WW);
    ccl_roots_kernel<<<wg, wb, 0, st>>>(ctx->open_bits, ctx->parent2, ctx->root_bits, ctx->rowcnt, H, W, WW);
    row_scan_kernel<<<batch, 1024, 0, st>>>(ctx->rowcnt, ctx->rowoff, ctx->d_ncont, H);
    rank_roots_kernel<1><<<gr, 256, 0, st>>>(ctx->root_bits, ctx->rowoff, ctx->d_ncont, nullptr, ctx->croot, H, W, WW, M, nrows, ctx->d_status);
    holes_kernel<<<(batch + 127) / 128, 128, 0, st>>>(ctx->euler4, ctx->d_ncont, ctx->holes, batch);
    // background pass: every thread of a hole-free frame returns at once
    ccl_init_kernel<false, true><<<wg, wb, 0, st>>>(ctx->open_bits, ctx->parent2, ctx->holes, H, W, WW);
    ccl_merge_kernel<false, true, true><<<wg, wb, 0, st>>>(ctx->open_bits, ctx->parent2, ctx->holes, H, W, WW);
    ccl_flatten_bg_kernel<<<wg, wb, 0, st>>>(ctx->open_bits, ctx->parent2, ctx->holes, H, W, WW);
    ctx->launches += 17;
    return cudaGetLastError();
}
