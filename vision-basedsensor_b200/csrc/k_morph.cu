// k_morph.cu - K3a/K4a on 1 bit/pixel images (32 pixels per word, funnel shifts + AND/OR):
//   maxima  = mask & ~(every in-image pixel of the [-n/2, n/2-1]^2 window is set)      MD:170-174
//             (== (mask == maximum_filter) & ((max - min) > 0) for a 0/1 mask)
//   opened  = dilate5x5(erode5x5(area_mask)), erode: outside = 1, dilate: outside = 0    MD:194-195
// plus pack (uint8 image -> bits) and unpack (bits / labels -> uint8 / int32 stage images).
#include "vbs_ctx.h"

namespace {

constexpr int TH = 32;     // output rows per CTA
constexpr int TY = 8;      // thread rows

// bits [x+lo .. x+hi] ANDed for every pixel x of `cur`  (lo <= 0 <= hi, |lo|,|hi| < 32)
template <int LO, int HI>
__device__ __forceinline__ uint32_t hand(uint32_t prev, uint32_t cur, uint32_t next) {
    uint32_t r = cur;
#pragma unroll
    for (int d = LO; d <= HI; ++d) {
        if (d < 0) r &= __funnelshift_r(prev, cur, 32 + d);
        else if (d > 0) r &= __funnelshift_r(cur, next, d);
    }
    return r;
}
template <int LO, int HI>
__device__ __forceinline__ uint32_t hor(uint32_t prev, uint32_t cur, uint32_t next) {
    uint32_t r = cur;
#pragma unroll
    for (int d = LO; d <= HI; ++d) {
        if (d < 0) r |= __funnelshift_r(prev, cur, 32 + d);
        else if (d > 0) r |= __funnelshift_r(cur, next, d);
    }
    return r;
}

__device__ __forceinline__ uint32_t valid_mask(int wx, int W) {
    const int rem = W - 32 * wx;
    return rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
}

// word of an image row with "outside = 1" semantics
__device__ __forceinline__ uint32_t ld_ones(const uint32_t *img, int y, int wx, int H, int W, int WW) {
    if (y < 0 || y >= H || wx < 0 || wx >= WW) return 0xffffffffu;
    return __ldg(img + (size_t)y * WW + wx) | ~valid_mask(wx, W);
}

template <int NB>
__global__ void __launch_bounds__(32 * TY) maxima_kernel(const uint32_t *__restrict__ mask_bits, uint32_t *__restrict__ max_bits,
                                                          int H, int W, int WW) {
    constexpr int LO = -(NB / 2), HI = NB - 1 - NB / 2;
    constexpr int ROWS = TH + NB - 1;
    __shared__ uint32_t eh[ROWS][33];
    const int f = blockIdx.z;
    const uint32_t *img = mask_bits + (size_t)f * H * WW;
    const int wx = blockIdx.x * 32 + threadIdx.x;
    const int y0 = blockIdx.y * TH;
    for (int r = threadIdx.y; r < ROWS; r += TY) {
        const int y = y0 + LO + r;
        eh[r][threadIdx.x] = hand<LO, HI>(ld_ones(img, y, wx - 1, H, W, WW), ld_ones(img, y, wx, H, W, WW),
                                          ld_ones(img, y, wx + 1, H, W, WW));
    }
    __syncthreads();
    if (wx >= WW) return;
    for (int r = threadIdx.y; r < TH; r += TY) {
        const int y = y0 + r;
        if (y >= H) break;
        uint32_t e = 0xffffffffu;
#pragma unroll
        for (int d = 0; d < NB; ++d) e &= eh[r + d][threadIdx.x];
        const uint32_t m = __ldg(img + (size_t)y * WW + wx);
        max_bits[((size_t)f * H + y) * WW + wx] = m & ~e;
    }
}

// 30 output words per CTA row + one halo word on each side (lanes 0 and 31)
__global__ void __launch_bounds__(32 * TY) open5_kernel(const uint32_t *__restrict__ area_bits, uint32_t *__restrict__ open_bits,
                                                         int H, int W, int WW) {
    constexpr int RA = TH + 8, RE = TH + 4;
    __shared__ uint32_t a[RA][33];     // horizontal erode, rows y0-4 .. y0+TH+3
    __shared__ uint32_t e[RE][33];     // eroded image,     rows y0-2 .. y0+TH+1
    __shared__ uint32_t dh[RE][33];    // horizontal dilate of the eroded image
    const int f = blockIdx.z;
    const uint32_t *img = area_bits + (size_t)f * H * WW;
    const int lane = threadIdx.x;
    const int wx = blockIdx.x * 30 + lane - 1;
    const int y0 = blockIdx.y * TH;
    for (int r = threadIdx.y; r < RA; r += TY) {
        const int y = y0 - 4 + r;
        a[r][lane] = hand<-2, 2>(ld_ones(img, y, wx - 1, H, W, WW), ld_ones(img, y, wx, H, W, WW), ld_ones(img, y, wx + 1, H, W, WW));
    }
    __syncthreads();
    for (int r = threadIdx.y; r < RE; r += TY) {
        const int y = y0 - 2 + r;
        uint32_t v = 0;
        if (y >= 0 && y < H && wx >= 0 && wx < WW) {
            v = a[r][lane] & a[r + 1][lane] & a[r + 2][lane] & a[r + 3][lane] & a[r + 4][lane];
            v &= valid_mask(wx, W);            // dilate treats everything outside the image as background
        }
        e[r][lane] = v;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < RE; r += TY) {
        const uint32_t p = lane > 0 ? e[r][lane - 1] : 0u, n = lane < 31 ? e[r][lane + 1] : 0u;
        dh[r][lane] = hor<-2, 2>(p, e[r][lane], n);
    }
    __syncthreads();
    if (lane == 0 || lane == 31 || wx >= WW) return;
    for (int r = threadIdx.y; r < TH; r += TY) {
        const int y = y0 + r;
        if (y >= H) break;
        const uint32_t v = dh[r][lane] | dh[r + 1][lane] | dh[r + 2][lane] | dh[r + 3][lane] | dh[r + 4][lane];
        open_bits[((size_t)f * H + y) * WW + wx] = v & valid_mask(wx, W);
    }
}

// uint8 image -> bits (nonzero = 1); one warp per output word
__global__ void pack_kernel(const uint8_t *__restrict__ src, uint32_t *__restrict__ dst, int H, int W, int WW, size_t nwords,
                            uint32_t *__restrict__ count = nullptr) {
    const size_t w = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= nwords) return;
    const int lane = threadIdx.x & 31;
    const int wx = (int)(w % WW);
    const size_t fy = w / WW;
    const int x = 32 * wx + lane;
    const bool on = x < W && src[fy * W + x] != 0;
    const uint32_t word = __ballot_sync(0xffffffffu, on);
    if (lane == 0) {
        dst[w] = word;
        if (count && word) atomicAdd(count + fy / H, (uint32_t)__popc(word));   // per-frame set pixels (mean of area_mask, MD:153)
    }
}

__global__ void unpack_kernel(const uint32_t *__restrict__ bits, uint8_t *__restrict__ dst, int W, int WW, size_t npx, uint8_t on_value) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npx) return;
    const int x = (int)(i % W);
    const size_t fy = i / W;
    dst[i] = ((bits[fy * WW + (x >> 5)] >> (x & 31)) & 1u) ? on_value : 0;
}

// label image: every maxima pixel -> 1 + raster rank of its component.  After k_ccl's passes
// parent[segment start] is the root index (>= 0) or the root's code -2 - slot; slot2label ranks the slots.
__global__ void unpack_labels_kernel(const uint32_t *__restrict__ bits, const int32_t *__restrict__ parent,
                                     const int32_t *__restrict__ slot2label, int32_t *__restrict__ dst, int H, int W, int WW, int M, size_t npx) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npx) return;
    const int x = (int)(i % W);
    const size_t fy = i / W;
    const size_t f = fy / H;
    const uint32_t w = bits[fy * WW + (x >> 5)];
    const int b = x & 31;
    int32_t out = 0;
    if ((w >> b) & 1u) {
        const uint32_t t = ~w & ((2u << b) - 1u);
        const int s = t ? 32 - __clz(t) : 0;
        const int32_t *par = parent + f * (size_t)H * W;
        const int idx = (int)((fy % H) * W) + (x - b + s);
        int32_t p = par[idx];
        for (int q = idx; p >= 0 && p != q; p = par[q]) q = p;     // segment -> tile root -> ... -> coded root
        const int slot = -2 - p;
        out = (p < 0 && slot < M) ? slot2label[f * M + slot] + 1 : 0;
    }
    dst[i] = out;
}

}  // namespace

// which: bit 0 = ring maxima of the NCC mask (needs K2), bit 1 = 5x5 open of the area mask (needs K1 only)
cudaError_t vbs_launch_morph(vbs_ctx *ctx, int batch, int which) {
    VbsRange range("vbs:morphology");
    const dim3 block(32, TY);
    if (which & 1) {
        const dim3 gm((ctx->WW + 31) / 32, (ctx->H + TH - 1) / TH, batch);
        if (ctx->br.nb == 14) maxima_kernel<14><<<gm, block, 0, ctx->stream>>>(ctx->mask_bits, ctx->max_bits, ctx->H, ctx->W, ctx->WW);
        else maxima_kernel<8><<<gm, block, 0, ctx->stream>>>(ctx->mask_bits, ctx->max_bits, ctx->H, ctx->W, ctx->WW);
        ctx->launches += 1;
    }
    if (which & 2) {
        const dim3 go((ctx->WW + 29) / 30, (ctx->H + TH - 1) / TH, batch);
        open5_kernel<<<go, block, 0, ctx->stream>>>(ctx->area_bits, ctx->open_bits, ctx->H, ctx->W, ctx->WW);
        ctx->launches += 1;
    }
    return cudaGetLastError();
}

cudaError_t vbs_launch_pack_masks(vbs_ctx *ctx, const uint8_t *mask, const uint8_t *area, int batch) {
    const size_t nwords = (size_t)batch * ctx->H * ctx->WW;
    const int wpb = 8;
    const unsigned grid = (unsigned)((nwords + wpb - 1) / wpb);
    pack_kernel<<<grid, wpb * 32, 0, ctx->stream>>>(mask, ctx->mask_bits, ctx->H, ctx->W, ctx->WW, nwords);
    pack_kernel<<<grid, wpb * 32, 0, ctx->stream>>>(area, ctx->area_bits, ctx->H, ctx->W, ctx->WW, nwords);
    ctx->launches += 2;
    return cudaGetLastError();
}

// area mask supplied by the caller -> area_bits + per-frame popcount, i.e. the state K1 leaves behind for K2
cudaError_t vbs_launch_pack_area(vbs_ctx *ctx, const uint8_t *area, int batch) {
    const size_t nwords = (size_t)batch * ctx->H * ctx->WW;
    const int wpb = 8;
    const unsigned grid = (unsigned)((nwords + wpb - 1) / wpb);
    cudaError_t e = cudaMemsetAsync(ctx->area_count, 0, sizeof(uint32_t) * batch, ctx->stream);
    if (e != cudaSuccess) return e;
    pack_kernel<<<grid, wpb * 32, 0, ctx->stream>>>(area, ctx->area_bits, ctx->H, ctx->W, ctx->WW, nwords, ctx->area_count);
    ctx->launches += 1;
    return cudaGetLastError();
}

cudaError_t vbs_launch_unpack(vbs_ctx *ctx, int stage, void *dst, int batch) {
    const size_t npx = (size_t)batch * ctx->H * ctx->W;
    const unsigned grid = (unsigned)((npx + 255) / 256);
    switch (stage) {
    case VBS_STAGE_AREA_MASK: unpack_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->area_bits, (uint8_t *)dst, ctx->W, ctx->WW, npx, 255); break;
    case VBS_STAGE_MASK: unpack_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->mask_bits, (uint8_t *)dst, ctx->W, ctx->WW, npx, 1); break;
    case VBS_STAGE_MAXIMA: unpack_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->max_bits, (uint8_t *)dst, ctx->W, ctx->WW, npx, 1); break;
    case VBS_STAGE_OPENED: unpack_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->open_bits, (uint8_t *)dst, ctx->W, ctx->WW, npx, 255); break;
    case VBS_STAGE_LABELS:
        unpack_labels_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->max_bits, ctx->parent, ctx->slot2label, (int32_t *)dst, ctx->H, ctx->W, ctx->WW, ctx->M, npx);
        break;
    default: return cudaErrorInvalidValue;
    }
    ctx->launches += 1;
    return cudaGetLastError();
}
