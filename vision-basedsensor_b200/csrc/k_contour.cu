// k_contour.cu - K4c/K5a: one thread per opened blob: externality test, border following,
// ellipse fit (four re-traces, nothing stored), then centre <-> ellipse matching and compaction
// of the marker list in the reference's output order.  Replaces MD:196-243.
#include "vbs_ctx.h"

namespace {

using namespace vbs;

__device__ __forceinline__ uint32_t valid_mask(int wx, int W) {
    const int rem = W - 32 * wx;
    return rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
}

struct BitImage {                 // foreground test on a bit-packed frame, false outside the image
    const uint32_t *img; int H, W, WW;
    __device__ __forceinline__ bool operator()(int x, int y) const {
        if ((unsigned)x >= (unsigned)W || (unsigned)y >= (unsigned)H) return false;
        return (__ldg(img + (size_t)y * WW + (x >> 5)) >> (x & 31)) & 1u;
    }
};

// cell = {cx, cy, major, minor, angle, valid}
__global__ void contour_fit_kernel(const uint32_t *__restrict__ open_bits, const int32_t *__restrict__ parent2,
                                   const int32_t *__restrict__ croot, const int32_t *__restrict__ ncont, double *__restrict__ cell,
                                   int H, int W, int WW, int M, size_t total, uint32_t *status) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const size_t f = i / M;
    const int slot = (int)(i % M);
    double *out = cell + i * 6;
    out[5] = 0.0;
    if (slot >= min(ncont[f], M)) return;
    const int idx = croot[i];
    const int y0 = idx / W, x0 = idx - y0 * W;
    const uint32_t *img = open_bits + f * (size_t)H * WW;
    // external <=> the background left of the start pixel is 4-connected to the outside
    if (x0 > 0) {
        const int xb = x0 - 1, wx = xb >> 5, b = xb & 31;
        const uint32_t bg = ~__ldg(img + (size_t)y0 * WW + wx) & valid_mask(wx, W);
        const uint32_t t = ~bg & ((2u << b) - 1u);
        const int s = t ? 32 - __clz(t) : 0;
        const int32_t *par = parent2 + f * (size_t)H * W;
        const int bidx = y0 * W + 32 * wx + s;
        int p = par[bidx];
        if (p >= 0 && p != bidx) p = par[p];
        if (p >= 0) return;                       // enclosed by another blob: RETR_EXTERNAL drops it
    }
    BitImage fg{img, H, W, WW};
    int npts = 0;
    const EllipseResult e = fit_ellipse_traced(fg, x0, y0, 8LL * H * W + 16, npts);
    if (npts < 0) { atomicOr(status, VBS_DEV_TRACE_GUARD); return; }
    if (npts < 5 || !e.ok) return;                // MD:204
    double major, minor, ang;
    if (e.w > e.h) { major = e.w; minor = e.h; ang = (double)e.angle; }
    else { major = e.h; minor = e.w; ang = (double)e.angle + 90.0; }     // MD:212-217 (float64 sum)
    if (minor < 5.0) return;                      // MD:219
    out[0] = e.cx; out[1] = e.cy; out[2] = major; out[3] = minor; out[4] = ang; out[5] = 1.0;
}

// nearest centroid inside the contour polygon and inside the (minor/10)^2 gate (MD:222-237)
__global__ void match_kernel(const uint32_t *__restrict__ open_bits, const int32_t *__restrict__ croot,
                             const double *__restrict__ cell, const double *__restrict__ centres,
                             const int32_t *__restrict__ nlabels, int32_t *__restrict__ cmatch, int32_t *__restrict__ claim,
                             int H, int W, int WW, int M, size_t total, uint32_t *status) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const size_t f = i / M;
    cmatch[i] = -1;
    const double *c = cell + i * 6;
    if (c[5] == 0.0) return;
    const double ecx = c[0], ecy = c[1];
    const double tenth = c[3] / 10.0;
    const double gate = mul_rn(tenth, tenth);
    const int n = min(nlabels[f], M);
    const int idx = croot[i];
    const int y0 = idx / W, x0 = idx - y0 * W;
    BitImage fg{open_bits + f * (size_t)H * WW, H, W, WW};
    int best = -1;
    double best_d = INFINITY;
    const double *cen = centres + f * (size_t)M * 2;
    for (int j = 0; j < n; ++j) {
        const double y = cen[2 * j], x = cen[2 * j + 1];
        const double dx = x - ecx, dy = y - ecy;
        const double d = add_rn(mul_rn(dx, dx), mul_rn(dy, dy));
        if (d < gate && d < best_d) {
            PointPolygon pp; pp.init(x, y);
            trace_external_simple(fg, x0, y0, 8LL * H * W + 16, pp);
            if (pp.result() >= 0) { best = j; best_d = d; }
        }
    }
    cmatch[i] = best;
    if (best >= 0 && atomicAdd(claim + f * M + best, 1) > 0) atomicOr(status, VBS_DEV_MATCH_CONFLICT);
}

// marker list in contour order (one warp per frame)
__global__ void compact_kernel(const double *__restrict__ cell, const int32_t *__restrict__ cmatch, const double *__restrict__ centres,
                               const int32_t *__restrict__ ncont, int32_t *__restrict__ nmarkers, double *__restrict__ marker_xy,
                               double *__restrict__ marker_axes, int M, int batch) {
    const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (f >= batch) return;
    const int lane = threadIdx.x & 31;
    const int n = min(ncont[f], M);
    int base = 0;
    for (int s0 = 0; s0 < n; s0 += 32) {
        const int slot = s0 + lane;
        const int j = slot < n ? cmatch[(size_t)f * M + slot] : -1;
        const uint32_t has = __ballot_sync(0xffffffffu, j >= 0);
        if (j >= 0) {
            const int k = base + __popc(has & ((1u << lane) - 1u));
            const double *c = cell + ((size_t)f * M + slot) * 6;
            const double *cen = centres + ((size_t)f * M + j) * 2;
            marker_xy[((size_t)f * M + k) * 2 + 0] = cen[1];      // x = col  (MD:199)
            marker_xy[((size_t)f * M + k) * 2 + 1] = cen[0];      // y = row
            marker_axes[((size_t)f * M + k) * 3 + 0] = c[2];
            marker_axes[((size_t)f * M + k) * 3 + 1] = c[3];
            marker_axes[((size_t)f * M + k) * 3 + 2] = c[4];
        }
        base += __popc(has);
    }
    if (lane == 0) nmarkers[f] = base;
}

}  // namespace

cudaError_t vbs_launch_contours(vbs_ctx *ctx, int batch) {
    const size_t total = (size_t)batch * ctx->M;
    const unsigned g = (unsigned)((total + 127) / 128);
    cudaError_t e = cudaMemsetAsync(ctx->claim, 0, sizeof(int32_t) * total, ctx->stream);
    if (e != cudaSuccess) return e;
    contour_fit_kernel<<<g, 128, 0, ctx->stream>>>(ctx->open_bits, ctx->parent2, ctx->croot, ctx->d_ncont, ctx->cell, ctx->H, ctx->W,
                                                   ctx->WW, ctx->M, total, ctx->d_status);
    match_kernel<<<g, 128, 0, ctx->stream>>>(ctx->open_bits, ctx->croot, ctx->cell, ctx->centres, ctx->d_nlabels, ctx->cmatch, ctx->claim,
                                             ctx->H, ctx->W, ctx->WW, ctx->M, total, ctx->d_status);
    compact_kernel<<<(batch + 3) / 4, 128, 0, ctx->stream>>>(ctx->cell, ctx->cmatch, ctx->centres, ctx->d_ncont, ctx->d_nmarkers,
                                                             ctx->marker_xy, ctx->marker_axes, ctx->M, batch);
    ctx->launches += 3;
    return cudaGetLastError();
}
