// k_contour.cu - K4c/K5a: one thread per opened blob: externality test, border following,
// ellipse fit (four re-traces, nothing stored), then centre <-> ellipse matching and compaction
// of the marker list in the reference's output order.  Replaces MD:196-243.
#include "vbs_ctx.h"
#include "vbs_bin.cuh"

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

constexpr int PCAP = 128;          // stored vertices per contour; longer contours are re-traced instead

// ---- 1. follow every external border once; keep its CHAIN_APPROX_SIMPLE vertices -----------------
struct StoreVisitor {
    uint32_t *dst; int n;
    __device__ __forceinline__ void operator()(int x, int y) {
        if (n < PCAP) dst[n] = (uint32_t)x | ((uint32_t)y << 16);
        ++n;
    }
};

// cpn[slot] = number of kept vertices (0: not an external contour, <0: trace guard tripped)
__global__ void contour_trace_kernel(const uint32_t *__restrict__ open_bits, const int32_t *__restrict__ parent2,
                                     const int32_t *__restrict__ croot, const int32_t *__restrict__ ncont,
                                     const int32_t *__restrict__ holes, uint32_t *__restrict__ cpts, int32_t *__restrict__ cpn,
                                     int H, int W, int WW, int M, uint32_t *status) {
    const int f = blockIdx.y;
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= M) return;
    const size_t i = (size_t)f * M + slot;
    cpn[i] = 0;
    if (slot >= min(ncont[f], M)) return;
    const int idx = croot[i];
    const int y0 = idx / W, x0 = idx - y0 * W;
    const uint32_t *img = open_bits + (size_t)f * H * WW;
    // external <=> the background left of the start pixel is 4-connected to the outside.  Frames
    // whose opened image has no hole at all (Euler number == component count) skip the test.
    if (x0 > 0 && holes[f] != 0) {
        const int xb = x0 - 1, wx = xb >> 5, b = xb & 31;
        const uint32_t bg = ~__ldg(img + (size_t)y0 * WW + wx) & valid_mask(wx, W);
        const uint32_t t = ~bg & ((2u << b) - 1u);
        const int s = t ? 32 - __clz(t) : 0;
        const int32_t *par = parent2 + (size_t)f * H * W;
        const int bidx = y0 * W + 32 * wx + s;
        int p = par[bidx];
        if (p >= 0 && p != bidx) p = par[p];
        if (p >= 0) return;                       // enclosed by another blob: RETR_EXTERNAL drops it
    }
    BitImage fg{img, H, W, WW};
    StoreVisitor sv{cpts + i * PCAP, 0};
    const int n = trace_external_simple(fg, x0, y0, 8LL * H * W + 16, sv);
    if (n < 0) atomicOr(status, VBS_DEV_TRACE_GUARD);
    cpn[i] = n;
}

// ---- 2. ellipse fit per contour (MD:203-220) -------------------------------------------------------
// cell = {cx, cy, major, minor, angle, valid}
__global__ void contour_fit_kernel(const uint32_t *__restrict__ open_bits, const int32_t *__restrict__ croot,
                                   const uint32_t *__restrict__ cpts, const int32_t *__restrict__ cpn, double *__restrict__ cell,
                                   int H, int W, int WW, int M) {
    const int f = blockIdx.y;
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= M) return;
    const size_t i = (size_t)f * M + slot;
    double *out = cell + i * 6;
    out[5] = 0.0;
    const int n = cpn[i];
    if (n < 5) return;                            // MD:204 (also: not external / guard)
    int npts = 0;
    EllipseResult e;
    if (n <= PCAP) {
        StoredSource src{cpts + i * PCAP, n};
        e = fit_ellipse_from(src, npts);
    } else {                                      // long contour: replay by following the border again
        const int idx = croot[i];
        const int y0 = idx / W, x0 = idx - y0 * W;
        BitImage fg{open_bits + (size_t)f * H * WW, H, W, WW};
        e = fit_ellipse_traced(fg, x0, y0, 8LL * H * W + 16, npts);
    }
    if (!e.ok) return;
    double major, minor, ang;
    if (e.w > e.h) { major = e.w; minor = e.h; ang = (double)e.angle; }
    else { major = e.h; minor = e.w; ang = (double)e.angle + 90.0; }     // MD:212-217 (float64 sum)
    if (minor < 5.0) return;                      // MD:219
    out[0] = e.cx; out[1] = e.cy; out[2] = major; out[3] = minor; out[4] = ang; out[5] = 1.0;
}

// ---- 3. nearest centroid inside the contour polygon and inside the (minor/10)^2 gate (MD:222-237) --
// One thread per contour slot.  The gate d^2 < (minor/10)^2 only admits centroids within minor/10 of the ellipse
// centre, so the frame's centroids are binned into cells (vbs_bin.cuh) and a slot whose gate radius fits a cell
// looks at the 3 x 3 cells around its centre; larger blobs scan all centroids.  Pass 1 keeps the centroids inside
// the gate (rarely more than one), pass 2 runs the polygon tests on them in index order with the reference's
// running minimum - testing inside the scan would serialise the 32 slots of a warp.
constexpr int MATCH_T = 128;
__global__ void __launch_bounds__(MATCH_T) match_kernel(const uint32_t *__restrict__ open_bits, const int32_t *__restrict__ croot,
                             const uint32_t *__restrict__ cpts, const int32_t *__restrict__ cpn,
                             const double *__restrict__ cell, const double *__restrict__ centres,
                             const int32_t *__restrict__ nlabels, const int32_t *__restrict__ bin_start, const int32_t *__restrict__ bin_items,
                             BinGrid g, int32_t *__restrict__ cmatch, int32_t *__restrict__ claim,
                             int H, int W, int WW, int M, uint32_t *status) {
    const int f = blockIdx.y;
    const int slot = blockIdx.x * MATCH_T + threadIdx.x;
    if (slot >= M) return;
    const size_t i = (size_t)f * M + slot;
    const double *c = cell + i * 6;
    cmatch[i] = -1;
    if (c[5] == 0.0) return;
    const double ecx = c[0], ecy = c[1];
    const double tenth = c[3] / 10.0;
    const double gate = mul_rn(tenth, tenth);
    const int np = cpn[i];
    const int n = min(nlabels[f], M);
    const double2 *cen = reinterpret_cast<const double2 *>(centres + (size_t)f * M * 2);     // (row, col)
    constexpr int NC = 4;
    int ck[NC]; double cd[NC];
    int ncand = 0;
    auto consider = [&](int j) {
        const double y = cen[j].x, x = cen[j].y;
        const double dx = x - ecx, dy = y - ecy;
        const double d = add_rn(mul_rn(dx, dx), mul_rn(dy, dy));
        if (d < gate) {
#pragma unroll
            for (int q = 0; q < NC; ++q)
                if (q == ncand) { ck[q] = j; cd[q] = d; }
            ++ncand;
        }
    };
    if (tenth <= g.cell) {
        const int32_t *start = bin_start + (size_t)f * (BIN_MAX_CELLS + 1);
        const int32_t *items = bin_items + (size_t)f * M;
        const int cx = bin_coord(ecx, g.inv_cell, g.gx), cy = bin_coord(ecy, g.inv_cell, g.gy);
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.gx - 1), y0 = max(cy - 1, 0), y1 = min(cy + 1, g.gy - 1);
        for (int y = y0; y <= y1; ++y) {
            const int i0 = start[y * g.gx + x0], i1 = start[y * g.gx + x1 + 1];      // the cells of one grid row are contiguous
            for (int k = i0; k < i1; ++k) consider(items[k]);
        }
        // bins hold the indices in arbitrary order: the reference visits the centroids by ascending index
#pragma unroll
        for (int a = 1; a < NC; ++a)
#pragma unroll
            for (int b = NC - 1; b >= a; --b)
                if (b < ncand && ck[b] < ck[b - 1]) {
                    const int tk = ck[b]; ck[b] = ck[b - 1]; ck[b - 1] = tk;
                    const double td = cd[b]; cd[b] = cd[b - 1]; cd[b - 1] = td;
                }
    } else {
        for (int j = 0; j < n; ++j) consider(j);
    }
    auto inside = [&](int j) {
        PointPolygon pp; pp.init(cen[j].y, cen[j].x);
        if (np <= PCAP) {
            StoredSource src{cpts + i * PCAP, np};
            src(pp);
        } else {
            const int idx = croot[i];
            const int y0 = idx / W, x0 = idx - y0 * W;
            BitImage fg{open_bits + (size_t)f * H * WW, H, W, WW};
            trace_external_simple(fg, x0, y0, 8LL * H * W + 16, pp);
        }
        return pp.result() >= 0;
    };
    int best = -1;
    double best_d = INFINITY;
    if (ncand <= NC) {
#pragma unroll
        for (int q = 0; q < NC; ++q)
            if (q < ncand && cd[q] < best_d && inside(ck[q])) { best = ck[q]; best_d = cd[q]; }
    } else {                                        // more candidates than registers: the literal loop
        for (int j = 0; j < n; ++j) {
            const double y = cen[j].x, x = cen[j].y;
            const double dx = x - ecx, dy = y - ecy;
            const double d = add_rn(mul_rn(dx, dx), mul_rn(dy, dy));
            if (d < gate && d < best_d && inside(j)) { best = j; best_d = d; }
        }
    }
    cmatch[i] = best;
    if (best >= 0 && atomicAdd(claim + (size_t)f * M + best, 1) > 0) atomicOr(status, VBS_DEV_MATCH_CONFLICT);
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

// which: bit 1 = follow + fit the blobs of the opened mask (open branch), bit 2 = match centroids to
// ellipses and compact the marker list (needs both branches)
cudaError_t vbs_launch_contours(vbs_ctx *ctx, int batch, int which) {
    VbsRange range("vbs:contours");
    const dim3 grid((ctx->M + 63) / 64, batch);
    if (which & 2) {
        contour_trace_kernel<<<grid, 64, 0, ctx->stream>>>(ctx->open_bits, ctx->parent2, ctx->croot, ctx->d_ncont, ctx->holes, ctx->cpts,
                                                           ctx->cpn, ctx->H, ctx->W, ctx->WW, ctx->M, ctx->d_status);
        contour_fit_kernel<<<grid, 64, 0, ctx->stream>>>(ctx->open_bits, ctx->croot, ctx->cpts, ctx->cpn, ctx->cell, ctx->H, ctx->W, ctx->WW,
                                                         ctx->M);
        ctx->launches += 2;
    }
    if (which & 4) {
        const BinGrid g = make_bin_grid(ctx->W, ctx->H, 32.0);
        bin_kernel<true><<<batch, 256, 0, ctx->stream>>>(ctx->centres, ctx->d_nlabels, g, ctx->cbin_start, ctx->cbin_items, ctx->M);
        match_kernel<<<dim3((ctx->M + MATCH_T - 1) / MATCH_T, batch), MATCH_T, 0, ctx->stream>>>(ctx->open_bits, ctx->croot, ctx->cpts, ctx->cpn, ctx->cell, ctx->centres,
                                                   ctx->d_nlabels, ctx->cbin_start, ctx->cbin_items, g, ctx->cmatch, ctx->claim, ctx->H, ctx->W, ctx->WW, ctx->M,
                                                   ctx->d_status);
        ctx->launches += 1;
        compact_kernel<<<(batch + 3) / 4, 128, 0, ctx->stream>>>(ctx->cell, ctx->cmatch, ctx->centres, ctx->d_ncont, ctx->d_nmarkers,
                                                                 ctx->marker_xy, ctx->marker_axes, ctx->M, batch);
        ctx->launches += 2;
    }
    return cudaGetLastError();
}
