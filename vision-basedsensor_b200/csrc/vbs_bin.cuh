// vbs_bin.cuh - points of one frame counting-sorted into square cells (shared by the nearest-marker match of
// k_track3d.cu and the centre <-> ellipse match of k_contour.cu).  A query that can only accept points within a
// distance d <= cell side looks at the 3 x 3 cells around its own; cell indices are clamped to the grid, which
// never moves two points further apart than they are, so points outside the image are handled too.
#pragma once
#include <math.h>
#include <stdint.h>

constexpr int BIN_MAX_CELLS = 8192;
struct BinGrid { double cell, inv_cell; int gx, gy; };

// cells of side >= min_cell, few enough for the shared-memory histogram of bin_kernel
inline BinGrid make_bin_grid(int W, int H, double min_cell) {
    double cell = min_cell > 32.0 ? min_cell : 32.0;
    if (!(cell < 1e300)) cell = 1e300;
    BinGrid g;
    for (;;) {
        const double gx = ceil(W / cell), gy = ceil(H / cell);
        g.gx = gx < 1.0 ? 1 : (int)gx; g.gy = gy < 1.0 ? 1 : (int)gy;
        if ((long long)g.gx * g.gy <= BIN_MAX_CELLS) break;
        cell *= 2.0;
    }
    g.cell = cell; g.inv_cell = 1.0 / cell;
    return g;
}

__device__ __forceinline__ int bin_coord(double v, double inv_cell, int g) {
    const double c = floor(v * inv_cell);
    return c < 0.0 ? 0 : (c >= (double)g ? g - 1 : (int)c);       // NaN compares false twice -> (int)NaN is 0 on the GPU
}

// one CTA (256 threads) per frame.  pts: [frames][M] pairs, (x, y) or - YX - (y, x); cell_start: [frames][BIN_MAX_CELLS + 1]
// first item of every cell (row-major cells, so the cells of one grid row are contiguous); cell_items: [frames][M]
// point indices sorted by cell (order inside a cell is arbitrary)
template <bool YX>
__global__ void __launch_bounds__(256) bin_kernel(const double *__restrict__ pts, const int32_t *__restrict__ npts, BinGrid g,
                                                   int32_t *__restrict__ cell_start, int32_t *__restrict__ cell_items, int M) {
    __shared__ int32_t cnt[BIN_MAX_CELLS];
    __shared__ int32_t part[256];
    const int f = blockIdx.x, tid = threadIdx.x;
    const int nc = g.gx * g.gy;
    const int n = min(npts[f], M);
    const double2 *p2 = reinterpret_cast<const double2 *>(pts + (size_t)f * M * 2);
    auto cell_of = [&](int k) {
        const double2 p = p2[k];
        return bin_coord(YX ? p.x : p.y, g.inv_cell, g.gy) * g.gx + bin_coord(YX ? p.y : p.x, g.inv_cell, g.gx);
    };
    for (int c = tid; c < nc; c += 256) cnt[c] = 0;
    __syncthreads();
    for (int k = tid; k < n; k += 256) atomicAdd(&cnt[cell_of(k)], 1);
    __syncthreads();
    // exclusive scan: every thread owns a contiguous slice of cells
    const int per = (nc + 255) / 256, c0 = tid * per, c1 = min(nc, c0 + per);
    int sum = 0;
    for (int c = c0; c < c1; ++c) sum += cnt[c];
    part[tid] = sum;
    __syncthreads();
    if (tid == 0) { int run = 0; for (int i = 0; i < 256; ++i) { const int v = part[i]; part[i] = run; run += v; } }
    __syncthreads();
    int32_t *start = cell_start + (size_t)f * (BIN_MAX_CELLS + 1);
    int run = part[tid];
    for (int c = c0; c < c1; ++c) { const int v = cnt[c]; start[c] = run; cnt[c] = run; run += v; }     // cnt becomes the fill cursor
    if (tid == 255) start[nc] = n;
    __syncthreads();
    int32_t *items = cell_items + (size_t)f * M;
    for (int k = tid; k < n; k += 256) items[atomicAdd(&cnt[cell_of(k)], 1)] = k;
}
