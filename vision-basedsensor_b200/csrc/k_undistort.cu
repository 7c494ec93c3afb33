// k_undistort.cu - K0 (optional): lens correction of the cropped frame before detection, MD:93-109
// (getOptimalNewCameraMatrix alpha = 0 -> initUndistortRectifyMap CV_16SC2 -> remap INTER_LINEAR).
// The reference recomputes the maps for every frame; they depend on (K, D, size) only, so they are
// built once per vbs_set_undistort: the new camera matrix on the host (81 scalar undistortions,
// vbs_geom.h), the maps by one kernel in float64 with OpenCV's operation order, kept as the source
// position in 1/32 px, {iu, iv} int32 per pixel (CV_16SC2 is the same information split into int16
// integer parts and a 10-bit fraction word).  The remap itself is HBM-shaped: per output pixel 8 bytes
// of map, 4 gathered source bytes per channel (neighbouring outputs share them through L1/L2) and one
// store per channel.
#include "vbs_ctx.h"

namespace {

__global__ void __launch_bounds__(256) rectify_map_kernel(vbs::LensF64 lens, double nk0, double nk1, double nk2, double nk3,
                                                           int2 *__restrict__ map, int H, int W) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= W) return;
    const double nk[4] = {nk0, nk1, nk2, nk3};
    int iu, iv;
    vbs::rectify_source_q5(lens, nk, i, j, iu, iv);
    map[(size_t)i * W + j] = make_int2(iu, iv);
}

// CV_16SC2 view of the map for tests / debugging: map1 = (iu >> 5, iv >> 5) int16, map2 = fractions
__global__ void __launch_bounds__(256) export_maps_kernel(const int2 *__restrict__ map, short2 *__restrict__ map1, uint16_t *__restrict__ map2, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int2 m = map[i];
    map1[i] = make_short2((short)(m.x >> 5), (short)(m.y >> 5));
    map2[i] = (uint16_t)((m.y & 31) * 32 + (m.x & 31));
}

// one thread per output pixel; grid (ceil(W / 256), H, batch)
template <int C>
__global__ void __launch_bounds__(256) remap_kernel(const uint8_t *__restrict__ frames, int64_t frame_stride, int64_t row_pitch,
                                                     const int2 *__restrict__ map, uint8_t *__restrict__ out, int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const int2 m = __ldg(map + (size_t)y * W + x);
    const vbs::RemapTap t = vbs::remap_tap(m.x, m.y);
    const uint8_t *src = frames + (size_t)blockIdx.z * frame_stride;
    uint8_t *dst = out + ((size_t)blockIdx.z * H + y) * W * C + (size_t)x * C;
    const bool x0 = t.sx >= 0 && t.sx < W, x1 = t.sx + 1 >= 0 && t.sx + 1 < W;
    const bool y0 = t.sy >= 0 && t.sy < H, y1 = t.sy + 1 >= 0 && t.sy + 1 < H;
    const uint8_t *r0 = src + (int64_t)t.sy * row_pitch + (int64_t)t.sx * C;
    const uint8_t *r1 = r0 + row_pitch;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int p00 = (x0 && y0) ? r0[c] : 0, p01 = (x1 && y0) ? r0[C + c] : 0;
        const int p10 = (x0 && y1) ? r1[c] : 0, p11 = (x1 && y1) ? r1[C + c] : 0;
        dst[c] = (uint8_t)((p00 * t.w00 + p01 * t.w01 + p10 * t.w10 + p11 * t.w11 + (1 << 14)) >> 15);
    }
}

}  // namespace

// K: 3x3 row-major, D: nd in {4, 5, 8} coefficients (k1 k2 p1 p2 [k3 [k4 k5 k6]])
cudaError_t vbs_undistort_setup(vbs_ctx *ctx, const double *K, const double *D, int nd) {
    vbs::LensF64 c;
    c.fx = K[0]; c.fy = K[4]; c.cx = K[2]; c.cy = K[5];
    double d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < nd; ++i) d[i] = D[i];
    c.k1 = d[0]; c.k2 = d[1]; c.p1 = d[2]; c.p2 = d[3]; c.k3 = d[4]; c.k4 = d[5]; c.k5 = d[6]; c.k6 = d[7];
    vbs::optimal_new_camera_alpha0(c, ctx->W, ctx->H, ctx->new_k);
    rectify_map_kernel<<<dim3((ctx->W + 255) / 256, ctx->H), 256, 0, ctx->stream>>>(c, ctx->new_k[0], ctx->new_k[1], ctx->new_k[2], ctx->new_k[3],
                                                                                    ctx->undist_map, ctx->H, ctx->W);
    ctx->launches += 1;
    return cudaGetLastError();
}

cudaError_t vbs_launch_export_maps(vbs_ctx *ctx, int16_t *map1, uint16_t *map2) {
    const int n = ctx->H * ctx->W;
    export_maps_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->undist_map, reinterpret_cast<short2 *>(map1), map2, n);
    ctx->launches += 1;
    return cudaGetLastError();
}

cudaError_t vbs_launch_remap(vbs_ctx *ctx, const uint8_t *frames, int batch, int64_t frame_stride, int64_t row_pitch, uint8_t *out) {
    const dim3 grid((ctx->W + 255) / 256, ctx->H, batch);
    if (ctx->C == 3) remap_kernel<3><<<grid, 256, 0, ctx->stream>>>(frames, frame_stride, row_pitch, ctx->undist_map, out, ctx->H, ctx->W);
    else remap_kernel<1><<<grid, 256, 0, ctx->stream>>>(frames, frame_stride, row_pitch, ctx->undist_map, out, ctx->H, ctx->W);
    ctx->launches += 1;
    return cudaGetLastError();
}
