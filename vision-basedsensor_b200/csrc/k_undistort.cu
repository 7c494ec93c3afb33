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

// One thread = RPX consecutive output pixels of one row, for RFR frames in turn: the map entries are read once
// per RFR frames and stay in registers; the RPX * C result bytes leave as 32-bit words when the row allows it.
// grid (ceil(W / (RPX * 128)), H, ceil(batch / RFR)), block 128
constexpr int RPX = 4, RFR = 8;
template <int C>
__global__ void __launch_bounds__(128) remap_kernel(const uint8_t *__restrict__ frames, int64_t frame_stride, int64_t row_pitch,
                                                     const int2 *__restrict__ map, uint8_t *__restrict__ out, int H, int W, int batch) {
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * RPX, y = blockIdx.y;
    if (x >= W) return;
    const int npx = min(RPX, W - x);
    vbs::RemapTap t[RPX];
    uint32_t ok[RPX];                                    // bit 0..3: sample (0,0) (0,1) (1,0) (1,1) lies inside the image
    int64_t off[RPX];
#pragma unroll
    for (int k = 0; k < RPX; ++k) {
        const int2 m = __ldg(map + (size_t)y * W + min(x + k, W - 1));
        t[k] = vbs::remap_tap(m.x, m.y);
        const bool x0 = t[k].sx >= 0 && t[k].sx < W, x1 = t[k].sx + 1 >= 0 && t[k].sx + 1 < W;
        const bool y0 = t[k].sy >= 0 && t[k].sy < H, y1 = t[k].sy + 1 >= 0 && t[k].sy + 1 < H;
        ok[k] = (x0 && y0 ? 1u : 0u) | (x1 && y0 ? 2u : 0u) | (x0 && y1 ? 4u : 0u) | (x1 && y1 ? 8u : 0u);
        off[k] = (int64_t)t[k].sy * row_pitch + (int64_t)t[k].sx * C;
    }
    const size_t row_bytes = (size_t)W * C;
    const bool words = npx == RPX && (row_bytes & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0;   // x * C is a multiple of 4 already
    const int f0 = blockIdx.z * RFR, f1 = min(batch, f0 + RFR);
    for (int f = f0; f < f1; ++f) {
        const uint8_t *src = frames + (size_t)f * frame_stride;
        uint8_t res[RPX * C];
#pragma unroll
        for (int k = 0; k < RPX; ++k) {
            const uint8_t *r0 = src + off[k], *r1 = r0 + row_pitch;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int p00 = (ok[k] & 1u) ? r0[c] : 0, p01 = (ok[k] & 2u) ? r0[C + c] : 0;
                const int p10 = (ok[k] & 4u) ? r1[c] : 0, p11 = (ok[k] & 8u) ? r1[C + c] : 0;
                res[k * C + c] = (uint8_t)((p00 * t[k].w00 + p01 * t[k].w01 + p10 * t[k].w10 + p11 * t[k].w11 + (1 << 14)) >> 15);
            }
        }
        uint8_t *dst = out + ((size_t)f * H + y) * row_bytes + (size_t)x * C;
        if (words) {
#pragma unroll
            for (int q = 0; q < RPX * C / 4; ++q)
                reinterpret_cast<uint32_t *>(dst)[q] = res[4 * q] | (res[4 * q + 1] << 8) | (res[4 * q + 2] << 16) | ((uint32_t)res[4 * q + 3] << 24);
        } else {
            for (int q = 0; q < npx * C; ++q) dst[q] = res[q];
        }
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
    VbsRange range("vbs:undistort");
    const dim3 grid((ctx->W + RPX * 128 - 1) / (RPX * 128), ctx->H, (batch + RFR - 1) / RFR);
    if (ctx->C == 3) remap_kernel<3><<<grid, 128, 0, ctx->stream>>>(frames, frame_stride, row_pitch, ctx->undist_map, out, ctx->H, ctx->W, batch);
    else remap_kernel<1><<<grid, 128, 0, ctx->stream>>>(frames, frame_stride, row_pitch, ctx->undist_map, out, ctx->H, ctx->W, batch);
    ctx->launches += 1;
    return cudaGetLastError();
}
