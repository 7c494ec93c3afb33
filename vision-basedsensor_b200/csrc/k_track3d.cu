// k_track3d.cu - K5b/K6/K7: batched per-marker kernels.
//   track_kernel  : nearest detection per reference entry, first minimum, <= min distance   MD:349-396
//   pos3d_kernel  : row filter + warm-up, undistortPoints (5 iterations), depth from the
//                   apparent diameter, world transform                                       R3:172-176,185-238,255-260
//   disp_kernel   : last-seen displacement along the frame axis                              R3:262-314
//   plane_kernel  : deviation end points + least-squares plane + tilt                        FD:196-204,219-232,141-159
#include "vbs_ctx.h"
#include "vbs_bin.cuh"

namespace {

using namespace vbs;

// ---- MarkerTracker._track_markers (MD:349-396): nearest marker per reference entry --------------------
// cdist + argmin (MD:369-371) compares float64 square roots, keeps the FIRST minimum and then rejects it when it is
// farther than min_marker_distance (MD:372).  Only markers within that distance can therefore be returned, so the
// markers of a frame are binned into square cells of side >= min_marker_distance and a reference entry looks at
// the 3 x 3 cells around its own (cell indices are clamped, which never separates two points by more than their
// distance).  Inside the scan the comparison runs on squared distances: beyond a relative margin of 2^-40 the
// correctly rounded roots are strictly ordered like the squares; when the runner-up is within that margin
// (practically never) the candidates are compared again by their rounded roots, smallest index first.
__global__ void __launch_bounds__(128) track_kernel(const double *__restrict__ ref_xy, const double *__restrict__ marker_xy,
                             const double *__restrict__ marker_axes, const int32_t *__restrict__ cell_start, const int32_t *__restrict__ cell_items,
                             BinGrid g, int32_t *__restrict__ row_det, double *__restrict__ row_cxy, double *__restrict__ row_axes,
                             int R, int M, double min_dist) {
    const size_t f = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const double ox = ref_xy[2 * r], oy = ref_xy[2 * r + 1];
    const double2 *mk = reinterpret_cast<const double2 *>(marker_xy + f * (size_t)M * 2);
    const int32_t *start = cell_start + f * (BIN_MAX_CELLS + 1);
    const int32_t *items = cell_items + f * M;
    const int cx = bin_coord(ox, g.inv_cell, g.gx), cy = bin_coord(oy, g.inv_cell, g.gy);
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.gx - 1), y0 = max(cy - 1, 0), y1 = min(cy + 1, g.gy - 1);
    int best = -1;
    double best_sq = INFINITY, second_sq = INFINITY;
    for (int y = y0; y <= y1; ++y) {
        const int i0 = start[y * g.gx + x0], i1 = start[y * g.gx + x1 + 1];      // the cells of one row are contiguous
        for (int i = i0; i < i1; ++i) {
            const int k = items[i];
            const double dx = ox - mk[k].x, dy = oy - mk[k].y;
            const double sq = add_rn(mul_rn(dx, dx), mul_rn(dy, dy));
            if (sq < best_sq || (sq == best_sq && k < best)) { second_sq = best_sq; best_sq = sq; best = k; }
            else second_sq = fmin(second_sq, sq);
        }
    }
    double best_d = __dsqrt_rn(best_sq);
    if (best >= 0 && second_sq <= mul_rn(best_sq, 1.0 + 0x1p-40)) {       // near-tie: compare the rounded roots, first index wins
        best = -1; best_d = INFINITY;
        for (int y = y0; y <= y1; ++y) {
            const int i0 = start[y * g.gx + x0], i1 = start[y * g.gx + x1 + 1];
            for (int i = i0; i < i1; ++i) {
                const int k = items[i];
                const double dx = ox - mk[k].x, dy = oy - mk[k].y;
                const double d = __dsqrt_rn(add_rn(mul_rn(dx, dx), mul_rn(dy, dy)));
                if (best < 0 || d < best_d || (d == best_d && k < best)) { best_d = d; best = k; }
            }
        }
    }
    if (best >= 0 && best_d > min_dist) best = -1;              // MD:372
    const size_t i = f * R + r;
    row_det[i] = best;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (best >= 0) {
        const double2 c = mk[best];
        row_cxy[2 * i] = c.x; row_cxy[2 * i + 1] = c.y;
        const double *ax = marker_axes + (f * (size_t)M + best) * 3;
        row_axes[3 * i] = ax[0]; row_axes[3 * i + 1] = ax[1]; row_axes[3 * i + 2] = ax[2];
    } else {
        row_cxy[2 * i] = nan; row_cxy[2 * i + 1] = nan;
        row_axes[3 * i] = nan; row_axes[3 * i + 1] = nan; row_axes[3 * i + 2] = nan;
    }
}

__global__ void pos3d_kernel(CameraF64 cam, const int32_t *__restrict__ row_det, const double *__restrict__ row_cxy,
                             const double *__restrict__ row_axes, double *__restrict__ obs, double *__restrict__ pos3d,
                             uint8_t *__restrict__ flags, int R, int64_t frameno0, int64_t first_kept, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const size_t f = i / R;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    uint8_t fl = 0;
    double P[3] = {nan, nan, nan};
    double u = nan, v = nan, d = nan;
    const int64_t frameno = frameno0 + (int64_t)f;
    if (row_det[i] >= 0 && row_axes[3 * i] >= cam.min_size && frameno >= first_kept) {   // R3:173, R3:255-256
        d = row_axes[3 * i];                                                             // diameter = major_axis (R3:273)
        undistort5(cam, row_cxy[2 * i], row_cxy[2 * i + 1], u, v);
        fl = 1;
        if (position3d(cam, u, v, d, P)) fl |= 2;
        else { P[0] = P[1] = P[2] = nan; }
    }
    obs[3 * i] = u; obs[3 * i + 1] = v; obs[3 * i + 2] = d;
    double *o = pos3d + 7 * i;
    o[0] = P[0]; o[1] = P[1]; o[2] = P[2]; o[3] = nan; o[4] = nan; o[5] = nan; o[6] = nan;
    flags[i] = fl;
}

// one warp per reference entry: lanes take 32 consecutive frames, the "previous observation" of a
// frame is found with a ballot (nearest earlier lane) or carried over from the previous 32 frames
__global__ void disp_kernel(CameraF64 cam, const double *__restrict__ obs, double *__restrict__ pos3d, uint8_t *__restrict__ flags,
                            double *__restrict__ last_seen, int R, int batch, int64_t frameno0) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= R) return;
    const int lane = threadIdx.x & 31;
    // carry = most recent observation before the current 32-frame window: -2 none, -1 the last-seen table, >= 0 a frame of this batch
    int carry = last_seen[4 * r + 3] >= 0.0 ? -1 : -2;
    double Pl[3] = {0, 0, 0};
    bool okl = false;
    if (carry == -1) okl = position3d(cam, last_seen[4 * r], last_seen[4 * r + 1], last_seen[4 * r + 2], Pl);
    for (int f0 = 0; f0 < batch; f0 += 32) {
        const int f = f0 + lane;
        const size_t i = (size_t)(f < batch ? f : batch - 1) * R + r;
        const uint8_t fl = f < batch ? flags[i] : 0;
        const uint32_t has = __ballot_sync(0xffffffffu, fl & 1);
        if (fl & 1) {
            const uint32_t below = has & ((1u << lane) - 1u);
            const int prev = below ? f0 + (31 - __clz(below)) : carry;
            bool okp = false;
            double Pp[3] = {0, 0, 0};
            if (prev >= 0) {
                const size_t j = (size_t)prev * R + r;
                okp = flags[j] & 2;
                Pp[0] = pos3d[7 * j]; Pp[1] = pos3d[7 * j + 1]; Pp[2] = pos3d[7 * j + 2];
            } else if (prev == -1) {
                okp = okl; Pp[0] = Pl[0]; Pp[1] = Pl[1]; Pp[2] = Pl[2];
            }
            if (prev != -2 && okp && (fl & 2)) {
                double *o = pos3d + 7 * i;
                const double dx = o[0] - Pp[0], dy = o[1] - Pp[1], dz = o[2] - Pp[2];
                const double nrm = sqrt(dx * dx + dy * dy + dz * dz);
                if (!(nrm > cam.max_disp)) {                                              // R3:293
                    o[3] = dx; o[4] = dy; o[5] = dz; o[6] = nrm;
                    flags[i] = fl | 4;                    // bit 2 is only read by the host
                }
            }
        }
        if (has) carry = f0 + (31 - __clz(has));
        __syncwarp();
    }
    if (lane == 0 && carry >= 0) {
        const size_t j = (size_t)carry * R + r;
        last_seen[4 * r] = obs[3 * j]; last_seen[4 * r + 1] = obs[3 * j + 1]; last_seen[4 * r + 2] = obs[3 * j + 2];
        last_seen[4 * r + 3] = (double)(frameno0 + carry);
    }
}

// frame-sharded runs: a shard is first processed with an empty last-seen table, so the FIRST
// observation of every reference entry in the shard has no displacement row yet.  Given the
// last-seen table that arrives from the preceding shards, emit exactly that missing row.
__global__ void fix_displacement_kernel(CameraF64 cam, double *__restrict__ pos3d, uint8_t *__restrict__ flags,
                                        const double *__restrict__ incoming, int R, long long nframes) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    if (!(incoming[4 * r + 3] >= 0.0)) return;
    double Pp[3];
    const bool okp = position3d(cam, incoming[4 * r], incoming[4 * r + 1], incoming[4 * r + 2], Pp);
    for (long long f = 0; f < nframes; ++f) {
        const size_t i = (size_t)f * R + r;
        const uint8_t fl = flags[i];
        if (!(fl & 1)) continue;
        if (okp && (fl & 2)) {
            double *o = pos3d + 7 * i;
            const double dx = o[0] - Pp[0], dy = o[1] - Pp[1], dz = o[2] - Pp[2];
            const double nrm = sqrt(dx * dx + dy * dy + dz * dz);
            if (!(nrm > cam.max_disp)) { o[3] = dx; o[4] = dy; o[5] = dz; o[6] = nrm; flags[i] = fl | 4; }
        }
        return;                                    // only the first observation was missing its predecessor
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per frame
__global__ void plane_kernel(const double *__restrict__ pos3d, const uint8_t *__restrict__ flags, const double *__restrict__ ref,
                             const double *__restrict__ start, const double *__restrict__ dvert, const uint8_t *__restrict__ use,
                             int shell, double scale, double *__restrict__ plane, int32_t *__restrict__ plane_n, int R, int batch) {
    const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (f >= batch) return;
    const int lane = threadIdx.x & 31;
    auto point = [&](int r, double &X, double &Y, double &Z) -> bool {
        const size_t i = (size_t)f * R + r;
        if (!use[r] || !(flags[i] & 2)) return false;
        const double *p = pos3d + 7 * i;
        const double dx = (p[0] - start[3 * r]) - dvert[3 * r];           // d_tilt - d_vert (FD:196-203)
        const double dy = (p[1] - start[3 * r + 1]) - dvert[3 * r + 1];
        const double dz = (p[2] - start[3 * r + 2]) - dvert[3 * r + 2];
        X = ref[3 * r] + dx * scale;                                      // FD:225-232
        Y = ref[3 * r + 1] + dy * scale;
        Z = (shell ? ref[3 * r + 2] : 0.0) + dz * scale;
        return true;
    };
    double n = 0, sx = 0, sy = 0, sz = 0;
    for (int r = lane; r < R; r += 32) {
        double X, Y, Z;
        if (point(r, X, Y, Z)) { n += 1; sx += X; sy += Y; sz += Z; }
    }
    n = warp_sum(n); sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
    const double mx = sx / n, my = sy / n, mz = sz / n;
    double sxx = 0, sxy = 0, syy = 0, sxz = 0, syz = 0;
    for (int r = lane; r < R; r += 32) {
        double X, Y, Z;
        if (point(r, X, Y, Z)) {
            const double x = X - mx, y = Y - my, z = Z - mz;
            sxx += x * x; sxy += x * y; syy += y * y; sxz += x * z; syz += y * z;
        }
    }
    sxx = warp_sum(sxx); sxy = warp_sum(sxy); syy = warp_sum(syy); sxz = warp_sum(sxz); syz = warp_sum(syz);
    if (lane == 0) {
        double out[4];
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        if (!plane_solve(n, mx, my, mz, sxx, sxy, syy, sxz, syz, out)) out[0] = out[1] = out[2] = out[3] = nan;
        plane[4 * f] = out[0]; plane[4 * f + 1] = out[1]; plane[4 * f + 2] = out[2]; plane[4 * f + 3] = out[3];
        plane_n[f] = (int32_t)n;
    }
}

__global__ void undistort_kernel(CameraF64 cam, const double *__restrict__ uv, double *__restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    undistort5(cam, uv[2 * i], uv[2 * i + 1], out[2 * i], out[2 * i + 1]);
}
__global__ void position_kernel(CameraF64 cam, const double *__restrict__ uvd, double *__restrict__ P, uint8_t *__restrict__ ok, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double p[3];
    const bool good = position3d(cam, uvd[3 * i], uvd[3 * i + 1], uvd[3 * i + 2], p);
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    P[3 * i] = good ? p[0] : nan; P[3 * i + 1] = good ? p[1] : nan; P[3 * i + 2] = good ? p[2] : nan;
    ok[i] = good ? 1 : 0;
}
// plain plane fit of n points (one warp)
__global__ void plane_points_kernel(const double *__restrict__ X, const double *__restrict__ Y, const double *__restrict__ Z, int n,
                                    double *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    double sx = 0, sy = 0, sz = 0;
    for (int i = lane; i < n; i += 32) { sx += X[i]; sy += Y[i]; sz += Z[i]; }
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
    const double mx = sx / n, my = sy / n, mz = sz / n;
    double sxx = 0, sxy = 0, syy = 0, sxz = 0, syz = 0;
    for (int i = lane; i < n; i += 32) {
        const double x = X[i] - mx, y = Y[i] - my, z = Z[i] - mz;
        sxx += x * x; sxy += x * y; syy += y * y; sxz += x * z; syz += y * z;
    }
    sxx = warp_sum(sxx); sxy = warp_sum(sxy); syy = warp_sum(syy); sxz = warp_sum(sxz); syz = warp_sum(syz);
    if (lane == 0) {
        double r[4];
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        if (!plane_solve((double)n, mx, my, mz, sxx, sxy, syy, sxz, syz, r)) r[0] = r[1] = r[2] = r[3] = nan;
        out[0] = r[0]; out[1] = r[1]; out[2] = r[2]; out[3] = r[3];
    }
}

}  // namespace

cudaError_t vbs_launch_fix_displacement(vbs_ctx *ctx, double *pos3d, uint8_t *flags, const double *incoming_dev, long long nframes) {
    fix_displacement_kernel<<<(ctx->R + 63) / 64, 64, 0, ctx->stream>>>(ctx->cam, pos3d, flags, incoming_dev, ctx->R, nframes);
    ctx->launches += 1;
    return cudaGetLastError();
}
cudaError_t vbs_launch_undistort(vbs_ctx *ctx, const double *uv, double *out, int n) {
    undistort_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->cam, uv, out, n);
    ctx->launches += 1;
    return cudaGetLastError();
}
cudaError_t vbs_launch_position(vbs_ctx *ctx, const double *uvd, double *P, uint8_t *ok, int n) {
    position_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->cam, uvd, P, ok, n);
    ctx->launches += 1;
    return cudaGetLastError();
}
cudaError_t vbs_launch_plane_points(vbs_ctx *ctx, const double *X, const double *Y, const double *Z, int n, double *out) {
    plane_points_kernel<<<1, 32, 0, ctx->stream>>>(X, Y, Z, n, out);
    ctx->launches += 1;
    return cudaGetLastError();
}

cudaError_t vbs_launch_track(vbs_ctx *ctx, int batch, int64_t frameno0) {
    VbsRange range("vbs:track");
    const int R = ctx->R;
    if (R <= 0) return cudaSuccess;
    const BinGrid g = make_bin_grid(ctx->W, ctx->H, ctx->min_dist);      // cells of side >= min_marker_distance
    bin_kernel<false><<<batch, 256, 0, ctx->stream>>>(ctx->marker_xy, ctx->d_nmarkers, g, ctx->cell_start, ctx->cell_items, ctx->M);
    track_kernel<<<dim3((R + 127) / 128, batch), 128, 0, ctx->stream>>>(ctx->ref_xy, ctx->marker_xy, ctx->marker_axes, ctx->cell_start, ctx->cell_items, g,
                                                                          ctx->row_det, ctx->row_cxy, ctx->row_axes, R, ctx->M, ctx->min_dist);
    ctx->launches += 1;
    ctx->launches += 2;
    return vbs_launch_reconstruct(ctx, batch, frameno0);
}

// R3:240-316 + FD:138-162 on the tracking rows currently in ctx->row_* (batch x R)
cudaError_t vbs_launch_reconstruct(vbs_ctx *ctx, int batch, int64_t frameno0) {
    VbsRange range("vbs:reconstruct");
    const int R = ctx->R;
    if (R <= 0) return cudaSuccess;
    const size_t total = (size_t)batch * R;
    const unsigned g = (unsigned)((total + 127) / 128);
    if (ctx->have_cam) {
        if (!ctx->have_first) { ctx->first_frame = frameno0; ctx->have_first = 1; }
        const int64_t first_kept = ctx->first_frame + (ctx->warmup > 0 ? ctx->warmup : 0);
        pos3d_kernel<<<g, 128, 0, ctx->stream>>>(ctx->cam, ctx->row_det, ctx->row_cxy, ctx->row_axes, ctx->obs, ctx->pos3d, ctx->pos_flags,
                                                 R, frameno0, first_kept, total);
        disp_kernel<<<(R + 3) / 4, 128, 0, ctx->stream>>>(ctx->cam, ctx->obs, ctx->pos3d, ctx->pos_flags, ctx->last_seen, R, batch, frameno0);
        ctx->launches += 2;
        if (ctx->have_plane) {
            plane_kernel<<<(batch + 3) / 4, 128, 0, ctx->stream>>>(ctx->pos3d, ctx->pos_flags, ctx->pl_ref, ctx->pl_start, ctx->pl_dvert,
                                                                   ctx->pl_use, ctx->shell, ctx->pscale, ctx->plane, ctx->plane_n, R, batch);
            ctx->launches += 1;
        }
    }
    return cudaGetLastError();
}
