// hostcheck.cpp - TEST INFRASTRUCTURE ONLY.  Compiles the host/device-shared scalar geometry of
// csrc/vbs_geom.h for the CPU so the `not gpu` tests can pin it against cv2 / numpy without a
// GPU.  The product never loads this library.
#include <cstdint>
#include <cstring>
#include "../../vision-basedsensor_b200/csrc/vbs_geom.h"
#include "../../vision-basedsensor_b200/csrc/vbs_segplan.h"

using namespace vbs;

struct ByteImage {
    const uint8_t *p; int h, w;
    bool operator()(int x, int y) const { return x >= 0 && y >= 0 && x < w && y < h && p[(size_t)y * w + x] != 0; }
};
struct Collect {
    int32_t *out; int cap; int n;
    void operator()(int x, int y) { if (n < cap) { out[2 * n] = x; out[2 * n + 1] = y; } ++n; }
};

extern "C" {

int hc_trace(const uint8_t *img, int h, int w, int x0, int y0, int32_t *pts, int cap) {
    ByteImage b{img, h, w};
    Collect c{pts, cap, 0};
    return trace_external_simple(b, x0, y0, 8LL * h * w + 16, c);
}

int hc_fit_ellipse(const uint8_t *img, int h, int w, int x0, int y0, float out[5], int *n_pts) {
    ByteImage b{img, h, w};
    int n = 0;
    EllipseResult e = fit_ellipse_traced(b, x0, y0, 8LL * h * w + 16, n);
    out[0] = e.cx; out[1] = e.cy; out[2] = e.w; out[3] = e.h; out[4] = e.angle;
    *n_pts = n;
    return e.ok;
}

int hc_point_polygon(const int32_t *pts, int n, double x, double y) {
    PointPolygon pp; pp.init(x, y);
    for (int i = 0; i < n; ++i) pp(pts[2 * i], pts[2 * i + 1]);
    return pp.result();
}

static CameraF64 make_cam(const float *K, const float *D, const float *R, const float *T, double diam) {
    CameraF64 c;
    c.fx = K[0]; c.fy = K[4]; c.cx = K[2]; c.cy = K[5];
    c.k1 = D[0]; c.k2 = D[1]; c.p1 = D[2]; c.p2 = D[3]; c.k3 = D[4];
    for (int i = 0; i < 9; ++i) c.R[i] = R[i];
    for (int i = 0; i < 3; ++i) c.T[i] = T[i];
    volatile float favg = (K[0] + K[4]) / 2.0f;
    volatile float ratio = (float)diam / favg;
    volatile float favg_sq = favg * favg;
    c.f_avg = favg; c.ratio = ratio; c.f_avg_sq = favg_sq; c.min_size = 5; c.max_disp = 50;
    return c;
}

void hc_undistort(const float *K, const float *D, const double *uv, int n, double *out) {
    float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, T[3] = {0, 0, 0};
    CameraF64 c = make_cam(K, D, R, T, 2.0);
    for (int i = 0; i < n; ++i) undistort5(c, uv[2 * i], uv[2 * i + 1], out[2 * i], out[2 * i + 1]);
}

int hc_position3d(const float *K, const float *R, const float *T, double diam_mm, double u, double v, double d, double *P) {
    float D[5] = {0, 0, 0, 0, 0};
    CameraF64 c = make_cam(K, D, R, T, diam_mm);
    return position3d(c, u, v, d, P) ? 1 : 0;
}

int hc_plane(const double *X, const double *Y, const double *Z, int n, double out[4]) {
    double mx = 0, my = 0, mz = 0;
    for (int i = 0; i < n; ++i) { mx += X[i]; my += Y[i]; mz += Z[i]; }
    mx /= n; my /= n; mz /= n;
    double sxx = 0, sxy = 0, syy = 0, sxz = 0, syz = 0;
    for (int i = 0; i < n; ++i) {
        const double x = X[i] - mx, y = Y[i] - my, z = Z[i] - mz;
        sxx += x * x; sxy += x * y; syy += y * y; sxz += x * z; syz += y * z;
    }
    return plane_solve((double)n, mx, my, mz, sxx, sxy, syy, sxz, syz, out) ? 1 : 0;
}

}  // extern "C"

extern "C" {

static LensF64 make_lens(const double *K, const double *D, int nd) {
    LensF64 c;
    c.fx = K[0]; c.fy = K[4]; c.cx = K[2]; c.cy = K[5];
    double d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < nd && i < 8; ++i) d[i] = D[i];
    c.k1 = d[0]; c.k2 = d[1]; c.p1 = d[2]; c.p2 = d[3]; c.k3 = d[4]; c.k4 = d[5]; c.k5 = d[6]; c.k6 = d[7];
    return c;
}

// MD:93-109 geometry: new camera matrix {fx', fy', cx', cy'} and the CV_16SC2 maps
void hc_optimal_new_camera(const double *K, const double *D, int nd, int w, int h, double out[4]) {
    optimal_new_camera_alpha0(make_lens(K, D, nd), w, h, out);
}

void hc_rectify_maps(const double *K, const double *D, int nd, const double nk[4], int w, int h, int16_t *map1, uint16_t *map2) {
    const LensF64 c = make_lens(K, D, nd);
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j) {
            int iu, iv;
            rectify_source_q5(c, nk, i, j, iu, iv);
            map1[2 * ((size_t)i * w + j)] = (int16_t)(iu >> 5);
            map1[2 * ((size_t)i * w + j) + 1] = (int16_t)(iv >> 5);
            map2[(size_t)i * w + j] = (uint16_t)((iv & 31) * 32 + (iu & 31));
        }
}

void hc_undistort_normalized(const double *K, const double *D, int nd, const double *uv, int n, double *out) {
    const LensF64 c = make_lens(K, D, nd);
    for (int i = 0; i < n; ++i) undistort_normalized(c, uv[2 * i], uv[2 * i + 1], out[2 * i], out[2 * i + 1]);
}

// grid plan of the strip-marching kernels: out = {n_full, vsegs, seg_rows, ctas}
void hc_seg_plan(int H, long long items, int slots, int lead, int rb, double lead_cost, int mixed, int out[4]) {
    const VbsSegPlan p = vbs_seg_plan(H, items, slots, lead, rb, lead_cost, mixed != 0);
    out[0] = p.n_full; out[1] = p.vsegs; out[2] = p.seg_rows; out[3] = p.ctas;
}
}  // extern "C"
