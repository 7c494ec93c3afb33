// vbs_geom.h - per-marker geometry shared by the CUDA kernels (device) and the host unit
// tests (tests/hostcheck).  Everything here is scalar float64/float32 arithmetic written so
// that it rounds like the library routine it stands in for.  Citations: MD = reference
// code/Marker_Tracking/marker_detection.py, R3 = code/Marker_Calibration/3d_reconstruction.py,
// FD = code/ForceDistribution/ForceDistribution.py.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define VBS_HD __host__ __device__ __forceinline__
#else
#define VBS_HD inline
#endif

namespace vbs {

// ---- exact (non-contracted) float64 helpers -------------------------------------------------
// nvcc contracts a*b+c into fma by default; the reference's NumPy / SciPy / OpenCV builds do
// not (x86-64 baseline).  Where the last bit matters we spell the roundings out.
VBS_HD double mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b; return r;
#endif
}
VBS_HD double add_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
}
VBS_HD double sub_rn(double a, double b) { return add_rn(a, -b); }
VBS_HD float fadd_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b; return r;
#endif
}
VBS_HD float fsub_rn(float a, float b) { return fadd_rn(a, -b); }

// ---- border following (cv2.findContours, RETR_EXTERNAL + CHAIN_APPROX_SIMPLE; MD:196) -------
// Bits: functor  bool operator()(int x, int y)  -> foreground test, false outside the image.
// Visitor: functor void operator()(int x, int y) called for every kept vertex, in contour order.
// Start pixel = topmost-leftmost pixel of an 8-connected blob.  Returns the number of kept
// vertices, or -1 if the step guard tripped (never on a consistent bit image).
template <class Bits, class Visitor>
VBS_HD int trace_external_simple(const Bits &fg, int x0, int y0, long long max_steps, Visitor &visit) {
    // chain codes 0..7 = E,NE,N,NW,W,SW,S,SE; (dx+1, dy+1) packed 2 bits per code so the lookup is
    // two shifts instead of an indexed local array
    //   dx+1 = {2,2,1,0,0,0,1,2}   dy+1 = {1,0,0,0,1,2,2,2}
    const unsigned PX = 2u | (2u << 2) | (1u << 4) | (0u << 6) | (0u << 8) | (0u << 10) | (1u << 12) | (2u << 14);
    const unsigned PY = 1u | (0u << 2) | (0u << 4) | (0u << 6) | (1u << 8) | (2u << 10) | (2u << 12) | (2u << 14);
#define VBS_DX(s_) ((int)((PX >> (2 * (s_))) & 3u) - 1)
#define VBS_DY(s_) ((int)((PY >> (2 * (s_))) & 3u) - 1)
    int s = 4;
    const int s_first_end = 4;
    bool found = false;
    // clockwise search for the "previous" border pixel
    for (int it = 0; it < 8; ++it) {
        s = (s - 1) & 7;
        if (fg(x0 + VBS_DX(s), y0 + VBS_DY(s))) { found = true; break; }
        if (s == s_first_end) break;
    }
    if (!found) { visit(x0, y0); return 1; }          // isolated pixel
    const int x1 = x0 + VBS_DX(s), y1 = y0 + VBS_DY(s);
    int x3 = x0, y3 = y0;
    int prev_step = -1;
    int kept = 0;
    for (long long guard = 0; guard < max_steps; ++guard) {
        int x4 = x3, y4 = y3;
        for (int it = 0; it < 8; ++it) {              // counter-clockwise search for the next pixel
            s = (s + 1) & 7;
            x4 = x3 + VBS_DX(s); y4 = y3 + VBS_DY(s);
            if (fg(x4, y4)) break;
        }
        // (x3,y3) is a border point leaving with step s; the start point is always kept
        if (prev_step < 0 || prev_step != s) { visit(x3, y3); ++kept; }
        prev_step = s;
        if (x4 == x0 && y4 == y0 && x3 == x1 && y3 == y1) return kept;
        x3 = x4; y3 = y4;
        s = (s + 4) & 7;
    }
    return -1;
#undef VBS_DX
#undef VBS_DY
}

// ---- streaming least squares by Givens rotations ---------------------------------------------
// Rows arrive one at a time (the contour is re-traced, never stored); R is N x N upper
// triangular, d the rotated right-hand side.  Replaces cv::solve(DECOMP_SVD) / SVBackSubst in
// cv2.fitEllipse for full-rank systems (agreement ~1e-13 relative, outputs are float32).
template <int N>
struct GivensLsq {
    double R[N][N];
    double d[N];
    VBS_HD void reset() {
        for (int i = 0; i < N; ++i) { d[i] = 0.0; for (int j = 0; j < N; ++j) R[i][j] = 0.0; }
    }
    VBS_HD void add_row(double a[N], double beta) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < N; ++j) {
            const double aj = a[j];
            if (aj != 0.0) {
                const double rjj = R[j][j];
                // explicit fma everywhere: host (-ffp-contract=off) and device round alike.  The device takes the
                // reciprocal square root (1 ulp, a third of the instructions of sqrt + divide); the fit only has to
                // agree with cv2's SVD to ~1e-13 before its float32 outputs are rounded.
                const double q = fma(aj, aj, rjj * rjj);
#if defined(__CUDA_ARCH__)
                const double inv = rsqrt(q);
                const double h = q * inv;
#else
                const double h = sqrt(q);
                const double inv = 1.0 / h;
#endif
                const double c = rjj * inv, sn = aj * inv;
                R[j][j] = h;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int k = j + 1; k < N; ++k) {
                    const double t = R[j][k];
                    R[j][k] = fma(c, t, sn * a[k]);
                    a[k] = fma(c, a[k], -(sn * t));
                }
                const double t = d[j];
                d[j] = fma(c, t, sn * beta);
                beta = fma(c, beta, -(sn * t));
            }
        }
    }
    // returns false when R is singular
    VBS_HD bool solve(double x[N]) const {
        for (int i = N - 1; i >= 0; --i) {
            double acc = d[i];
            for (int k = i + 1; k < N; ++k) acc = fma(-R[i][k], x[k], acc);
            if (R[i][i] == 0.0) return false;
            x[i] = acc / R[i][i];
        }
        return true;
    }
};

// ---- cv2.fitEllipse (MD:208) as four passes over the contour ---------------------------------
struct EllipseResult {
    float cx, cy, w, h, angle;   // exactly the RotatedRect cv2 returns
    int ok;                      // 0 when the fit is degenerate (singular system / non-finite)
};

struct EllipsePassCentroid {     // pass 1: n and the float32 running sum in contour order
    int n; float sx, sy;
    VBS_HD void reset() { n = 0; sx = 0.f; sy = 0.f; }
    VBS_HD void operator()(int x, int y) { sx = fadd_rn(sx, (float)x); sy = fadd_rn(sy, (float)y); ++n; }
};
struct EllipsePassScale {        // pass 2: s = sum |qx| + |qy| in float64 of float32 offsets
    float cx, cy; double s;
    VBS_HD void operator()(int x, int y) {
        const float qx = fsub_rn((float)x, cx), qy = fsub_rn((float)y, cy);
        s = add_rn(s, add_rn(fabs((double)qx), fabs((double)qy)));
    }
};
struct EllipsePassConic {        // pass 3: [-px^2, -py^2, -px py, px, py] g = 10000
    float cx, cy; double scale; GivensLsq<5> q;
    VBS_HD void operator()(int x, int y) {
        const double px = mul_rn((double)fsub_rn((float)x, cx), scale);
        const double py = mul_rn((double)fsub_rn((float)y, cy), scale);
        double a[5] = {-mul_rn(px, px), -mul_rn(py, py), -mul_rn(px, py), px, py};
        q.add_row(a, 10000.0);
    }
};
struct EllipsePassAxes {         // pass 4: [(px-r0)^2, (py-r1)^2, (px-r0)(py-r1)] h = 1
    float cx, cy; double scale, r0, r1; GivensLsq<3> q;
    VBS_HD void operator()(int x, int y) {
        const double px = mul_rn((double)fsub_rn((float)x, cx), scale);
        const double py = mul_rn((double)fsub_rn((float)y, cy), scale);
        const double ex = sub_rn(px, r0), ey = sub_rn(py, r1);
        double a[3] = {mul_rn(ex, ex), mul_rn(ey, ey), mul_rn(ex, ey)};
        q.add_row(a, 1.0);
    }
};

// Final stage of fitEllipse from h = (h0,h1,h2), the centre estimate and the scale.
VBS_HD EllipseResult ellipse_from_fit(const double h[3], double r0, double r1, double scale, float cx, float cy) {
    const double min_eps = 1e-8;
    const double PI = 3.1415926535897932384626433832795;
    EllipseResult e;
    const double th = -0.5 * atan2(h[2], h[1] - h[0]);
    double t;
    if (fabs(h[2]) > min_eps) t = h[2] / sin(-2.0 * th);
    else t = h[1] - h[0];
    double a = fabs(h[0] + h[1] - t);
    if (a > min_eps) a = sqrt(2.0 / a);
    double b = fabs(h[0] + h[1] + t);
    if (b > min_eps) b = sqrt(2.0 / b);
    e.cx = fadd_rn((float)(r0 / scale), cx);
    e.cy = fadd_rn((float)(r1 / scale), cy);
    e.w = (float)(a * 2 / scale);
    e.h = (float)(b * 2 / scale);
    e.angle = 0.f;                                   // cv2 assigns the angle only when it swaps
    if (e.w > e.h) {
        const float tmp = e.w; e.w = e.h; e.h = tmp;
        e.angle = (float)(90 + th * 180 / PI);
    }
    if (e.angle < -180) e.angle += 360;
    if (e.angle > 360) e.angle -= 360;
    e.ok = (isfinite(e.cx) && isfinite(e.cy) && isfinite(e.w) && isfinite(e.h) && isfinite(e.angle)) ? 1 : 0;
    return e;
}

// 2x2 solve for the centre: [[2g0, g2],[g2, 2g1]] r = [g3, g4]  (cv::solve DECOMP_SVD in cv2)
VBS_HD bool ellipse_centre_solve(const double g[5], double &r0, double &r1) {
    const double a = 2 * g[0], b = g[2], c = g[2], d = 2 * g[1];
    // Gaussian elimination with partial pivoting
    if (fabs(a) >= fabs(c)) {
        if (a == 0.0) return false;
        const double m = c / a;
        const double dd = d - m * b;
        if (dd == 0.0) return false;
        r1 = (g[4] - m * g[3]) / dd;
        r0 = (g[3] - b * r1) / a;
    } else {
        const double m = a / c;
        const double bb = b - m * d;
        if (bb == 0.0) return false;
        r1 = (g[3] - m * g[4]) / bb;
        r0 = (g[4] - d * r1) / c;
    }
    return true;
}

// Whole fit.  `Source` replays the kept vertices of one contour in order:
//   template <class V> int operator()(V &visitor) const   -> number of vertices (or -1)
// (re-tracing the border, or reading back a stored vertex list).  n_out = kept vertices.
template <class Source>
VBS_HD EllipseResult fit_ellipse_from(const Source &src, int &n_out) {
    EllipseResult bad; bad.cx = bad.cy = bad.w = bad.h = bad.angle = 0.f; bad.ok = 0;
    EllipsePassCentroid p1; p1.reset();
    n_out = src(p1);
    if (n_out < 5) return bad;                       // MD:204 len(contour) < 5 (and the guard case)
    const float cx = p1.sx / (float)p1.n, cy = p1.sy / (float)p1.n;
    EllipsePassScale p2; p2.cx = cx; p2.cy = cy; p2.s = 0.0;
    src(p2);
    const double FLT_EPS = 1.1920928955078125e-07;
    const double scale = 100.0 / (p2.s > FLT_EPS ? p2.s : FLT_EPS);
    EllipsePassConic p3; p3.cx = cx; p3.cy = cy; p3.scale = scale; p3.q.reset();
    src(p3);
    double g[5];
    if (!p3.q.solve(g)) return bad;
    double r0, r1;
    if (!ellipse_centre_solve(g, r0, r1)) return bad;
    EllipsePassAxes p4; p4.cx = cx; p4.cy = cy; p4.scale = scale; p4.r0 = r0; p4.r1 = r1; p4.q.reset();
    src(p4);
    double h[3];
    if (!p4.q.solve(h)) return bad;
    return ellipse_from_fit(h, r0, r1, scale, cx, cy);
}

template <class Bits> struct TraceSource {            // replay by following the border again
    const Bits &fg; int x0, y0; long long max_steps;
    template <class V> VBS_HD int operator()(V &v) const { return trace_external_simple(fg, x0, y0, max_steps, v); }
};
struct StoredSource {                                 // replay a stored vertex list (x | y << 16)
    const uint32_t *pts; int n;
    template <class V> VBS_HD int operator()(V &v) const {
        for (int i = 0; i < n; ++i) { const uint32_t p = pts[i]; v((int)(p & 0xffffu), (int)(p >> 16)); }
        return n;
    }
};

template <class Bits>
VBS_HD EllipseResult fit_ellipse_traced(const Bits &fg, int x0, int y0, long long max_steps, int &n_out) {
    TraceSource<Bits> src{fg, x0, y0, max_steps};
    return fit_ellipse_from(src, n_out);
}

// ---- cv2.pointPolygonTest(contour, (x, y), False) (MD:228) as an edge visitor ----------------
// Feed the polygon vertices in order; finish() closes it.  result(): +1 inside, 0 on the border,
// -1 outside.  The query point is rounded to float32 first (cv::Point2f), then either the
// integer branch (int64 cross products) or the float branch (double cross products) is used.
struct PointPolygon {
    float fx, fy; bool is_int; long long ix, iy;
    int first_x, first_y, prev_x, prev_y; int nv; int counter; bool on_edge;
    VBS_HD void init(double x, double y) {
        fx = (float)x; fy = (float)y;
        const double rx = rint((double)fx), ry = rint((double)fy);
        is_int = (rx == (double)fx) && (ry == (double)fy);
        ix = (long long)rx; iy = (long long)ry;
        nv = 0; counter = 0; on_edge = false;
        first_x = first_y = prev_x = prev_y = 0;
    }
    VBS_HD void edge(int v0x, int v0y, int vx, int vy) {
        if (is_int) {
            if ((v0y <= iy && vy <= iy) || (v0y > iy && vy > iy) || (v0x < ix && vx < ix)) {
                if (iy == vy && (ix == vx || (iy == v0y && ((v0x <= ix && ix <= vx) || (vx <= ix && ix <= v0x)))))
                    on_edge = true;
                return;
            }
            long long dist = (iy - v0y) * (long long)(vx - v0x) - (ix - v0x) * (long long)(vy - v0y);
            if (dist == 0) { on_edge = true; return; }
            if (vy < v0y) dist = -dist;
            counter += dist > 0;
        } else {
            const float a0x = (float)v0x, a0y = (float)v0y, ax = (float)vx, ay = (float)vy;
            if ((a0y <= fy && ay <= fy) || (a0y > fy && ay > fy) || (a0x < fx && ax < fx)) {
                if (fy == ay && (fx == ax || (fy == a0y && ((a0x <= fx && fx <= ax) || (ax <= fx && fx <= a0x)))))
                    on_edge = true;
                return;
            }
            // (double)(pt.y - v0.y)*(v.x - v0.x) - (double)(pt.x - v0.x)*(v.y - v0.y): float32
            // differences, first factor widened, products and difference in float64
            double dist = sub_rn(mul_rn((double)fsub_rn(fy, a0y), (double)fsub_rn(ax, a0x)),
                                 mul_rn((double)fsub_rn(fx, a0x), (double)fsub_rn(ay, a0y)));
            if (dist == 0) { on_edge = true; return; }
            if (ay < a0y) dist = -dist;
            counter += dist > 0;
        }
    }
    VBS_HD void operator()(int x, int y) {
        if (nv == 0) { first_x = x; first_y = y; }
        else edge(prev_x, prev_y, x, y);
        prev_x = x; prev_y = y; ++nv;
    }
    VBS_HD int result() {
        if (nv == 0) return -1;
        edge(prev_x, prev_y, first_x, first_y);       // closing edge (cv2 starts with it)
        if (on_edge) return 0;
        return (counter & 1) ? 1 : -1;
    }
};

// ---- cv2.undistortPoints(pts, K, D, None, K) (R3:187-193): exactly 5 iterations ---------------
struct CameraF64 {
    double fx, fy, cx, cy;          // float32 values widened
    double k1, k2, p1, p2, k3;
    double R[9], T[3];              // world->cam, float32 values widened
    double f_avg;                   // float32((fx+fy)/2) widened        (R3:211)
    double f_avg_sq;                // float32(f_avg * f_avg) widened: `f_avg**2` on an np.float32 scalar stays float32 (R3:219)
    double ratio;                   // float32(diam_mm / f_avg) widened  (R3:219)
    double min_size, max_disp;
};

VBS_HD void undistort5(const CameraF64 &c, double u, double v, double &uo, double &vo) {
    const double x0 = mul_rn(u - c.cx, 1.0 / c.fx), y0 = mul_rn(v - c.cy, 1.0 / c.fy);   // cv2 multiplies by 1/f (bit-exact this way)
    double x = x0, y = y0;
    for (int it = 0; it < 5; ++it) {
        const double r2 = add_rn(mul_rn(x, x), mul_rn(y, y));
        const double poly = add_rn(1.0, mul_rn(add_rn(mul_rn(add_rn(mul_rn(c.k3, r2), c.k2), r2), c.k1), r2));
        const double icd = 1.0 / poly;
        if (icd < 0) { x = x0; y = y0; break; }   // cv2 bails out on a negative factor
        const double dx = add_rn(mul_rn(mul_rn(mul_rn(2.0, c.p1), x), y), mul_rn(c.p2, add_rn(r2, mul_rn(mul_rn(2.0, x), x))));
        const double dy = add_rn(mul_rn(c.p1, add_rn(r2, mul_rn(mul_rn(2.0, y), y))), mul_rn(mul_rn(mul_rn(2.0, c.p2), x), y));
        x = mul_rn(sub_rn(x0, dx), icd);
        y = mul_rn(sub_rn(y0, dy), icd);
    }
    uo = add_rn(mul_rn(x, c.fx), c.cx);
    vo = add_rn(mul_rn(y, c.fy), c.cy);
}

// ---- MarkerTracker._undistort_frame (MD:93-109): lens model of the frame remap ----------------
// float64 throughout (MD:96-97 builds K and D with np.array on Python lists); rational model k4..k6 = 0
// unless 8 coefficients are given.
struct LensF64 {
    double fx, fy, cx, cy;
    double k1, k2, p1, p2, k3, k4, k5, k6;
};

// cv2.undistortPoints(pt, K, D): 5 fixed-point iterations, normalised coordinates out
VBS_HD void undistort_normalized(const LensF64 &c, double u, double v, double &xo, double &yo) {
    const double x0 = mul_rn(u - c.cx, 1.0 / c.fx), y0 = mul_rn(v - c.cy, 1.0 / c.fy);
    double x = x0, y = y0;
    for (int it = 0; it < 5; ++it) {
        const double r2 = add_rn(mul_rn(x, x), mul_rn(y, y));
        const double num = add_rn(1.0, mul_rn(add_rn(mul_rn(add_rn(mul_rn(c.k6, r2), c.k5), r2), c.k4), r2));
        const double den = add_rn(1.0, mul_rn(add_rn(mul_rn(add_rn(mul_rn(c.k3, r2), c.k2), r2), c.k1), r2));
        const double icd = num / den;
        if (icd < 0) { x = x0; y = y0; break; }
        const double dx = add_rn(mul_rn(mul_rn(mul_rn(2.0, c.p1), x), y), mul_rn(c.p2, add_rn(r2, mul_rn(mul_rn(2.0, x), x))));
        const double dy = add_rn(mul_rn(c.p1, add_rn(r2, mul_rn(mul_rn(2.0, y), y))), mul_rn(mul_rn(mul_rn(2.0, c.p2), x), y));
        x = mul_rn(sub_rn(x0, dx), icd);
        y = mul_rn(sub_rn(y0, dy), icd);
    }
    xo = x; yo = y;
}

// cv2.getOptimalNewCameraMatrix(K, D, (w,h), alpha = 0, (w,h)): the largest axis-aligned rectangle of
// normalised coordinates that a 9 x 9 grid of undistorted image points proves to be free of invalid
// pixels is stretched over the full image.  out = {fx', fy', cx', cy'}
inline void optimal_new_camera_alpha0(const LensF64 &c, int w, int h, double out[4]) {
    const int N = 9;
    double ix0 = -1e300, ix1 = 1e300, iy0 = -1e300, iy1 = 1e300;
    for (int y = 0; y < N; ++y)
        for (int x = 0; x < N; ++x) {
            double px, py;
            undistort_normalized(c, (double)x * (w - 1) / (N - 1), (double)y * (h - 1) / (N - 1), px, py);
            if (x == 0 && px > ix0) ix0 = px;
            if (x == N - 1 && px < ix1) ix1 = px;
            if (y == 0 && py > iy0) iy0 = py;
            if (y == N - 1 && py < iy1) iy1 = py;
        }
    const double fx0 = (w - 1) / (ix1 - ix0), fy0 = (h - 1) / (iy1 - iy0);
    out[0] = fx0; out[1] = fy0; out[2] = -fx0 * ix0; out[3] = -fy0 * iy0;
}

// cv2.initUndistortRectifyMap(K, D, None, newK, size, CV_16SC2) for destination pixel (column j, row i):
// source position in 1/32 px; nk = {fx', fy', cx', cy'}
VBS_HD void rectify_source_q5(const LensF64 &c, const double nk[4], int i, int j, int &iu, int &iv) {
    const double irx = 1.0 / nk[0], iry = 1.0 / nk[1];                     // inverse of [[fx',0,cx'],[0,fy',cy'],[0,0,1]]
    const double x = add_rn(mul_rn((double)j, irx), -nk[2] / nk[0]);
    const double y = add_rn(mul_rn((double)i, iry), -nk[3] / nk[1]);
    const double x2 = mul_rn(x, x), y2 = mul_rn(y, y), r2 = add_rn(x2, y2), xy2 = mul_rn(mul_rn(2.0, x), y);
    const double kr = add_rn(1.0, mul_rn(add_rn(mul_rn(add_rn(mul_rn(c.k3, r2), c.k2), r2), c.k1), r2)) /
                      add_rn(1.0, mul_rn(add_rn(mul_rn(add_rn(mul_rn(c.k6, r2), c.k5), r2), c.k4), r2));
    const double xd = add_rn(add_rn(mul_rn(x, kr), mul_rn(c.p1, xy2)), mul_rn(c.p2, add_rn(r2, mul_rn(2.0, x2))));
    const double yd = add_rn(add_rn(mul_rn(y, kr), mul_rn(c.p1, add_rn(r2, mul_rn(2.0, y2)))), mul_rn(c.p2, xy2));
    const double u = add_rn(mul_rn(c.fx, xd), c.cx), v = add_rn(mul_rn(c.fy, yd), c.cy);
    const double su = mul_rn(u, 32.0), sv = mul_rn(v, 32.0);               // cvRound(u * INTER_TAB_SIZE), saturating
    iu = su >= 2147483647.0 ? 2147483647 : su <= -2147483648.0 ? (int)(-2147483647 - 1) : (int)rint(su);
    iv = sv >= 2147483647.0 ? 2147483647 : sv <= -2147483648.0 ? (int)(-2147483647 - 1) : (int)rint(sv);
}

// cv2.remap(INTER_LINEAR, BORDER_CONSTANT 0) on uint8 with CV_16SC2 maps: 5-bit fractions, weights scaled
// to 2^15 (exact products of 1/32 steps, so no table fix-up ever applies), (sum + 2^14) >> 15
struct RemapTap { int sx, sy; int w00, w01, w10, w11; };
VBS_HD RemapTap remap_tap(int iu, int iv) {
    RemapTap t;
    t.sx = (int)(short)(iu >> 5); t.sy = (int)(short)(iv >> 5);           // map1 is int16: cv2 casts, it does not saturate
    const int fx = iu & 31, fy = iv & 31;
    t.w00 = (32 - fx) * (32 - fy) * 32; t.w01 = fx * (32 - fy) * 32; t.w10 = (32 - fx) * fy * 32; t.w11 = fx * fy * 32;
    return t;
}

// ---- MarkerAnalysis._calculate_3d_position (R3:195-238) ---------------------------------------
// returns false when the reference would raise (R < 1e-6 or non-finite result)
VBS_HD bool position3d(const CameraF64 &c, double u, double v, double diam, double P[3]) {
    const double du = u - c.cx, dv = v - c.cy;
    const double rad = sqrt(add_rn(mul_rn(du, du), mul_rn(dv, dv)));
    if (rad < 1e-6) return false;
    const double d_eff = mul_rn(c.ratio, sqrt(add_rn(mul_rn(rad, rad), c.f_avg_sq)));
    const double h = mul_rn(c.f_avg, d_eff / diam);
    const double pc0 = sub_rn(mul_rn(h, du) / c.fx, c.T[0]);
    const double pc1 = sub_rn(mul_rn(h, dv) / c.fy, c.T[1]);
    const double pc2 = sub_rn(h, c.T[2]);
    for (int i = 0; i < 3; ++i)       // R^T (Pc - T): column i of R
        P[i] = add_rn(add_rn(mul_rn(c.R[0 * 3 + i], pc0), mul_rn(c.R[1 * 3 + i], pc1)), mul_rn(c.R[2 * 3 + i], pc2));
    return isfinite(P[0]) && isfinite(P[1]) && isfinite(P[2]);
}

// ---- fit_plane_least_squares (FD:141-159): centred normal equations ---------------------------
struct PlaneSums {
    double n, sx, sy, sz;
    VBS_HD void reset() { n = sx = sy = sz = 0.0; }
};
VBS_HD bool plane_solve(double n, double mx, double my, double mz, double sxx, double sxy, double syy,
                        double sxz, double syz, double out[4]) {
    const double PI = 3.1415926535897932384626433832795;
    const double det = sxx * syy - sxy * sxy;
    if (!(n >= 3.0) || det == 0.0 || !isfinite(det)) return false;
    const double a = (sxz * syy - syz * sxy) / det;
    const double b = (syz * sxx - sxz * sxy) / det;
    out[0] = a; out[1] = b; out[2] = mz - a * mx - b * my;
    out[3] = atan(sqrt(a * a + b * b)) * (180.0 / PI);
    return true;
}

}  // namespace vbs
