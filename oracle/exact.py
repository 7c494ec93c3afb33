"""Exact restatements of the third-party arithmetic on the hot path.  TEST INFRASTRUCTURE ONLY.

``oracle/port.py`` calls OpenCV / SciPy like the reference does.  This module restates
what those calls *compute* (integer fixed-point blur, the NCC closed form, window
maxima, border following, the ellipse fit, the point-in-polygon rule, the 5-iteration
undistort, the float32 roundings in the depth formula, the plane fit) in plain NumPy,
because that is what the CUDA kernels implement.  Each restatement is pinned against
the library it restates in ``tests/test_oracle_exact.py`` (cv2 4.13.0, scipy 1.18.1,
numpy 2.3.5 - the versions in the image; the reference pins none).

Third-party algorithms restated here (none is vendored in /root/reference):
  OpenCV 4.13.0 : GaussianBlur (8.8 fixed-point path for CV_8U), morphologyEx(OPEN),
                  findContours (Suzuki-Abe border following, CHAIN_APPROX_SIMPLE),
                  fitEllipse (LIN fit, Fitzgibbon-style general conic), pointPolygonTest,
                  undistortPoints (5 fixed-point iterations)
  SciPy 1.18.1  : fftconvolve-based NCC, maximum/minimum_filter, ndimage.label,
                  center_of_mass, cdist
Reference call sites: marker_detection.py:114-133,170-243,369; 3d_reconstruction.py:187-228;
ForceDistribution.py:141-159.
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------------
# a3: fixed-point Gaussian blur (cv2.GaussianBlur on uint8; call sites MD:118-119,123-124)
# ----------------------------------------------------------------------------------
def fixed_gaussian_kernel(ksize: int, sigma: float) -> np.ndarray:
    """8.8 fixed-point taps (sum 256): error diffusion from the outside in, remainder at the centre."""
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) / 2.0
    k = np.exp(-(x * x) / (2.0 * sigma * sigma))
    k /= k.sum()
    out = np.zeros(ksize, dtype=np.int64)
    err = 0.0
    half = ksize // 2
    for i in range(half):
        adj = k[i] * 256.0 + err
        v = int(np.rint(adj))            # round half to even, like cvRound
        err = adj - v
        out[i] = out[ksize - 1 - i] = v
    out[half] = 256 - 2 * int(out[:half].sum())
    return out


def blur_fixed(gray: np.ndarray, taps: np.ndarray) -> np.ndarray:
    """Separable integer blur, BORDER_REFLECT_101, one rounding at the end: (acc + 32768) >> 16."""
    r = len(taps) // 2
    p = np.pad(gray.astype(np.int64), r, mode="reflect")
    h, w = gray.shape
    hp = np.zeros((h + 2 * r, w), dtype=np.int64)
    for i, t in enumerate(taps):
        if t:
            hp += int(t) * p[:, i : i + w]
    acc = np.zeros((h, w), dtype=np.int64)
    for i, t in enumerate(taps):
        if t:
            acc += int(t) * hp[i : i + h, :]
    return ((acc + 32768) >> 16).astype(np.uint8)


def area_mask_exact(gray: np.ndarray) -> np.ndarray:
    """uint8 {0,255} area mask: wrapping DoG + inRange (MD:117-129)."""
    from .port import branch_constants

    c = branch_constants(gray.shape[0])
    small = blur_fixed(gray, fixed_gaussian_kernel(c["k_small"], c["s_small"]))
    large = blur_fixed(gray, fixed_gaussian_kernel(c["k_large"], c["s_large"]))
    dog = (large.astype(np.int32) - small.astype(np.int32) + 15) & 255
    return np.where((dog >= c["lo"]) & (dog <= c["hi"]), 255, 0).astype(np.uint8)


def gray_exact(bgr: np.ndarray) -> np.ndarray:
    """cvtColor(BGR2GRAY) on uint8: 15-bit fixed point (MD:114)."""
    b = bgr[..., 0].astype(np.int64)
    g = bgr[..., 1].astype(np.int64)
    r = bgr[..., 2].astype(np.int64)
    return ((3735 * b + 19235 * g + 9798 * r + 16384) >> 15).astype(np.uint8)


# ----------------------------------------------------------------------------------
# a5-a7: NCC against the Gaussian template in closed form (MD:120,125,132-133,138-164)
# ----------------------------------------------------------------------------------
def template_1d(l: int, sig: float) -> np.ndarray:
    """1-D factor n of the l x l template (template == outer(n, n) up to rounding)."""
    ax = np.linspace(-(l - 1) / 2.0, (l - 1) / 2.0, l)
    e = np.exp(-0.5 * np.square(ax) / np.square(sig))
    return e / e.sum()


def _window_sum_1d(a: np.ndarray, weights: np.ndarray, off: int, axis: int) -> np.ndarray:
    """out[i] = sum_j weights[j] * a[i - off + j], zero outside (correlation, 'same' size)."""
    l = len(weights)
    pad = [(0, 0)] * a.ndim
    pad[axis] = (off, l - 1 - off)
    p = np.pad(a, pad)
    n = a.shape[axis]
    out = np.zeros(a.shape, dtype=np.float64)
    for j in range(l):
        sl = [slice(None)] * a.ndim
        sl[axis] = slice(j, j + n)
        out += weights[j] * p[tuple(sl)]
    return out


def ncc_closed_form(area_mask: np.ndarray, l: int, sig: float):
    """(ncc float64 HxW, mask uint8) from the separable/integer decomposition.

    Window of output pixel i on each axis: [i - off, i - off + l - 1], off = (l-1) - (l-1)//2.
    """
    h, w = area_mask.shape
    b = (area_mask > 0).astype(np.float64)
    mu = float(np.mean(area_mask))
    n = template_1d(l, sig)
    off = (l - 1) - (l - 1) // 2
    ones = np.ones(l)
    gb = _window_sum_1d(_window_sum_1d(b, n, off, 1), n, off, 0)
    s = _window_sum_1d(_window_sum_1d(b, ones, off, 1), ones, off, 0)
    inside = np.ones((h, w))
    g1 = _window_sum_1d(_window_sum_1d(inside, n, off, 1), n, off, 0)
    a = _window_sum_1d(_window_sum_1d(inside, ones, off, 1), ones, off, 0)
    l2 = float(l * l)
    sep = 255.0 * gb - mu * g1
    box1 = 255.0 * s - mu * a
    box2 = 255.0 * 255.0 * s - 510.0 * mu * s + mu * mu * a
    num = sep - box1 / l2
    isq = np.maximum(box2 - box1 * box1 / l2, 0.0)
    st2 = float(np.sum(n * n)) ** 2 - 1.0 / l2
    with np.errstate(divide="ignore", invalid="ignore"):
        ncc = num / np.sqrt(isq * st2)
    ncc[~np.isfinite(ncc)] = 0.0
    return ncc, (ncc > 0.1).astype(np.uint8)


# ----------------------------------------------------------------------------------
# a8-a10: ring maxima, 4-connected labels in raster order, centroids (MD:170-181)
# ----------------------------------------------------------------------------------
def window_all_ones(mask: np.ndarray, size: int) -> np.ndarray:
    """True where every in-image pixel of the window [i - size//2, i + size//2 - 1]^2 is set."""
    lo = size // 2
    hi = size - 1 - lo
    h, w = mask.shape
    p = np.pad(mask > 0, ((lo, hi), (lo, hi)), constant_values=True)
    out = np.ones((h, w), dtype=bool)
    for dy in range(size):
        for dx in range(size):
            out &= p[dy : dy + h, dx : dx + w]
    return out


def ring_maxima_exact(mask: np.ndarray) -> np.ndarray:
    size = 8 if mask.shape[0] <= 480 else 14
    return (mask > 0) & ~window_all_ones(mask, size)


def label4_raster(img: np.ndarray):
    """4-connected labels numbered by the raster index of each component's first pixel."""
    h, w = img.shape
    parent = np.arange(h * w, dtype=np.int64)

    def find(i):
        while parent[i] != i:
            parent[i] = parent[parent[i]]
            i = parent[i]
        return i

    ys, xs = np.nonzero(img)
    for y, x in zip(ys, xs):
        i = y * w + x
        for ny, nx in ((y, x - 1), (y - 1, x)):
            if ny >= 0 and nx >= 0 and img[ny, nx]:
                a, b = find(i), find(ny * w + nx)
                if a != b:
                    parent[max(a, b)] = min(a, b)
    lab = np.zeros((h, w), dtype=np.int32)
    ids = {}
    for y, x in zip(ys, xs):
        r = find(y * w + x)
        if r not in ids:
            ids[r] = len(ids) + 1
    # roots are minimal raster indices, so sorting them gives the raster numbering
    order = {r: k + 1 for k, r in enumerate(sorted(ids))}
    for y, x in zip(ys, xs):
        lab[y, x] = order[find(y * w + x)]
    return lab, len(order)


def centroids_from_labels(lab: np.ndarray, n: int) -> np.ndarray:
    """(row, col) means: exact integer sums, one float64 division each."""
    ys, xs = np.nonzero(lab)
    l = lab[ys, xs] - 1
    cnt = np.bincount(l, minlength=n).astype(np.float64)
    sy = np.bincount(l, weights=ys.astype(np.float64), minlength=n)
    sx = np.bincount(l, weights=xs.astype(np.float64), minlength=n)
    return np.stack([sy / cnt, sx / cnt], axis=1)


# ----------------------------------------------------------------------------------
# a11: 5x5 open, external contours by border following (MD:194-196)
# ----------------------------------------------------------------------------------
def open5(area: np.ndarray) -> np.ndarray:
    """erode (outside = foreground) then dilate (outside = background), 5x5, anchor centre."""
    b = area > 0
    h, w = b.shape
    p = np.pad(b, 2, constant_values=True)
    er = np.ones((h, w), dtype=bool)
    for dy in range(5):
        for dx in range(5):
            er &= p[dy : dy + h, dx : dx + w]
    p = np.pad(er, 2, constant_values=False)
    di = np.zeros((h, w), dtype=bool)
    for dy in range(5):
        for dx in range(5):
            di |= p[dy : dy + h, dx : dx + w]
    return di


_DX = (1, 1, 0, -1, -1, -1, 0, 1)
_DY = (0, -1, -1, -1, 0, 1, 1, 1)


def trace_border(fg: np.ndarray, x0: int, y0: int):
    """Full 8-connected border chain from start pixel (x0, y0) (SURVEY A.6 rule); list of (x, y)."""
    h, w = fg.shape

    def at(x, y):
        return 0 <= x < w and 0 <= y < h and fg[y, x]

    s = s_end = 4
    while True:
        s = (s - 1) & 7
        if at(x0 + _DX[s], y0 + _DY[s]) or s == s_end:
            break
    if s == s_end and not at(x0 + _DX[s], y0 + _DY[s]):
        return [(x0, y0)], [None]
    x1, y1 = x0 + _DX[s], y0 + _DY[s]
    pts, steps = [], []
    x3, y3 = x0, y0
    while True:
        s_end = s
        while True:
            s = (s + 1) & 7
            x4, y4 = x3 + _DX[s], y3 + _DY[s]
            if at(x4, y4):
                break
        pts.append((x3, y3))
        steps.append(s)
        if (x4, y4) == (x0, y0) and (x3, y3) == (x1, y1):
            break
        x3, y3 = x4, y4
        s = (s + 4) & 7
    return pts, steps


def approx_simple(pts, steps):
    """CHAIN_APPROX_SIMPLE: keep points whose incoming step differs from the outgoing step (cyclic)."""
    n = len(pts)
    if n == 1:
        return list(pts)
    return [pts[i] for i in range(n) if steps[i - 1] != steps[i]]


def label8_min(fg: np.ndarray):
    """8-connected components; returns dict root_index -> list of pixel raster indices (root = min index)."""
    from scipy import ndimage

    lab, n = ndimage.label(fg, structure=np.ones((3, 3)))
    return lab, n


def external_contours(opened: np.ndarray):
    """Contours like findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE): descending start raster index.

    A component is external iff the background next to its start pixel is 4-connected to the
    outside of the image (components inside another component's hole get no contour).
    """
    from scipy import ndimage

    fg = opened > 0
    h, w = fg.shape
    lab, n = ndimage.label(fg, structure=np.ones((3, 3)))
    # outer background: 4-connected background components touching the (virtual) outside
    bgp = np.pad(~fg, 1, constant_values=True)
    bl, _ = ndimage.label(bgp)
    outer = bl == bl[0, 0]
    starts = ndimage.minimum(np.arange(h * w).reshape(h, w), lab, range(1, n + 1)) if n else []
    out = []
    for s in sorted((int(v) for v in starts), reverse=True):
        y0, x0 = divmod(s, w)
        if not outer[y0 + 1, x0]:        # padded coords: pixel (x0-1, y0) -> [y0+1, x0]
            continue
        pts, steps = trace_border(fg, x0, y0)
        out.append(np.array(approx_simple(pts, steps), dtype=np.int32).reshape(-1, 1, 2))
    return out


# ----------------------------------------------------------------------------------
# a12: cv2.fitEllipse restated (MD:208)
# ----------------------------------------------------------------------------------
def fit_ellipse_exact(contour: np.ndarray):
    """((cx, cy), (w, h), angle) float32 values as Python floats, like cv2.fitEllipse."""
    p = contour.reshape(-1, 2).astype(np.float32)
    n = len(p)
    c = np.zeros(2, dtype=np.float32)
    for q in p:                           # sequential float32 accumulation, contour order
        c = (c + q).astype(np.float32)
    c = (c / np.float32(n)).astype(np.float32)
    q = (p - c).astype(np.float32)
    s = float(np.sum(np.abs(q[:, 0].astype(np.float64)) + np.abs(q[:, 1].astype(np.float64))))
    scale = 100.0 / max(s, float(np.finfo(np.float32).eps))
    px = q[:, 0].astype(np.float64) * scale
    py = q[:, 1].astype(np.float64) * scale
    A = np.stack([-px * px, -py * py, -px * py, px, py], axis=1)
    g = np.linalg.lstsq(A, np.full(n, 10000.0), rcond=None)[0]
    M = np.array([[2 * g[0], g[2]], [g[2], 2 * g[1]]])
    rp = np.linalg.lstsq(M, np.array([g[3], g[4]]), rcond=None)[0]
    B = np.stack([(px - rp[0]) ** 2, (py - rp[1]) ** 2, (px - rp[0]) * (py - rp[1])], axis=1)
    hh = np.linalg.lstsq(B, np.ones(n), rcond=None)[0]
    th = -0.5 * np.arctan2(hh[2], hh[1] - hh[0])
    t = hh[2] / np.sin(-2.0 * th) if abs(hh[2]) > 1e-8 else hh[1] - hh[0]
    a = abs(hh[0] + hh[1] - t)
    a = np.sqrt(2.0 / a) if a > 1e-8 else a
    b = abs(hh[0] + hh[1] + t)
    b = np.sqrt(2.0 / b) if b > 1e-8 else b
    cx = np.float32(np.float32(rp[0] / scale) + c[0])
    cy = np.float32(np.float32(rp[1] / scale) + c[1])
    wv = np.float32(a * 2 / scale)
    hv = np.float32(b * 2 / scale)
    ang = np.float32(0.0)                 # only the swap branch assigns the angle (t > 0 there)
    if wv > hv:
        wv, hv = hv, wv
        ang = np.float32(90.0 + th * 180.0 / np.pi)
    if ang < -180:
        ang = np.float32(ang + np.float32(360))
    if ang > 360:
        ang = np.float32(ang - np.float32(360))
    return (float(cx), float(cy)), (float(wv), float(hv)), float(ang)


# ----------------------------------------------------------------------------------
# a13: cv2.pointPolygonTest(contour, (x, y), False) restated (MD:228)
# ----------------------------------------------------------------------------------
def point_polygon_sign(contour: np.ndarray, x: float, y: float) -> int:
    """+1 inside, 0 on the border, -1 outside; the query point is rounded to float32 first."""
    pts = contour.reshape(-1, 2)
    fx = np.float32(x)
    fy = np.float32(y)
    ix, iy = int(np.rint(fx)), int(np.rint(fy))
    if ix == fx and iy == fy:
        qx, qy = ix, iy                   # integer branch: int64 cross products
    else:
        qx, qy = float(fx), float(fy)     # float branch: double cross products of float32 values
    counter = 0
    vx, vy = int(pts[-1][0]), int(pts[-1][1])
    for k in range(len(pts)):
        v0x, v0y = vx, vy
        vx, vy = int(pts[k][0]), int(pts[k][1])
        if (v0y <= qy and vy <= qy) or (v0y > qy and vy > qy) or (v0x < qx and vx < qx):
            if qy == vy and (qx == vx or (qy == v0y and ((v0x <= qx <= vx) or (vx <= qx <= v0x)))):
                return 0
            continue
        dist = (qy - v0y) * (vx - v0x) - (qx - v0x) * (vy - v0y)
        if dist == 0:
            return 0
        if vy < v0y:
            dist = -dist
        counter += dist > 0
    return -1 if counter % 2 == 0 else 1


# ----------------------------------------------------------------------------------
# a15-a16: undistortPoints (5 iterations) and the depth formula with its float32 roundings
# ----------------------------------------------------------------------------------
def undistort_exact(K, D, pts):
    """cv2.undistortPoints(pts, K, D, None, K): exactly 5 fixed-point iterations in float64 (R3:187-193)."""
    K = np.asarray(K, dtype=np.float32).astype(np.float64)
    k1, k2, p1, p2, k3 = (float(v) for v in np.asarray(D, dtype=np.float32).astype(np.float64)[:5])
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
    n = undistort_normalized_exact(K, [k1, k2, p1, p2, k3], pts)
    return np.stack([n[:, 0] * fx + cx, n[:, 1] * fy + cy], axis=1)


def _lens(K, D):
    K = np.asarray(K, dtype=np.float64)
    d = np.zeros(8)
    D = np.asarray(D, dtype=np.float64).ravel()
    d[: min(len(D), 8)] = D[:8]
    return K[0, 0], K[1, 1], K[0, 2], K[1, 2], d


def undistort_normalized_exact(K, D, pts):
    """cv2.undistortPoints(pts, K, D) (normalised output), bit for bit: (u - cx) * (1 / fx), exactly 5
    iterations, rational factor, and the bail-out to the start value when the factor turns negative."""
    fx, fy, cx, cy, (k1, k2, p1, p2, k3, k4, k5, k6) = _lens(K, D)
    ifx, ify = 1.0 / fx, 1.0 / fy
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
    out = np.empty_like(pts)
    for i, (u, v) in enumerate(pts):
        x0, y0 = (u - cx) * ifx, (v - cy) * ify
        x, y = x0, y0
        for _ in range(5):
            r2 = x * x + y * y
            icd = (1 + ((k6 * r2 + k5) * r2 + k4) * r2) / (1 + ((k3 * r2 + k2) * r2 + k1) * r2)
            if icd < 0:
                x, y = x0, y0
                break
            dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
            dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
            x, y = (x0 - dx) * icd, (y0 - dy) * icd
        out[i] = (x, y)
    return out


# ---- MarkerTracker._undistort_frame (MD:93-109) ------------------------------------------------------
def optimal_new_camera_matrix_exact(K, D, w: int, h: int) -> np.ndarray:
    """cv2.getOptimalNewCameraMatrix(K, D, (w, h), 0, (w, h))[0]: inner rectangle of a 9 x 9 undistorted grid
    (normalised coordinates) stretched over the image."""
    n = 9
    grid = np.array([[x * (w - 1) / (n - 1), y * (h - 1) / (n - 1)] for y in range(n) for x in range(n)], dtype=np.float64)
    g = undistort_normalized_exact(K, D, grid).reshape(n, n, 2)
    ix0, ix1 = g[:, 0, 0].max(), g[:, n - 1, 0].min()
    iy0, iy1 = g[0, :, 1].max(), g[n - 1, :, 1].min()
    fx0, fy0 = (w - 1) / (ix1 - ix0), (h - 1) / (iy1 - iy0)
    return np.array([[fx0, 0.0, -fx0 * ix0], [0.0, fy0, -fy0 * iy0], [0.0, 0.0, 1.0]])


def rectify_maps_exact(K, D, new_k, w: int, h: int):
    """cv2.initUndistortRectifyMap(K, D, None, new_k, (w, h), CV_16SC2): (map1 int16 [h,w,2], map2 uint16 [h,w])."""
    fx, fy, u0, v0, (k1, k2, p1, p2, k3, k4, k5, k6) = _lens(K, D)
    nk = np.asarray(new_k, dtype=np.float64)
    irx, iry, ox, oy = 1.0 / nk[0, 0], 1.0 / nk[1, 1], -nk[0, 2] / nk[0, 0], -nk[1, 2] / nk[1, 1]
    j = np.arange(w, dtype=np.float64)[None, :]
    i = np.arange(h, dtype=np.float64)[:, None]
    x = j * irx + ox + 0 * i
    y = i * iry + oy + 0 * j
    x2, y2 = x * x, y * y
    r2, xy2 = x2 + y2, 2 * x * y
    kr = (1 + ((k3 * r2 + k2) * r2 + k1) * r2) / (1 + ((k6 * r2 + k5) * r2 + k4) * r2)
    xd = x * kr + p1 * xy2 + p2 * (r2 + 2 * x2)
    yd = y * kr + p1 * (r2 + 2 * y2) + p2 * xy2
    iu = np.rint((fx * xd + u0) * 32).astype(np.int64)
    iv = np.rint((fy * yd + v0) * 32).astype(np.int64)
    map1 = np.stack([(iu >> 5).astype(np.int16), (iv >> 5).astype(np.int16)], axis=2)
    map2 = ((iv & 31) * 32 + (iu & 31)).astype(np.uint16)
    return map1, map2


def remap_linear_exact(img: np.ndarray, map1: np.ndarray, map2: np.ndarray) -> np.ndarray:
    """cv2.remap(img, map1, map2, INTER_LINEAR) on uint8 (BORDER_CONSTANT 0): 2^15 fixed-point bilinear weights from the
    5-bit fractions, (sum + 2^14) >> 15, samples outside the image count as 0."""
    H, W = img.shape[:2]
    src = img.reshape(H, W, -1).astype(np.int64)
    sx, sy = map1[..., 0].astype(np.int64), map1[..., 1].astype(np.int64)
    fx, fy = (map2 & 31).astype(np.int64), ((map2 >> 5) & 31).astype(np.int64)
    acc = 0
    for dy, dx, wgt in ((0, 0, (32 - fx) * (32 - fy)), (0, 1, fx * (32 - fy)), (1, 0, (32 - fx) * fy), (1, 1, fx * fy)):
        yy, xx = sy + dy, sx + dx
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        acc = acc + (src[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)] * ok[..., None]) * (wgt * 32)[..., None]
    out = ((acc + (1 << 14)) >> 15).astype(np.uint8)
    return out.reshape(map2.shape + img.shape[2:])


def position_3d_exact(K, R, T, u, v, diameter_px, marker_diameter_mm=2.0):
    """R3:209-228 with NumPy-2 promotion written out: f_avg, f_avg**2 and 2.0/f_avg are float32, the rest float64."""
    K32 = np.asarray(K, dtype=np.float32)
    fx32, fy32 = K32[0, 0], K32[1, 1]
    f_avg32 = np.float32(np.float32(fx32 + fy32) / np.float32(2))
    ratio32 = np.float32(np.float32(marker_diameter_mm) / f_avg32)
    fx, fy, cx, cy = float(fx32), float(fy32), float(K32[0, 2]), float(K32[1, 2])
    f_avg, ratio = float(f_avg32), float(ratio32)
    f_avg_sq = float(np.float32(f_avg32 * f_avg32))          # `f_avg**2` on an np.float32 scalar is a float32 product
    du, dv = u - cx, v - cy
    rad = np.sqrt(du ** 2 + dv ** 2)
    if rad < 1e-6:
        return None
    d_eff = ratio * np.sqrt(rad ** 2 + f_avg_sq)
    h = f_avg * (d_eff / diameter_px)
    pc = np.array([h * du / fx, h * dv / fy, h]) - np.asarray(T, dtype=np.float32).astype(np.float64).reshape(3)
    Rm = np.asarray(R, dtype=np.float32).astype(np.float64)
    pw = np.array([Rm[0, i] * pc[0] + Rm[1, i] * pc[1] + Rm[2, i] * pc[2] for i in range(3)])
    if not np.all(np.isfinite(pw)):
        return None
    return pw


# ----------------------------------------------------------------------------------
# a19: plane fit by centred normal equations (FD:141-159)
# ----------------------------------------------------------------------------------
def plane_tilt_exact(X, Y, Z):
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    Z = np.asarray(Z, dtype=np.float64)
    mx, my, mz = X.mean(), Y.mean(), Z.mean()
    x, y, z = X - mx, Y - my, Z - mz
    sxx, sxy, syy = np.sum(x * x), np.sum(x * y), np.sum(y * y)
    sxz, syz = np.sum(x * z), np.sum(y * z)
    det = sxx * syy - sxy * sxy
    a = (sxz * syy - syz * sxy) / det
    b = (syz * sxx - sxz * sxy) / det
    c = mz - a * mx - b * my
    return a, b, c, np.degrees(np.arctan(np.sqrt(a * a + b * b)))
