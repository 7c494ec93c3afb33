"""CPU port of the reference's per-frame marker pipeline.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product path
(``vision-basedsensor_b200``) never does: it calls the CUDA library and fails
loudly when that is missing.

This is a restatement of the reference algorithm that calls the same third-party
routines (OpenCV / SciPy / NumPy) the reference calls, in the same order and with
the same constants, so it can run on a box where ``/root/reference`` is absent.
It is *pinned* against the unmodified reference executed in the build container:
``oracle/make_golden.py`` imports the real modules from ``/root/reference`` and
checks this port against them output by output before writing ``tests/golden``.

Abbreviations for citations (all paths under /root/reference/code):
  MD = Marker_Tracking/marker_detection.py
  R3 = Marker_Calibration/3d_reconstruction.py
  FD = ForceDistribution/ForceDistribution.py
"""
from __future__ import annotations

import numpy as np
import cv2
from scipy import ndimage
from scipy.signal import fftconvolve
from scipy.spatial.distance import cdist


# ----------------------------------------------------------------------------------
# constants that switch on the frame height (MD:117-126, MD:129, MD:170)
# ----------------------------------------------------------------------------------
def branch_constants(height: int) -> dict:
    """Detection constants of the two resolution branches (MD:117-126,129,170)."""
    if height <= 480:
        return dict(k_small=21, s_small=4.56, k_large=35, s_large=11.4, tmpl=33, tmpl_sigma=7.4,
                    lo=35, hi=180, nbhd=8)
    return dict(k_small=39, s_small=8.0, k_large=101, s_large=20.0, tmpl=80, tmpl_sigma=13.0,
                lo=20, hi=200, nbhd=14)


def crop_box(width: int, height: int, ratios) -> tuple[int, int, int, int]:
    """(left, right, top, bottom) pixel bounds of the crop (MD:81-84)."""
    left = int(width * ratios[0])
    right = width - int(width * ratios[1])
    top = int(height * ratios[2])
    bottom = height - int(height * ratios[3])
    return left, right, top, bottom


def undistort_frame(frame: np.ndarray, camera_matrix, dist_coeffs, taps: dict | None = None) -> np.ndarray:
    """Optional lens correction of the cropped frame (MD:93-109): new camera matrix with alpha = 0, CV_16SC2 maps,
    bilinear remap.  The reference recomputes the maps for every frame; they only depend on (K, D, size)."""
    K = np.array(camera_matrix)
    D = np.array(dist_coeffs)
    h, w = frame.shape[:2]
    new_k, _roi = cv2.getOptimalNewCameraMatrix(K, D, (w, h), 0, (w, h))
    map1, map2 = cv2.initUndistortRectifyMap(K, D, None, new_k, (w, h), cv2.CV_16SC2)
    if taps is not None:
        taps.update(new_camera_matrix=new_k, map1=map1, map2=map2)
    return cv2.remap(frame, map1, map2, cv2.INTER_LINEAR)


def to_gray(frame: np.ndarray) -> np.ndarray:
    """BGR -> gray as MD:114; a 2-D frame is already gray (gray-replicated BGR maps to itself)."""
    if frame.ndim == 2:
        return frame
    return cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)


def gaussian_template(l: int, sig: float) -> np.ndarray:
    """l x l normalised Gaussian on linspace(-(l-1)/2, (l-1)/2, l) (MD:138-143)."""
    ax = np.linspace(-(l - 1) / 2.0, (l - 1) / 2.0, l)
    gx, gy = np.meshgrid(ax, ax)
    k = np.exp(-0.5 * (np.square(gx) + np.square(gy)) / np.square(sig))
    return k / np.sum(k)


def ncc_same(template: np.ndarray, image: np.ndarray) -> np.ndarray:
    """Normalised cross-correlation, 'same' size, three FFT convolutions (MD:146-164)."""
    t = template - np.mean(template)
    im = image - np.mean(image)
    flipped = np.flipud(np.fliplr(t))
    num = fftconvolve(im, flipped.conj(), mode="same")
    ones = np.ones(t.shape)
    energy = fftconvolve(np.square(im), ones, mode="same")
    energy -= np.square(fftconvolve(im, ones, mode="same")) / np.prod(t.shape)
    energy[energy < 0] = 0
    with np.errstate(divide="ignore", invalid="ignore"):
        out = num / np.sqrt(energy * np.sum(np.square(t)))
    out[np.logical_not(np.isfinite(out))] = 0
    return out


def detect_masks(frame: np.ndarray, taps: dict | None = None):
    """(mask, area_mask) of one frame (MD:111-135).  ``taps`` collects intermediates."""
    gray = to_gray(frame)
    c = branch_constants(gray.shape[0])
    blur_small = cv2.GaussianBlur(gray, (c["k_small"], c["k_small"]), c["s_small"])
    blur_large = cv2.GaussianBlur(gray, (c["k_large"], c["k_large"]), c["s_large"])
    dog = blur_large - blur_small + 15          # uint8, wraps (MD:128)
    area_mask = cv2.inRange(dog, c["lo"], c["hi"])
    ncc = ncc_same(gaussian_template(c["tmpl"], c["tmpl_sigma"]), area_mask)
    mask = (ncc > 0.1).astype("uint8")
    if taps is not None:
        taps.update(gray=gray, blur_small=blur_small, blur_large=blur_large, dog=dog, ncc=ncc)
    return mask, area_mask


def ring_maxima(mask: np.ndarray) -> np.ndarray:
    """Boolean 'maxima' image: mask pixels with a zero inside the filter window (MD:170-174)."""
    size = 8 if mask.shape[0] <= 480 else 14
    mx = ndimage.maximum_filter(mask, size)
    maxima = mask == mx
    maxima[((mx - ndimage.minimum_filter(mask, size)) > 0) == 0] = 0
    return maxima


def label_centres(mask: np.ndarray, taps: dict | None = None):
    """(labeled int32 image, n, centres[n,2] as (row, col) float64) (MD:176-185)."""
    maxima = ring_maxima(mask)
    labeled, n = ndimage.label(maxima)
    if taps is not None:
        taps.update(maxima=maxima, labeled=labeled, n_labels=n)
    if n == 0:
        return labeled, 0, np.zeros((0, 2))
    centres = np.array(ndimage.center_of_mass(mask, labeled, range(1, n + 1)))
    if centres.ndim == 1 and n == 1:
        centres = centres.reshape(1, -1)
    return labeled, n, centres


def opened_contours(area_mask: np.ndarray, taps: dict | None = None):
    """5x5 open then external simple contours (MD:188-196)."""
    if np.max(area_mask) > 1:
        a8 = area_mask.astype(np.uint8)
    else:
        a8 = (area_mask * 255).astype(np.uint8)
    opened = cv2.morphologyEx(a8, cv2.MORPH_OPEN, np.ones((5, 5), np.uint8))
    contours, _ = cv2.findContours(opened, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    if taps is not None:
        taps.update(opened=opened, contours=contours)
    return contours


def ellipse_axes(contour: np.ndarray):
    """(cx, cy, major, minor, angle) with the reference's axis ordering (MD:208-217)."""
    (cx, cy), (w, h), angle = cv2.fitEllipse(contour)
    if w > h:
        return cx, cy, w, h, angle
    return cx, cy, h, w, angle + 90


def marker_center(mask: np.ndarray, area_mask: np.ndarray, taps: dict | None = None) -> list[dict]:
    """Marker list of one frame (MD:166-249), without the drawing side effect."""
    _, n, centres = label_centres(mask, taps)
    if n == 0 or centres.size == 0:
        return []
    contours = opened_contours(area_mask, taps)
    pending = [(i, (c[1], c[0])) for i, c in enumerate(centres)]   # (x, y) = (col, row), MD:199
    out = []
    ellipses = []
    for contour in contours:
        if len(contour) < 5:
            ellipses.append(None)
            continue
        cx, cy, major, minor, ang = ellipse_axes(contour)
        ellipses.append((cx, cy, major, minor, ang))
        if minor < 5:
            continue
        gate = (minor / 10) ** 2
        best, best_d = -1, float("inf")
        for j, (_, (x, y)) in enumerate(pending):
            if cv2.pointPolygonTest(contour, (x, y), False) < 0:
                continue
            d = (x - cx) ** 2 + (y - cy) ** 2
            if d < gate and d < best_d:
                best, best_d = j, d
        if best != -1:
            _, (x, y) = pending.pop(best)
            out.append({"center": (x, y), "major_axis": float(major), "minor_axis": float(minor),
                        "angle": float(ang)})
    if taps is not None:
        taps.update(centres=centres, ellipses=ellipses)
    return out


def find_markers_frame(frame: np.ndarray, taps: dict | None = None) -> list[dict]:
    """Frame -> marker list: `_find_markers` then `_marker_center` (MD:441-442)."""
    mask, area_mask = detect_masks(frame, taps)
    if taps is not None:
        taps.update(mask=mask, area_mask=area_mask)
    return marker_center(mask, area_mask, taps)


# ----------------------------------------------------------------------------------
# ID assignment by nearest neighbour to the reference-state array (MD:349-396)
# ----------------------------------------------------------------------------------
def track_rows(ref_keys, ref_xy, markers: list[dict], frameno: int, min_dist: float = 20) -> list[dict]:
    """Tracking rows of one frame (MD:349-396).

    ref_keys: list of (row, col) keys in reference-dict order; ref_xy: [R,2] (Ox, Oy).
    """
    if len(ref_keys) == 0 or not markers:
        return []
    by_pos = {tuple(m["center"]): m for m in markers}
    cur = np.array([m["center"] for m in markers])
    rows = []
    for (layer, angle), (ox, oy) in zip(ref_keys, ref_xy):
        d = cdist([np.array([ox, oy])], cur)[0]
        j = np.argmin(d)
        if d[j] > min_dist:
            continue
        m = by_pos.get(tuple(cur[j]))
        if m:
            rows.append({"frameno": frameno, "row": layer, "col": angle, "Ox": ox, "Oy": oy,
                         "Cx": m["center"][0], "Cy": m["center"][1], "major_axis": m["major_axis"],
                         "minor_axis": m["minor_axis"], "angle": m["angle"]})
    return rows


# ----------------------------------------------------------------------------------
# 3D reconstruction (R3:185-316)
# ----------------------------------------------------------------------------------
class Camera:
    """K (3x3), D (5), R (3x3), T (3x1), all float32 like R3:87-124 builds them."""

    def __init__(self, K, D, R, T):
        self.matrix = np.asarray(K, dtype=np.float32)
        self.dist_coeffs = np.asarray(D, dtype=np.float32)
        self.R_world_to_cam = np.asarray(R, dtype=np.float32)
        self.T_world_to_cam = np.asarray(T, dtype=np.float32).reshape(3, 1)


def undistort_points(cam: Camera, pts: np.ndarray) -> np.ndarray:
    """cv2.undistortPoints with P = K (R3:185-193)."""
    pts = np.asarray(pts, dtype=np.float64)
    if pts.size == 0:
        return pts.reshape(0, 2)
    return cv2.undistortPoints(pts.reshape(-1, 1, 2), cam.matrix, cam.dist_coeffs, None, cam.matrix).reshape(-1, 2)


def position_3d(cam: Camera, u: float, v: float, diameter_px: float, marker_diameter_mm: float = 2.0):
    """World position of one marker or None when the reference would raise (R3:195-238).

    The float32 scalars ``fx, fy, cx, cy`` keep NumPy-2 promotion exactly as in the reference:
    ``f_avg`` and ``marker_diameter_mm / f_avg`` round to float32.
    """
    fx, fy = cam.matrix[0, 0], cam.matrix[1, 1]
    cx, cy = cam.matrix[0, 2], cam.matrix[1, 2]
    f_avg = (fx + fy) / 2
    with np.errstate(all="ignore"):
        rad = np.sqrt((u - cx) ** 2 + (v - cy) ** 2)
        if rad < 1e-6:
            return None
        d_eff = (marker_diameter_mm / f_avg) * np.sqrt(rad ** 2 + f_avg ** 2)
        h = f_avg * (d_eff / diameter_px)
        p_cam = np.array([h * (u - cx) / fx, h * (v - cy) / fy, h]).reshape(3, 1)
        p_world = (cam.R_world_to_cam.T @ (p_cam - cam.T_world_to_cam)).flatten()
    if not np.all(np.isfinite(p_world)):
        return None
    return p_world


def displacement_rows(cam: Camera, table: dict, warmup_frames: int = 100, marker_diameter_mm: float = 2.0,
                      min_marker_size_px: float = 5.0, max_displacement: float = 50.0) -> list[dict]:
    """3D rows with last-seen displacement (R3:172-176 filter + R3:240-316).

    ``table`` has equal-length arrays frameno,row,col,Cx,Cy,major_axis sorted by frame
    (stable), like the CSV the tracker writes.
    """
    fr = np.asarray(table["frameno"])
    keep = np.asarray(table["major_axis"], dtype=np.float64) >= min_marker_size_px
    order = np.argsort(fr[keep], kind="stable")
    sel = np.flatnonzero(keep)[order]
    fr = fr[sel]
    row = np.asarray(table["row"])[sel]
    col = np.asarray(table["col"])[sel]
    uv = np.stack([np.asarray(table["Cx"], dtype=np.float64)[sel], np.asarray(table["Cy"], dtype=np.float64)[sel]], axis=1)
    diam = np.asarray(table["major_axis"], dtype=np.float64)[sel]
    if len(fr) == 0:
        return []
    if warmup_frames > 0:
        m = fr >= fr.min() + warmup_frames
        fr, row, col, uv, diam = fr[m], row[m], col[m], uv[m], diam[m]
    uv = undistort_points(cam, uv)
    seen: dict = {}
    out = []
    i = 0
    n = len(fr)
    while i < n:
        j = i
        current = {}
        while j < n and fr[j] == fr[i]:
            key = (float(row[j]), float(col[j]))
            current[key] = (uv[j, 0], uv[j, 1], diam[j])
            if key in seen:
                prev = position_3d(cam, *seen[key], marker_diameter_mm)
                cur = position_3d(cam, uv[j, 0], uv[j, 1], diam[j], marker_diameter_mm) if prev is not None else None
                if prev is not None and cur is not None:
                    d = cur - prev
                    dn = np.linalg.norm(d)
                    if not dn > max_displacement:
                        out.append({"frameno": fr[i], "row": float(row[j]), "col": float(col[j]),
                                    "X": cur[0], "Y": cur[1], "Z": cur[2], "dX": d[0], "dY": d[1], "dZ": d[2],
                                    "displacement": dn})
            j += 1
        seen.update(current)
        i = j
    return out


# ----------------------------------------------------------------------------------
# contact-plane tilt (FD:138-162) and deviation arithmetic (FD:196-204, 219-232)
# ----------------------------------------------------------------------------------
def plane_tilt(X, Y, Z):
    """(a, b, c, tilt_deg) of the least-squares plane Z = aX + bY + c (FD:141,144,159)."""
    X = np.asarray(X, dtype=np.float64)
    A = np.vstack([X, Y, np.ones(len(X))]).T
    coeff, _, _, _ = np.linalg.lstsq(A, np.asarray(Z, dtype=np.float64), rcond=None)
    a, b, c = coeff
    return a, b, c, np.degrees(np.arctan(np.sqrt(a ** 2 + b ** 2)))


def deviation_endpoints(ref_xyz, d_tilt, d_vert, shell: bool = False, scale: float = 1.0):
    """End points fed to the plane fit: ref + (d_tilt - d_vert), Z from 0 in 'plane' mode (FD:196-204,219-232)."""
    ref_xyz = np.asarray(ref_xyz, dtype=np.float64)
    dev = np.asarray(d_tilt, dtype=np.float64) - np.asarray(d_vert, dtype=np.float64)
    z0 = ref_xyz[:, 2] if shell else np.zeros_like(ref_xyz[:, 2])
    return ref_xyz[:, 0] + dev[:, 0] * scale, ref_xyz[:, 1] + dev[:, 1] * scale, z0 + dev[:, 2] * scale
