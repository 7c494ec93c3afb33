"""Synthetic marker-array frames and sequences (workload generation, host side).

The reference ships no data (its demo videos are absent from the checkout), so
every parity case and every bench line runs on frames rendered here.  The
recipes follow SURVEY.md section 8(d): dark anti-aliased disks (fg 40) on a
bright gel (bg 170), Gaussian PSF sigma 1, additive N(0, 2^2) noise, per-marker
jitter, ``np.random.default_rng(seed=frame_index)``.

Layouts:
  * ``grid_layout``  - rows x cols grid, centred, isotropic pitch (configs 2, 4)
  * ``ring_layout``  - the 65-marker concentric layout the sensor uses
                       (reference table: code/ForceDistribution/ForceDistribution.py:29-95)
"""
from __future__ import annotations

import numpy as np

BG = 170.0
FG = 40.0

# Ring radii (mm), marker counts and Z (mm) of the 65-marker bonnet layout.  Values
# re-derived from the table at ForceDistribution.py:29-95 (1/6/12/18/24/4 markers).
RING_SPEC = (
    (0.00, 1, 0.00),
    (3.49, 6, 0.23),
    (6.92, 12, 0.90),
    (10.23, 18, 2.01),
    (13.37, 24, 3.55),
    (16.29, 4, 5.47),
)


def ring_layout_mm() -> np.ndarray:
    """(65, 3) array of X, Y, Z in mm in the id order of ForceDistribution.py:29-95."""
    pts = [(0.0, 0.0, 0.0)]
    # ring 1: starts at 150 deg going clockwise in 60 deg steps
    for k in range(6):
        a = np.deg2rad(150.0 - 60.0 * k)
        pts.append((3.49 * np.cos(a), 3.49 * np.sin(a), 0.23))
    for k in range(12):
        a = np.deg2rad(120.0 - 30.0 * k)
        pts.append((6.92 * np.cos(a), 6.92 * np.sin(a), 0.90))
    for k in range(18):
        a = np.deg2rad(130.0 - 20.0 * k)
        pts.append((10.23 * np.cos(a), 10.23 * np.sin(a), 2.01))
    for k in range(24):
        a = np.deg2rad(135.0 - 15.0 * k)
        pts.append((13.37 * np.cos(a), 13.37 * np.sin(a), 3.55))
    for k in range(4):
        a = np.deg2rad(90.0 - 90.0 * k)
        pts.append((16.29 * np.cos(a), 16.29 * np.sin(a), 5.47))
    return np.asarray(pts, dtype=np.float64)


def ring_layout(height: int, width: int, px_per_mm: float = 11.0, dy: float = 15.0) -> np.ndarray:
    """(65, 2) pixel centres (x, y) of the ring layout, centred, shifted ``dy`` px down."""
    mm = ring_layout_mm()
    x = width / 2.0 + mm[:, 0] * px_per_mm
    y = height / 2.0 + dy - mm[:, 1] * px_per_mm
    return np.stack([x, y], axis=1)


def grid_layout(height: int, width: int, rows: int, cols: int, pitch: float) -> np.ndarray:
    """(rows*cols, 2) pixel centres (x, y) of a centred grid in ascending raster order."""
    x0 = (width - 1) / 2.0 - (cols - 1) / 2.0 * pitch
    y0 = (height - 1) / 2.0 - (rows - 1) / 2.0 * pitch
    ys, xs = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    return np.stack([x0 + xs.ravel() * pitch, y0 + ys.ravel() * pitch], axis=1)


def _psf_kernel(sigma: float) -> np.ndarray:
    r = int(np.ceil(4 * sigma))
    ax = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (ax / sigma) ** 2)
    return k / k.sum()


def render_frame(
    height: int,
    width: int,
    centres_xy: np.ndarray,
    radii: np.ndarray | float,
    seed: int,
    jitter: float = 1.5,
    noise_sigma: float = 2.0,
    psf_sigma: float = 1.0,
) -> np.ndarray:
    """Render one uint8 grayscale frame.  Deterministic in ``seed``."""
    rng = np.random.default_rng(seed)
    centres = np.asarray(centres_xy, dtype=np.float64)
    n = centres.shape[0]
    radii = np.broadcast_to(np.asarray(radii, dtype=np.float64), (n,))
    jit = rng.uniform(-jitter, jitter, size=(n, 2)) if jitter > 0 else np.zeros((n, 2))
    cov = np.zeros((height, width), dtype=np.float64)
    for (cx, cy), r in zip(centres + jit, radii):
        x_lo = max(int(np.floor(cx - r - 2)), 0)
        x_hi = min(int(np.ceil(cx + r + 2)) + 1, width)
        y_lo = max(int(np.floor(cy - r - 2)), 0)
        y_hi = min(int(np.ceil(cy + r + 2)) + 1, height)
        if x_lo >= x_hi or y_lo >= y_hi:
            continue
        yy, xx = np.mgrid[y_lo:y_hi, x_lo:x_hi]
        d = np.sqrt((xx - cx) ** 2 + (yy - cy) ** 2)
        patch = np.clip(r + 0.5 - d, 0.0, 1.0)
        np.maximum(cov[y_lo:y_hi, x_lo:x_hi], patch, out=cov[y_lo:y_hi, x_lo:x_hi])
    img = BG + (FG - BG) * cov
    if psf_sigma > 0:
        k = _psf_kernel(psf_sigma)
        pad = len(k) // 2
        tmp = np.pad(img, ((0, 0), (pad, pad)), mode="edge")
        img = sum(k[i] * tmp[:, i : i + width] for i in range(len(k)))
        tmp = np.pad(img, ((pad, pad), (0, 0)), mode="edge")
        img = sum(k[i] * tmp[i : i + height, :] for i in range(len(k)))
    if noise_sigma > 0:
        img = img + rng.normal(0.0, noise_sigma, size=img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


# Named workloads (BASELINE.json configs) -----------------------

WORKLOADS = {
    # name: (H, W, rows, cols, pitch, radius)
    "1080p_20x20": (1080, 1920, 20, 20, 52.0, 11.0),
    "4k_40x72": (2160, 3840, 40, 72, 52.6, 10.0),
    "small_6x8": (560, 640, 6, 8, 60.0, 11.0),  # smallest case that takes the >480 branch
    "tiny_4x5": (300, 360, 4, 5, 56.0, 6.0),   # <=480 branch
}


def workload_frames(name: str, n_unique: int, seed0: int = 0, jitter: float = 1.5) -> np.ndarray:
    """[n_unique, H, W] uint8 frames of a named grid workload, seeds seed0..seed0+n-1."""
    h, w, rows, cols, pitch, radius = WORKLOADS[name]
    centres = grid_layout(h, w, rows, cols, pitch)
    return np.stack(
        [render_frame(h, w, centres, radius, seed=seed0 + i, jitter=jitter) for i in range(n_unique)]
    )


def compression_sequence(
    height: int,
    width: int,
    centres_xy: np.ndarray,
    radius: float,
    n_frames: int,
    tilt: float = 0.0,
    depth: float = 1.0,
    seed0: int = 1000,
    noise_sigma: float = 2.0,
) -> np.ndarray:
    """Vertical / tilted compression style sequence (config 3).

    Marker centres move along a smooth radial field growing linearly with time
    (vertical press) plus a component linear in x (tilt); marker radius scales by up
    to +-5 % (drives ``major_axis`` and hence the depth estimate).
    """
    centres = np.asarray(centres_xy, dtype=np.float64)
    c0 = centres.mean(axis=0)
    rel = centres - c0
    rmax = np.max(np.linalg.norm(rel, axis=1)) + 1e-9
    frames = []
    for t in range(n_frames):
        s = depth * t / max(n_frames - 1, 1)
        radial = rel * (0.02 * s) * (1.0 - (np.linalg.norm(rel, axis=1, keepdims=True) / rmax) ** 2 * 0.5)
        lin = np.zeros_like(rel)
        lin[:, 0] = tilt * s * 2.0 * rel[:, 0] / rmax
        scale = 1.0 + 0.05 * s * (1.0 - np.linalg.norm(rel, axis=1) / rmax) + 0.03 * tilt * s * rel[:, 0] / rmax
        frames.append(
            render_frame(
                height, width, centres + radial + lin, radius * scale, seed=seed0 + t,
                jitter=0.0, noise_sigma=noise_sigma,
            )
        )
    return np.stack(frames)


# Synthetic camera of SURVEY.md section 8d (config 3) -- all float32 like R3:87-124 would make them.
def synthetic_camera():
    K = np.array([[1450.3, 0, 962.1], [0, 1448.7, 541.9], [0, 0, 1]], dtype=np.float32)
    D = np.array([-0.31, 0.12, 7e-4, -4e-4, -0.02], dtype=np.float32)
    ax, ay, az = np.deg2rad([3.0, -2.0, 1.0])
    Rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
    Ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
    Rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
    R = (Rz @ Ry @ Rx).astype(np.float32)
    T = np.array([1.5, -0.7, 42.0], dtype=np.float32)
    return K, D, R, T
