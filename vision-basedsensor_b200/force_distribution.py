"""Drop-in mirror of the plane-fit part of code/ForceDistribution/ForceDistribution.py ("FD"):
``fit_plane_least_squares`` (FD:138-162) and the deviation arithmetic (FD:196-204, 219-232).

The reference only PRINTS the tilt (FD:160); ``fit_plane_tilt`` adds the missing return path.
Plotting (matplotlib surfaces, quivers) is outside the path: ``ax`` may be None.
"""
from __future__ import annotations

import numpy as np

from . import pipeline as _pl

_ctx = None


def _context():
    global _ctx
    if _ctx is None:
        _ctx = _pl.MarkerPipeline(8, 8, 1, max_batch=1, max_markers=1, max_refs=1)
    return _ctx


def fit_plane_tilt(X, Y, Z):
    """(a, b, c, tilt_deg) of the least-squares plane Z = aX + bY + c, computed on the GPU."""
    return _context().fit_plane(np.asarray(X, dtype=np.float64), np.asarray(Y, dtype=np.float64), np.asarray(Z, dtype=np.float64))


def fit_plane_least_squares(ax, X, Y, Z, color="orange", label="Fitted Plane"):
    """Same call and console output as FD:138-162; returns (legend_patch_or_None, label)."""
    a, b, c, tilt = fit_plane_tilt(X, Y, Z)
    if not np.isfinite(tilt):
        print(f"Warning: Plane fitting failed for {label}")
        return None, None
    patch = None
    if ax is not None:
        X = np.asarray(X); Y = np.asarray(Y)
        XX, YY = np.meshgrid(np.linspace(X.min(), X.max(), 10), np.linspace(Y.min(), Y.max(), 10))
        ax.plot_surface(XX, YY, a * XX + b * YY + c, color=color, alpha=0.3, linewidth=0)
        try:
            from matplotlib.patches import Patch
            patch = Patch(color=color, alpha=0.3)
        except Exception:
            patch = None
    print(f"-> Plane Fit ({label}): Tilt Angle = {tilt:.2f} degrees")
    return patch, label


def deviation_endpoints(ref_xyz, d_tilt, d_vert, shell=False, scale=1.0):
    """X_end, Y_end, Z_end fed to the plane fit (FD:196-204, 219-232); pure indexing / adds on host arrays."""
    ref_xyz = np.asarray(ref_xyz, dtype=np.float64)
    dev = np.asarray(d_tilt, dtype=np.float64) - np.asarray(d_vert, dtype=np.float64)
    z0 = ref_xyz[:, 2] if shell else np.zeros_like(ref_xyz[:, 2])
    return ref_xyz[:, 0] + dev[:, 0] * scale, ref_xyz[:, 1] + dev[:, 1] * scale, z0 + dev[:, 2] * scale
