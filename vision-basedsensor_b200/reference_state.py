"""Reference-state arrays (the tracker's ``first_frame_markers``, MD:31,289-347).

Host-side, once per sequence.  ``grid_ids`` is the ID layout used for the synthetic grid
workloads (detections of frame 0 in ascending raster order, ids
``(i // cols, i % cols)``).
"""
from __future__ import annotations

import numpy as np


def grid_ids(marker_xy: np.ndarray, cols: int, row_gap: float = 20.0):
    """(keys, xy): detections sorted into raster order; a new grid row starts where the sorted y
    coordinates jump by more than ``row_gap`` px."""
    pts = np.asarray(marker_xy, dtype=np.float64).reshape(-1, 2)
    by_y = pts[np.argsort(pts[:, 1], kind="stable")]
    row_id = np.concatenate([[0], np.cumsum(np.diff(by_y[:, 1]) > row_gap)]) if len(by_y) else np.zeros(0, int)
    order = np.lexsort((by_y[:, 0], row_id))
    pts = by_y[order]
    keys = [(i // cols, i % cols) for i in range(len(pts))]
    return keys, pts


def _kmeans_1d(values: np.ndarray, k: int) -> np.ndarray:
    """Optimal 1-D k-means labels by dynamic programming (deterministic).  The reference calls
    sklearn KMeans(n_clusters, n_init=10) with no random_state (MD:308); for well separated rings
    both give the same partition, and this one needs no RNG."""
    v = np.sort(values)
    n = len(v)
    k = min(k, n)
    ps, ps2 = np.concatenate([[0], np.cumsum(v)]), np.concatenate([[0], np.cumsum(v * v)])

    def cost(i, j):          # sum of squared deviations of v[i:j]
        s, s2, m = ps[j] - ps[i], ps2[j] - ps2[i], j - i
        return s2 - s * s / m

    D = np.full((k + 1, n + 1), np.inf); D[0, 0] = 0.0
    arg = np.zeros((k + 1, n + 1), dtype=int)
    for c in range(1, k + 1):
        for j in range(c, n + 1):
            best, bi = np.inf, c - 1
            for i in range(c - 1, j):
                t = D[c - 1, i] + cost(i, j)
                if t < best:
                    best, bi = t, i
            D[c, j], arg[c, j] = best, bi
    bounds, j = [], n
    for c in range(k, 0, -1):
        i = arg[c, j]; bounds.append((i, j)); j = i
    bounds.reverse()
    edges = [v[i] for i, _ in bounds[1:]]
    return np.searchsorted(np.asarray(edges), values, side="right")       # 0 = innermost cluster


def ring_ids(markers: list, num_layers: int = 5, full: bool = False) -> dict:
    """First-frame identities of the concentric layout (MD:275-347).

    ``full=False`` reproduces the reference literally: every marker of a ring is stored under the
    placeholder key ``(layer, -1)`` (MD:321), so only the LAST one survives and the tracker follows
    at most ``1 + num_layers`` ids.  ``full=True`` is the evidently intended behaviour: every marker
    gets ``(layer, angle_index)``, angle index 0 at the marker nearest 0 rad, increasing with angle.
    """
    centres = np.array([m["center"] for m in markers], dtype=np.float64)
    ci = int(np.argmin(np.linalg.norm(centres - centres.mean(axis=0), axis=1)))
    cm = markers[ci]
    out = {(0, 0): {**cm, "Ox": cm["center"][0], "Oy": cm["center"][1]}}
    rest = [m for i, m in enumerate(markers) if i != ci]
    if not rest:
        return out
    vec = np.array([m["center"] for m in rest], dtype=np.float64) - np.asarray(cm["center"], dtype=np.float64)
    dist = np.linalg.norm(vec, axis=1)
    ang = np.arctan2(vec[:, 1], vec[:, 0])
    layer = _kmeans_1d(dist, num_layers) + 1
    for lay in range(1, num_layers + 1):
        idx = [i for i in range(len(rest)) if layer[i] == lay]
        if not idx:
            continue
        if not full:
            i = idx[-1]                                  # the overwrite at MD:321 keeps the last marker of the ring
            out[(lay, 0)] = {**rest[i], "angle_rad": ang[i], "Ox": rest[i]["center"][0], "Oy": rest[i]["center"][1]}
            continue
        idx.sort(key=lambda i: ang[i])
        start = int(np.argmin([abs(ang[i]) for i in idx]))
        for pos, i in enumerate(idx):
            out[(lay, (pos - start) % len(idx))] = {**rest[i], "angle_rad": ang[i], "Ox": rest[i]["center"][0], "Oy": rest[i]["center"][1]}
    return out
