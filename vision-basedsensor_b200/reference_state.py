"""Reference-state arrays (the tracker's ``first_frame_markers``, MD:31,289-347).

Host-side, once per sequence.  ``grid_ids`` is the ID layout used for the synthetic grid
workloads (detections of frame 0 in ascending raster order, ids
``(i // cols, i % cols)``).
"""
from __future__ import annotations

import numpy as np


def grid_ids(marker_xy: np.ndarray, cols: int, row_quantum: float = 20.0):
    """(keys, xy): detections sorted into raster order (rows bucketed by ``row_quantum`` px)."""
    pts = np.asarray(marker_xy, dtype=np.float64).reshape(-1, 2)
    order = np.lexsort((pts[:, 0], np.round(pts[:, 1] / row_quantum)))
    pts = pts[order]
    keys = [(i // cols, i % cols) for i in range(len(pts))]
    return keys, pts
