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
