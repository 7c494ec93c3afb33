"""Drop-in mirror of code/Marker_Calibration/3d_reconstruction.py ("R3"): ``Config``,
``CameraParameters`` and ``MarkerAnalysis`` with ``_undistort_points`` (R3:185),
``_calculate_3d_position`` (R3:195) and ``_track_markers`` (R3:240) - arithmetic on the GPU.

The reference module cannot be imported as shipped (mutable dataclass default at R3:28-32, log
file opened at import, R3:42); this mirror fixes only that.  File loading (XLSX / CSV parsing,
R3:70-183), plots and the XLSX writer (R3:318-442) are host glue outside the path: camera
parameters are set on ``analysis.camera`` directly, as R3:87-124 would leave them.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, Tuple

import numpy as np

from . import pipeline as _pl


@dataclass
class Config:
    marker_diameter_mm: float = 2.0
    warmup_frames: int = 100
    min_marker_size_px: float = 5.0
    max_displacement_px: float = 50.0
    data_dir: Path = Path("Results/data")
    output_dir: Path = Path("Results/data/results")
    plots_dir: Path = Path("Results/data/results/Displacement_Analysis_Plots")
    column_mapping: Dict[str, str] = field(default_factory=lambda: {"Cx": "u", "Cy": "v", "major_axis": "major_axis"})


class CameraParameters:
    def __init__(self):
        self.matrix = None
        self.dist_coeffs = None
        self.R_world_to_cam = None
        self.T_world_to_cam = None
        self.resolution: Tuple[int, int] = None


class MarkerAnalysis:
    def __init__(self, config: Config, max_frames_per_call: int = 4096):
        self.config = config
        self.camera = CameraParameters()
        self._max_frames = int(max_frames_per_call)
        self._pipe = None
        self._pipe_key = None

    def _context(self, n_refs: int = 1):
        c = self.camera
        if c.matrix is None or c.dist_coeffs is None or c.R_world_to_cam is None or c.T_world_to_cam is None:
            raise ValueError("camera parameters not set")
        K = np.asarray(c.matrix, dtype=np.float32)
        if K[0, 0] <= 0 or K[1, 1] <= 0:
            raise ValueError("Focal lengths must be positive")                      # R3:94-95
        key = (K.tobytes(), np.asarray(c.dist_coeffs, np.float32).tobytes(), np.asarray(c.R_world_to_cam, np.float32).tobytes(),
               np.asarray(c.T_world_to_cam, np.float32).tobytes(), self.config.marker_diameter_mm, self.config.min_marker_size_px,
               self.config.max_displacement_px, self.config.warmup_frames, n_refs)
        if self._pipe is None or self._pipe_key != key:
            if self._pipe is not None:
                self._pipe.close()
            self._pipe = _pl.MarkerPipeline(8, 8, 1, max_batch=self._max_frames, max_markers=1, max_refs=max(n_refs, 1))
            self._pipe_key = key
            self._refs_set = 0
        if self._refs_set != n_refs:
            z = np.zeros(n_refs)
            self._pipe.set_reference(np.zeros(n_refs, np.int32), np.arange(n_refs, dtype=np.int32), z, z, 20.0)
            self._refs_set = n_refs
        self._pipe.set_camera(K, c.dist_coeffs, c.R_world_to_cam, np.asarray(c.T_world_to_cam).reshape(3), self.config.marker_diameter_mm,
                              self.config.min_marker_size_px, self.config.max_displacement_px, self.config.warmup_frames)
        return self._pipe

    def _undistort_points(self, points: np.ndarray) -> np.ndarray:
        return self._context().undistort_points(np.asarray(points, dtype=np.float64).reshape(-1, 2))

    def _calculate_3d_position(self, u: float, v: float, diameter_px: float) -> np.ndarray:
        P, ok = self._context().position_3d([[u, v, diameter_px]])
        if not ok[0]:
            raise ValueError("Marker too close to principal point" if np.hypot(u - float(np.float32(self.camera.matrix[0][2])),
                                                                                v - float(np.float32(self.camera.matrix[1][2]))) < 1e-6
                             else "Non-finite coordinates calculated")            # R3:216-217, 231-232
        return P[0]

    def _track_markers(self, df):
        """DataFrame with frameno,row,col,u,v,major_axis (as load_marker_data returns it) ->
        DataFrame frameno,row,col,X,Y,Z,dX,dY,dZ,displacement (R3:296-307), rows in frame order."""
        import pandas as pd
        cols = ["frameno", "row", "col", "X", "Y", "Z", "dX", "dY", "dZ", "displacement"]
        if len(df) == 0:
            return pd.DataFrame([])
        fr = df["frameno"].to_numpy()
        keys = list(dict.fromkeys(zip(df["row"].to_numpy().tolist(), df["col"].to_numpy().tolist())))
        kidx = {k: i for i, k in enumerate(keys)}
        R = len(keys)
        f0, f1 = int(fr.min()), int(fr.max())
        F = f1 - f0 + 1
        det = np.full((F, R), -1, np.int32); cxy = np.zeros((F, R, 2)); axes = np.zeros((F, R, 3))
        order = np.zeros((F, R), np.int64)
        for pos, (f, r, c, u, v, m) in enumerate(zip(fr, df["row"].to_numpy(), df["col"].to_numpy(), df["u"].to_numpy(), df["v"].to_numpy(),
                                                      df["major_axis"].to_numpy())):
            i, j = int(f) - f0, kidx[(r, c)]
            det[i, j] = 0; cxy[i, j] = (u, v); axes[i, j, 0] = m; order[i, j] = pos
        pipe = self._context(R)
        pipe.reset_sequence()
        pipe.set_first_frame(f0)
        rows = []
        for s in range(0, F, self._max_frames):
            e = min(F, s + self._max_frames)
            pos3d, flags, _ = pipe.reconstruct_rows(det[s:e], cxy[s:e], axes[s:e], frameno0=f0 + s)
            for i in range(e - s):
                js = [j for j in np.flatnonzero(flags[i] & 4)]
                js.sort(key=lambda j: order[s + i, j])
                for j in js:
                    rows.append([f0 + s + i, float(keys[j][0]), float(keys[j][1]), *pos3d[i, j]])
        return pd.DataFrame(rows, columns=cols) if rows else pd.DataFrame([])
