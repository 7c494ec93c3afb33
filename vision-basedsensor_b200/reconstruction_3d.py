"""Drop-in mirror of code/Marker_Calibration/3d_reconstruction.py ("R3"): ``Config``,
``CameraParameters`` and ``MarkerAnalysis`` with ``_undistort_points`` (R3:185),
``_calculate_3d_position`` (R3:195) and ``_track_markers`` (R3:240) - arithmetic on the GPU.

The reference module cannot be imported as shipped (mutable dataclass default at R3:28-32, log
file opened at import, R3:42); this mirror fixes only that.  The data formats either side of the
path are kept: ``load_parameters`` (Parameter/Value tables, R3:70-130; .xlsx needs openpyxl, .csv
always works), ``load_marker_data`` (the tracking CSV, R3:132-183) and ``run_analysis``
(R3:405-442), which writes ``marker_3d_coordinates.csv`` (and ``.xlsx`` when openpyxl is
installed).  The plots of R3:318-403 are not part of the path.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, Tuple

import logging

import numpy as np

from . import pipeline as _pl

logger = logging.getLogger(__name__)


def _read_parameter_table(path):
    """Parameter / Value table -> Series (R3:82-83, 106-107); .csv is read directly, anything else by read_excel."""
    import pandas as pd
    path = Path(path)
    df = pd.read_csv(path) if path.suffix.lower() == ".csv" else pd.read_excel(path)
    return df.set_index("Parameter")["Value"]


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
    def __init__(self, config: Config, max_frames_per_call: int = 1024):
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
        # One table-only context (the C library allocates per-pixel scratch lazily, so this one never does), kept
        # while it is large enough: alternating per-point calls (n_refs = 1) and _track_markers (n_refs = R)
        # only re-sets the reference array; it is rebuilt when a larger R arrives.
        if self._pipe is None or self._pipe.Rcap < n_refs:
            if self._pipe is not None:
                self._pipe.close()
            self._pipe = _pl.MarkerPipeline(8, 8, 1, max_batch=self._max_frames, max_markers=1, max_refs=max(n_refs, 64))
            self._refs_set = 0
        if self._refs_set != n_refs:
            z = np.zeros(n_refs)
            self._pipe.set_reference(np.zeros(n_refs, np.int32), np.arange(n_refs, dtype=np.int32), z, z, 20.0)
            self._refs_set = n_refs
        self._pipe.set_camera(K, c.dist_coeffs, c.R_world_to_cam, np.asarray(c.T_world_to_cam).reshape(3), self.config.marker_diameter_mm,
                              self.config.min_marker_size_px, self.config.max_displacement_px, self.config.warmup_frames)
        return self._pipe

    # -- file formats either side of the path ---------------------------------------------------------
    def _validate_paths(self) -> None:
        for path in (self.config.data_dir, self.config.output_dir, self.config.plots_dir):     # R3:65-68
            Path(path).mkdir(parents=True, exist_ok=True)

    def load_parameters(self, intrinsic_path, extrinsic_path) -> None:
        """Camera matrix, 5 distortion coefficients, R and T as float32, validated like R3:70-130."""
        p = _read_parameter_table(intrinsic_path)
        K = np.array([[p["fx"], p.get("skew", 0), p["cx"]], [0, p["fy"], p["cy"]], [0, 0, 1]], dtype=np.float32)
        if K[0, 0] <= 0 or K[1, 1] <= 0:
            raise ValueError("Focal lengths must be positive")                                     # R3:94-95
        D = np.array([p.get(k, 0) for k in ("k1", "k2", "p1", "p2", "k3")], dtype=np.float32)    # R3:98-102
        e = _read_parameter_table(extrinsic_path)
        R = np.array([[e[f"R_wc_{i}{j}"] for j in range(1, 4)] for i in range(1, 4)], dtype=np.float32)
        if not np.allclose(R @ R.T, np.eye(3), atol=1e-6):
            raise ValueError("Rotation matrix is not orthogonal")                                  # R3:115-117
        T = np.array([e["Tx_wc"], e["Ty_wc"], e["Tz_wc"]], dtype=np.float32).reshape(3, 1)
        self.camera.matrix, self.camera.dist_coeffs, self.camera.R_world_to_cam, self.camera.T_world_to_cam = K, D, R, T

    def load_marker_data(self, filepath):
        """Tracking CSV (MD:380-391 columns) -> DataFrame with u, v, major_axis, too-small markers
        dropped, sorted by frame (R3:132-183).  chardet is optional: without it the file is read as utf-8."""
        import pandas as pd
        filepath = Path(filepath)
        if not filepath.exists():
            raise FileNotFoundError(f"Marker data file not found: {filepath}")
        encoding = "utf-8"
        try:
            import chardet
            with open(filepath, "rb") as f:
                encoding = chardet.detect(f.read(30000))["encoding"] or "utf-8"
        except ImportError:
            pass
        df = pd.read_csv(filepath, sep=r"\s+|,|\t", encoding=encoding, engine="python", skipinitialspace=True)
        df.columns = [c.strip() for c in df.columns]
        missing = (set(self.config.column_mapping.keys()) | {"frameno", "row", "col"}) - set(df.columns)
        if missing:
            raise ValueError(f"Missing required columns: {missing}")
        df = df.rename(columns=self.config.column_mapping)
        valid = df["major_axis"] >= self.config.min_marker_size_px
        if not valid.all():
            logger.warning(f"Filtered {len(df) - valid.sum()} markers for being too small")
            df = df[valid].copy()
        return df.sort_values("frameno").reset_index(drop=True)

    def run_analysis(self, input_csv, intrinsic_path=None, extrinsic_path=None):
        """CSV in -> 3D table out (R3:405-442).  Camera files default to the reference's locations;
        pass them, or set ``self.camera`` beforehand and leave both None to skip loading."""
        self._validate_paths()
        if intrinsic_path is None and extrinsic_path is None and self.camera.matrix is None:
            intrinsic_path = Path(self.config.data_dir) / "PreprocessPara" / "IntrinsicParameters.xlsx"
            extrinsic_path = Path(self.config.data_dir) / "PreprocessPara" / "ExtrinsicParameters.xlsx"
        if intrinsic_path is not None:
            self.load_parameters(intrinsic_path, extrinsic_path)
        marker_df = self.load_marker_data(input_csv)
        logger.info(f"Loaded {len(marker_df)} marker detections")
        results_df = self._track_markers(marker_df)
        if results_df.empty:
            raise ValueError("No valid 3D positions calculated")                                   # R3:422-423
        out = Path(self.config.output_dir) / "marker_3d_coordinates.csv"
        results_df.to_csv(out, index=False)
        try:
            import openpyxl  # noqa: F401
            results_df.to_excel(out.with_suffix(".xlsx"), index=False)                             # R3:428-429
        except ImportError:
            logger.info("openpyxl not installed: wrote %s only", out)
        return results_df

    # -- arithmetic -----------------------------------------------------------------------------------
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
