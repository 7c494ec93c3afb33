"""Drop-in mirror of the reference's ``marker_detection`` module (code/Marker_Tracking/
marker_detection.py, "MD"): same class, method names, arguments, return types, CSV columns and
exception types - every per-frame computation runs in the CUDA library (no CPU fallback).

    from vbs_b200.marker_detection import MarkerTracker, find_marker, marker_center

``find_marker`` / ``marker_center`` are the module-level names code/Marker_Tracking/tracking.py:7
imports (they do not exist in the reference, so that script never ran).

Optional frame undistortion (MD:93-109, ``config['calibration_params']``) runs on the GPU too, both in
``_undistort_frame`` and inside the batched ``process()``.  Out of scope here (host glue
either side of the path): drawing on frames (MD:251-273,398-427) and the XVID video writer (MD:69-76,453).
``frame=`` arguments are accepted and ignored.
"""
from __future__ import annotations

import os

import numpy as np

from . import pipeline as _pl
from . import reference_state as _rs

_pipes: dict = {}
_undistort_pipes: dict = {}


def _pipe_for(h: int, w: int, c: int, max_markers: int = 4096) -> "_pl.MarkerPipeline":
    """One small context per frame geometry for the static-method mirrors (created on first use)."""
    key = (h, w, c, _pl._current_device())
    if key not in _pipes:
        _pipes[key] = _pl.MarkerPipeline(h, w, c, max_batch=1, max_markers=max_markers, max_refs=1)
    return _pipes[key]


class MarkerTracker:
    """Marker tracking for video analysis (MD:12).  ``config`` keys as MD:15-31; additionally
    ``ids``: 'reference' (default; reproduces MD:316-347, which keeps ONE marker per ring),
    'full' (every marker gets (layer, angle_index)) or 'grid' (raster ids, needs ``grid_cols``);
    ``batch``: frames per GPU call (default 64)."""

    def __init__(self, config):
        self.config = config
        self._validate_config()
        self._setup_paths()
        self.frame_count = 0
        self.first_frame_markers = {}
        self._track_pipe = None

    # -- MD:33-48 --------------------------------------------------------------------------------
    def _validate_config(self):
        for key in ("video_path", "output_dir", "crop_ratios"):
            if key not in self.config:
                raise ValueError(f"Missing required config key: {key}")
        if not os.path.exists(self.config["video_path"]):
            raise FileNotFoundError(f"Video file not found: {self.config['video_path']}")

    def _setup_paths(self):
        os.makedirs(self.config["output_dir"], exist_ok=True)
        name = os.path.splitext(os.path.basename(self.config["video_path"]))[0]
        self.output_csv = os.path.join(self.config["output_dir"], f"{name}_markers.csv")
        self.output_video = os.path.join(self.config["output_dir"], f"{name}_tracked.avi")

    # -- MD:78-91: the crop is a view; at the C boundary it is a pointer + pitch --------------------
    def _crop_box(self):
        r = self.config["crop_ratios"]
        left = int(self.width * r[0]); right = self.width - int(self.width * r[1])
        top = int(self.height * r[2]); bottom = self.height - int(self.height * r[3])
        return left, right, top, bottom

    def _preprocess_frame(self, frame):
        left, right, top, bottom = self._crop_box()
        cropped = frame[top:bottom, left:right]
        if "calibration_params" in self.config:                  # MD:88-89
            cropped = self._undistort_frame(cropped)
        return cropped

    # -- MD:93-109: lens correction (new camera matrix with alpha = 0, CV_16SC2 maps, bilinear remap) ---
    def _undistort_frame(self, frame):
        import torch
        frame = np.ascontiguousarray(frame)
        h, w = frame.shape[:2]
        c = 1 if frame.ndim == 2 else frame.shape[2]
        cal = self.config["calibration_params"]
        K = np.array(cal["camera_matrix"], dtype=np.float64); D = np.array(cal["dist_coeffs"], dtype=np.float64).ravel()
        key = (h, w, c, K.tobytes(), D.tobytes(), _pl._current_device())
        if key not in _undistort_pipes:                          # the maps are built once per (K, D, size)
            pipe = _pl.MarkerPipeline(h, w, c, max_batch=1, max_markers=1, max_refs=1)
            pipe.set_undistort(K, D)
            _undistort_pipes[key] = pipe
        pipe = _undistort_pipes[key]
        out = pipe.undistort_frames(torch.from_numpy(frame[None]).cuda(pipe.device))
        pipe.sync()
        return out[0].cpu().numpy()

    # -- MD:111-135 --------------------------------------------------------------------------------
    @staticmethod
    def _find_markers(frame):
        """(mask uint8 {0,1}, area_mask uint8 {0,255}) of one BGR or gray frame."""
        import torch
        frame = np.ascontiguousarray(frame)
        h, w = frame.shape[:2]
        c = 1 if frame.ndim == 2 else frame.shape[2]
        pipe = _pipe_for(h, w, c)
        mask, area = pipe.find_markers(torch.from_numpy(frame[None]).cuda(pipe.device))
        return mask[0].cpu().numpy(), area[0].cpu().numpy()

    @staticmethod
    def _gkern(l=5, sig=1.0):
        """Gaussian template (MD:138-143); a host-side constant, kept for API compatibility."""
        ax = np.linspace(-(l - 1) / 2.0, (l - 1) / 2.0, l)
        xx, yy = np.meshgrid(ax, ax)
        k = np.exp(-0.5 * (np.square(xx) + np.square(yy)) / np.square(sig))
        return k / np.sum(k)

    # -- MD:166-249 --------------------------------------------------------------------------------
    @staticmethod
    def _marker_center(mask, area_mask, frame=None):
        """List of {'center': (x, y), 'major_axis', 'minor_axis', 'angle'} in the reference's order."""
        import torch
        mask = np.ascontiguousarray(mask); area_mask = np.ascontiguousarray(area_mask)
        h, w = mask.shape
        pipe = _pipe_for(h, w, 1)
        res = pipe.marker_center(torch.from_numpy(mask[None].astype(np.uint8)).cuda(pipe.device),
                                 torch.from_numpy(area_mask[None].astype(np.uint8)).cuda(pipe.device))
        pipe.sync()
        return res.markers(0)

    # -- MD:275-347 (host side, once per video) ------------------------------------------------------
    def _process_first_frame(self, markers):
        if not markers:
            raise ValueError("No markers detected in first frame!")
        mode = self.config.get("ids", "reference")
        if mode == "grid":
            keys, xy = _rs.grid_ids(np.array([m["center"] for m in markers]), int(self.config["grid_cols"]))
            by_pos = {tuple(np.asarray(m["center"], dtype=np.float64)): m for m in markers}
            self.first_frame_markers = {k: {**by_pos[tuple(p)], "Ox": p[0], "Oy": p[1]} for k, p in zip(keys, xy)}
        else:
            self.first_frame_markers = _rs.ring_ids(markers, self.config.get("num_layers", 5), full=(mode == "full"))
        self._track_pipe = None

    # -- MD:349-396 --------------------------------------------------------------------------------
    def _ensure_track_pipe(self, n_markers):
        if self._track_pipe is None:
            keys = list(self.first_frame_markers)
            self._track_pipe = _pl.MarkerPipeline(8, 8, 1, max_batch=1, max_markers=max(4096, n_markers), max_refs=max(len(keys), 1))
            self._track_pipe.set_reference([k[0] for k in keys], [k[1] for k in keys],
                                           [self.first_frame_markers[k]["Ox"] for k in keys],
                                           [self.first_frame_markers[k]["Oy"] for k in keys],
                                           self.config.get("min_marker_distance", 20))
        return self._track_pipe

    def _track_markers(self, frame, markers):
        if not self.first_frame_markers or not markers:
            return []
        det, cxy, axes = self._ensure_track_pipe(len(markers)).track_markers(markers)
        rows = []
        for r, ((layer, angle), ref) in enumerate(self.first_frame_markers.items()):
            if det[r] < 0:
                continue
            rows.append({"frameno": self.frame_count, "row": layer, "col": angle, "Ox": ref["Ox"], "Oy": ref["Oy"],
                         "Cx": cxy[r, 0], "Cy": cxy[r, 1], "major_axis": float(axes[r, 0]), "minor_axis": float(axes[r, 1]),
                         "angle": float(axes[r, 2])})
        return rows

    # -- MD:429-474 --------------------------------------------------------------------------------
    def process(self):
        """Decode the video in chunks, run every chunk through the batched CUDA path, write the CSV.
        Chunk i+1 is decoded while chunk i is on the GPU (``vbs_submit_host`` / ``vbs_wait_host``, two pinned
        staging buffers); only the per-frame tracking rows come back (compact output block)."""
        import cv2
        import torch
        cap = cv2.VideoCapture(self.config["video_path"])
        if not cap.isOpened():
            raise IOError(f"Could not open video: {self.config['video_path']}")
        self.cap = cap
        self.fps = cap.get(cv2.CAP_PROP_FPS)
        self.width = int(cap.get(cv2.CAP_PROP_FRAME_WIDTH))
        self.height = int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
        left, right, top, bottom = self._crop_box()
        self.crop_width, self.crop_height = right - left, bottom - top
        B = int(self.config.get("batch", 64))
        pipe = None
        data = []
        stagings = [torch.empty((B, self.height, self.width, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
        fs, rp = self.height * self.width * 3, self.width * 3
        keys, outs = [], None

        def emit(n, frame0, res):
            for f in range(n):
                for r, k in enumerate(keys):
                    if res.row_det[f, r] >= 0:
                        ref = self.first_frame_markers[k]
                        data.append({"frameno": frame0 + f, "row": k[0], "col": k[1], "Ox": ref["Ox"], "Oy": ref["Oy"],
                                     "Cx": res.row_cxy[f, r, 0], "Cy": res.row_cxy[f, r, 1], "major_axis": float(res.row_axes[f, r, 0]),
                                     "minor_axis": float(res.row_axes[f, r, 1]), "angle": float(res.row_axes[f, r, 2])})

        pending = None                   # (frames, first frame number, result block) of the chunk on the GPU
        chunk = 0
        while True:
            slot = chunk & 1
            stage_np = stagings[slot].numpy()
            n = 0
            while n < B:
                ret, frame = cap.read()
                if not ret:
                    break
                stage_np[n] = frame
                n += 1
            if n == 0:
                break
            if pipe is None:                 # first chunk: establish identities from frame 0 (MD:445-446)
                first = MarkerTracker._marker_center(*MarkerTracker._find_markers(self._preprocess_frame(stage_np[0])))
                self._process_first_frame(first)
                keys = list(self.first_frame_markers)
                pipe = _pl.MarkerPipeline(self.crop_height, self.crop_width, 3, max_batch=B, max_markers=4096, max_refs=len(keys))
                pipe.set_reference([k[0] for k in keys], [k[1] for k in keys], [self.first_frame_markers[k]["Ox"] for k in keys],
                                   [self.first_frame_markers[k]["Oy"] for k in keys], self.config.get("min_marker_distance", 20))
                if "calibration_params" in self.config:          # MD:88-89, inside the batched path
                    cal = self.config["calibration_params"]
                    pipe.set_undistort(cal["camera_matrix"], cal["dist_coeffs"])
                outs = [pipe.alloc_outputs(B, False, compact=True) for _ in range(2)]
            res = pipe.submit_host_ptr(stagings[slot].data_ptr() + top * rp + left * 3, n, fs, rp, self.frame_count, outs[slot])
            if pending is not None:          # the previous chunk finishes while this one was being decoded
                pipe.wait_host()
                emit(*pending)
            pending = (n, self.frame_count, res)
            before = self.frame_count
            self.frame_count += n
            chunk += 1
            if self.frame_count // 100 != before // 100:
                print(f"Processed frame {self.frame_count // 100 * 100}")
        if pending is not None:
            pipe.wait_host()
            emit(*pending)
        self._save_results(data)
        self._cleanup()
        if pipe is not None:
            pipe.close()
        return data

    def _save_results(self, data):
        import pandas as pd
        pd.DataFrame(data).to_csv(self.output_csv, index=False)
        print(f"Saved tracking data to {self.output_csv}")

    def _cleanup(self):
        self.cap.release()          # the reference also calls cv2.destroyAllWindows(), which raises on headless OpenCV (MD:474)


find_marker = MarkerTracker._find_markers        # names expected by tracking.py:7,99-100
marker_center = MarkerTracker._marker_center
