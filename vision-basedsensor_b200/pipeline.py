"""Batched host-side driver of the CUDA marker pipeline (thin: no arithmetic in Python).

``MarkerPipeline`` owns one C context (= one GPU, one stream).  Frames go in as
``[B, H, W]`` / ``[B, H, W, 3]`` uint8, either a CUDA ``torch`` tensor (nothing is
copied) or a host ``numpy`` array (the H2D / D2H copies are part of the call); results
come back as a :class:`BatchResult` of arrays on the same side.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import capi


@dataclass
class BatchResult:
    """SoA outputs of one batch (device torch tensors or host numpy arrays)."""
    frameno0: int
    n_labels: object = None     # [B] int32
    centres: object = None      # [B, M, 2] float64 (row, col)        (None in compact blocks)
    n_markers: object = None    # [B] int32
    marker_xy: object = None    # [B, M, 2] float64 (x, y)            (None in compact blocks)
    marker_axes: object = None  # [B, M, 3] float64 major, minor, angle (None in compact blocks)
    row_det: object = None      # [B, R] int32
    row_cxy: object = None      # [B, R, 2]
    row_axes: object = None     # [B, R, 3]
    pos3d: object = None        # [B, R, 7]
    pos_flags: object = None    # [B, R] uint8
    plane: object = None        # [B, 4]
    plane_n: object = None      # [B]

    def to_host(self) -> "BatchResult":
        def cv(a):
            return a.cpu().numpy() if hasattr(a, "cpu") else a          # None stays None
        return BatchResult(self.frameno0, *[cv(getattr(self, k)) for k in
                                            ("n_labels", "centres", "n_markers", "marker_xy", "marker_axes", "row_det", "row_cxy",
                                             "row_axes", "pos3d", "pos_flags", "plane", "plane_n")])

    def markers(self, f: int) -> list:
        """Marker dicts of frame ``f`` exactly as ``_marker_center`` returns them (MD:238-243)."""
        h = self if isinstance(self.n_markers, np.ndarray) else self.to_host()
        n = int(h.n_markers[f])
        return [{"center": (h.marker_xy[f, k, 0], h.marker_xy[f, k, 1]), "major_axis": float(h.marker_axes[f, k, 0]),
                 "minor_axis": float(h.marker_axes[f, k, 1]), "angle": float(h.marker_axes[f, k, 2])} for k in range(n)]


def _current_device() -> int:
    """The GPU a context is created on when none is named: torch's current device (a rank of a sharded job has
    called ``torch.cuda.set_device(local_rank)``), so helper contexts never pile up on GPU 0."""
    try:
        import torch
        if torch.cuda.is_available():
            return int(torch.cuda.current_device())
    except ImportError:
        pass
    return 0


class MarkerPipeline:
    def __init__(self, height: int, width: int, channels: int = 1, max_batch: int = 32, max_markers: int = 1024,
                 max_refs: int = 1024, device: Optional[int] = None):
        self.H, self.W, self.C = int(height), int(width), int(channels)
        self.B, self.M, self.Rcap = int(max_batch), int(max_markers), int(max_refs)
        self.device = _current_device() if device is None else int(device)
        self.R = 0
        self.have_cam = False
        self.have_plane = False
        self._ctx = C.c_void_p()
        cfg = capi.VbsConfig(self.device, self.H, self.W, self.C, self.B, self.M, self.Rcap)
        code = capi.lib.vbs_create(C.byref(self._ctx), C.byref(cfg))
        if code != capi.VBS_OK:
            ctx, self._ctx = self._ctx, C.c_void_p()
            try:
                capi.check(ctx if ctx else None, code)
            finally:
                if ctx:
                    capi.lib.vbs_destroy(ctx)
        self.ref_rows = self.ref_cols = None

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            capi.lib.vbs_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def sync(self):
        capi.check(self._ctx, capi.lib.vbs_sync(self._ctx))

    def use_stream(self, cuda_stream_ptr: int):
        capi.check(self._ctx, capi.lib.vbs_set_stream(self._ctx, C.c_void_p(cuda_stream_ptr)))

    @property
    def kernel_launches(self) -> int:
        return int(capi.lib.vbs_kernel_launches(self._ctx))

    @property
    def tma_launches(self) -> int:
        return int(capi.lib.vbs_tma_launches(self._ctx))

    @property
    def tc_launches(self) -> int:
        return int(capi.lib.vbs_tc_launches(self._ctx))

    def set_blur_tc(self, on: bool):
        """Opt-in experiment: the two Gaussian blurs as int8 GEMMs on the tensor cores (tcgen05);
        default off = the integer-dot-product kernel.  Same bits out."""
        capi.check(self._ctx, capi.lib.vbs_set_blur_tc(self._ctx, int(bool(on))))

    STAGES = ("blur_dog_area", "ncc_mask", "morphology", "components", "contours_ellipse", "track_3d_plane", "output_copies")

    def set_profiling(self, on: bool):
        capi.check(self._ctx, capi.lib.vbs_set_profiling(self._ctx, int(bool(on))))

    def stage_ms(self):
        """(dict stage -> accumulated ms, batches accumulated) since profiling was switched on."""
        ms = (C.c_double * 7)()
        calls = C.c_int64()
        capi.check(self._ctx, capi.lib.vbs_get_stage_ms(self._ctx, ms, C.byref(calls)))
        return dict(zip(self.STAGES, list(ms))), int(calls.value)

    # -- state ------------------------------------------------------------------------------
    def set_reference(self, rows, cols, ox, oy, min_marker_distance: float = 20.0):
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        ox = np.ascontiguousarray(ox, dtype=np.float64)
        oy = np.ascontiguousarray(oy, dtype=np.float64)
        n = len(rows)
        capi.check(self._ctx, capi.lib.vbs_set_reference(self._ctx, n, rows.ctypes.data, cols.ctypes.data, ox.ctypes.data,
                                                         oy.ctypes.data, float(min_marker_distance)))
        self.R = n
        self.ref_rows, self.ref_cols, self.ref_ox, self.ref_oy = rows, cols, ox, oy
        self.have_plane = False

    def set_camera(self, K, D, R, T, marker_diameter_mm: float = 2.0, min_marker_size_px: float = 5.0,
                   max_displacement: float = 50.0, warmup_frames: int = 100):
        K = np.ascontiguousarray(K, dtype=np.float32).reshape(9)
        D = np.ascontiguousarray(np.asarray(D, dtype=np.float32).reshape(-1)[:5])
        R = np.ascontiguousarray(R, dtype=np.float32).reshape(9)
        T = np.ascontiguousarray(T, dtype=np.float32).reshape(3)
        capi.check(self._ctx, capi.lib.vbs_set_camera(self._ctx, K.ctypes.data, D.ctypes.data, R.ctypes.data, T.ctypes.data,
                                                      float(marker_diameter_mm), float(min_marker_size_px),
                                                      float(max_displacement), int(warmup_frames)))
        self.have_cam = True

    def set_plane(self, ref_xyz, start_xyz, d_vert=None, use=None, shell: bool = False, scale: float = 1.0):
        ref_xyz = np.ascontiguousarray(ref_xyz, dtype=np.float64)
        start_xyz = np.ascontiguousarray(start_xyz, dtype=np.float64)
        dv = None if d_vert is None else np.ascontiguousarray(d_vert, dtype=np.float64)
        us = None if use is None else np.ascontiguousarray(use, dtype=np.uint8)
        capi.check(self._ctx, capi.lib.vbs_set_plane(self._ctx, len(ref_xyz), ref_xyz.ctypes.data, start_xyz.ctypes.data,
                                                     dv.ctypes.data if dv is not None else None,
                                                     us.ctypes.data if us is not None else None, int(bool(shell)), float(scale)))
        self.have_plane = True

    def set_undistort(self, camera_matrix=None, dist_coeffs=None):
        """Lens correction of every frame before detection, like ``_undistort_frame`` (MD:93-109);
        ``set_undistort(None)`` switches it off.  K 3x3 and 4/5/8 distortion coefficients, float64."""
        if camera_matrix is None:
            capi.check(self._ctx, capi.lib.vbs_set_undistort(self._ctx, None, None, 0))
            return
        K = np.ascontiguousarray(np.array(camera_matrix), dtype=np.float64)          # MD:96-97
        D = np.ascontiguousarray(np.array(dist_coeffs), dtype=np.float64).ravel()
        if K.shape != (3, 3):
            raise ValueError("camera_matrix must be 3x3")
        self._follow_torch_stream()
        capi.check(self._ctx, capi.lib.vbs_set_undistort(self._ctx, K.ctypes.data, D.ctypes.data, len(D)))

    def undistort_maps(self):
        """(new_camera_matrix [3,3] float64 host, map1 int16 [H,W,2], map2 uint16 [H,W] device) in OpenCV's CV_16SC2 layout."""
        import torch
        dev = torch.device("cuda", self.device)
        nk = np.zeros((3, 3))
        m1 = torch.empty((self.H, self.W, 2), dtype=torch.int16, device=dev)
        m2 = torch.empty((self.H, self.W), dtype=torch.uint16, device=dev)
        self._follow_torch_stream()
        capi.check(self._ctx, capi.lib.vbs_get_undistort_maps(self._ctx, nk.ctypes.data, m1.data_ptr(), m2.data_ptr()))
        self.sync()
        return nk, m1, m2

    def undistort_frames(self, frames):
        """Device frames [B,H,W(,3)] uint8 -> corrected frames, like ``_preprocess_frame`` after the crop (MD:88-91)."""
        import torch
        batch = self._geometry(frames)
        fr = frames.contiguous()
        out = torch.empty_like(fr)
        self._follow_torch_stream()
        rowb = self.W * self.C
        capi.check(self._ctx, capi.lib.vbs_undistort_frames(self._ctx, fr.data_ptr(), batch, rowb * self.H, rowb, out.data_ptr()))
        return out

    def set_overlap(self, on: bool):
        """Two-stream chunk pipelining of device-resident batches inside process() (off by default: measured slower on
        B200; the branch-level overlap of the open-mask kernels with the NCC is always on unless VBS_BRANCH_OVERLAP=0)."""
        capi.check(self._ctx, capi.lib.vbs_set_overlap(self._ctx, int(bool(on))))

    def set_host_chunk(self, frames_per_chunk: int):
        """Frames per chunk of the copy/compute overlap in the host entry point (0 = default 64)."""
        capi.check(self._ctx, capi.lib.vbs_set_host_chunk(self._ctx, int(frames_per_chunk)))

    def reset_sequence(self):
        capi.check(self._ctx, capi.lib.vbs_reset_sequence(self._ctx))

    def get_last_seen(self) -> np.ndarray:
        t = np.zeros((self.R, 4), dtype=np.float64)
        capi.check(self._ctx, capi.lib.vbs_get_last_seen(self._ctx, t.ctypes.data))
        return t

    def set_last_seen(self, table):
        t = np.ascontiguousarray(table, dtype=np.float64)
        assert t.shape == (self.R, 4)
        capi.check(self._ctx, capi.lib.vbs_set_last_seen(self._ctx, t.ctypes.data))

    def set_first_frame(self, frameno: int):
        capi.check(self._ctx, capi.lib.vbs_set_first_frame(self._ctx, int(frameno)))

    # -- output allocation --------------------------------------------------------------------
    def _alloc(self, batch: int, on_device: bool, compact: bool = False, pinned: bool = True):
        B, M, R = batch, self.M, self.R
        shapes = {"n_labels": ((B,), "int32"), "n_markers": ((B,), "int32")}
        if not compact:          # the padded per-frame marker lists: [B][M] arrays, 56 B per slot
            shapes.update({"centres": ((B, M, 2), "float64"), "marker_xy": ((B, M, 2), "float64"), "marker_axes": ((B, M, 3), "float64")})
        if R > 0:
            shapes.update({"row_det": ((B, R), "int32"), "row_cxy": ((B, R, 2), "float64"), "row_axes": ((B, R, 3), "float64")})
            if self.have_cam:
                shapes.update({"pos3d": ((B, R, 7), "float64"), "pos_flags": ((B, R), "uint8")})
                if self.have_plane:
                    shapes.update({"plane": ((B, 4), "float64"), "plane_n": ((B,), "int32")})
        arrays, out = {}, capi.VbsOutputs()
        if on_device:
            import torch
            dev = torch.device("cuda", self.device)
            for k, (shp, dt) in shapes.items():
                arrays[k] = torch.empty(shp, dtype=getattr(torch, dt), device=dev)
                setattr(out, k, arrays[k].data_ptr())
        else:
            # pinned host memory: D2H copies into pageable memory would block the host per chunk and
            # defeat the copy/compute overlap of vbs_process_host.  The numpy view keeps its tensor alive.
            import torch
            for k, (shp, dt) in shapes.items():
                t = torch.empty(shp, dtype=getattr(torch, dt), pin_memory=pinned and torch.cuda.is_available())
                arrays[k] = t.numpy()
                setattr(out, k, arrays[k].ctypes.data)
        return arrays, out

    # -- hot path -------------------------------------------------------------------------------
    def _geometry(self, frames):
        shp = tuple(frames.shape)
        want = (self.H, self.W) if self.C == 1 else (self.H, self.W, self.C)
        if len(shp) == len(want):
            shp = (1,) + shp
        if shp[1:] != want:
            raise ValueError(f"frames must be [B,{','.join(map(str, want))}] uint8, got {tuple(frames.shape)}")
        if shp[0] > self.B:
            raise ValueError(f"batch {shp[0]} exceeds max_batch {self.B}")
        return shp[0]

    def process(self, frames, frameno0: int = 0, out: Optional[tuple] = None) -> BatchResult:
        """Run the whole path on one batch.  CUDA tensor in -> device results (asynchronous, call
        :meth:`sync` before trusting them); numpy in -> host results (synchronous)."""
        batch = self._geometry(frames)
        rowb = self.W * self.C
        if isinstance(frames, np.ndarray):
            if frames.dtype != np.uint8:
                raise ValueError("frames must be uint8")
            fr = np.ascontiguousarray(frames)
            arrays, o = out if out is not None else self._alloc(batch, False)
            capi.check(self._ctx, capi.lib.vbs_process_host(self._ctx, fr.ctypes.data, batch, rowb * self.H, rowb, int(frameno0),
                                                            C.byref(o)))
        else:
            import torch
            if frames.dtype != torch.uint8 or not frames.is_cuda:
                raise ValueError("frames must be a CUDA uint8 tensor (or a host numpy array)")
            fr = frames.contiguous()
            self._follow_torch_stream()
            arrays, o = out if out is not None else self._alloc(batch, True)
            self._keep = fr
            capi.check(self._ctx, capi.lib.vbs_process_device(self._ctx, fr.data_ptr(), batch, rowb * self.H, rowb, int(frameno0),
                                                              C.byref(o)))
        return BatchResult(frameno0=int(frameno0), **arrays)

    def _follow_torch_stream(self):
        """Launch on torch's current stream so later torch ops on the results are stream-ordered."""
        import torch
        ptr = torch.cuda.current_stream(self.device).cuda_stream
        if ptr == 0:
            # torch's default stream is the LEGACY default stream; 0 means "own stream" at the C ABI, so
            # name it explicitly (cudaStreamLegacy).  Without this the kernels would run on the context's
            # non-blocking stream and could read frames whose .cuda() copy is still in flight.
            ptr = 1
        if getattr(self, "_stream_ptr", None) != ptr:
            self.use_stream(ptr)
            self._stream_ptr = ptr

    def alloc_outputs(self, batch: int, on_device: bool, compact: bool = False):
        """Pre-allocate an output block to reuse across calls (pass as ``out=``).  ``compact=True`` leaves out
        the padded marker lists (``centres``, ``marker_xy``, ``marker_axes``: 56 B x max_markers per frame) and keeps
        the per-frame records the reference writes to its tables: tracking rows, 3D rows, plane, counts - the
        algorithmic 96 B per reference entry + 32 B per frame (MD:380-391, R3:296-307, FD:141-159)."""
        return self._alloc(batch, on_device, compact)

    def process_host_ptr(self, ptr: int, batch: int, frame_stride: int, row_pitch: int, frameno0: int, out):
        """Host frames by raw pointer (pinned staging buffers, crop views): MD:85 crop is a pointer + pitch."""
        capi.check(self._ctx, capi.lib.vbs_process_host(self._ctx, C.c_void_p(ptr), batch, frame_stride, row_pitch, int(frameno0),
                                                        C.byref(out[1])))
        return BatchResult(frameno0=int(frameno0), **out[0])

    def submit_host_ptr(self, ptr: int, batch: int, frame_stride: int, row_pitch: int, frameno0: int, out):
        """Asynchronous host path: enqueue one batch (pinned frames by pointer, pinned outputs from
        ``alloc_outputs(batch, False)``) and return at once; at most two batches in flight."""
        capi.check(self._ctx, capi.lib.vbs_submit_host(self._ctx, C.c_void_p(ptr), batch, frame_stride, row_pitch, int(frameno0),
                                                       C.byref(out[1])))
        return BatchResult(frameno0=int(frameno0), **out[0])

    def wait_host(self):
        """Block until the oldest submitted batch has landed in its output arrays."""
        capi.check(self._ctx, capi.lib.vbs_wait_host(self._ctx))

    def ncc_mask(self, area_mask):
        """Device area masks [B,H,W] uint8 -> (mask uint8 {0,1} device tensor, float64 re-decisions per frame):
        ``(normxcorr2(gkern, area_mask) > 0.1)`` alone (MD:132-133, 146-164)."""
        import torch
        batch = area_mask.shape[0]
        if tuple(area_mask.shape[1:]) != (self.H, self.W) or batch > self.B:
            raise ValueError(f"area_mask must be [B<={self.B},{self.H},{self.W}], got {tuple(area_mask.shape)}")
        a = ((area_mask != 0).to(torch.uint8) * 255).contiguous()
        self._follow_torch_stream()
        capi.check(self._ctx, capi.lib.vbs_ncc_mask(self._ctx, a.data_ptr(), batch))
        return self.debug_stage(capi.STAGE_MASK, batch), self.debug_stage(capi.STAGE_RECHECKS, batch)

    def find_markers(self, frames):
        """Device frames -> (mask, area_mask) uint8 device tensors, like ``_find_markers`` (MD:111-135)."""
        import torch
        batch = self._geometry(frames)
        fr = frames.contiguous()
        self._follow_torch_stream()
        rowb = self.W * self.C
        capi.check(self._ctx, capi.lib.vbs_find_markers(self._ctx, fr.data_ptr(), batch, rowb * self.H, rowb))
        return self.debug_stage(capi.STAGE_MASK, batch), self.debug_stage(capi.STAGE_AREA_MASK, batch)

    def marker_center(self, mask, area_mask) -> BatchResult:
        """Device masks [B,H,W] uint8 -> marker lists, like ``_marker_center`` (MD:166-249)."""
        import torch
        batch = mask.shape[0]
        m = (mask != 0).to(torch.uint8).contiguous()
        a = (area_mask != 0).to(torch.uint8).contiguous()
        self._follow_torch_stream()
        arrays, o = self._alloc(batch, True)
        capi.check(self._ctx, capi.lib.vbs_marker_center(self._ctx, m.data_ptr(), a.data_ptr(), batch, C.byref(o)))
        keep = {k: arrays[k] for k in ("n_labels", "centres", "n_markers", "marker_xy", "marker_axes")}
        return BatchResult(frameno0=0, **keep)

    # -- table-level calls (host arrays, synchronous) -------------------------------------------
    def track_markers(self, markers: list):
        """One marker list -> (row_det [R], row_cxy [R,2], row_axes [R,3]) like MD:349-396."""
        n = len(markers)
        xy = np.ascontiguousarray([m["center"] for m in markers], dtype=np.float64).reshape(n, 2)
        ax = np.ascontiguousarray([[m["major_axis"], m["minor_axis"], m["angle"]] for m in markers], dtype=np.float64).reshape(n, 3)
        det = np.empty(self.R, np.int32); cxy = np.empty((self.R, 2)); axes = np.empty((self.R, 3))
        capi.check(self._ctx, capi.lib.vbs_track_markers(self._ctx, n, xy.ctypes.data, ax.ctypes.data, det.ctypes.data,
                                                         cxy.ctypes.data, axes.ctypes.data))
        return det, cxy, axes

    def reconstruct_rows(self, row_det, row_cxy, row_axes, frameno0: int = 0):
        """Dense tracking rows [B,R] -> (pos3d [B,R,7], flags [B,R], plane [B,4] | None), R3:240-316."""
        det = np.ascontiguousarray(row_det, dtype=np.int32)
        B = det.shape[0]
        cxy = np.ascontiguousarray(row_cxy, dtype=np.float64); axes = np.ascontiguousarray(row_axes, dtype=np.float64)
        pos = np.empty((B, self.R, 7)); fl = np.empty((B, self.R), np.uint8)
        plane = np.empty((B, 4)) if self.have_plane else None
        pn = np.empty(B, np.int32) if self.have_plane else None
        capi.check(self._ctx, capi.lib.vbs_reconstruct_rows(self._ctx, B, int(frameno0), det.ctypes.data, cxy.ctypes.data, axes.ctypes.data,
                                                            pos.ctypes.data, fl.ctypes.data, plane.ctypes.data if plane is not None else None,
                                                            pn.ctypes.data if pn is not None else None))
        return pos, fl, plane

    def undistort_points(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 2)
        out = np.empty_like(pts)
        capi.check(self._ctx, capi.lib.vbs_undistort_points(self._ctx, len(pts), pts.ctypes.data, out.ctypes.data))
        return out

    def position_3d(self, uvd):
        uvd = np.ascontiguousarray(uvd, dtype=np.float64).reshape(-1, 3)
        P = np.empty_like(uvd); ok = np.empty(len(uvd), np.uint8)
        capi.check(self._ctx, capi.lib.vbs_position_3d(self._ctx, len(uvd), uvd.ctypes.data, P.ctypes.data, ok.ctypes.data))
        return P, ok.astype(bool)

    def fit_plane(self, X, Y, Z):
        X = np.ascontiguousarray(X, dtype=np.float64); Y = np.ascontiguousarray(Y, dtype=np.float64); Z = np.ascontiguousarray(Z, dtype=np.float64)
        out = np.empty(4)
        capi.check(self._ctx, capi.lib.vbs_fit_plane(self._ctx, len(X), X.ctypes.data, Y.ctypes.data, Z.ctypes.data, out.ctypes.data))
        return tuple(out)

    def debug_stage(self, stage: int, batch: int):
        import torch
        dev = torch.device("cuda", self.device)
        if stage in (capi.STAGE_RECHECKS, capi.STAGE_NCONTOURS):
            t = torch.empty((batch,), dtype=torch.int32, device=dev)
        elif stage == capi.STAGE_ELLIPSES:
            t = torch.empty((batch, self.M, 6), dtype=torch.float64, device=dev)
        elif stage == capi.STAGE_LABELS:
            t = torch.empty((batch, self.H, self.W), dtype=torch.int32, device=dev)
        else:
            t = torch.empty((batch, self.H, self.W), dtype=torch.uint8, device=dev)
        capi.check(self._ctx, capi.lib.vbs_debug_stage(self._ctx, stage, t.data_ptr(), t.numel() * t.element_size()))
        self.sync()
        return t
