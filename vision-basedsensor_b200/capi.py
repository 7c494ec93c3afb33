"""ctypes binding of the C ABI in ``include/vbs.h`` (``csrc/libvbs_b200.so``).

There is deliberately no fallback: if the shared library is missing or does not load,
importing this module raises ``ImportError`` telling the user to build it.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libvbs_b200.so")

VBS_OK, VBS_ERR_BAD_ARG, VBS_ERR_CAPACITY, VBS_ERR_CUDA, VBS_ERR_STATE, VBS_ERR_INTERNAL = 0, -1, -2, -3, -4, -5
STAGE_AREA_MASK, STAGE_MASK, STAGE_MAXIMA, STAGE_LABELS, STAGE_OPENED, STAGE_RECHECKS, STAGE_ELLIPSES, STAGE_NCONTOURS = range(8)


class VbsConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("device", "height", "width", "channels", "max_batch", "max_markers", "max_refs")]


class VbsOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("n_labels", "centres", "n_markers", "marker_xy", "marker_axes", "row_det", "row_cxy",
                 "row_axes", "pos3d", "pos_flags", "plane", "plane_n")]


# every symbol include/vbs.h declares: (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "vbs_create": (C.c_int, [C.POINTER(_P), C.POINTER(VbsConfig)]),
    "vbs_destroy": (None, [_P]),
    "vbs_last_error": (C.c_char_p, [_P]),
    "vbs_set_stream": (C.c_int, [_P, _P]),
    "vbs_sync": (C.c_int, [_P]),
    "vbs_version": (C.c_char_p, []),
    "vbs_set_reference": (C.c_int, [_P, C.c_int32, _P, _P, _P, _P, C.c_double]),
    "vbs_set_camera": (C.c_int, [_P, _P, _P, _P, _P, C.c_double, C.c_double, C.c_double, C.c_int32]),
    "vbs_set_plane": (C.c_int, [_P, C.c_int32, _P, _P, _P, _P, C.c_int32, C.c_double]),
    "vbs_reset_sequence": (C.c_int, [_P]),
    "vbs_get_last_seen": (C.c_int, [_P, _P]),
    "vbs_set_last_seen": (C.c_int, [_P, _P]),
    "vbs_set_first_frame": (C.c_int, [_P, C.c_int64]),
    "vbs_fix_displacement": (C.c_int, [_P, _P, _P, C.c_int64, _P]),
    "vbs_process_device": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.POINTER(VbsOutputs)]),
    "vbs_process_host": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.POINTER(VbsOutputs)]),
    "vbs_set_host_chunk": (C.c_int, [_P, C.c_int32]),
    "vbs_set_overlap": (C.c_int, [_P, C.c_int32]),
    "vbs_submit_host": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.POINTER(VbsOutputs)]),
    "vbs_wait_host": (C.c_int, [_P]),
    "vbs_set_undistort": (C.c_int, [_P, _P, _P, C.c_int32]),
    "vbs_get_undistort_maps": (C.c_int, [_P, _P, _P, _P]),
    "vbs_undistort_frames": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_int64, _P]),
    "vbs_find_markers": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_int64]),
    "vbs_marker_center": (C.c_int, [_P, _P, _P, C.c_int32, C.POINTER(VbsOutputs)]),
    "vbs_ncc_mask": (C.c_int, [_P, _P, C.c_int32]),
    "vbs_track_markers": (C.c_int, [_P, C.c_int32, _P, _P, _P, _P, _P]),
    "vbs_reconstruct_rows": (C.c_int, [_P, C.c_int32, C.c_int64, _P, _P, _P, _P, _P, _P, _P]),
    "vbs_undistort_points": (C.c_int, [_P, C.c_int32, _P, _P]),
    "vbs_position_3d": (C.c_int, [_P, C.c_int32, _P, _P, _P]),
    "vbs_fit_plane": (C.c_int, [_P, C.c_int32, _P, _P, _P, _P]),
    "vbs_debug_stage": (C.c_int, [_P, C.c_int32, _P, C.c_size_t]),
    "vbs_kernel_launches": (C.c_int64, [_P]),
    "vbs_tma_launches": (C.c_int64, [_P]),
    "vbs_set_blur_tc": (C.c_int, [_P, C.c_int32]),
    "vbs_tc_launches": (C.c_int64, [_P]),
    "vbs_set_profiling": (C.c_int, [_P, C.c_int32]),
    "vbs_get_stage_ms": (C.c_int, [_P, _P, _P]),
}


def load_library(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  vbs_b200 has no CPU fallback."
        )
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load_library()


class VbsError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"vbs error {code}: {msg}")
        self.code = code


def check(ctx, code: int):
    """Map a C status to the exception type the reference raises for the same condition."""
    if code == VBS_OK:
        return
    msg = lib.vbs_last_error(ctx).decode() if ctx else "no context"
    if code == VBS_ERR_BAD_ARG:
        raise ValueError(msg)              # MD:38, R3:95,117
    raise VbsError(code, msg)
