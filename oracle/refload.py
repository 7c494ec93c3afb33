"""Loaders for the UNMODIFIED reference modules.  TEST INFRASTRUCTURE ONLY, build container only.

``/root/reference`` exists only in the build container; nothing that runs on the GPU
box may import this file.  It is used by ``oracle/make_golden.py`` and by the
``not gpu`` tests that pin ``oracle/port.py`` against the real reference (those tests
skip when the directory is absent).

Why loaders are needed:
  * Marker_Tracking/marker_detection.py imports cleanly; only
    ``cv2.destroyAllWindows`` (MD:474) has to be stubbed on headless OpenCV.
  * Marker_Calibration/3d_reconstruction.py cannot be imported as shipped:
    matplotlib/chardet are missing, the dataclass at R3:28-32 has a mutable default,
    and R3:42 opens a log file in ./Results/data/results at import time.  The loader
    stubs the two modules, wraps that one dict literal in ``field(default_factory=..)``
    and executes the source from a temp cwd; every method body is the reference's own.
  * ForceDistribution.py imports matplotlib at the top; ``fit_plane_least_squares`` is
    extracted by AST and executed with a dummy ``ax``/``Patch`` (it only prints the
    tilt, with 2 decimals, so we capture the coefficients through ``np.linalg.lstsq``).
"""
from __future__ import annotations

import ast
import contextlib
import importlib.util
import io
import os
import re
import sys
import tempfile
import types

REF_ROOT = os.environ.get("VBS_REFERENCE_ROOT", "/root/reference")
MD_PATH = os.path.join(REF_ROOT, "code/Marker_Tracking/marker_detection.py")
R3_PATH = os.path.join(REF_ROOT, "code/Marker_Calibration/3d_reconstruction.py")
FD_PATH = os.path.join(REF_ROOT, "code/ForceDistribution/ForceDistribution.py")


def available() -> bool:
    return os.path.exists(MD_PATH)


_cache: dict = {}


def marker_detection():
    """The reference ``marker_detection`` module (MarkerTracker class)."""
    if "md" not in _cache:
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spec = importlib.util.spec_from_file_location("ref_marker_detection", MD_PATH)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
        _cache["md"] = mod
    return _cache["md"]


def reconstruction_3d():
    """The reference ``3d_reconstruction`` module, loaded through the minimal source patch."""
    if "r3" in _cache:
        return _cache["r3"]
    for name in ("matplotlib", "matplotlib.pyplot", "chardet"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    src = open(R3_PATH, encoding="utf-8").read()
    patched, n = re.subn(
        r"(column_mapping: Dict\[str, str\] = )(\{.*?\})",
        lambda m: m.group(1) + "__import__('dataclasses').field(default_factory=lambda: " + m.group(2) + ")",
        src, count=1, flags=re.S,
    )
    assert n == 1, "reference layout changed: column_mapping literal not found"
    tmp = tempfile.mkdtemp(prefix="vbs_r3_")
    os.makedirs(os.path.join(tmp, "Results/data/results"), exist_ok=True)
    mod = types.ModuleType("ref_reconstruction_3d")
    mod.__file__ = R3_PATH
    sys.modules[mod.__name__] = mod          # dataclasses looks the module up by name
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        import logging

        root = logging.getLogger()
        saved = (list(root.handlers), root.level)
        with contextlib.redirect_stderr(io.StringIO()):
            exec(compile(patched, R3_PATH, "exec"), mod.__dict__)
        for h in list(root.handlers):
            if h not in saved[0]:
                root.removeHandler(h)
                with contextlib.suppress(Exception):
                    h.close()
        root.setLevel(saved[1])
        mod.logger.disabled = True
    finally:
        os.chdir(cwd)
    mod._tmpdir = tmp
    _cache["r3"] = mod
    return mod


def make_analysis(K, D, R, T, warmup_frames=100):
    """A reference ``MarkerAnalysis`` with camera parameters set as R3:87-124 would set them."""
    import numpy as np
    from pathlib import Path

    r3 = reconstruction_3d()
    base = Path(r3._tmpdir) / "Results" / "data"
    cfg = r3.Config(warmup_frames=warmup_frames, data_dir=base, output_dir=base / "results",
                    plots_dir=base / "results" / "plots")
    an = r3.MarkerAnalysis(cfg)
    an.camera.matrix = np.asarray(K, dtype=np.float32)
    an.camera.dist_coeffs = np.asarray(D, dtype=np.float32)
    an.camera.R_world_to_cam = np.asarray(R, dtype=np.float32)
    an.camera.T_world_to_cam = np.asarray(T, dtype=np.float32).reshape(3, 1)
    return an


def fit_plane_reference(X, Y, Z):
    """Run the reference's own ``fit_plane_least_squares`` (FD:138-162); return (a, b, c, printed_text).

    The function returns only a legend patch and prints the tilt; the coefficients are
    captured by wrapping ``np.linalg.lstsq`` in the namespace the function executes in.
    """
    import numpy as np

    tree = ast.parse(open(FD_PATH, encoding="utf-8").read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "fit_plane_least_squares")
    captured = {}

    class _LinalgTap:
        LinAlgError = np.linalg.LinAlgError

        @staticmethod
        def lstsq(*a, **k):
            r = np.linalg.lstsq(*a, **k)
            captured["coeff"] = r[0]
            return r

    class _NpTap:
        linalg = _LinalgTap

        def __getattr__(self, name):
            return getattr(np, name)

    class _Ax:
        def plot_surface(self, *a, **k):
            pass

    ns = {"np": _NpTap(), "Patch": lambda **k: None}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), FD_PATH, "exec"), ns)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ns["fit_plane_least_squares"](_Ax(), np.asarray(X), np.asarray(Y), np.asarray(Z))
    a, b, c = captured["coeff"]
    return a, b, c, buf.getvalue()


def bare_tracker(ref_markers: dict, min_dist=20, frame_count=0):
    """A reference ``MarkerTracker`` without a video, to exercise ``_track_markers`` (MD:349-396)."""
    md = marker_detection()
    tr = object.__new__(md.MarkerTracker)
    tr.config = {"min_marker_distance": min_dist}
    tr.first_frame_markers = ref_markers
    tr.frame_count = frame_count
    tr._draw_tracking = lambda *a, **k: None
    return tr
