"""GPU suite, part 2: the Python drop-in mirrors (same names / arguments / outputs as the
reference modules) against golden outputs of the unmodified reference and the oracle port."""
import io
import os
import contextlib
import warnings

import numpy as np
import pytest

import parity_util as pu
import vbs_b200  # noqa: F401
from vbs_b200 import marker_detection, reconstruction_3d, force_distribution, synth
from oracle import port

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_tracking_script_names_exist():
    """tracking.py:7 does `from marker_detection import find_marker, marker_center`."""
    assert callable(marker_detection.find_marker) and callable(marker_detection.marker_center)


def test_static_methods_match_oracle():
    fr = synth.workload_frames("small_6x8", 1, seed0=55)[0]
    bgr = np.repeat(fr[..., None], 3, axis=2)
    taps = {}
    want = port.find_markers_frame(bgr, taps)
    mask, area = marker_detection.MarkerTracker._find_markers(bgr)
    assert mask.dtype == np.uint8 and np.array_equal(mask, taps["mask"]) and np.array_equal(area, taps["area_mask"])
    got = marker_detection.MarkerTracker._marker_center(mask, area, bgr.copy())
    assert got == want
    assert np.array_equal(marker_detection.MarkerTracker._gkern(80, 13), port.gaussian_template(80, 13))


def test_process_reproduces_reference_csv(tmp_path):
    """MarkerTracker(config).process() on the FFV1 ring video == the CSV the unmodified reference wrote."""
    import pandas as pd
    cfg = {"video_path": os.path.join(GOLDEN, "ring_video.avi"), "output_dir": str(tmp_path), "crop_ratios": (1 / 8, 1 / 8, 1 / 16, 0),
           "num_layers": 5, "min_marker_distance": 20, "batch": 2}
    tr = marker_detection.MarkerTracker(cfg)
    with contextlib.redirect_stdout(io.StringIO()):
        tr.process()
    # pandas' default float parser is not correctly rounded (1-ulp errors): ask for the exact one
    got = pd.read_csv(tr.output_csv, float_precision="round_trip")
    want = pd.read_csv(os.path.join(GOLDEN, "ring_video_markers.csv"), float_precision="round_trip")
    assert open(tr.output_csv).read().split("\n")[0] == open(os.path.join(GOLDEN, "ring_video_markers.csv")).read().split("\n")[0]
    assert list(got.columns) == list(want.columns) == ["frameno", "row", "col", "Ox", "Oy", "Cx", "Cy", "major_axis", "minor_axis", "angle"]
    assert len(got) == len(want)
    for c in ("frameno", "row", "col", "Ox", "Oy", "Cx", "Cy"):
        assert np.array_equal(got[c].to_numpy(), want[c].to_numpy()), c                  # ids, order, centroids: bit-exact
    for c in ("major_axis", "minor_axis"):
        assert pu.f32_ulps(got[c].to_numpy(), want[c].to_numpy()).max() <= 2.0
    da = np.abs(got["angle"].to_numpy() - want["angle"].to_numpy()) % 180.0
    assert np.minimum(da, 180 - da).max() <= 1e-3
    # the intended behaviour (all 65 markers) behind a switch
    cfg["ids"] = "full"
    tr = marker_detection.MarkerTracker(cfg)
    with contextlib.redirect_stdout(io.StringIO()):
        rows = tr.process()
    assert len(tr.first_frame_markers) == 65 and len(rows) == 65 * 5


def _compare_csv(got_path, want_path):
    import pandas as pd
    got = pd.read_csv(got_path, float_precision="round_trip")
    want = pd.read_csv(want_path, float_precision="round_trip")
    assert list(got.columns) == list(want.columns) and len(got) == len(want)
    for c in ("frameno", "row", "col", "Ox", "Oy", "Cx", "Cy"):
        assert np.array_equal(got[c].to_numpy(), want[c].to_numpy()), c
    for c in ("major_axis", "minor_axis"):
        assert pu.f32_ulps(got[c].to_numpy(), want[c].to_numpy()).max() <= 2.0
    da = np.abs(got["angle"].to_numpy() - want["angle"].to_numpy()) % 180.0
    assert np.minimum(da, 180 - da).max() <= 1e-3


def test_process_with_calibration_params_reproduces_reference_csv(tmp_path):
    """config['calibration_params'] switches the lens correction of MD:88-109 on: corrected frame and CSV equal what the
    unmodified reference produced (tests/golden/ring_undistort.npz, ring_video_undistorted_markers.csv)."""
    import cv2
    g = np.load(os.path.join(GOLDEN, "ring_undistort.npz"))
    cfg = {"video_path": os.path.join(GOLDEN, "ring_video.avi"), "output_dir": str(tmp_path), "crop_ratios": (1 / 8, 1 / 8, 1 / 16, 0),
           "num_layers": 5, "min_marker_distance": 20, "batch": 3,
           "calibration_params": {"camera_matrix": g["K"].tolist(), "dist_coeffs": g["D"].tolist()}}
    tr = marker_detection.MarkerTracker(cfg)
    with contextlib.redirect_stdout(io.StringIO()):
        tr.process()
    _compare_csv(tr.output_csv, os.path.join(GOLDEN, "ring_video_undistorted_markers.csv"))
    cap = cv2.VideoCapture(cfg["video_path"])
    ok, fr = cap.read()
    cap.release()
    assert ok and np.array_equal(tr._preprocess_frame(fr), g["first_frame"])


LENSES = [  # (h, w, channels, K, D)
    (450, 480, 1, [[363.3, 0, 243.1], [0, 358.3, 222.7], [0, 0, 1]], [-0.31, 0.12, 7e-4, -4e-4, -0.02]),
    (480, 640, 3, [[483.3, 0, 323.1], [0, 478.3, 237.7], [0, 0, 1]], [-0.2, 0.05, 1e-3, 2e-3]),
    (480, 640, 1, [[483.3, 0, 323.1], [0, 478.3, 237.7], [0, 0, 1]], [0.12, -0.05, 1e-3, -2e-3, 0.01, 0.2, -0.03, 0.004]),
    (1080, 1920, 1, [[1450.3, 0, 962.1], [0, 1448.7, 541.9], [0, 0, 1]], [-0.31, 0.12, 7e-4, -4e-4, -0.02]),
]


@pytest.mark.parametrize("h,w,c,K,D", LENSES)
def test_undistort_maps_and_frames_equal_opencv(h, w, c, K, D):
    """New camera matrix, CV_16SC2 maps and remapped frames of the CUDA path == OpenCV's, bit for bit (MD:93-109)."""
    import cv2
    import torch
    from vbs_b200 import pipeline
    Kn, Dn = np.array(K), np.array(D)
    new_k, _ = cv2.getOptimalNewCameraMatrix(Kn, Dn, (w, h), 0, (w, h))
    m1, m2 = cv2.initUndistortRectifyMap(Kn, Dn, None, new_k, (w, h), cv2.CV_16SC2)
    rng = np.random.default_rng(h + c)
    frames = rng.integers(0, 256, (2, h, w) if c == 1 else (2, h, w, 3), dtype=np.uint8)
    with pipeline.MarkerPipeline(h, w, c, max_batch=2, max_markers=64, max_refs=1) as pipe:
        with pytest.raises(pipeline.capi.VbsError):
            pipe.undistort_maps()                                            # not configured yet
        with pytest.raises(ValueError):
            pipe.set_undistort(K, [0.1, 0.2])                                # 2 coefficients: not a cv2 model we carry
        pipe.set_undistort(K, D)
        nk, g1, g2 = pipe.undistort_maps()
        assert np.array_equal(nk, new_k)
        assert np.array_equal(g1.cpu().numpy(), m1) and np.array_equal(g2.cpu().numpy().astype(np.uint16), m2)
        out = pipe.undistort_frames(torch.from_numpy(frames).cuda())
        pipe.sync()
        for f in range(2):
            assert np.array_equal(out[f].cpu().numpy(), cv2.remap(frames[f], m1, m2, cv2.INTER_LINEAR))


def test_detection_behind_lens_correction_matches_oracle():
    """Whole path with the correction active == oracle on frames corrected by OpenCV (MD:88-89 then MD:111-249)."""
    import torch
    from vbs_b200 import pipeline
    K = [[900.3, 0, 323.1], [0, 898.3, 277.7], [0, 0, 1]]; D = [-0.10, 0.02, 4e-4, -3e-4, 0.0]
    frames = synth.workload_frames("small_6x8", 2, seed0=77)
    h, w = frames.shape[1:]
    want = [port.find_markers_frame(port.undistort_frame(f, K, D)) for f in frames]
    assert all(len(m) >= 40 for m in want)
    with pipeline.MarkerPipeline(h, w, 1, max_batch=2, max_markers=256, max_refs=1) as pipe:
        pipe.set_undistort(K, D)
        res = pipe.process(torch.from_numpy(frames).cuda(), 0); pipe.sync()
        assert [res.markers(f) for f in range(2)] == want
        pin = torch.from_numpy(frames).pin_memory()
        host = pipe.process_host_ptr(pin.data_ptr(), 2, h * w, w, 0, pipe.alloc_outputs(2, False))
        assert [host.markers(f) for f in range(2)] == want
        pipe.set_undistort(None)
        res = pipe.process(torch.from_numpy(frames).cuda(), 0); pipe.sync()
        assert [res.markers(f) for f in range(2)] == [port.find_markers_frame(f) for f in frames]


def test_config_errors_match_reference(tmp_path):
    with pytest.raises(ValueError, match="Missing required config key"):
        marker_detection.MarkerTracker({"video_path": "x"})
    with pytest.raises(FileNotFoundError):
        marker_detection.MarkerTracker({"video_path": str(tmp_path / "nope.avi"), "output_dir": str(tmp_path), "crop_ratios": (0, 0, 0, 0)})
    tr = object.__new__(marker_detection.MarkerTracker)
    tr.config = {}; tr.first_frame_markers = {}
    with pytest.raises(ValueError, match="No markers detected in first frame"):
        tr._process_first_frame([])


def test_marker_analysis_matches_reference_golden():
    import pandas as pd
    g = np.load(os.path.join(GOLDEN, "ring65_crop.npz"))
    an = reconstruction_3d.MarkerAnalysis(reconstruction_3d.Config(warmup_frames=0))
    an.camera.matrix, an.camera.dist_coeffs = g["K"], g["D"]
    an.camera.R_world_to_cam, an.camera.T_world_to_cam = g["R"], g["T"].reshape(3, 1)
    cam = port.Camera(g["K"], g["D"], g["R"], g["T"])
    pts = np.random.default_rng(0).uniform([0, 0], [480, 450], (200, 2))
    assert np.abs(an._undistort_points(pts) - port.undistort_points(cam, pts)).max() <= 1e-9
    # float64 arguments, as R3:279-287 passes them (pandas values); with bare Python floats NumPy-2
    # promotion would make the reference compute the whole formula in float32
    for u, v, d in ((100.3, 200.2, 13.1), (400.0, 50.5, 12.2)):
        u, v, d = np.float64(u), np.float64(v), np.float64(d)
        assert np.abs(an._calculate_3d_position(u, v, d) - port.position_3d(cam, u, v, d)).max() <= 1e-9
    with pytest.raises(ValueError):
        an._calculate_3d_position(float(np.float32(g["K"][0, 2])), float(np.float32(g["K"][1, 2])), 12.0)
    rows = g["rows"]
    df = pd.DataFrame({"frameno": rows[:, 0].astype(int), "row": rows[:, 1].astype(int), "col": rows[:, 2].astype(int),
                       "u": rows[:, 5], "v": rows[:, 6], "major_axis": rows[:, 7]})
    out = an._track_markers(df)
    want = g["rows3d"]
    assert list(out.columns) == ["frameno", "row", "col", "X", "Y", "Z", "dX", "dY", "dZ", "displacement"]
    assert len(out) == len(want)
    a = out.sort_values(["frameno", "row", "col"]).to_numpy(dtype=np.float64)
    b = want[np.lexsort((want[:, 2], want[:, 1], want[:, 0]))]
    assert np.array_equal(a[:, :3], b[:, :3]) and np.abs(a[:, 3:] - b[:, 3:]).max() <= 1e-9
    # warm-up: the first `warmup_frames` frames are dropped and the next one yields no rows (R3:255-256)
    an2 = reconstruction_3d.MarkerAnalysis(reconstruction_3d.Config(warmup_frames=1))
    an2.camera = an.camera
    out2 = an2._track_markers(df)
    assert set(out2["frameno"]) == {2}


def _write_analysis_inputs(tmp_path, g):
    import pandas as pd
    rows = g["rows"]
    cols = ["frameno", "row", "col", "Ox", "Oy", "Cx", "Cy", "major_axis", "minor_axis", "angle"]
    df = pd.DataFrame(rows[:, :10], columns=cols).astype({"frameno": int, "row": int, "col": int})
    csv = tmp_path / "markers.csv"
    df.to_csv(csv, index=False)                                                   # MD:466-467
    K, D, R, T = g["K"], g["D"], g["R"], g["T"].reshape(3)
    intr = {"fx": K[0, 0], "fy": K[1, 1], "cx": K[0, 2], "cy": K[1, 2], "k1": D[0], "k2": D[1], "p1": D[2], "p2": D[3], "k3": D[4]}
    extr = {f"R_wc_{i + 1}{j + 1}": R[i, j] for i in range(3) for j in range(3)}
    extr.update({"Tx_wc": T[0], "Ty_wc": T[1], "Tz_wc": T[2]})
    for name, d in (("intr.csv", intr), ("extr.csv", extr)):
        pd.DataFrame({"Parameter": list(d), "Value": [float(v) for v in d.values()]}).to_csv(tmp_path / name, index=False)
    return csv, tmp_path / "intr.csv", tmp_path / "extr.csv"


def test_run_analysis_csv_in_table_out(tmp_path):
    """The file formats either side of the 3D path: tracking CSV + Parameter/Value tables in, 3D table out (R3:405-442)."""
    import pandas as pd
    g = np.load(os.path.join(GOLDEN, "ring65_crop.npz"))
    csv, intr, extr = _write_analysis_inputs(tmp_path, g)
    cfg = reconstruction_3d.Config(warmup_frames=0, data_dir=tmp_path / "data", output_dir=tmp_path / "out", plots_dir=tmp_path / "out" / "plots")
    an = reconstruction_3d.MarkerAnalysis(cfg)
    out = an.run_analysis(csv, intr, extr)
    assert np.array_equal(an.camera.matrix, g["K"]) and an.camera.T_world_to_cam.shape == (3, 1)
    back = pd.read_csv(tmp_path / "out" / "marker_3d_coordinates.csv", float_precision="round_trip")
    assert list(back.columns) == ["frameno", "row", "col", "X", "Y", "Z", "dX", "dY", "dZ", "displacement"]
    want = g["rows3d"]
    a = back.sort_values(["frameno", "row", "col"]).to_numpy(dtype=np.float64)
    b = want[np.lexsort((want[:, 2], want[:, 1], want[:, 0]))]
    assert len(out) == len(want) and np.array_equal(a[:, :3], b[:, :3]) and np.abs(a[:, 3:] - b[:, 3:]).max() <= 1e-9
    with pytest.raises(FileNotFoundError):
        an.load_marker_data(tmp_path / "missing.csv")
    bad = tmp_path / "bad.csv"
    bad.write_text("frameno,row,col,Cx\n0,0,0,1.0\n")
    with pytest.raises(ValueError, match="Missing required columns"):
        an.load_marker_data(bad)


def test_fit_plane_least_squares_prints_like_reference():
    g = np.load(os.path.join(GOLDEN, "tiny_4x5.npz"))
    X, Y, Z = synth.ring_layout_mm().T
    th, az = np.deg2rad(15.0), np.deg2rad(30.0)
    Zp = np.tan(th) * (np.cos(az) * X + np.sin(az) * Y) + 0.7 + np.random.default_rng(3).normal(0, 0.01, 65)
    want = port.plane_tilt(X, Y, Zp)
    got = force_distribution.fit_plane_tilt(X, Y, Zp)
    assert abs(got[3] - want[3]) <= 1e-4 and np.abs(np.array(got[:3]) - np.array(want[:3])).max() <= 1e-9
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        patch, label = force_distribution.fit_plane_least_squares(None, X, Y, Zp, label="Tilted")
    assert label == "Tilted" and buf.getvalue() == f"-> Plane Fit (Tilted): Tilt Angle = {want[3]:.2f} degrees\n"
    assert "15.0" in buf.getvalue()                                   # SURVEY A.10 known answer: prints 15.01
