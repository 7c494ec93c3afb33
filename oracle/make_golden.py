#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python -m oracle.make_golden

For every case the real ``MarkerTracker._find_markers`` / ``_marker_center`` /
``_track_markers`` (marker_detection.py:111-249,349-396), the real
``MarkerAnalysis._track_markers`` (3d_reconstruction.py:240-316, loaded through
oracle/refload.py) and the real ``fit_plane_least_squares`` (ForceDistribution.py:138-162)
are executed on synthetic frames; ``oracle/port.py`` is asserted equal to them output by
output, and inputs + outputs are written as compressed fixtures so the GPU box (which has no
/root/reference) can check the CUDA path against the reference's own results.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vbs_b200  # noqa: E402,F401
from vbs_b200 import synth, reference_state  # noqa: E402
from oracle import port, refload  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def tap_reference(md, frame_bgr):
    """Run the reference's two static methods, capturing `labeled` and the contours by wrapping
    the callables in the reference module's namespace (its own lines still execute)."""
    cap = {}
    orig_label, orig_fc = md.ndimage.label, md.cv2.findContours

    def label_tap(*a, **k):
        r = orig_label(*a, **k)
        cap["labeled"], cap["n_labels"] = r
        return r

    class Cv2Tap:
        def __getattr__(self, name):
            return getattr(md_cv2, name)

        def findContours(self, img, *a, **k):
            cap["opened"] = img.copy()
            r = orig_fc(img, *a, **k)
            cap["contours"] = r[0]
            return r

    md_cv2 = md.cv2
    md.ndimage.label = label_tap
    md.cv2 = Cv2Tap()
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mask, area = md.MarkerTracker._find_markers(frame_bgr)
            markers = md.MarkerTracker._marker_center(mask, area, None)
    finally:
        md.ndimage.label = orig_label
        md.cv2 = md_cv2
    cap.update(mask=mask, area_mask=area, markers=markers)
    return cap


def run_case(name, frames, cols, cam_params, layout="grid"):
    md = refload.marker_detection()
    B = len(frames)
    per_frame = []
    for fr in frames:
        bgr = np.repeat(fr[..., None], 3, axis=2) if fr.ndim == 2 else fr
        ref = tap_reference(md, bgr)
        taps = {}
        mine = port.find_markers_frame(fr, taps)
        assert np.array_equal(ref["mask"], taps["mask"]) and np.array_equal(ref["area_mask"], taps["area_mask"]), name
        assert np.array_equal(ref.get("labeled", 0), taps.get("labeled", 0)), name
        assert ref["markers"] == mine, name
        per_frame.append(ref)
    # reference state: frame-0 detections; grid ids for grids, list order for the ring layout
    m0 = per_frame[0]["markers"]
    if layout == "grid":
        keys, xy = reference_state.grid_ids(np.array([m["center"] for m in m0]), cols)
    else:
        keys = [(0, i) for i in range(len(m0))]
        xy = np.array([m["center"] for m in m0])
    ref_markers = {k: {"Ox": p[0], "Oy": p[1]} for k, p in zip(keys, xy)}
    rows_all = []
    for f, ref in enumerate(per_frame):
        tr = refload.bare_tracker(ref_markers, 20, frame_count=f)
        rows = tr._track_markers(None, ref["markers"])
        assert rows == port.track_rows(keys, xy, ref["markers"], f, 20), name
        rows_all.append(rows)
    # 3D: the real MarkerAnalysis._track_markers on the rows as a DataFrame like load_marker_data returns
    import pandas as pd
    K, D, R, T = cam_params
    an = refload.make_analysis(K, D, R, T, warmup_frames=0)
    flat = [r for rows in rows_all for r in rows]
    df = pd.DataFrame(flat).rename(columns={"Cx": "u", "Cy": "v"})
    df = df[df["major_axis"] >= 5.0].sort_values("frameno", kind="stable").reset_index(drop=True)
    res3d = an._track_markers(df.copy())
    tab = {k: np.array([r[k] for r in flat]) for k in ("frameno", "row", "col", "Cx", "Cy", "major_axis")}
    mine3d = port.displacement_rows(port.Camera(K, D, R, T), tab, warmup_frames=0)
    cols3 = ["frameno", "row", "col", "X", "Y", "Z", "dX", "dY", "dZ", "displacement"]
    a = res3d[cols3].values if len(res3d) else np.zeros((0, 10))
    b = np.array([[r[c] for c in cols3] for r in mine3d]).reshape(-1, 10)
    assert np.array_equal(a, b), (name, a.shape, b.shape)
    # plane on the last frame's positions (reference function, full-precision coefficients via the lstsq tap)
    cam = port.Camera(K, D, R, T)
    last = rows_all[-1]
    uv = port.undistort_points(cam, np.array([[r["Cx"], r["Cy"]] for r in last]))
    P = np.array([port.position_3d(cam, u, v, r["major_axis"]) for r, (u, v) in zip(last, uv)])
    first = rows_all[0]
    uv0 = port.undistort_points(cam, np.array([[r["Cx"], r["Cy"]] for r in first]))
    P0 = {(r["row"], r["col"]): port.position_3d(cam, u, v, r["major_axis"]) for r, (u, v) in zip(first, uv0)}
    H, W = frames[0].shape[:2]
    ref_xyz = np.stack([(xy[:, 0] - W / 2) / 11.0, (xy[:, 1] - H / 2) / 11.0, np.zeros(len(xy))], 1)
    idx = [keys.index((r["row"], r["col"])) for r in last]
    start = np.array([P0[keys[i]] for i in idx])
    X, Y, Z = port.deviation_endpoints(ref_xyz[idx], P - start, np.zeros_like(start))
    pa, pb, pc, text = refload.fit_plane_reference(X, Y, Z)
    mine_p = port.plane_tilt(X, Y, Z)
    assert (pa, pb, pc) == tuple(mine_p[:3]), name
    assert f"{mine_p[3]:.2f}" in text, (text, mine_p)

    M = max(len(p["markers"]) for p in per_frame)
    pack = dict(
        frames=np.stack(frames), cols=np.int32(cols), K=K, D=D, R=R, T=T,
        ref_keys=np.array(keys, dtype=np.int32), ref_xy=xy,
        area_bits=np.stack([np.packbits(p["area_mask"] > 0, axis=1) for p in per_frame]),
        mask_bits=np.stack([np.packbits(p["mask"] > 0, axis=1) for p in per_frame]),
        opened_bits=np.stack([np.packbits(p["opened"] > 0, axis=1) for p in per_frame]),
        labeled=np.stack([p["labeled"].astype(np.uint16) for p in per_frame]),
        n_labels=np.array([p["n_labels"] for p in per_frame], dtype=np.int32),
        n_markers=np.array([len(p["markers"]) for p in per_frame], dtype=np.int32),
        marker_xy=np.stack([np.pad(np.array([m["center"] for m in p["markers"]]).reshape(-1, 2), ((0, M - len(p["markers"])), (0, 0))) for p in per_frame]),
        marker_axes=np.stack([np.pad(np.array([[m["major_axis"], m["minor_axis"], m["angle"]] for m in p["markers"]]).reshape(-1, 3), ((0, M - len(p["markers"])), (0, 0))) for p in per_frame]),
        rows=np.array([[r["frameno"], r["row"], r["col"], r["Ox"], r["Oy"], r["Cx"], r["Cy"], r["major_axis"], r["minor_axis"], r["angle"]] for r in flat]),
        rows3d=a, plane=np.array([pa, pb, pc, mine_p[3]]), plane_text=np.array(text),
        ref_xyz=ref_xyz,
    )
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, f"{name}.npz")
    np.savez_compressed(path, **pack)
    print(f"{name}: {B} frames, markers {pack['n_markers'].tolist()}, {len(flat)} rows, {len(a)} 3D rows, "
          f"tilt {mine_p[3]:.6f} deg -> {os.path.getsize(path) / 1e3:.0f} kB")


def main():
    assert refload.available(), "reference checkout not found"
    K, D, R, T = synth.synthetic_camera()

    def cam_for(h, w):
        Kc = K.copy(); Kc[0, 2] = np.float32(w / 2 + 3.1); Kc[1, 2] = np.float32(h / 2 - 2.3)
        return Kc, D, R, T

    # <=480 branch, grid
    h, w, rows, cols, pitch, rad = synth.WORKLOADS["tiny_4x5"]
    run_case("tiny_4x5", list(synth.workload_frames("tiny_4x5", 3, seed0=10)), cols, cam_for(h, w))
    # >480 branch, grid
    h, w, rows, cols, pitch, rad = synth.WORKLOADS["small_6x8"]
    run_case("small_6x8", list(synth.workload_frames("small_6x8", 2, seed0=20)), cols, cam_for(h, w))
    # the sensor's 65-marker ring layout on the cropped 640x480 frame (crop (1/8,1/8,1/16,0) -> 480 x 450, MD:481)
    full_h, full_w = 480, 640
    left, right, top, bottom = port.crop_box(full_w, full_h, (1 / 8, 1 / 8, 1 / 16, 0))
    centres = synth.ring_layout(full_h, full_w, px_per_mm=11.0, dy=15.0)
    seq = synth.compression_sequence(full_h, full_w, centres, 6.0, 3, tilt=0.6, depth=1.0, seed0=30)
    crops = [np.ascontiguousarray(f[top:bottom, left:right]) for f in seq]
    run_case("ring65_crop", crops, 0, cam_for(bottom - top, right - left), layout="list")


def video_case():
    """Config-1 substitute (the demo mp4 is absent from the checkout): a lossless FFV1 ring video run
    through the UNMODIFIED ``MarkerTracker(config).process()`` (MD:429-462); the CSV it writes is the golden."""
    import cv2
    import tempfile
    import pandas as pd
    md = refload.marker_detection()
    full_h, full_w = 480, 640
    centres = synth.ring_layout(full_h, full_w, px_per_mm=11.0, dy=15.0)
    seq = synth.compression_sequence(full_h, full_w, centres, 6.0, 5, tilt=0.5, depth=1.0, seed0=40, noise_sigma=1.0)
    os.makedirs(OUT, exist_ok=True)
    vpath = os.path.join(OUT, "ring_video.avi")
    wr = cv2.VideoWriter(vpath, cv2.VideoWriter_fourcc(*"FFV1"), 12.0, (full_w, full_h))
    for f in seq:
        wr.write(np.repeat(f[..., None], 3, axis=2))
    wr.release()
    cap = cv2.VideoCapture(vpath)
    for f in seq:                                   # lossless round trip or the golden is meaningless
        ok, fr = cap.read()
        assert ok and np.array_equal(fr[..., 0], f) and np.array_equal(fr[..., 1], f)
    cap.release()
    tmp = tempfile.mkdtemp(prefix="vbs_vid_")
    cfg = {"video_path": vpath, "output_dir": tmp, "crop_ratios": (1 / 8, 1 / 8, 1 / 16, 0), "num_layers": 5, "min_marker_distance": 20}
    orig = md.cv2.destroyAllWindows
    md.cv2.destroyAllWindows = lambda: None         # raises on headless OpenCV, after the CSV is saved (MD:474)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            md.MarkerTracker(cfg).process()
    finally:
        md.cv2.destroyAllWindows = orig
    import shutil
    shutil.copyfile(os.path.join(tmp, "ring_video_markers.csv"), os.path.join(OUT, "ring_video_markers.csv"))   # byte for byte
    df = pd.read_csv(os.path.join(OUT, "ring_video_markers.csv"), float_precision="round_trip")
    print(f"ring_video: {len(seq)} frames, {len(df)} CSV rows, keys {sorted(set(zip(df.row, df.col)))}, "
          f"video {os.path.getsize(vpath) / 1e3:.0f} kB")


UNDISTORT_LENS = {"camera_matrix": [[520.3, 0.0, 243.1], [0.0, 518.7, 222.7], [0.0, 0.0, 1.0]],
                  "dist_coeffs": [-0.12, 0.03, 5e-4, -3e-4, 0.0]}


def undistorted_video_case():
    """The same video through the UNMODIFIED reference with ``calibration_params`` configured, i.e. with the optional
    lens correction of MD:88-109 active: golden CSV + the corrected first frame as ``_preprocess_frame`` returns it."""
    import cv2
    import shutil
    import tempfile
    md = refload.marker_detection()
    vpath = os.path.join(OUT, "ring_video.avi")
    tmp = tempfile.mkdtemp(prefix="vbs_vid_u_")
    cfg = {"video_path": vpath, "output_dir": tmp, "crop_ratios": (1 / 8, 1 / 8, 1 / 16, 0), "num_layers": 5, "min_marker_distance": 20,
           "calibration_params": UNDISTORT_LENS}
    orig = md.cv2.destroyAllWindows
    md.cv2.destroyAllWindows = lambda: None
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            tr = md.MarkerTracker(cfg)
            tr.process()
            cap = cv2.VideoCapture(vpath)
            ok, fr = cap.read()
            cap.release()
            first = tr._preprocess_frame(fr)                      # width/height were set by process()
    finally:
        md.cv2.destroyAllWindows = orig
    assert np.array_equal(first, port.undistort_frame(fr[30:, 80:560], UNDISTORT_LENS["camera_matrix"], UNDISTORT_LENS["dist_coeffs"]))
    shutil.copyfile(os.path.join(tmp, "ring_video_markers.csv"), os.path.join(OUT, "ring_video_undistorted_markers.csv"))
    np.savez_compressed(os.path.join(OUT, "ring_undistort.npz"), K=np.array(UNDISTORT_LENS["camera_matrix"]),
                        D=np.array(UNDISTORT_LENS["dist_coeffs"]), first_frame=first)
    import pandas as pd
    df = pd.read_csv(os.path.join(OUT, "ring_video_undistorted_markers.csv"), float_precision="round_trip")
    print(f"ring_video (undistorted): {len(df)} CSV rows, keys {sorted(set(zip(df.row, df.col)))}, first frame {first.shape}")


if __name__ == "__main__":
    import sys
    if "--undistort-only" not in sys.argv:
        main()
        video_case()
    undistorted_video_case()
