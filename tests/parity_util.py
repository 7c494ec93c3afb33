"""Shared parity machinery: run the CUDA path (through the C ABI) and the oracle on the same
frames and report, stage by stage, how they differ.  Used by tests/test_gpu_parity.py and by
tools/gpu_diag.py (which prints the whole report instead of stopping at the first failure)."""
from __future__ import annotations

import numpy as np

import vbs_b200  # noqa: F401  (import shim)
from vbs_b200 import capi, pipeline, synth
from oracle import port


def f32_ulps(a, b):
    """|a-b| in units of float32 ulp at max(|a|,|b|) (values are float32-representable or close)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.spacing(np.maximum(np.abs(a), np.abs(b)).astype(np.float32)).astype(np.float64)
    return np.abs(a - b) / scale


def angle_ulps(a, b):
    """Angle difference mod 180 deg in float32 ulps at the angle's magnitude (the reference's angle is a float32
    from cv2.fitEllipse, +90 in float64: MD:213-217).  SURVEY 8d: <= 2 ulp, about 1.5e-5 deg."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    da = np.abs(a - b) % 180.0
    d = np.minimum(da, 180.0 - da)
    scale = np.spacing(np.maximum(np.maximum(np.abs(a), np.abs(b)), 1.0).astype(np.float32)).astype(np.float64)
    return d / scale


ANGLE_TOL_ULP = 2.0


def grid_reference(markers0, cols):
    """Reference-state array = detections of frame 0 in ascending raster order, ids (i//cols, i%cols)."""
    from vbs_b200 import reference_state
    return reference_state.grid_ids(np.array([m["center"] for m in markers0]), cols)


def oracle_frames(frames, ref_keys=None, ref_xy=None, min_dist=20.0, frameno0=0):
    """Run the oracle port on every frame; returns list of dicts with taps, markers and rows."""
    out = []
    for i, fr in enumerate(frames):
        taps = {}
        markers = port.find_markers_frame(fr, taps)
        rec = {"taps": taps, "markers": markers}
        if ref_keys is not None:
            rec["rows"] = port.track_rows(ref_keys, ref_xy, markers, frameno0 + i, min_dist)
        out.append(rec)
    return out


def compare_detection(pipe: pipeline.MarkerPipeline, frames_np, res, oracle, stages=True):
    """Stage-by-stage comparison.  Returns dict name -> (ok, detail)."""
    import torch

    B = len(frames_np)
    rep = {}
    h = res.to_host()
    if stages:
        for name, stage, key, conv in (
            ("area_mask", capi.STAGE_AREA_MASK, "area_mask", lambda a: a),
            ("mask", capi.STAGE_MASK, "mask", lambda a: a),
            ("maxima", capi.STAGE_MAXIMA, "maxima", lambda a: a.astype(np.uint8)),
            ("opened", capi.STAGE_OPENED, "opened", lambda a: a),
            ("labeled", capi.STAGE_LABELS, "labeled", lambda a: a.astype(np.int32)),
        ):
            got = pipe.debug_stage(stage, B).cpu().numpy()
            bad_total, first = 0, None
            for f in range(B):
                if key not in oracle[f]["taps"]:
                    # the reference returns early when nothing was labelled (MD:177-178); the stage
                    # image is still defined, so derive it with the same library call
                    t = {}
                    port.opened_contours(oracle[f]["taps"]["area_mask"], t)
                    oracle[f]["taps"].setdefault("opened", t["opened"])
                want = conv(oracle[f]["taps"][key])
                bad = np.argwhere(got[f] != want)
                bad_total += len(bad)
                if first is None and len(bad):
                    y, x = bad[0]
                    first = (f, int(y), int(x), int(got[f][y, x]), int(want[y, x]), len(bad))
            rep[name] = (bad_total == 0, f"{bad_total} mismatching px" + (f"; first (frame,y,x,got,want,count)={first}" if first else ""))
        rc = pipe.debug_stage(capi.STAGE_RECHECKS, B).cpu().numpy()
        rep["rechecks"] = (True, f"float64 rechecks per frame: {rc.tolist()}")
    # labels / centres
    nl_want = [o["taps"].get("n_labels", 0) for o in oracle]
    rep["n_labels"] = (h.n_labels.tolist() == nl_want, f"got {h.n_labels.tolist()} want {nl_want}")
    cbad, cmax = 0, 0.0
    for f in range(B):
        n = min(int(h.n_labels[f]), nl_want[f])
        if n and "centres" in oracle[f]["taps"]:
            d = np.abs(h.centres[f, :n] - oracle[f]["taps"]["centres"][:n])
            cbad += int((d != 0).sum())
            cmax = max(cmax, float(d.max()))
    rep["centres"] = (cbad == 0, f"{cbad} non-identical coordinates, max |d| = {cmax:.3e} px")
    # markers
    nm_want = [len(o["markers"]) for o in oracle]
    rep["n_markers"] = (h.n_markers.tolist() == nm_want, f"got {h.n_markers.tolist()} want {nm_want}")
    xy_bad, ax_ulps, ang_max, ang_ulp = 0, 0.0, 0.0, 0.0
    for f in range(B):
        n = min(int(h.n_markers[f]), nm_want[f])
        for k in range(n):
            m = oracle[f]["markers"][k]
            xy_bad += int(h.marker_xy[f, k, 0] != m["center"][0]) + int(h.marker_xy[f, k, 1] != m["center"][1])
            ax_ulps = max(ax_ulps, float(f32_ulps(h.marker_axes[f, k, 0], m["major_axis"])), float(f32_ulps(h.marker_axes[f, k, 1], m["minor_axis"])))
            da = abs(h.marker_axes[f, k, 2] - m["angle"]) % 180.0
            ang_max = max(ang_max, min(da, 180.0 - da))
            ang_ulp = max(ang_ulp, float(angle_ulps(h.marker_axes[f, k, 2], m["angle"])))
    rep["marker_order_xy"] = (xy_bad == 0, f"{xy_bad} centre coordinates differ (order or value)")
    rep["marker_axes"] = (ax_ulps <= 2.0, f"max axis error {ax_ulps:.2f} float32 ulp")
    rep["marker_angle"] = (ang_ulp <= ANGLE_TOL_ULP, f"max angle error {ang_max:.3e} deg (mod 180) = {ang_ulp:.2f} float32 ulp")
    return rep


def compare_rows(res, oracle, ref_keys):
    """Tracking rows (MD:349-396): same refs matched, same Cx/Cy, axes and angle within 2 float32 ulp."""
    h = res.to_host()
    B, R = len(oracle), len(ref_keys)
    id_bad, xy_bad, ax_ulps, ang_ulp, nrows = 0, 0, 0.0, 0.0, 0
    for f in range(B):
        want = {(r["row"], r["col"]): r for r in oracle[f]["rows"]}
        nrows += len(want)
        for r, key in enumerate(ref_keys):
            got = h.row_det[f, r] >= 0
            if got != (key in want):
                id_bad += 1
                continue
            if got:
                w = want[key]
                xy_bad += int(h.row_cxy[f, r, 0] != w["Cx"]) + int(h.row_cxy[f, r, 1] != w["Cy"])
                ax_ulps = max(ax_ulps, float(f32_ulps(h.row_axes[f, r, 0], w["major_axis"])),
                              float(f32_ulps(h.row_axes[f, r, 1], w["minor_axis"])))
                ang_ulp = max(ang_ulp, float(angle_ulps(h.row_axes[f, r, 2], w["angle"])))
    return {"row_ids": (id_bad == 0, f"{id_bad} (frame, ref) pairs matched differently; {nrows} oracle rows"),
            "row_xy": (xy_bad == 0, f"{xy_bad} Cx/Cy values differ"),
            "row_axes": (ax_ulps <= 2.0, f"max axis error {ax_ulps:.2f} float32 ulp"),
            "row_angle": (ang_ulp <= ANGLE_TOL_ULP, f"max angle error {ang_ulp:.2f} float32 ulp (mod 180)")}


def oracle_3d(oracle, cam: port.Camera, warmup, frameno0=0, marker_diameter_mm=2.0, min_size=5.0, max_disp=50.0):
    """Feed the oracle's tracking rows through the 3D port (R3:172-176, 240-316); returns
    (rows keyed by (frameno,row,col), positions keyed the same for every observation that enters R3)."""
    tab = {k: [] for k in ("frameno", "row", "col", "Cx", "Cy", "major_axis")}
    for o in oracle:
        for r in o["rows"]:
            for k in tab:
                tab[k].append(r[k])
    tab = {k: np.asarray(v) for k, v in tab.items()}
    rows = port.displacement_rows(cam, tab, warmup_frames=warmup, marker_diameter_mm=marker_diameter_mm,
                                  min_marker_size_px=min_size, max_displacement=max_disp)
    by_key = {(int(r["frameno"]), int(r["row"]), int(r["col"])): r for r in rows}
    # positions of every kept observation (for the plane fit)
    pos = {}
    if len(tab["frameno"]):
        keep = tab["major_axis"] >= min_size
        first = tab["frameno"][keep].min() if keep.any() else 0
        sel = keep & (tab["frameno"] >= first + max(warmup, 0))
        uv = port.undistort_points(cam, np.stack([tab["Cx"][sel], tab["Cy"][sel]], axis=1))
        for (fr, rw, cl, dm), (u, v) in zip(zip(tab["frameno"][sel], tab["row"][sel], tab["col"][sel], tab["major_axis"][sel]), uv):
            pos[(int(fr), int(rw), int(cl))] = port.position_3d(cam, u, v, dm, marker_diameter_mm)
    return by_key, pos


def compare_3d(res, ref_keys, rows3d, pos, frameno0=0, tol=1e-9):
    h = res.to_host()
    B, R = h.pos_flags.shape
    miss, extra, dmax, pmax, pbad = 0, 0, 0.0, 0.0, 0
    for f in range(B):
        for r, key in enumerate(ref_keys):
            k = (frameno0 + f, key[0], key[1])
            fl = int(h.pos_flags[f, r])
            want_pos = pos.get(k)
            if (want_pos is not None) != bool(fl & 2):
                pbad += 1
            elif want_pos is not None:
                pmax = max(pmax, float(np.abs(h.pos3d[f, r, :3] - want_pos).max()))
            w = rows3d.get(k)
            if (w is not None) != bool(fl & 4):
                if w is not None:
                    miss += 1
                else:
                    extra += 1
            elif w is not None:
                want = np.array([w[c] for c in ("X", "Y", "Z", "dX", "dY", "dZ", "displacement")])
                dmax = max(dmax, float(np.abs(h.pos3d[f, r] - want).max()))
    return {"pos3d_presence": (pbad == 0, f"{pbad} (frame, ref) pairs with different position validity"),
            "pos3d_value": (pmax <= tol, f"max |dP| = {pmax:.3e} mm over {len(pos)} observations"),
            "disp_rows": (miss == 0 and extra == 0, f"{miss} missing / {extra} extra displacement rows of {len(rows3d)}"),
            "disp_value": (dmax <= tol, f"max |d| = {dmax:.3e} mm")}


def oracle_plane(pos, ref_keys, frames, ref_xyz, start_xyz, d_vert, shell=False, scale=1.0):
    """Per-frame plane fit on the oracle positions (FD:196-204,219-232,141-159)."""
    out = {}
    for fr in frames:
        idx = [i for i, k in enumerate(ref_keys) if pos.get((fr, k[0], k[1])) is not None]
        if len(idx) < 3:
            out[fr] = None
            continue
        P = np.array([pos[(fr, ref_keys[i][0], ref_keys[i][1])] for i in idx])
        X, Y, Z = port.deviation_endpoints(ref_xyz[idx], P - start_xyz[idx], d_vert[idx], shell=shell, scale=scale)
        out[fr] = port.plane_tilt(X, Y, Z)
    return out


def compare_plane(res, planes, frameno0=0, tol_deg=1e-4):
    h = res.to_host()
    dmax, cmax, bad = 0.0, 0.0, 0
    for f in range(h.plane.shape[0]):
        w = planes.get(frameno0 + f)
        if w is None:
            bad += int(np.isfinite(h.plane[f, 3]))
            continue
        if not np.isfinite(h.plane[f, 3]):
            bad += 1
            continue
        dmax = max(dmax, abs(h.plane[f, 3] - w[3]))
        cmax = max(cmax, float(np.abs(h.plane[f, :3] - np.array(w[:3])).max()))
    return {"plane_presence": (bad == 0, f"{bad} frames with different fit availability"),
            "plane_tilt": (dmax <= tol_deg, f"max |d tilt| = {dmax:.3e} deg, max |d coeff| = {cmax:.3e}")}


def format_report(rep):
    return "\n".join(f"  [{'ok' if ok else 'FAIL'}] {name}: {detail}" for name, (ok, detail) in rep.items())


def assert_report(rep):
    failed = {k: v for k, v in rep.items() if not v[0]}
    assert not failed, "parity failures:\n" + format_report(failed) + "\nfull report:\n" + format_report(rep)
