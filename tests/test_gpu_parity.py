"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against (1) the committed golden
outputs of the unmodified reference, (2) the oracle port on seeded frames incl. edge cases, and
(3) size-independent properties at BASELINE.json's full batch size.

Tolerances (stated once): masks, labels, counts, IDs, marker order, centroids: bit-exact.
Ellipse axes and angle: <= 2 float32 ulp (angle mod 180, about 1.5e-5 deg).  3D positions /
displacements: <= 1e-9 mm.  Plane tilt: <= 1e-4 deg (north_star), observed ~1e-14.
"""
import os
import warnings

import numpy as np
import pytest

import parity_util as pu
import vbs_b200  # noqa: F401
from vbs_b200 import capi, pipeline, synth, reference_state
from oracle import port

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def torch_cuda(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def unpack(bits, W):
    return np.unpackbits(bits, axis=-1)[..., :W]


# ---------------------------------------------------------------------------------------------
# 1. golden vectors produced by the reference itself
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["tiny_4x5", "small_6x8", "ring65_crop"])
def test_cuda_reproduces_reference_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    frames = g["frames"]
    B, H, W = frames.shape
    keys = [tuple(k) for k in g["ref_keys"].tolist()]
    R = len(keys)
    with pipeline.MarkerPipeline(H, W, 1, max_batch=B, max_markers=256, max_refs=R) as pipe:
        pipe.set_reference([k[0] for k in keys], [k[1] for k in keys], g["ref_xy"][:, 0], g["ref_xy"][:, 1], 20.0)
        pipe.set_camera(g["K"], g["D"], g["R"], g["T"], 2.0, 5.0, 50.0, warmup_frames=0)
        res = pipe.process(torch_cuda(frames), 0)
        pipe.sync()
        h = res.to_host()
        for stage, key, on in ((capi.STAGE_AREA_MASK, "area_bits", 255), (capi.STAGE_MASK, "mask_bits", 1), (capi.STAGE_OPENED, "opened_bits", 255)):
            got = pipe.debug_stage(stage, B).cpu().numpy()
            assert np.array_equal(got, unpack(g[key], W) * on), key
        assert np.array_equal(pipe.debug_stage(capi.STAGE_LABELS, B).cpu().numpy(), g["labeled"].astype(np.int32))
        assert h.n_labels.tolist() == g["n_labels"].tolist()
        assert h.n_markers.tolist() == g["n_markers"].tolist()
        for f in range(B):
            n = int(g["n_markers"][f])
            assert np.array_equal(h.marker_xy[f, :n], g["marker_xy"][f, :n])                     # order + bit-exact centroids
            assert pu.f32_ulps(h.marker_axes[f, :n, :2], g["marker_axes"][f, :n, :2]).max() <= 2.0
            assert pu.angle_ulps(h.marker_axes[f, :n, 2], g["marker_axes"][f, :n, 2]).max() <= pu.ANGLE_TOL_ULP
        # tracking rows (CSV contract MD:380-391): same (frameno,row,col) set, same Cx/Cy
        want = {(int(r[0]), int(r[1]), int(r[2])): r for r in g["rows"]}
        got_rows = 0
        for f in range(B):
            for r, k in enumerate(keys):
                key = (f, k[0], k[1])
                assert (h.row_det[f, r] >= 0) == (key in want), key
                if key in want:
                    got_rows += 1
                    assert h.row_cxy[f, r, 0] == want[key][5] and h.row_cxy[f, r, 1] == want[key][6]
        assert got_rows == len(want)
        # 3D rows (R3:296-307)
        want3 = {(int(r[0]), int(r[1]), int(r[2])): r[3:] for r in g["rows3d"]}
        n3 = 0
        for f in range(B):
            for r, k in enumerate(keys):
                key = (f, k[0], k[1])
                emitted = bool(h.pos_flags[f, r] & 4)
                assert emitted == (key in want3), key
                if emitted:
                    n3 += 1
                    assert np.abs(h.pos3d[f, r] - want3[key]).max() <= 1e-9
        assert n3 == len(want3)
        # plane of the last frame against the reference's own fit_plane_least_squares (tilt <= 1e-4 deg)
        start = np.nan_to_num(h.pos3d[0, :, :3])
        pipe.set_plane(g["ref_xyz"], start, None)
        pipe.reset_sequence()
        res = pipe.process(torch_cuda(frames), 0)
        pipe.sync()
        pl = res.to_host().plane[B - 1]
        assert abs(pl[3] - g["plane"][3]) <= 1e-4 and np.abs(pl[:3] - g["plane"][:3]).max() <= 1e-9


# ---------------------------------------------------------------------------------------------
# 2. seeded frames against the oracle port, incl. ragged sizes and BGR input
# ---------------------------------------------------------------------------------------------
def run_against_oracle(frames, cols, channels=1, max_markers=512, with3d=True):
    H, W = frames.shape[1:3]
    ora0 = pu.oracle_frames(frames[:1])
    m0 = ora0[0]["markers"]
    keys, xy = pu.grid_reference(m0, max(cols, 1)) if m0 else ([], np.zeros((0, 2)))
    oracle = pu.oracle_frames(frames, keys, xy, 20.0)
    with pipeline.MarkerPipeline(H, W, channels, max_batch=len(frames), max_markers=max_markers, max_refs=max(len(keys), 1)) as pipe:
        if keys:
            pipe.set_reference([k[0] for k in keys], [k[1] for k in keys], xy[:, 0], xy[:, 1], 20.0)
        K, D, R, T = synth.synthetic_camera()
        K = K.copy(); K[0, 2] = W / 2 + 3.1; K[1, 2] = H / 2 - 2.3
        if keys and with3d:
            pipe.set_camera(K, D, R, T, 2.0, 5.0, 50.0, warmup_frames=0)
        res = pipe.process(torch_cuda(frames), 0)
        pipe.sync()
        rep = pu.compare_detection(pipe, frames, res, oracle)
        if keys:
            rep.update(pu.compare_rows(res, oracle, keys))
            if with3d:
                rows3d, pos = pu.oracle_3d(oracle, port.Camera(K, D, R, T), warmup=0)
                rep.update(pu.compare_3d(res, keys, rows3d, pos))
        pu.assert_report(rep)
    return rep


@pytest.mark.parametrize("name,n", [("tiny_4x5", 3), ("small_6x8", 3)])
def test_grid_frames_match_oracle(name, n):
    run_against_oracle(synth.workload_frames(name, n, seed0=100), synth.WORKLOADS[name][3])


def test_bgr_input_matches_oracle():
    fr = synth.workload_frames("small_6x8", 2, seed0=7)
    bgr = np.repeat(fr[..., None], 3, axis=3).copy()
    bgr[..., 0] = np.clip(bgr[..., 0].astype(int) + 9, 0, 255)
    bgr[..., 2] = np.clip(bgr[..., 2].astype(int) - 11, 0, 255)
    run_against_oracle(bgr, 8, channels=3)


@pytest.mark.parametrize("h,w,rows,cols,pitch,rad", [(301, 357, 4, 5, 56.0, 6.0), (563, 645, 6, 8, 60.0, 11.0), (481, 130, 5, 1, 62.0, 11.0)])
def test_ragged_sizes_match_oracle(h, w, rows, cols, pitch, rad):
    """H not a multiple of 8, W not a multiple of 32/128, a strip narrower than the halo."""
    centres = synth.grid_layout(h, w, rows, cols, pitch)
    frames = np.stack([synth.render_frame(h, w, centres, rad, seed=s) for s in (1, 2)])
    run_against_oracle(frames, cols)


def test_edge_cases_match_oracle():
    h, w = 520, 600
    rng = np.random.default_rng(5)
    blank = np.full((h, w), 170, np.uint8)
    dark = np.full((h, w), 12, np.uint8)
    single = synth.render_frame(h, w, np.array([[300.0, 260.0]]), 11.0, seed=3)
    border = synth.render_frame(h, w, np.array([[4.0, 5.0], [595.0, 300.0], [300.0, 516.0], [120.0, 130.0], [2.0, 300.0]]), 11.0, seed=4, jitter=0)
    noise = np.clip(rng.normal(128, 60, (h, w)), 0, 255).astype(np.uint8)
    bright = synth.render_frame(h, w, synth.grid_layout(h, w, 4, 5, 90.0), 11.0, seed=6)
    bright = (255 - bright.astype(int)).clip(0, 255).astype(np.uint8)          # bright markers: exercises the uint8 wrap (MD:128)
    touching = synth.render_frame(h, w, np.array([[200.0, 200.0], [222.0, 200.0], [400.0, 300.0], [400.0, 323.0]]), 11.0, seed=8, jitter=0)
    frames = np.stack([blank, dark, single, border, noise, bright, touching])
    run_against_oracle(frames, 1, max_markers=4096, with3d=False)


def test_capacity_overflow_is_reported():
    h, w = 520, 600
    noise = np.clip(np.random.default_rng(5).normal(128, 60, (1, h, w)), 0, 255).astype(np.uint8)
    n = len(port.find_markers_frame(noise[0], t := {})) or t.get("n_labels", 0)
    if t.get("n_labels", 0) < 9:
        pytest.skip("noise frame did not produce enough components")
    with pipeline.MarkerPipeline(h, w, 1, max_batch=1, max_markers=8, max_refs=1) as pipe:
        pipe.process(torch_cuda(noise), 0)
        with pytest.raises(capi.VbsError, match="exceed max_markers"):
            pipe.sync()


def test_bad_arguments_raise_value_error():
    with pytest.raises(ValueError):
        pipeline.MarkerPipeline(4, 4)
    with pipeline.MarkerPipeline(560, 640, 1, max_batch=1, max_markers=64, max_refs=4) as pipe:
        with pytest.raises(ValueError):
            pipe.set_reference(np.arange(9), np.arange(9), np.zeros(9), np.zeros(9))
        with pytest.raises(ValueError):
            pipe.process(np.zeros((2, 560, 640), np.uint8))
        with pytest.raises(ValueError, match="Focal lengths must be positive"):
            pipe.set_camera(np.zeros((3, 3)), np.zeros(5), np.eye(3), np.zeros(3))


# ---------------------------------------------------------------------------------------------
# 3. stage entry points and the crop view
# ---------------------------------------------------------------------------------------------
def test_stage_entry_points_and_crop_pitch():
    import torch
    full = synth.workload_frames("small_6x8", 2, seed0=11)
    H, W = full.shape[1:]
    taps = {}
    want = port.find_markers_frame(full[0], taps)
    with pipeline.MarkerPipeline(H, W, 1, max_batch=2, max_markers=256, max_refs=1) as pipe:
        mask, area = pipe.find_markers(torch_cuda(full))
        assert np.array_equal(mask[0].cpu().numpy(), taps["mask"]) and np.array_equal(area[0].cpu().numpy(), taps["area_mask"])
        res = pipe.marker_center(torch_cuda(np.stack([taps["mask"]] * 2)), torch_cuda(np.stack([taps["area_mask"]] * 2)))
        pipe.sync()
        assert res.markers(1) == want
    # crop (MD:81-85) is a pointer + pitch at the boundary: no copy of the frame is made on the host
    left, right, top, bottom = port.crop_box(W, H, (1 / 8, 1 / 8, 1 / 16, 0))
    ch, cw = bottom - top, right - left
    crop = np.ascontiguousarray(full[:, top:bottom, left:right])
    want_c = [port.find_markers_frame(c) for c in crop]
    with pipeline.MarkerPipeline(ch, cw, 1, max_batch=2, max_markers=256, max_refs=1) as pipe:
        pin = torch.from_numpy(full).pin_memory()
        outs = pipe.alloc_outputs(2, False)
        res = pipe.process_host_ptr(pin.data_ptr() + top * W + left, 2, H * W, W, 0, outs)
        assert [res.markers(f) for f in range(2)] == want_c


def _disk(img, cx, cy, r, v=1):
    yy, xx = np.mgrid[:img.shape[0], :img.shape[1]]
    img[(xx - cx) ** 2 + (yy - cy) ** 2 <= r * r] = v


def nested_masks(h=520, w=640):
    """Hand-made masks: a ring with a blob inside its hole (dropped by RETR_EXTERNAL), a blob inside
    an OPEN cavity (kept), a contour with > 128 vertices (re-trace path), blobs touching the frame."""
    area = np.zeros((h, w), np.uint8)
    _disk(area, 150, 150, 60); _disk(area, 150, 150, 38, 0); _disk(area, 150, 150, 12)        # ring + nested disk
    _disk(area, 440, 200, 75)                                                                  # long contour
    area[330:430, 100:110] = 1; area[330:430, 190:200] = 1; area[420:430, 100:200] = 1          # U shape
    _disk(area, 150, 370, 14)                                                                  # blob in the open cavity
    _disk(area, 3, 250, 16); _disk(area, 636, 500, 15); _disk(area, 320, 517, 13)               # frame contact
    _disk(area, 560, 420, 20); _disk(area, 560, 420, 6, 0)                                     # small hole, nothing inside
    mask = np.zeros((h, w), np.uint8)
    for cx, cy, r in ((150, 150, 22), (440, 200, 28), (150, 370, 22), (3, 250, 20), (636, 500, 20), (320, 517, 20), (560, 420, 24), (150, 92, 20)):
        _disk(mask, cx, cy, r)
    return mask, area * 255


def test_marker_center_nested_holes_and_long_contours():
    mask, area = nested_masks()
    taps = {}
    want = port.marker_center(mask, area, taps)
    assert len(taps["contours"]) >= 7 and max(len(c) for c in taps["contours"]) > 128
    with pipeline.MarkerPipeline(mask.shape[0], mask.shape[1], 1, max_batch=2, max_markers=64, max_refs=1) as pipe:
        res = pipe.marker_center(torch_cuda(np.stack([mask, mask])), torch_cuda(np.stack([area, area])))
        pipe.sync()
        assert np.array_equal(pipe.debug_stage(capi.STAGE_OPENED, 2).cpu().numpy()[0], taps["opened"])
        assert np.array_equal(pipe.debug_stage(capi.STAGE_LABELS, 2).cpu().numpy()[1], taps["labeled"])
        for f in range(2):
            got = res.markers(f)
            assert len(got) == len(want) and len(want) >= 3
            for a, b in zip(got, want):
                assert a["center"] == b["center"]
                assert pu.f32_ulps(a["major_axis"], b["major_axis"]) <= 2 and pu.f32_ulps(a["minor_axis"], b["minor_axis"]) <= 2
                assert pu.angle_ulps(a["angle"], b["angle"]) <= pu.ANGLE_TOL_ULP


# ---------------------------------------------------------------------------------------------
# 4. BASELINE.json full size: oracle on unique frames + size-independent properties at batch 256
# ---------------------------------------------------------------------------------------------
def test_full_1080p_batch256_properties():
    import torch
    name, U, B = "1080p_20x20", 32, 256                      # >= 32 unique oracle-checked frames at 1080p
    H, W, rows, cols, _, _ = synth.WORKLOADS[name]
    uniq = synth.workload_frames(name, U, seed0=0)
    ora0 = pu.oracle_frames(uniq[:1])
    keys, xy = pu.grid_reference(ora0[0]["markers"], cols)
    oracle = pu.oracle_frames(uniq, keys, xy, 20.0)
    K, D, R, T = synth.synthetic_camera()
    cam = port.Camera(K, D, R, T)
    rows3d, pos = pu.oracle_3d(oracle, cam, warmup=0)
    start = np.array([pos[(0, k[0], k[1])] for k in keys])
    ref_xyz = np.stack([(xy[:, 0] - W / 2) / 11.0, (xy[:, 1] - H / 2) / 11.0, np.zeros(len(xy))], 1)
    d_vert = np.zeros_like(start)
    planes = pu.oracle_plane(pos, keys, range(U), ref_xyz, start, d_vert)
    batch = torch_cuda(np.tile(uniq, (B // U, 1, 1)))
    with pipeline.MarkerPipeline(H, W, 1, max_batch=B, max_markers=1024, max_refs=len(keys)) as pipe:
        pipe.set_reference([k[0] for k in keys], [k[1] for k in keys], xy[:, 0], xy[:, 1], 20.0)
        pipe.set_camera(K, D, R, T, 2.0, 5.0, 50.0, warmup_frames=0)
        pipe.set_plane(ref_xyz, start, d_vert)
        res = pipe.process(batch, 0)
        pipe.sync()
        h = res.to_host()
        # (a) the unique frames against the oracle (stage images of 4 of them; markers, rows, 3D field and plane of all)
        first = pipeline.BatchResult(0, *[getattr(h, k)[:U] if getattr(h, k) is not None else None for k in
                                          ("n_labels", "centres", "n_markers", "marker_xy", "marker_axes", "row_det", "row_cxy", "row_axes", "pos3d", "pos_flags", "plane", "plane_n")])
        rep = pu.compare_detection(pipe, uniq, first, oracle, stages=False)
        rep.update(pu.compare_rows(first, oracle, keys))
        rep.update(pu.compare_3d(first, keys, rows3d, pos))
        rep.update(pu.compare_plane(first, planes))
        for stage, key, on in ((capi.STAGE_AREA_MASK, "area_mask", 1), (capi.STAGE_MASK, "mask", 1), (capi.STAGE_LABELS, "labeled", 1)):
            got = pipe.debug_stage(stage, U).cpu().numpy()
            bad = sum(int((got[f] != oracle[f]["taps"][key]).sum()) for f in (0, 7, 19, 31))
            rep["stage_" + key] = (bad == 0, f"{bad} mismatching px in 4 full-size stage images")
        pu.assert_report(rep)
        # (b) tiling invariance: frame i of the batch == frame i mod U, bit for bit
        for k in ("n_labels", "n_markers", "marker_xy", "marker_axes", "row_det", "row_cxy"):
            a = getattr(h, k)
            assert np.array_equal(a, np.tile(a[:U], (B // U,) + (1,) * (a.ndim - 1)), equal_nan=True), k
        assert (h.n_markers == rows * cols).all()
        # (c) idempotence and batch-split invariance incl. the last-seen carry (R3:277,314)
        pipe.reset_sequence()
        again = pipe.process(batch, 0); pipe.sync(); again = again.to_host()
        assert np.array_equal(again.pos3d, h.pos3d, equal_nan=True) and np.array_equal(again.pos_flags, h.pos_flags)
        pipe.reset_sequence()
        a = pipe.process(batch[:100], 0); pipe.sync(); a = a.to_host()
        b = pipe.process(batch[100:], 100); pipe.sync(); b = b.to_host()
        assert np.array_equal(np.concatenate([a.pos3d, b.pos3d]), h.pos3d, equal_nan=True)
        assert np.array_equal(np.concatenate([a.pos_flags, b.pos_flags]), h.pos_flags)
        # (d) displacement linearity: dX of frame f = X(f) - X(f-1) wherever both are valid
        ok = (h.pos_flags[1:] & 4).astype(bool) & (h.pos_flags[:-1] & 2).astype(bool)
        d = h.pos3d[1:, :, :3] - h.pos3d[:-1, :, :3]
        assert np.abs((d - h.pos3d[1:, :, 3:6])[ok]).max() <= 1e-12
        # (e) host entry point gives the same bytes as the device entry point
        pipe.reset_sequence()
        hh = pipe.process(np.ascontiguousarray(batch[:32].cpu().numpy()), 0)
        assert np.array_equal(hh.marker_xy, h.marker_xy[:32]) and np.array_equal(hh.pos3d, h.pos3d[:32], equal_nan=True)


# ---------------------------------------------------------------------------------------------
# 5. frame sharding: N shards processed independently + last-seen exchange == one sequential run
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_run_is_byte_identical_to_single_run(world):
    import torch
    from vbs_b200 import sharding
    h, w, rows, cols = 560, 640, 6, 8
    centres = synth.grid_layout(h, w, rows, cols, 60.0)
    seq = synth.compression_sequence(h, w, centres, 11.0, 19, tilt=0.4, depth=1.0, seed0=500)
    # dropouts that straddle shard boundaries: paint two markers out for a few frames
    for f in (4, 5, 9, 10, 11):
        x, y = centres[13].astype(int); seq[f, y - 20:y + 20, x - 20:x + 20] = 170
    for f in range(7, 15):
        x, y = centres[30].astype(int); seq[f, y - 20:y + 20, x - 20:x + 20] = 170
    K, D, R, T = synth.synthetic_camera()
    K = K.copy(); K[0, 2] = w / 2 + 3.1; K[1, 2] = h / 2 - 2.3

    def make():
        p = pipeline.MarkerPipeline(h, w, 1, max_batch=len(seq), max_markers=256, max_refs=rows * cols)
        p.set_reference([k[0] for k in keys], [k[1] for k in keys], xy[:, 0], xy[:, 1], 20.0)
        p.set_camera(K, D, R, T, 2.0, 5.0, 50.0, warmup_frames=2)
        p.set_first_frame(0)
        return p

    keys, xy = pu.grid_reference(port.find_markers_frame(seq[0]), cols)
    dev = torch_cuda(seq)
    with make() as p:
        single = p.process(dev, 0); p.sync(); single = single.to_host()
    assert (single.pos_flags & 4).any() and not (single.pos_flags[:, 13] & 1)[[4, 5, 9, 10, 11]].any()
    pipes, results, tails = [], [], []
    for r in range(world):
        lo, hi = sharding.shard_bounds(len(seq), r, world)
        p = make()
        res = p.process(dev[lo:hi], lo); p.sync()
        pipes.append(p); results.append(res); tails.append(p.get_last_seen())
    for r in range(world):
        inc = sharding.incoming_last_seen(np.stack(tails), r)
        capi.check(pipes[r]._ctx, capi.lib.vbs_fix_displacement(pipes[r]._ctx, results[r].pos3d.data_ptr(), results[r].pos_flags.data_ptr(),
                                                                results[r].pos3d.shape[0], inc.ctypes.data))
    hosts = [r.to_host() for r in results]
    for k in ("n_markers", "row_det", "row_cxy", "pos_flags", "pos3d"):
        cat = np.concatenate([getattr(x, k) for x in hosts])
        assert np.array_equal(cat, getattr(single, k), equal_nan=True), k
    for p in pipes:
        p.close()


# ---------------------------------------------------------------------------------------------
# 6. BASELINE.json config 4 geometry: 4K frame with the densest array the reference detects (40 x 72)
# ---------------------------------------------------------------------------------------------
def test_4k_dense_array_matches_oracle():
    name, U = "4k_40x72", 8                                   # >= 8 unique oracle-checked frames at 4K
    H, W, rows, cols, _, _ = synth.WORKLOADS[name]
    uniq = synth.workload_frames(name, U, seed0=0, jitter=1.0)
    ora0 = pu.oracle_frames(uniq[:1])
    assert len(ora0[0]["markers"]) == rows * cols          # 2880/2880 at this pitch
    keys, xy = pu.grid_reference(ora0[0]["markers"], cols)
    oracle = pu.oracle_frames(uniq, keys, xy, 20.0)
    K, D, R, T = synth.synthetic_camera()
    K = K.copy(); K[0, 2] = W / 2 + 3.1; K[1, 2] = H / 2 - 2.3; K[0, 0] *= 2; K[1, 1] *= 2
    rows3d, pos = pu.oracle_3d(oracle, port.Camera(K, D, R, T), warmup=0)
    batch = torch_cuda(uniq)
    with pipeline.MarkerPipeline(H, W, 1, max_batch=U, max_markers=4096, max_refs=len(keys)) as pipe:
        pipe.set_reference([k[0] for k in keys], [k[1] for k in keys], xy[:, 0], xy[:, 1], 20.0)
        pipe.set_camera(K, D, R, T, 2.0, 5.0, 50.0, warmup_frames=0)
        res = pipe.process(batch, 0)
        pipe.sync()
        h = res.to_host()
        got = pipe.debug_stage(capi.STAGE_AREA_MASK, U).cpu().numpy()
        assert all(np.array_equal(got[f], oracle[f]["taps"]["area_mask"]) for f in range(U))
        got = pipe.debug_stage(capi.STAGE_MASK, U).cpu().numpy()
        assert all(np.array_equal(got[f], oracle[f]["taps"]["mask"]) for f in range(U))
        got = pipe.debug_stage(capi.STAGE_LABELS, U).cpu().numpy()
        assert np.array_equal(got[1], oracle[1]["taps"]["labeled"]) and np.array_equal(got[6], oracle[6]["taps"]["labeled"])
        rep = pu.compare_detection(pipe, uniq, h, oracle, stages=False)
        rep.update(pu.compare_rows(h, oracle, keys))
        rep.update(pu.compare_3d(h, keys, rows3d, pos))
        pu.assert_report(rep)
        assert (h.n_markers >= rows * cols - 2).all()


# ---------------------------------------------------------------------------------------------
# 7. stress: random blob masks (holes, frame contact, nesting, ragged borders) through the
#    labelling / border following / ellipse / matching kernels against cv2 + scipy (oracle port)
# ---------------------------------------------------------------------------------------------
def random_blob_masks(seed, h=520, w=608):
    import cv2
    rng = np.random.default_rng(seed)
    field = cv2.GaussianBlur(rng.normal(0, 1, (h, w)).astype(np.float32), (0, 0), 6.0 + 3.0 * rng.random())
    area = (field > np.quantile(field, 0.62 + 0.2 * rng.random())).astype(np.uint8)
    field2 = cv2.GaussianBlur(rng.normal(0, 1, (h, w)).astype(np.float32), (0, 0), 9.0)
    mask = ((field2 > np.quantile(field2, 0.7)) | (area > 0) & (rng.random((h, w)) < 0.5)).astype(np.uint8)
    mask = cv2.morphologyEx(mask, cv2.MORPH_OPEN, np.ones((3, 3), np.uint8))
    return mask, area * 255


def random_ellipse_masks(seed, h=600, w=720, n=60):
    """Random filled ellipses (some overlapping, some cut by the frame); the mask holds the same
    ellipses 1.7x larger, so the ring maxima surround the area blobs and most of them match."""
    import cv2
    rng = np.random.default_rng(seed)
    area = np.zeros((h, w), np.uint8)
    mask = np.zeros((h, w), np.uint8)
    for _ in range(n):
        c = (int(rng.uniform(-10, w + 10)), int(rng.uniform(-10, h + 10)))
        ax = (int(rng.uniform(7, 24)), int(rng.uniform(7, 24)))
        ang = float(rng.uniform(0, 180))
        cv2.ellipse(area, c, ax, ang, 0, 360, 255, -1)
        cv2.ellipse(mask, c, (int(ax[0] * 1.7) + 6, int(ax[1] * 1.7) + 6), ang, 0, 360, 1, -1)
    return mask, area


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6, 101, 102, 103, 201, 202])
def test_random_blob_masks_match_oracle(seed):
    if seed > 200:                                   # three 1024-px labelling bands, five tile rows
        mask, area = random_blob_masks(seed, h=300, w=2100)
    else:
        mask, area = random_blob_masks(seed) if seed < 100 else random_ellipse_masks(seed)
    taps = {}
    want = port.marker_center(mask, area, taps)
    n_contours = len(taps.get("contours", ()))
    with pipeline.MarkerPipeline(mask.shape[0], mask.shape[1], 1, max_batch=1, max_markers=2048, max_refs=1) as pipe:
        res = pipe.marker_center(torch_cuda(mask[None]), torch_cuda(area[None]))
        pipe.sync()
        if "labeled" in taps:
            assert np.array_equal(pipe.debug_stage(capi.STAGE_MAXIMA, 1).cpu().numpy()[0], taps["maxima"].astype(np.uint8))
            assert np.array_equal(pipe.debug_stage(capi.STAGE_LABELS, 1).cpu().numpy()[0], taps["labeled"])
            assert np.array_equal(pipe.debug_stage(capi.STAGE_OPENED, 1).cpu().numpy()[0], taps["opened"])
            h = res.to_host()
            assert int(h.n_labels[0]) == taps["n_labels"]
            assert np.array_equal(h.centres[0, : taps["n_labels"]], taps["centres"])
        got = res.markers(0)
        assert len(got) == len(want), (len(got), len(want), n_contours)
        for a, b in zip(got, want):
            assert a["center"] == b["center"]
            assert pu.f32_ulps(a["major_axis"], b["major_axis"]) <= 2 and pu.f32_ulps(a["minor_axis"], b["minor_axis"]) <= 2
            assert pu.angle_ulps(a["angle"], b["angle"]) <= pu.ANGLE_TOL_ULP


# ---------------------------------------------------------------------------------------------
# 8. the TMA tile path and the generic loader give the same bits
# ---------------------------------------------------------------------------------------------
def test_tma_path_is_used_and_equals_generic_loader(monkeypatch):
    frames = synth.workload_frames("small_6x8", 3, seed0=31)
    H, W = frames.shape[1:]
    x = torch_cuda(frames)
    with pipeline.MarkerPipeline(H, W, 1, max_batch=3, max_markers=256, max_refs=1) as pipe:
        a = pipe.process(x, 0); pipe.sync()
        assert pipe.tma_launches >= 1                      # aligned gray frames: tiles staged by cp.async.bulk.tensor
        area_tma = pipe.debug_stage(capi.STAGE_AREA_MASK, 3).cpu().numpy()
        a = a.to_host()
    monkeypatch.setenv("VBS_NO_TMA", "1")
    with pipeline.MarkerPipeline(H, W, 1, max_batch=3, max_markers=256, max_refs=1) as pipe:
        b = pipe.process(x, 0); pipe.sync()
        assert pipe.tma_launches == 0
        assert np.array_equal(pipe.debug_stage(capi.STAGE_AREA_MASK, 3).cpu().numpy(), area_tma)
        b = b.to_host()
    assert np.array_equal(a.n_markers, b.n_markers)
    for f in range(3):                                 # entries beyond n_markers are padding, not output
        n = int(a.n_markers[f])
        assert np.array_equal(a.marker_xy[f, :n], b.marker_xy[f, :n]) and np.array_equal(a.marker_axes[f, :n], b.marker_axes[f, :n])
    # a view whose first pixel is not 16-byte aligned must fall back silently and still be right
    import torch
    big = torch.zeros((3, H, W + 16), dtype=torch.uint8, device="cuda")
    big[:, :, 3:W + 3] = x
    with pipeline.MarkerPipeline(H, W, 1, max_batch=3, max_markers=256, max_refs=1) as pipe:
        outs = pipe.alloc_outputs(3, True)
        capi.check(pipe._ctx, capi.lib.vbs_process_device(pipe._ctx, big.data_ptr() + 3, 3, H * (W + 16), W + 16, 0, __import__("ctypes").byref(outs[1])))
        pipe.sync()
        assert pipe.tma_launches == 0
        got_n, got_xy = outs[0]["n_markers"].cpu().numpy(), outs[0]["marker_xy"].cpu().numpy()
        assert np.array_equal(got_n, a.n_markers)
        for f in range(3):
            assert np.array_equal(got_xy[f, : got_n[f]], a.marker_xy[f, : got_n[f]])


# ---------------------------------------------------------------------------------------------
# 9. asynchronous host entry point (two batches in flight) == synchronous calls, incl. the last-seen carry
# ---------------------------------------------------------------------------------------------
def test_async_host_api_equals_synchronous_calls():
    import torch
    h, w, rows, cols = 560, 640, 6, 8
    centres = synth.grid_layout(h, w, rows, cols, 60.0)
    seq = synth.compression_sequence(h, w, centres, 11.0, 12, tilt=0.4, depth=1.0, seed0=700)
    keys, xy = pu.grid_reference(port.find_markers_frame(seq[0]), cols)
    K, D, R, T = synth.synthetic_camera()
    K = K.copy(); K[0, 2] = w / 2 + 3.1; K[1, 2] = h / 2 - 2.3
    pin = torch.from_numpy(seq).pin_memory()
    B = 4
    with pipeline.MarkerPipeline(h, w, 1, max_batch=B, max_markers=256, max_refs=len(keys)) as pipe:
        pipe.set_reference([k[0] for k in keys], [k[1] for k in keys], xy[:, 0], xy[:, 1], 20.0)
        pipe.set_camera(K, D, R, T, 2.0, 5.0, 50.0, warmup_frames=0)
        sync_res = []
        for s in range(3):
            o = pipe.alloc_outputs(B, False)
            r = pipe.process_host_ptr(pin.data_ptr() + s * B * h * w, B, h * w, w, s * B, o)
            sync_res.append({k: np.copy(v) for k, v in o[0].items()})
        pipe.reset_sequence()
        outs = [pipe.alloc_outputs(B, False) for _ in range(3)]
        pipe.submit_host_ptr(pin.data_ptr(), B, h * w, w, 0, outs[0])
        pipe.submit_host_ptr(pin.data_ptr() + B * h * w, B, h * w, w, B, outs[1])
        with pytest.raises(capi.VbsError, match="two batches already in flight"):
            pipe.submit_host_ptr(pin.data_ptr(), B, h * w, w, 0, outs[2])
        pipe.wait_host()
        pipe.submit_host_ptr(pin.data_ptr() + 2 * B * h * w, B, h * w, w, 2 * B, outs[2])
        pipe.wait_host(); pipe.wait_host()
        with pytest.raises(capi.VbsError, match="no batch in flight"):
            pipe.wait_host()
        for s in range(3):
            for k, v in sync_res[s].items():
                assert np.array_equal(outs[s][0][k], v, equal_nan=True), (s, k)
        assert (sync_res[2]["pos_flags"] & 4).any()


def test_chunked_async_batches_with_changing_chunk_size():
    """Chunked submit_host with batches of different sizes in flight: the chunk size follows the batch (at most 8 chunks
    per batch), so chunk c of the new batch covers another scratch range than chunk c of the one before and the
    per-chunk-index events alone do not order the reuse (enqueue_host_chunks waits for every live chunk when the size
    changes)."""
    import torch
    h, w, rows, cols = 560, 640, 6, 8
    sizes = (12, 24, 16)                                           # chunk sizes 2, 3, 2 with vbs_set_host_chunk(2)
    centres = synth.grid_layout(h, w, rows, cols, 60.0)
    seq = synth.compression_sequence(h, w, centres, 11.0, sum(sizes), tilt=0.3, depth=1.0, seed0=710)
    keys, xy = pu.grid_reference(port.find_markers_frame(seq[0]), cols)
    K, D, R, T = synth.synthetic_camera()
    pin = torch.from_numpy(seq).pin_memory()
    starts = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    with pipeline.MarkerPipeline(h, w, 1, max_batch=max(sizes), max_markers=256, max_refs=len(keys)) as pipe:
        pipe.set_reference([k[0] for k in keys], [k[1] for k in keys], xy[:, 0], xy[:, 1], 20.0)
        pipe.set_camera(K, D, R, T, 2.0, 5.0, 50.0, warmup_frames=0)
        pipe.set_host_chunk(2)
        sync_res = []
        for n, f0 in zip(sizes, starts):
            o = pipe.alloc_outputs(n, False)
            pipe.process_host_ptr(pin.data_ptr() + int(f0) * h * w, n, h * w, w, int(f0), o)
            sync_res.append({k: np.copy(v) for k, v in o[0].items()})
        for rep in range(4):
            pipe.reset_sequence()
            outs = [pipe.alloc_outputs(n, False) for n in sizes]
            for s, (n, f0) in enumerate(zip(sizes, starts)):
                if s == 2:
                    pipe.wait_host()
                pipe.submit_host_ptr(pin.data_ptr() + int(f0) * h * w, n, h * w, w, int(f0), outs[s])
            pipe.wait_host(); pipe.wait_host()
            for s in range(3):
                for k, v in sync_res[s].items():
                    assert np.array_equal(outs[s][0][k], v, equal_nan=True), (rep, s, k)


# ---------------------------------------------------------------------------------------------
# 10. BASELINE.json config 3: vertical / tilted compression sequences on the sensor's 65-marker ring
#     layout -> tracking -> 3D displacement -> deviation against the vertical baseline -> plane tilt
# ---------------------------------------------------------------------------------------------
def test_compression_sequences_full_pipeline_matches_oracle():
    full_h, full_w, n = 480, 640, 128                          # sequences of >= 128 frames
    left, right, top, bottom = port.crop_box(full_w, full_h, (1 / 8, 1 / 8, 1 / 16, 0))
    H, W = bottom - top, right - left
    centres = synth.ring_layout(full_h, full_w, px_per_mm=11.0, dy=15.0)
    vert = synth.compression_sequence(full_h, full_w, centres, 6.0, n, tilt=0.0, depth=1.0, seed0=900, noise_sigma=1.0)
    tilt = synth.compression_sequence(full_h, full_w, centres, 6.0, n, tilt=0.7, depth=1.0, seed0=2900, noise_sigma=1.0)
    crop = lambda s: np.ascontiguousarray(s[:, top:bottom, left:right])
    vert, tilt = crop(vert), crop(tilt)
    K, D, R, T = synth.synthetic_camera()
    K = K.copy(); K[0, 2] = W / 2 + 3.1; K[1, 2] = H / 2 - 2.3
    cam = port.Camera(K, D, R, T)
    m0 = port.find_markers_frame(vert[0])
    assert len(m0) == 65
    keys = [(0, i) for i in range(65)]
    xy = np.array([m["center"] for m in m0])
    ref_xyz = np.stack([(xy[:, 0] - W / 2) / 11.0, (xy[:, 1] - H / 2) / 11.0, np.zeros(65)], 1)

    def oracle_positions(frames):
        orc = pu.oracle_frames(frames, keys, xy, 20.0)
        rows3d, pos = pu.oracle_3d(orc, cam, warmup=0)
        return orc, rows3d, pos

    ov, rv, pv = oracle_positions(vert)
    ot, rt, pt = oracle_positions(tilt)
    start_v = np.array([pv[(0, 0, i)] for i in range(65)])
    end_v = np.array([pv[(n - 1, 0, i)] for i in range(65)])
    d_vert = end_v - start_v                                                    # FD:197-198
    start_t = np.array([pt[(0, 0, i)] for i in range(65)])
    planes = pu.oracle_plane(pt, keys, range(n), ref_xyz, start_t, d_vert)
    with pipeline.MarkerPipeline(H, W, 1, max_batch=n, max_markers=256, max_refs=65) as pipe:
        pipe.set_reference([k[0] for k in keys], [k[1] for k in keys], xy[:, 0], xy[:, 1], 20.0)
        pipe.set_camera(K, D, R, T, 2.0, 5.0, 50.0, warmup_frames=0)
        rvg = pipe.process(torch_cuda(vert), 0); pipe.sync(); rvg = rvg.to_host()
        rep = pu.compare_rows(rvg, ov, keys)
        rep.update(pu.compare_3d(rvg, keys, rv, pv))
        gs, ge = rvg.pos3d[0, :, :3], rvg.pos3d[n - 1, :, :3]
        assert np.abs((ge - gs) - d_vert).max() <= 1e-9
        pipe.reset_sequence()
        pipe.set_plane(ref_xyz, start_t, d_vert)
        rtg = pipe.process(torch_cuda(tilt), 0); pipe.sync()
        rep2 = pu.compare_rows(rtg, ot, keys)
        rep2.update(pu.compare_3d(rtg, keys, rt, pt))
        rep2.update(pu.compare_plane(rtg, planes))
        pu.assert_report(rep); pu.assert_report(rep2)
        tilts = rtg.to_host().plane[:, 3]
        assert np.isfinite(tilts).all() and tilts[-1] > 1.0         # the tilted press really tilts the fitted plane


# ---------------------------------------------------------------------------------------------
# 11. labelling tiles whose label table fills up (closed early, k_ccl.cu) and links across tile edges
# ---------------------------------------------------------------------------------------------
def dense_masks(h=300, w=1100):
    """Thousands of one-pixel components inside one 64 x 1024 labelling tile, lines that cross the tile
    edges (rows 64, 128, ...; column 1024) in 4- and 8-connected ways, and a dense array of small blobs."""
    mask = np.zeros((h, w), np.uint8)
    mask[20:110:2, 100:380:2] = 1                       # 45 x 140 isolated dots
    mask[5:250, 301] = 1                                # vertical line through the dot field (joins its neighbours)
    mask[150, 900:1090] = 1                             # horizontal line across the band edge at x = 1024
    for i in range(120):                                # 4-connected staircase across rows 64 / 128 and x = 1024
        mask[40 + i, 960 + i] = 1; mask[40 + i, 961 + i] = 1
    area = np.zeros((h, w), np.uint8)
    blob = np.ones((7, 7), np.uint8); blob[0, 0] = blob[0, 6] = blob[6, 0] = blob[6, 6] = 0
    for y in range(20, 200, 9):
        for x in range(50, 1085, 9):
            area[y:y + 7, x:x + 7] = blob
            mask[y + 3, x + 3] = 1
    for i in range(150):                                # 8-connected diagonal band, 5 px thick so that it survives the open
        area[210 + i // 3: 215 + i // 3, 940 + i: 945 + i] = 1
    _disk(area, 300, 255, 38)                           # one large blob with seven centroids inside its distance gate
    for y, x in ((252, 296), (252, 300), (254, 298), (256, 302), (256, 296), (258, 300), (254, 304)):
        mask[y, x] = 1
    return mask, area * 255


def test_dense_components_overflow_the_tile_label_table():
    mask, area = dense_masks()
    taps = {}
    want = port.marker_center(mask, area, taps)
    assert 5000 < taps["n_labels"] <= 8192 and len(taps["contours"]) > 2000 and len(want) > 100
    with pipeline.MarkerPipeline(mask.shape[0], mask.shape[1], 1, max_batch=2, max_markers=8192, max_refs=1) as pipe:
        res = pipe.marker_center(torch_cuda(np.stack([mask, mask])), torch_cuda(np.stack([area, area])))
        pipe.sync()
        assert np.array_equal(pipe.debug_stage(capi.STAGE_LABELS, 2).cpu().numpy()[1], taps["labeled"])
        assert np.array_equal(pipe.debug_stage(capi.STAGE_OPENED, 2).cpu().numpy()[0], taps["opened"])
        h = res.to_host()
        for f in range(2):
            assert int(h.n_labels[f]) == taps["n_labels"]
            assert np.array_equal(h.centres[f, : taps["n_labels"]], taps["centres"])
            got = res.markers(f)
            assert len(got) == len(want)
            for a, b in zip(got, want):
                assert a["center"] == b["center"]
                assert pu.f32_ulps(a["major_axis"], b["major_axis"]) <= 2 and pu.f32_ulps(a["minor_axis"], b["minor_axis"]) <= 2
                assert pu.angle_ulps(a["angle"], b["angle"]) <= pu.ANGLE_TOL_ULP


# ---------------------------------------------------------------------------------------------
# 12. nearest-marker match (MD:369-372): exact ties and near-ties resolve like cdist + argmin
# ---------------------------------------------------------------------------------------------
def test_track_markers_ties_match_cdist_argmin():
    from scipy.spatial.distance import cdist
    rng = np.random.default_rng(9)
    refs = np.array([[100.0, 100.0], [300.25, 200.5], [50.0, 400.0], [600.0, 20.0], [333.3, 333.3]])
    mk = [[97.0, 100.0], [103.0, 100.0],                       # exact tie for ref 0: first wins
          [300.25, 205.5], [305.25, 200.5], [300.25, 195.5],   # three-way exact tie for ref 1
          [50.0 + 7.0 * (1 + 2e-16), 400.0], [50.0, 407.0],    # near-tie (one ulp apart) for ref 2
          [640.0, 20.0],                                       # farther than min_marker_distance from ref 3
          [333.3 + 3e-13, 336.3], [333.3, 336.3 + 1e-13]]      # near-tie for ref 4
    mk += rng.uniform(700, 900, (300, 2)).tolist()
    mk = np.array(mk)
    markers = [{"center": (float(x), float(y)), "major_axis": 12.0 + i, "minor_axis": 11.0, "angle": 90.0} for i, (x, y) in enumerate(mk)]
    with pipeline.MarkerPipeline(64, 64, 1, max_batch=1, max_markers=512, max_refs=8) as pipe:
        pipe.set_reference(np.arange(5), np.zeros(5, int), refs[:, 0], refs[:, 1], 20.0)
        det, cxy, axes = pipe.track_markers(markers)
    d = cdist(refs, mk)
    for r in range(5):
        j = int(np.argmin(d[r]))
        if d[r, j] > 20.0:
            assert det[r] == -1
        else:
            assert det[r] == j and tuple(cxy[r]) == tuple(mk[j]) and axes[r, 0] == 12.0 + j
    assert det[3] == -1 and det[0] == 0 and det[1] == 2


# ---------------------------------------------------------------------------------------------
# 13. the two-stream schedule (open-mask branch beside the NCC) and the sequential one give the same tables
# ---------------------------------------------------------------------------------------------
def test_branch_overlap_equals_sequential_schedule(monkeypatch):
    frames = synth.workload_frames("small_6x8", 4, seed0=61)
    H, W = frames.shape[1:]
    x = torch_cuda(frames)
    keys, xy = pu.grid_reference(port.find_markers_frame(frames[0]), 8)
    K, D, R, T = synth.synthetic_camera()

    def run():
        with pipeline.MarkerPipeline(H, W, 1, max_batch=4, max_markers=256, max_refs=len(keys)) as pipe:
            pipe.set_reference([k[0] for k in keys], [k[1] for k in keys], xy[:, 0], xy[:, 1], 20.0)
            pipe.set_camera(K, D, R, T, 2.0, 5.0, 50.0, warmup_frames=0)
            res = pipe.process(x, 0); pipe.sync()
            opened = pipe.debug_stage(capi.STAGE_OPENED, 4).cpu().numpy()
            labels = pipe.debug_stage(capi.STAGE_LABELS, 4).cpu().numpy()
            return res.to_host(), opened, labels

    a, a_open, a_lab = run()                                   # default: overlapped
    monkeypatch.setenv("VBS_BRANCH_OVERLAP", "0")
    b, b_open, b_lab = run()
    assert np.array_equal(a_open, b_open) and np.array_equal(a_lab, b_lab)
    for k in ("n_labels", "n_markers", "row_det", "pos_flags"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    for f in range(4):
        n = int(a.n_markers[f])
        assert n >= 40 and np.array_equal(a.marker_xy[f, :n], b.marker_xy[f, :n]) and np.array_equal(a.marker_axes[f, :n], b.marker_axes[f, :n])
    assert np.array_equal(a.row_cxy, b.row_cxy, equal_nan=True) and np.array_equal(a.pos3d, b.pos3d, equal_nan=True)
    want = [port.find_markers_frame(f) for f in frames]
    assert [len(w) for w in want] == a.n_markers.tolist()


@pytest.mark.parametrize("h,w,n", [(300, 1000, 230), (500, 900, 230)])
def test_mixed_segment_plan_equals_uniform_segments(monkeypatch, h, w, n):
    """The strip-marching kernels (blur, NCC) run the first (frame, strip) items whole-height and only the tail of the
    grid in row segments (VbsSegPlan, vbs_ctx.h).  Enough items here to have both kinds (>= 3 waves of 4 x 148 CTAs); the masks
    must equal those of the uniform plan (VBS_SEG_PLAN=0, every item in the same segments) and cv2 / the port."""
    import cv2
    rng = np.random.default_rng(h)
    radius = 6.0 if h <= 480 else 11.0
    centres = synth.grid_layout(h, w, 4, 12, 60.0)
    base = np.stack([synth.render_frame(h, w, centres, radius, seed=300 + i) for i in range(6)])
    frames = np.ascontiguousarray(base[np.arange(n) % len(base)])
    frames[:, ::7, ::5] ^= rng.integers(0, 64, frames[:, ::7, ::5].shape, dtype=np.uint8)      # make every frame different
    x = torch_cuda(frames)

    def run():
        with pipeline.MarkerPipeline(h, w, 1, max_batch=n, max_markers=256, max_refs=1) as pipe:
            pipe.process(x, 0); pipe.sync()
            return pipe.debug_stage(capi.STAGE_AREA_MASK, n).cpu().numpy(), pipe.debug_stage(capi.STAGE_MASK, n).cpu().numpy()

    a_area, a_mask = run()
    monkeypatch.setenv("VBS_SEG_PLAN", "0")
    b_area, b_mask = run()
    assert np.array_equal(a_area, b_area) and np.array_equal(a_mask, b_mask)
    for f in (0, n // 2, n - 1):                                  # first item (whole height), middle, last (row segments)
        want_mask, want_area = port.detect_masks(frames[f])[:2]
        assert np.array_equal(a_area[f] != 0, want_area != 0) and np.array_equal(a_mask[f] != 0, want_mask != 0), f


# ---------------------------------------------------------------------------------------------
# 14. K2 in isolation (vbs_ncc_mask): arbitrary area masks against the reference's three-FFT normxcorr2
#     (MD:146-164) > 0.1, bit for bit - random blob fields, border-heavy content, near-empty and near-full
#     frames, both template sizes.  The float32 filter bands (k_ncc.cu BAND / BAND_BORDER) are what is on trial:
#     every pixel the filter cannot decide must have been re-decided in float64.
# ---------------------------------------------------------------------------------------------
def ncc_stress_masks(h, w, seed):
    import cv2
    rng = np.random.default_rng(seed)
    kind = seed % 10
    a = np.zeros((h, w), np.uint8)
    if kind in (0, 1, 2):                                  # random blob field, three densities
        field = cv2.GaussianBlur(rng.normal(0, 1, (h, w)).astype(np.float32), (0, 0), 3.0 + 6.0 * rng.random())
        a = (field > np.quantile(field, (0.5, 0.8, 0.95)[kind])).astype(np.uint8)
    elif kind == 3:                                        # border-heavy: frames, stripes and blobs hugging every edge
        t = int(rng.integers(1, 30))
        a[:t] = 1; a[-int(rng.integers(1, 30)):] = 1; a[:, :int(rng.integers(1, 30))] = 1; a[:, -int(rng.integers(1, 30)):] = 1
        for _ in range(12):
            cx, cy = (int(rng.integers(0, w)), int(rng.choice([0, h - 1]))) if rng.random() < 0.5 else (int(rng.choice([0, w - 1])), int(rng.integers(0, h)))
            cv2.circle(a, (cx, cy), int(rng.integers(5, 40)), 1, -1)
    elif kind == 4:                                        # near-empty: a few pixels / one small blob (mean ~ 1e-4)
        for _ in range(int(rng.integers(1, 6))):
            a[int(rng.integers(0, h)), int(rng.integers(0, w))] = 1
        if rng.random() < 0.5:
            cv2.circle(a, (int(rng.integers(0, w)), int(rng.integers(0, h))), 3, 1, -1)
    elif kind == 5:                                        # near-full: everything set but a few holes
        a[:] = 1
        for _ in range(int(rng.integers(1, 8))):
            cv2.circle(a, (int(rng.integers(0, w)), int(rng.integers(0, h))), int(rng.integers(1, 25)), 0, -1)
    elif kind == 6:                                        # marker-like disks on a jittered grid, some clipped by the frame
        for y in range(-10, h + 30, 52):
            for x in range(-10, w + 30, 52):
                cv2.circle(a, (x + int(rng.integers(-4, 5)), y + int(rng.integers(-4, 5))), int(rng.integers(7, 14)), 1, -1)
    elif kind == 7:                                        # salt noise over blobs (many run transitions per window)
        field = cv2.GaussianBlur(rng.normal(0, 1, (h, w)).astype(np.float32), (0, 0), 8.0)
        a = ((field > 0.2 * field.std()) ^ (rng.random((h, w)) < 0.15)).astype(np.uint8)
    elif kind == 8:                                        # checkerboards and fine stripes: worst case for the run-based pass
        p = int(rng.integers(1, 5))
        yy, xx = np.mgrid[:h, :w]
        a = ((((xx // p) + (yy // p)) & 1) if rng.random() < 0.5 else ((xx // p) & 1)).astype(np.uint8)
        a[int(rng.integers(0, h)):, :] = 0
    else:                                                  # one half set, ragged diagonal edge
        yy, xx = np.mgrid[:h, :w]
        a = ((xx * float(rng.uniform(0.2, 2.0)) + yy + rng.integers(-3, 4, (h, w))) > (h + w) / 2).astype(np.uint8)
    return a * 255


@pytest.mark.parametrize("h,w,seed0", [(520, 600, 0), (520, 600, 30), (563, 645, 60), (300, 357, 100), (301, 450, 130)])
def test_ncc_random_area_masks(h, w, seed0):
    """>= 100 masks in total over the five parametrisations (5 x 30 = 150)."""
    from oracle import exact
    n = 30
    masks = np.stack([ncc_stress_masks(h, w, seed0 + i) for i in range(n)])
    c = port.branch_constants(h)
    tmpl = port.gaussian_template(c["tmpl"], c["tmpl_sigma"])
    with pipeline.MarkerPipeline(h, w, 1, max_batch=n, max_markers=64, max_refs=1) as pipe:
        got, rechecks = pipe.ncc_mask(torch_cuda(masks))
        got = got.cpu().numpy(); rechecks = rechecks.cpu().numpy()
    bad_total, explained = 0, 0
    for i in range(n):
        want = (port.ncc_same(tmpl, masks[i]) > 0.1).astype(np.uint8)
        bad = np.argwhere(got[i] != want)
        if len(bad):                                       # only a pixel within 1e-9 of the threshold may differ (FFT rounding)
            ncc64, _ = exact.ncc_closed_form(masks[i], c["tmpl"], c["tmpl_sigma"])
            near = np.abs(ncc64[bad[:, 0], bad[:, 1]] - 0.1) < 1e-9
            explained += int(near.sum())
            bad_total += int((~near).sum())
    print(f"ncc stress {h}x{w} seeds {seed0}..{seed0 + n - 1}: float64 re-decisions per mask min/median/max = "
          f"{int(rechecks.min())}/{int(np.median(rechecks))}/{int(rechecks.max())}, FFT-threshold ties: {explained}")
    assert bad_total == 0 and explained == 0
    assert rechecks.max() < 16384                          # the recheck list never overflowed


@pytest.mark.parametrize("h,w", [(520, 600), (300, 357)])
def test_ncc_kernel_variants_agree(monkeypatch, h, w):
    """The three matched-filter kernels (VBS_NCC_VARIANT = 0: round 1, 1: thread per column, 2: thread per column with
    horizontal / vertical warp roles, the default) decide every pixel of the stress masks alike and queue the same number
    of float64 re-decisions."""
    n = 24
    masks = torch_cuda(np.stack([ncc_stress_masks(h, w, 500 + i) for i in range(n)]))
    res = {}
    for v in ("2", "1", "0"):
        monkeypatch.setenv("VBS_NCC_VARIANT", v)
        with pipeline.MarkerPipeline(h, w, 1, max_batch=n, max_markers=64, max_refs=1) as pipe:
            got, rechecks = pipe.ncc_mask(masks)
            res[v] = (got.cpu().numpy(), rechecks.cpu().numpy())
    for v in ("1", "0"):
        assert np.array_equal(res["2"][0], res[v][0]), v
    assert np.array_equal(res["2"][1], res["1"][1])        # same float32 filter in both thread-per-column kernels


# ---------------------------------------------------------------------------------------------
# 15. 3D position with a camera whose f_avg^2 is NOT exactly representable in float32: the reference squares
#     an np.float32 scalar (R3:219), so that term is rounded to float32 before it meets the float64 radius
# ---------------------------------------------------------------------------------------------
def test_position_3d_float32_square_of_f_avg():
    K = np.array([[900.3, 0, 640.2], [0, 898.3, 360.4], [0, 0, 1]], dtype=np.float32)
    _, D, R, T = synth.synthetic_camera()
    favg = np.float32((K[0, 0] + K[1, 1]) / np.float32(2))
    assert float(np.float32(favg * favg)) != float(favg) * float(favg)        # the rounding really happens
    cam = port.Camera(K, D, R, T)
    rng = np.random.default_rng(4)
    uvd = np.concatenate([rng.uniform([0, 0, 8], [1280, 720, 40], (400, 3)), [[float(K[0, 2]), float(K[1, 2]), 20.0]]])   # last: the principal point
    with pipeline.MarkerPipeline(64, 64, 1, max_batch=1, max_markers=8, max_refs=1) as pipe:
        pipe.set_camera(K, D, R, T, 2.0, 5.0, 50.0, warmup_frames=0)
        P, ok = pipe.position_3d(uvd)
    worst = 0.0
    for (u, v, d), p, o in zip(uvd, P, ok):
        want = port.position_3d(cam, u, v, d)
        assert (want is not None) == bool(o)
        if want is not None:
            worst = max(worst, float(np.abs(p - want).max()))
    assert not ok[-1] and worst <= 1e-12, worst            # all-float64 arithmetic would be off by ~1.5e-6 mm


# ---------------------------------------------------------------------------------------------
# 16. OPT-IN tensor-core blur (k_blur_tc.cu: tcgen05.mma kind::i8, TMEM, TMA): area_mask must equal
#     cv2's fixed-point GaussianBlur -> uint8 DoG -> inRange (MD:114-129) and the default integer-dot-product kernel,
#     bit for bit - marker frames, full-range noise (saturated 65280 sums, the sign-flip edge), both height branches,
#     heights that are not multiples of 64 / 128, widths that are not multiples of 64 (padded pitch), image edges.
# ---------------------------------------------------------------------------------------------
def cv2_area_mask(gray):
    import cv2
    c = port.branch_constants(gray.shape[0])
    small = cv2.GaussianBlur(gray, (c["k_small"], c["k_small"]), c["s_small"])
    large = cv2.GaussianBlur(gray, (c["k_large"], c["k_large"]), c["s_large"])
    return cv2.inRange(large - small + 15, c["lo"], c["hi"])


def tc_blur_cases():
    rng = np.random.default_rng(77)
    cases = {"markers_560x640": synth.workload_frames("small_6x8", 3, seed0=41)}
    h, w = 563, 645                                              # H, W not multiples of 64 / 128
    cases["ragged_563x645"] = np.stack([synth.render_frame(h, w, synth.grid_layout(h, w, 6, 8, 60.0), 11.0, seed=s) for s in (1, 2)])
    noise = rng.integers(0, 256, (4, 520, 704), dtype=np.uint8)
    noise[1] = 255                                               # saturated: horizontal sums of 65280, high byte 255
    noise[2, :, :352] = 0; noise[2, :, 352:] = 255               # a step edge through strip and block boundaries
    noise[3, ::2] = 255                                          # row stripes: vertical pass at full swing
    cases["noise_520x704"] = noise
    small = rng.integers(0, 256, (3, 300, 368), dtype=np.uint8)  # <= 480 branch: 21 / 35 taps
    small[2] = synth.workload_frames("tiny_4x5", 1, seed0=3)[0][:, :368] if synth.WORKLOADS["tiny_4x5"][1] >= 368 else small[2]
    cases["small_branch_300x368"] = small
    blobs = np.clip(rng.normal(128, 70, (2, 700, 130)), 0, 255).astype(np.uint8)      # narrower than the halo on both sides
    cases["narrow_700x130"] = blobs
    return cases


@pytest.mark.parametrize("name", ["markers_560x640", "ragged_563x645", "noise_520x704", "small_branch_300x368", "narrow_700x130"])
def test_tensor_core_blur_equals_cv2_and_default_kernel(name):
    import ctypes
    import torch
    frames = tc_blur_cases()[name]
    B, H, W = frames.shape
    Wp = (W + 15) // 16 * 16                                      # the TMA unit needs 16-byte pitches: pad the rows, pass the pitch
    buf = torch.zeros((B, H, Wp), dtype=torch.uint8, device="cuda")
    buf[:, :, :W] = torch_cuda(frames)
    want = np.stack([cv2_area_mask(f) for f in frames])
    got = {}
    for tc in (0, 1):
        with pipeline.MarkerPipeline(H, W, 1, max_batch=B, max_markers=4096, max_refs=1) as pipe:
            pipe.set_blur_tc(bool(tc))
            pipe._follow_torch_stream()
            capi.check(pipe._ctx, capi.lib.vbs_find_markers(pipe._ctx, buf.data_ptr(), B, H * Wp, Wp))
            pipe.sync()
            got[tc] = pipe.debug_stage(capi.STAGE_AREA_MASK, B).cpu().numpy()
            assert (pipe.tc_launches >= 1) == bool(tc)
    for tc in (0, 1):
        bad = np.argwhere(got[tc] != want)
        assert len(bad) == 0, (name, "tc" if tc else "idp", len(bad), bad[:8].tolist(), bad[-4:].tolist())


@pytest.mark.parametrize("h,w,n,persist", [(300, 368, 130, "1"), (100, 256, 200, "1"), (563, 645, 40, "1"), (300, 368, 130, "0")])
def test_tensor_core_blur_persistent_ctas_many_items(monkeypatch, h, w, n, persist):
    """More (frame, strip) items than SMs: every persistent CTA runs several items back to back (running stage / ring /
    block counters, operator-matrix reload when its strip changes; 100 rows = fewer chunk slots per item than the ring
    has).  All frames equal the default kernel, three of them cv2.  persist = "0": one CTA per item (VBS_TC_PERSIST=0)."""
    import torch
    rng = np.random.default_rng(h * 1000 + w)
    frames = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
    frames[::3] = np.clip(rng.normal(120, 60, (len(frames[::3]), h, w)), 0, 255).astype(np.uint8)
    frames[1::5, : h // 2] = 255
    Wp = (w + 15) // 16 * 16
    buf = torch.zeros((n, h, Wp), dtype=torch.uint8, device="cuda")
    buf[:, :, :w] = torch_cuda(frames)
    monkeypatch.setenv("VBS_TC_PERSIST", persist)
    got = {}
    for tc in (0, 1):
        with pipeline.MarkerPipeline(h, w, 1, max_batch=n, max_markers=4096, max_refs=1) as pipe:
            pipe.set_blur_tc(bool(tc))
            pipe._follow_torch_stream()
            capi.check(pipe._ctx, capi.lib.vbs_find_markers(pipe._ctx, buf.data_ptr(), n, h * Wp, Wp))
            pipe.sync()
            got[tc] = pipe.debug_stage(capi.STAGE_AREA_MASK, n).cpu().numpy()
            assert (pipe.tc_launches >= 1) == bool(tc)
    bad = np.argwhere(got[0] != got[1])
    assert len(bad) == 0, (len(bad), bad[:8].tolist(), bad[-4:].tolist())
    for f in (0, n // 2, n - 1):
        assert np.array_equal(got[1][f], cv2_area_mask(frames[f])), f


def test_tensor_core_blur_whole_pipeline_1080p(monkeypatch):
    """VBS_BLUR_TC=1 end to end at the headline geometry: every table equals the default path's."""
    name, U = "1080p_20x20", 6
    H, W, rows, cols, _, _ = synth.WORKLOADS[name]
    uniq = synth.workload_frames(name, U, seed0=0)
    x = torch_cuda(uniq)

    def run():
        with pipeline.MarkerPipeline(H, W, 1, max_batch=U, max_markers=1024, max_refs=1) as pipe:
            res = pipe.process(x, 0); pipe.sync()
            return res.to_host(), pipe.debug_stage(capi.STAGE_AREA_MASK, U).cpu().numpy(), pipe.tc_launches

    a, a_area, a_tc = run()
    monkeypatch.setenv("VBS_BLUR_TC", "1")
    b, b_area, b_tc = run()
    assert a_tc == 0 and b_tc >= 1
    assert np.array_equal(a_area, b_area)
    assert np.array_equal(a_area[0], cv2_area_mask(uniq[0])) and np.array_equal(b_area[U - 1], cv2_area_mask(uniq[U - 1]))
    assert np.array_equal(a.n_markers, b.n_markers) and (a.n_markers == rows * cols).all()
    assert np.array_equal(a.marker_xy[:, : rows * cols], b.marker_xy[:, : rows * cols])
    assert np.array_equal(a.marker_axes[:, : rows * cols], b.marker_axes[:, : rows * cols])


# ---------------------------------------------------------------------------------------------
# 16b. column-sum form of the 101-tap vertical pass (k_blur.cu: blur_area_cs_kernel, the default for gray frames of
#      the > 480 branch) against cv2 and the all-dot-product kernel (VBS_BLUR_VARIANT=0): the tensor-core blur's cases
#      where that branch applies, TMA and generic loader, padded and unaligned pitches, more items than resident
#      CTAs (whole-height CTAs + row segments)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["markers_560x640", "ragged_563x645", "noise_520x704", "narrow_700x130"])
def test_column_sum_blur_equals_cv2(monkeypatch, name):
    import torch
    frames = tc_blur_cases()[name]
    B, H, W = frames.shape
    want = np.stack([cv2_area_mask(f) for f in frames])
    monkeypatch.delenv("VBS_BLUR_VARIANT", raising=False)
    for pitch, no_tma in (((W + 15) // 16 * 16, "0"), (W + 3, "0"), ((W + 15) // 16 * 16, "1")):     # TMA where a tile is interior / generic loader
        monkeypatch.setenv("VBS_NO_TMA", no_tma)
        buf = torch.zeros((B, H, pitch), dtype=torch.uint8, device="cuda")
        buf[:, :, :W] = torch_cuda(frames)
        with pipeline.MarkerPipeline(H, W, 1, max_batch=B, max_markers=4096, max_refs=1) as pipe:
            pipe._follow_torch_stream()
            capi.check(pipe._ctx, capi.lib.vbs_find_markers(pipe._ctx, buf.data_ptr(), B, H * pitch, pitch))
            pipe.sync()
            got = pipe.debug_stage(capi.STAGE_AREA_MASK, B).cpu().numpy()
        bad = np.argwhere(got != want)
        assert len(bad) == 0, (name, pitch, no_tma, len(bad), bad[:8].tolist(), bad[-4:].tolist())


@pytest.mark.parametrize("h,w,n", [(563, 645, 160), (1080, 1920, 64)])
def test_column_sum_blur_many_items_equals_dot_product_kernel(monkeypatch, h, w, n):
    """Grids of several waves: whole-height CTAs and row segments (VbsSegPlan with two resident CTAs per SM), segment
    starts inside the image (the column sums start from zero at row ys - 50), frames of full-range noise and of markers."""
    import torch
    rng = np.random.default_rng(h + w + n)
    frames = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
    frames[::3] = np.clip(rng.normal(120, 60, (len(frames[::3]), h, w)), 0, 255).astype(np.uint8)
    frames[1::5, : h // 2] = 255
    if (h, w) == (1080, 1920):
        frames[2::4] = synth.workload_frames("1080p_20x20", len(frames[2::4]), seed0=5)
    Wp = (w + 15) // 16 * 16
    buf = torch.zeros((n, h, Wp), dtype=torch.uint8, device="cuda")
    buf[:, :, :w] = torch_cuda(frames)
    got = {}
    for v in ("0", "1"):
        monkeypatch.setenv("VBS_BLUR_VARIANT", v)
        with pipeline.MarkerPipeline(h, w, 1, max_batch=n, max_markers=4096, max_refs=1) as pipe:
            pipe._follow_torch_stream()
            capi.check(pipe._ctx, capi.lib.vbs_find_markers(pipe._ctx, buf.data_ptr(), n, h * Wp, Wp))
            pipe.sync()
            got[v] = pipe.debug_stage(capi.STAGE_AREA_MASK, n).cpu().numpy()
    bad = np.argwhere(got["0"] != got["1"])
    assert len(bad) == 0, (len(bad), bad[:8].tolist(), bad[-4:].tolist())
    for f in (0, 2, n - 1):
        assert np.array_equal(got["1"][f], cv2_area_mask(frames[f])), f


# ---------------------------------------------------------------------------------------------
# 17. every ellipse the GPU fits (VBS_STAGE_ELLIPSES: one per external contour of the opened mask, cv2.findContours
#     order) against cv2.fitEllipse - on the shapes that stress its conditioning (thin slivers, diagonal staircases, L / T /
#     plus shapes, rings, frame contact) and on random opened blobs, incl. the 5-vertex contours cv2 routes through
#     fitEllipseDirect.  See tests/test_oracle_cpu.py::test_fit_ellipse_retry_is_unreachable_and_five_point_contours.
# ---------------------------------------------------------------------------------------------
def test_ellipse_stage_equals_cv2_on_degenerate_and_random_shapes():
    import cv2
    from test_oracle_cpu import degenerate_shape_masks, random_opened_masks
    H = W = 144
    shapes = degenerate_shape_masks() + list(random_opened_masks(150, 23))
    masks = np.zeros((len(shapes), H, W), np.uint8)
    for i, m in enumerate(shapes):
        masks[i, :m.shape[0], :m.shape[1]] = m
    B = len(masks)
    with pipeline.MarkerPipeline(H, W, 1, max_batch=B, max_markers=256, max_refs=1) as pipe:
        pipe.marker_center(torch_cuda(np.zeros_like(masks)), torch_cuda(masks))
        pipe.sync()
        cells = pipe.debug_stage(capi.STAGE_ELLIPSES, B).cpu().numpy()
        ncont = pipe.debug_stage(capi.STAGE_NCONTOURS, B).cpu().numpy()
        opened = pipe.debug_stage(capi.STAGE_OPENED, B).cpu().numpy()
    checked = skipped = 0
    for f in range(B):
        assert np.array_equal(opened[f], cv2.morphologyEx(masks[f], cv2.MORPH_OPEN, np.ones((5, 5), np.uint8)))
        contours, _ = cv2.findContours(opened[f], cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        assert ncont[f] >= len(contours)                        # the GPU also numbers blobs nested in holes (no ellipse, like RETR_EXTERNAL)
        want = []
        for cnt in contours:
            if len(cnt) < 5:                                    # MD:204
                continue
            refs = [cv2.fitEllipse(cnt) for _ in range(4 if len(cnt) == 5 else 1)]
            (cx, cy), (ww, hh), ang = refs[0]
            unstable = any(r != refs[0] for r in refs)          # cv2 re-fits randomly perturbed points: it does not agree with itself
            if unstable:
                want.append(None)
            elif min(ww, hh) >= 5:                              # MD:219
                want.append((cx, cy, max(ww, hh), min(ww, hh), ang if ww > hh else ang + 90.0))
        got = [tuple(c[:5]) for c in cells[f, : int(ncont[f])] if c[5] != 0.0]
        if any(w is None for w in want):
            skipped += 1
            continue
        assert len(got) == len(want), (f, len(got), len(want))
        for g, w in zip(got, want):
            assert pu.f32_ulps(np.array(g[:4]), np.array(w[:4])).max() <= 2.0, (f, g, w)
            assert pu.angle_ulps(g[4], w[4]) <= pu.ANGLE_TOL_ULP, (f, g, w)
            checked += 1
    assert checked > 400 and skipped < 10
