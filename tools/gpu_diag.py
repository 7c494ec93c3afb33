#!/usr/bin/env python
"""Stage-by-stage parity report of the CUDA path against the oracle (prints everything; never
stops at the first mismatch).  Run on a GPU box:  python tools/gpu_diag.py [case ...]"""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

import parity_util as pu
from vbs_b200 import pipeline, synth
from oracle import port


def run_case(name, n_frames=3, bgr=False, max_batch=4):
    print(f"=== case {name} bgr={bgr} frames={n_frames}", flush=True)
    h, w, rows, cols, pitch, radius = synth.WORKLOADS[name]
    frames = synth.workload_frames(name, n_frames, seed0=0)
    src = np.repeat(frames[..., None], 3, axis=3) if bgr else frames
    if bgr:   # make the channels differ so the gray conversion is exercised
        src = src.copy(); src[..., 0] = np.clip(src[..., 0].astype(int) + 7, 0, 255); src[..., 2] = np.clip(src[..., 2].astype(int) - 9, 0, 255)
    t0 = time.time()
    ora0 = pu.oracle_frames(src[:1])
    keys, xy = pu.grid_reference(ora0[0]["markers"], cols)
    oracle = pu.oracle_frames(src, keys, xy, 20.0)
    print(f"  oracle: {time.time() - t0:.2f}s, markers/frame {[len(o['markers']) for o in oracle]}", flush=True)
    pipe = pipeline.MarkerPipeline(h, w, channels=3 if bgr else 1, max_batch=max_batch, max_markers=max(64, 2 * rows * cols), max_refs=max(64, rows * cols))
    pipe.set_reference([k[0] for k in keys], [k[1] for k in keys], xy[:, 0], xy[:, 1], 20.0)
    K, D, R, T = synth.synthetic_camera()
    if name != "1080p_20x20" and name != "4k_40x72":
        K = K.copy(); K[0, 2] = w / 2 + 3.1; K[1, 2] = h / 2 - 2.3
    pipe.set_camera(K, D, R, T, 2.0, 5.0, 50.0, warmup_frames=0)
    cam = port.Camera(K, D, R, T)
    rows3d, pos = pu.oracle_3d(oracle, cam, warmup=0)
    ref_xyz = np.stack([(xy[:, 0] - w / 2) / 11.0, (xy[:, 1] - h / 2) / 11.0, np.zeros(len(xy))], axis=1)
    start = np.array([pos.get((0, k[0], k[1]), np.zeros(3)) if pos.get((0, k[0], k[1])) is not None else np.zeros(3) for k in keys])
    dvert = np.zeros_like(start)
    pipe.set_plane(ref_xyz, start, dvert)
    planes = pu.oracle_plane(pos, keys, range(n_frames), ref_xyz, start, dvert)
    dev = torch.from_numpy(src).cuda()
    t0 = time.time()
    res = pipe.process(dev, 0)
    try:
        pipe.sync()
    except Exception as e:
        print("  sync reported:", type(e).__name__, e, flush=True)
    print(f"  gpu first call {time.time() - t0:.3f}s", flush=True)
    rep = pu.compare_detection(pipe, src, res, oracle)
    rep.update(pu.compare_rows(res, oracle, keys))
    rep.update(pu.compare_3d(res, keys, rows3d, pos))
    rep.update(pu.compare_plane(res, planes))
    print(pu.format_report(rep), flush=True)
    # host entry point must give the same bytes
    pipe.reset_sequence()
    hres = pipe.process(np.ascontiguousarray(src), 0)
    dres = res.to_host()
    same = all(np.array_equal(getattr(hres, k), getattr(dres, k), equal_nan=True) for k in ("n_markers", "marker_xy", "marker_axes", "row_det", "pos3d", "plane"))
    print(f"  [{'ok' if same else 'FAIL'}] host entry point == device entry point", flush=True)
    ok = all(v[0] for v in rep.values()) and same
    pipe.close()
    return ok


if __name__ == "__main__":
    cases = sys.argv[1:] or ["tiny_4x5", "small_6x8", "small_6x8:bgr", "1080p_20x20"]
    allok = True
    for c in cases:
        try:
            nm, _, opt = c.partition(":")
            ok = run_case(nm, n_frames=2 if nm.startswith("1080p") else 3, bgr=(opt == "bgr"))
        except Exception:
            traceback.print_exc()
            ok = False
        allok &= ok
        print(f"=== {c}: {'PASS' if ok else 'FAIL'}", flush=True)
    sys.exit(0 if allok else 1)
