#!/usr/bin/env python
"""Times K1 alone (vbs_find_markers minus the NCC is not separable, so: per-stage events of vbs_process_device) for the
default integer-dot-product kernel and the opt-in tensor-core kernel; small enough to run under ncu.
    python tools/tc_probe.py [batch] [reps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vbs_b200  # noqa: F401
from vbs_b200 import pipeline, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
H, W = 1080, 1920
u = synth.workload_frames("1080p_20x20", 4, seed0=0)
x = torch.from_numpy(np.tile(u, (B // 4, 1, 1))).cuda()
REFS = os.environ.get("TC_PROBE_REFS", "0") == "1"
SAME = os.environ.get("TC_PROBE_SAME_CTX", "0") == "1"
shared = pipeline.MarkerPipeline(H, W, 1, max_batch=B, max_markers=1024, max_refs=512) if SAME else None
for tc in (0, 1):
    with (shared if SAME else pipeline.MarkerPipeline(H, W, 1, max_batch=B, max_markers=1024, max_refs=512)) as p:
        if SAME:
            p.__exit__ = lambda *a: None
        if REFS and (tc == 0 or not SAME):
            from vbs_b200 import reference_state
            r0 = p.process(x[:1], 0); p.sync(); h0 = r0.to_host()
            keys, xy = reference_state.grid_ids(h0.marker_xy[0, : int(h0.n_markers[0])], 20)
            p.set_reference([k[0] for k in keys], [k[1] for k in keys], xy[:, 0], xy[:, 1], 20.0)
            p.set_camera(*synth.synthetic_camera(), 2.0, 5.0, 50.0, warmup_frames=0)
        p.set_blur_tc(bool(tc))
        outs = p.alloc_outputs(B, True)
        for _ in range(2):
            p.process(x, 0, out=outs)
        p.sync()
        p.set_profiling(True)
        for _ in range(reps):
            p.process(x, 0, out=outs)
        p.sync()
        ms, calls = p.stage_ms()
        print("tc" if tc else "idp", "batch", B, {k: round(v / calls, 4) for k, v in ms.items()}, "tc_launches", p.tc_launches, flush=True)
