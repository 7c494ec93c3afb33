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
for tc in (0, 1):
    with pipeline.MarkerPipeline(H, W, 1, max_batch=B, max_markers=1024, max_refs=1) as p:
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
