#!/usr/bin/env python
"""Times the blur stage (per-stage events of vbs_process_device) for the given VBS_BLUR_VARIANT values (0 = all-dot-product
kernel, 1 = column-sum kernel, the default) and checks that the area masks of all of them are equal.
    python tools/blur_probe.py [batch] [reps] [variants, e.g. 01]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vbs_b200  # noqa: F401
from vbs_b200 import capi, pipeline, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
variants = sys.argv[3] if len(sys.argv) > 3 else "01"
H, W = 1080, 1920
u = synth.workload_frames("1080p_20x20", 8, seed0=0)
x = torch.from_numpy(np.tile(u, (B // 8, 1, 1))).cuda()
ref = None
for v in variants:
    os.environ["VBS_BLUR_VARIANT"] = v
    with pipeline.MarkerPipeline(H, W, 1, max_batch=B, max_markers=1024, max_refs=512) as p:
        outs = p.alloc_outputs(B, True)
        for _ in range(2):
            p.process(x, 0, out=outs)
        p.sync()
        area = p.debug_stage(capi.STAGE_AREA_MASK, 8).cpu().numpy()
        if ref is None:
            ref = area
        same = bool(np.array_equal(ref, area))
        p.set_profiling(True)
        for _ in range(reps):
            p.process(x, 0, out=outs)
        p.sync()
        ms, calls = p.stage_ms()
        tot = sum(ms.values()) / calls
        print("variant", v, "batch", B, "equal_to_first", same, {k: round(val / calls, 4) for k, val in ms.items()}, "sum", round(tot, 3), flush=True)
