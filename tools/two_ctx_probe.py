#!/usr/bin/env python
"""Experiment: two contexts fed alternately on two streams - how much would cross-batch overlap
(tail kernels of batch i beside the blur of batch i+1) buy over one context run back to back?"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vbs_b200  # noqa: F401
from vbs_b200 import pipeline, reference_state, synth

H, W, rows, cols, _, _ = synth.WORKLOADS["1080p_20x20"]
B = 256
uniq = synth.workload_frames("1080p_20x20", 16, seed0=0)
frames = torch.from_numpy(np.ascontiguousarray(np.tile(uniq, (16, 1, 1))[:B])).cuda()
K = synth.synthetic_camera()

def make():
    p = pipeline.MarkerPipeline(H, W, 1, max_batch=B, max_markers=1024, max_refs=400)
    r0 = p.process(frames[:1], 0); p.sync(); h0 = r0.to_host()
    keys, xy = reference_state.grid_ids(h0.marker_xy[0, : int(h0.n_markers[0])], cols)
    p.set_reference([k[0] for k in keys], [k[1] for k in keys], xy[:, 0], xy[:, 1], 20.0)
    p.set_camera(*K, 2.0, 5.0, 50.0, warmup_frames=0)
    return p, p.alloc_outputs(B, True)

def run(pipes, streams, steps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        p, outs = pipes[i % len(pipes)]
        with torch.cuda.stream(streams[i % len(streams)]):
            p.process(frames, 0, outs)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3

a, b = make(), make()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for n, (pp, ss) in {"one context, one stream": ([a], [s1]), "two contexts, two streams": ([a, b], [s1, s2])}.items():
    run(pp, ss, 6)
    print(f"{n}: {run(pp, ss, 40):.3f} ms per 256-frame batch")
