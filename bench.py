#!/usr/bin/env python
"""bench.py - frames/s of the per-frame marker pipeline (tracking -> 3D displacement -> plane tilt).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the whole hot path over one batch of synthetic frames per GPU
(BASELINE.json configs[1]: 1080p grayscale, 20x20 marker array, batch 256).  Frames are
independent, so ranks shard them with no data-path collective; the only NCCL traffic is the
gather of the per-frame records to rank 0 (inside the timed region).  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

METRIC = "frames/sec (centroids+IDs+3D field)"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="1080p_20x20", help="1080p_20x20 (BASELINE config 2, the headline), 4k_40x72 (config 4), "
                    "stream64k (config 5: a 65 536-frame 1080p sequence in contiguous shards, last-seen exchange, NCCL gather of records + tilt)")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--frames", type=int, default=65536, help="stream64k: frames of the whole job")
    ap.add_argument("--unique", type=int, default=32, help="distinct synthetic frames tiled to the batch (the 32 frames of seed 0..31 are the "
                    "ones tests/test_gpu_parity.py checks against the oracle, stage by stage)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU-baseline budget (bounded sample)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--overlap", action="store_true", help="two-stream chunk pipelining for device-resident batches (default off)")
    ap.add_argument("--undistort", action="store_true", help="side measurement: lens correction (MD:93-109) active on the CUDA arm; not the headline workload")
    ap.add_argument("--no-tc", action="store_true", help="skip the opt-in tensor-core blur arm (reported beside the default arm as `tc_blur`)")
    ap.add_argument("--no-balance", action="store_true", help="N > 1: skip the e2e mode that splits the global batch by measured host-to-device link rates")
    ap.add_argument("--host-chunk", type=int, default=0, help="frames per chunk of the host path copy/compute overlap (0 = library default)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_setup(name, unique):
    from vbs_b200 import synth
    h, w, rows, cols, pitch, radius = synth.WORKLOADS[name]
    frames = synth.workload_frames(name, unique, seed0=0)
    return frames, (h, w, rows, cols)


# ------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------
class Clocks:
    """One `nvidia-smi -lms` process streaming clock / throttle samples; samples taken between
    start() and stop() (the timed region) are summarised."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.t0, self.t1 = [], None, None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            time.sleep(0.5)                      # let the sampler come up before the timed region
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.lines.append((time.time(), line.strip()))

    def __enter__(self):
        self.t0 = time.time()
        return self

    def __exit__(self, *a):
        self.t1 = time.time()
        if self.p is not None:
            time.sleep(0.05)
            self.p.terminate()
            try:
                self.p.wait(timeout=3)
            except Exception:
                self.p.kill()

    def summary(self):
        rows = [l.split(",") for t, l in self.lines if self.t0 is not None and self.t0 <= t <= self.t1 + 0.05]
        rows = [[c.strip() for c in r] for r in rows if len(r) >= 7]
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# CPU legs (oracle port = the reference's own OpenCV/SciPy/NumPy path restated; oracle/port.py)
# ------------------------------------------------------------------------------------------
def _cpu_frame(args):
    frame, keys, xy, frameno, threads = args
    import cv2
    from oracle import port
    if threads:
        cv2.setNumThreads(threads)
    markers = port.find_markers_frame(frame)
    return port.track_rows(keys, xy, markers, frameno, 20.0)


def cpu_tail(rows_per_frame, cam_params, plane_params, keys):
    """3D + plane legs of the CPU path on the rows of a sample (R3:240-316, FD:138-162)."""
    from oracle import port
    cam = port.Camera(*cam_params)
    tab = {k: [] for k in ("frameno", "row", "col", "Cx", "Cy", "major_axis")}
    for rows in rows_per_frame:
        for r in rows:
            for k in tab:
                tab[k].append(r[k])
    tab = {k: np.asarray(v) for k, v in tab.items()}
    out = port.displacement_rows(cam, tab, warmup_frames=0)
    ref_xyz, start, dvert = plane_params
    index = {k: i for i, k in enumerate(keys)}
    for rows in rows_per_frame:
        if len(rows) < 3:
            continue
        uv = port.undistort_points(cam, np.array([[r["Cx"], r["Cy"]] for r in rows]))
        P, idx = [], []
        for r, (u, v) in zip(rows, uv):
            p = port.position_3d(cam, u, v, r["major_axis"])
            if p is not None:
                P.append(p); idx.append(index[(r["row"], r["col"])])
        if len(idx) >= 3:
            idx = np.array(idx)
            X, Y, Z = port.deviation_endpoints(ref_xyz[idx], np.array(P) - start[idx], dvert[idx])
            port.plane_tilt(X, Y, Z)
    return out


def cpu_single_process(frames, keys, xy, cam_params, plane_params, budget_s):
    """Variant A of SURVEY 8d: one process, OpenCV default threads.  Returns (fps, n_frames, threads)."""
    import cv2
    _cpu_frame((frames[0], keys, xy, 0, 0))                     # warm-up frame
    t0 = time.perf_counter()
    rows, n = [], 0
    while True:
        rows.append(_cpu_frame((frames[n % len(frames)], keys, xy, n, 0)))
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 64:
            break
    cpu_tail(rows, cam_params, plane_params, keys)
    dt = time.perf_counter() - t0
    return n / dt, n, cv2.getNumThreads()


def _cpu_tail_job(args):
    rows, cam_params, plane_params, keys = args
    cpu_tail(rows, cam_params, plane_params, keys)
    return len(rows)


def reference_arm(args, frames, keys, xy, cam_params, plane_params):
    """--impl reference: every host core (spawn pool, one OpenCV thread each; fork deadlocks after cv2 ran).
    The 3D + plane tail of a step runs in the pool as well, one slice of the step's frames per worker (each
    slice starts its last-seen table empty, so it emits one displacement row per marker fewer than a
    sequential pass - 1/len(slice) of the tail's work, in the CPU arm's favour)."""
    import multiprocessing as mp
    ncpu = os.cpu_count() or 1
    per_step = 2 * ncpu
    ctx = mp.get_context("spawn")
    with ctx.Pool(ncpu) as pool:
        def step(s):
            jobs = [(frames[(s * per_step + i) % len(frames)], keys, xy, s * per_step + i, 1) for i in range(per_step)]
            rows = pool.map(_cpu_frame, jobs, chunksize=1)
            k = max(1, len(rows) // ncpu)
            pool.map(_cpu_tail_job, [(rows[i:i + k], cam_params, plane_params, keys) for i in range(0, len(rows), k)], chunksize=1)
        for s in range(max(args.warmup, 1)):
            step(s)
        t0 = time.perf_counter()
        for s in range(args.steps):
            step(s)
        dt = time.perf_counter() - t0
    fps = per_step * args.steps / dt
    return fps, dt, per_step, ncpu


def periodic_ok(t, period, start=0):
    """Frames repeat with `period` (unique frames tiled): record rows of frame f and f + period must be
    byte-identical (NaN-safe: compared as raw bytes).  Size-independent check over the whole batch / stream."""
    import torch
    if t is None or t.shape[0] <= start + period:
        return True
    a = t[start:-period].contiguous().view(torch.uint8)
    b = t[start + period:].contiguous().view(torch.uint8)
    return bool(torch.equal(a, b))


def setup_pipe(args, frames_u, H, W, rows, cols, local, B, warmup_frames=0):
    """Context + reference state (detections of frame 0, GPU path) + camera + plane baseline."""
    import torch
    from vbs_b200 import pipeline, reference_state, synth
    dev = torch.device("cuda", local)
    n_markers = rows * cols
    pipe = pipeline.MarkerPipeline(H, W, 1, max_batch=B, max_markers=max(512, 2 * n_markers), max_refs=max(64, n_markers), device=local)
    r0 = pipe.process(torch.from_numpy(frames_u[:1]).to(dev), 0)
    pipe.sync()
    h0 = r0.to_host()
    keys, xy = reference_state.grid_ids(h0.marker_xy[0, : int(h0.n_markers[0])], cols)
    pipe.set_reference([k[0] for k in keys], [k[1] for k in keys], xy[:, 0], xy[:, 1], 20.0)
    cam_params = synth.synthetic_camera()
    pipe.set_camera(*cam_params, 2.0, 5.0, 50.0, warmup_frames=0)
    r0 = pipe.process(torch.from_numpy(frames_u[:1]).to(dev), 0)
    pipe.sync()
    start = np.nan_to_num(r0.to_host().pos3d[0, :, :3])
    ref_xyz = np.stack([(xy[:, 0] - W / 2) / 11.0, (xy[:, 1] - H / 2) / 11.0, np.zeros(len(xy))], 1)
    if warmup_frames:
        pipe.set_camera(*cam_params, 2.0, 5.0, 50.0, warmup_frames=warmup_frames)
    pipe.set_plane(ref_xyz, start, np.zeros_like(start))
    return pipe, keys, xy, (ref_xyz, start, np.zeros_like(start))


def init_dist(world, local):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        # rank 0 prints exactly one line: keep NCCL's version banner (written to fd 1 when the communicator comes up)
        # and any debug output off stdout
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dev = torch.device("cuda", local)
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            warm = torch.zeros(8, device=dev)
            dist.gather(warm, [torch.empty_like(warm) for _ in range(world)] if dist.get_rank() == 0 else None, dst=0)
            torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1)
            os.close(saved)


def stream_main(args, rank, world, local):
    """BASELINE config 5: `--frames` 1080p frames in contiguous shards (one per GPU), processed batch by batch,
    last-seen state patched across shard boundaries, records + plane tilt gathered to rank 0 over NCCL - all
    inside the timed region.  One step = one pass over the whole stream."""
    import torch
    import torch.distributed as dist
    from vbs_b200 import sharding, streaming
    geom = "1080p_20x20"
    frames_u, (H, W, rows, cols) = workload_setup(geom, args.unique)
    init_dist(world, local)
    dev = torch.device("cuda", local)
    B, U, N = args.batch, args.unique, args.frames
    assert B % U == 0 and all((sharding.shard_bounds(N, r, world)[0] % B) == 0 for r in range(world)), "shards must start on a batch boundary"
    pipe, keys, xy, plane_params = setup_pipe(args, frames_u, H, W, rows, cols, local, B, warmup_frames=100)   # R3:22 default warm-up
    R = pipe.R
    frames_d = torch.from_numpy(np.ascontiguousarray(np.tile(frames_u, (B // U, 1, 1)))).to(dev)      # frame g of the stream = unique frame g % U

    def frames_of(lo, hi):
        return frames_d[lo % B: lo % B + (hi - lo)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.current_stream()
    warm_frames = min(N, 2 * B * world)
    wsink = streaming.make_sink(pipe, warm_frames, rank, world)
    for _ in range(max(args.warmup, 1)):               # short warm-up streams: kernels, NCCL gather and all_gather paths
        streaming.run_stream(pipe, frames_of, warm_frames, B, rank, world, sink=wsink)
    del wsink
    sink = streaming.make_sink(pipe, N, rank, world)   # rank 0: preallocated landing area of the gathered record blobs
    pipe.sync()
    pipe.set_profiling(True)
    l0 = pipe.kernel_launches
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk_sampler = Clocks(local)
    steps = max(1, min(args.steps, 3))
    barrier()
    with clk_sampler as clk:
        e0.record(stream)
        for _ in range(steps):
            got, rec = streaming.run_stream(pipe, frames_of, N, B, rank, world, sink=sink)
        e1.record(stream)
        barrier()
    ms = e0.elapsed_time(e1)
    pipe.sync()
    stage, calls = pipe.stage_ms()
    pipe.set_profiling(False)
    launches = pipe.kernel_launches - l0
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    fps = N * steps / (ms * 1e-3)
    if rank == 0:
        # size-independent checks over the WHOLE gathered stream (shard boundaries included): past the warm-up window
        # every record repeats with the period of the unique frames - also the displacement rows, whose first one per
        # shard comes from the last-seen exchange (vbs_fix_displacement) rather than from the shard's own table
        chk = {k: periodic_ok(got[k], U, start=100 + U + 1) for k in ("row_det", "row_cxy", "row_axes", "pos3d", "pos_flags", "plane", "n_markers")}
        flags = got["pos_flags"]
        rows3d = int(((flags & 4) != 0).sum().item())
        tilt_ok = int(torch.isfinite(got["plane"][100:, 3]).sum().item())
        found = (int(got["n_markers"].min().item()), int(got["n_markers"].max().item()))
        peak, peak_src = peaks()
        b_alg = H * W + 96 * rows * cols + 32
        blur_ms = stage["blur_dog_area"] / max(calls, 1)
        frames_per_launch = min(B, sharding.shard_bounds(N, 0, world)[1])
        ach = frames_per_launch * b_alg / (blur_ms * 1e-3) / 1e9
        rec_bytes = sum(int(v.numel() * v.element_size()) for v in got.values())
        line = {"metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 1),
                "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u8 (integer blur/labels) + f32/f64 (NCC, geometry)", "data": "synthetic",
                "config": {"workload": f"stream64k: {N} frames of {H}x{W} gray u8 ({rows}x{cols} markers, {U} unique oracle-checked frames tiled on the "
                                       f"device), contiguous shards of {N // world} frames per GPU, batches of {B}; tracking+IDs -> 3D displacement "
                                       f"(warm-up 100 frames, last-seen exchange across shards) -> plane tilt; one NCCL gather of the record blobs + tilt to rank 0, shard boundaries patched there, all in the timed region",
                           "batch_per_gpu": B, "frames": N, "l2_policy": "inputs larger than L2 (batch of frames = %.0f MB/GPU)" % (B * H * W / 1e6),
                           "parallelism": f"frame-sharded x{world}", "gathered_bytes": rec_bytes},
                "roofline": {"bound": "hbm", "kernel": "blur_area_cs_kernel<39,101> (gray -> DoG -> area mask)", "achieved": ach, "peak": peak, "unit": "GB/s",
                             "frac": ach / peak, "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": frames_per_launch * b_alg,
                             "kernel_ms": blur_ms, "note": "ALU-bound path, see DESIGN.md"},
                "stage_ms_per_batch": {k: v / max(calls, 1) for k, v in stage.items()}, "gpu_launches": int(launches), "clocks": clk.summary(),
                "markers_per_frame": found, "checks": {"periodic_records": chk, "displacement_rows": rows3d, "frames_with_tilt": tilt_ok,
                                                        "frames_gathered": int(flags.shape[0])}}
        ok = all(chk.values()) and rows3d > 0 and flags.shape[0] == N and found == (rows * cols, rows * cols)
        line["checks"]["ok"] = bool(ok)
        print(json.dumps(line))
        if not ok:
            sys.exit("stream check failed: " + json.dumps(line["checks"]))
    pipe.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    import vbs_b200  # noqa: F401
    from vbs_b200 import synth, reference_state

    if args.workload == "stream64k" and args.impl != "reference":
        return stream_main(args, rank, world, local)
    geom = "1080p_20x20" if args.workload == "stream64k" else args.workload
    frames_u, (H, W, rows, cols) = workload_setup(geom, args.unique)
    n_markers = rows * cols
    b_alg = H * W * 1 + 96 * n_markers + 32             # SURVEY 8d algorithmic bytes per frame (gray input)
    cam_params = synth.synthetic_camera()
    config = {"workload": f"{args.workload}: {H}x{W} gray u8, {rows}x{cols} markers, batch {args.batch}/GPU, "
                          f"{args.unique} unique synthetic frames tiled; tracking+IDs -> 3D displacement -> plane tilt",
              "batch_per_gpu": args.batch, "l2_policy": "inputs larger than L2 (batch of frames = %.0f MB/GPU)" % (args.batch * H * W / 1e6),
              "parallelism": f"frame-sharded x{world}",
              "schedule": "sequential stages" if os.environ.get("VBS_BRANCH_OVERLAP", "")[:1] == "0" else
                          "open-mask branch (open, blobs, contours, ellipse fits) on a second stream beside the NCC; stage_ms_per_step "
                          "attributes that time to ncc_mask"}

    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import port
        m0 = port.find_markers_frame(frames_u[0])
        keys, xy = reference_state.grid_ids(np.array([m["center"] for m in m0]), cols)
        cam = port.Camera(*cam_params)
        start = np.array([port.position_3d(cam, *port.undistort_points(cam, np.array([p]))[0], 22.0) for p in xy])
        plane_params = (np.stack([(xy[:, 0] - W / 2) / 11.0, (xy[:, 1] - H / 2) / 11.0, np.zeros(len(xy))], 1), start, np.zeros_like(start))
        fps, dt, per_step, ncpu = reference_arm(args, frames_u, keys, xy, cam_params, plane_params)
        line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8/f64 (OpenCV, SciPy, NumPy)", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": fps, "unit": UNIT, "cores": ncpu, "kind": "port",
                                 "sample": f"{per_step} frames per step x {args.steps} steps, spawn pool of {ncpu} processes, cv2.setNumThreads(1); "
                                           "detection, ID match, 3D rows and plane fit all inside the pool"},
                "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from vbs_b200 import sharding

    init_dist(world, local)
    dev = torch.device("cuda", local)
    B = args.batch
    pipe, keys, xy, plane_params = setup_pipe(args, frames_u, H, W, rows, cols, local, B)
    stream = torch.cuda.current_stream()
    if args.overlap:
        pipe.set_overlap(True)
    # MarkerPipeline.process() launches on torch's current stream, so the events below bracket the work
    if args.undistort:                  # a nearly distortion-free lens: the frames stay detectable, the remap cost is the same
        pipe.set_undistort([[0.9 * W, 0, W / 2 + 0.3], [0, 0.9 * W, H / 2 - 0.2], [0, 0, 1]], [-2e-3, 1e-4, 1e-5, -1e-5, 0.0])
        config["workload"] += "; WITH optional lens correction (side measurement)"

    reps = (B + len(frames_u) - 1) // len(frames_u)
    host_frames = np.ascontiguousarray(np.tile(frames_u, (reps, 1, 1))[:B])
    frames_d = torch.from_numpy(host_frames).to(dev)
    outs = pipe.alloc_outputs(B, True)
    REC = ("pos3d", "pos_flags", "row_det", "row_cxy", "row_axes", "plane", "plane_n", "n_markers")

    def gather(arrays):
        """NCCL gather of the per-frame records (3D field, flags, IDs, tracking rows, plane) to rank 0."""
        if world > 1:
            return sharding.gather_records({k: arrays[k] for k in REC}, rank, world)
        return arrays

    def step(s):
        pipe.reset_sequence() if s == 0 else None
        res = pipe.process(frames_d, frameno0=(rank * args.steps + s) * B, out=outs)
        gather(outs[0])
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(args.warmup):
        step(s)
    pipe.sync()
    pipe.set_profiling(True)
    l0 = pipe.kernel_launches
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk_sampler = Clocks(local)
    barrier()
    with clk_sampler as clk:
        e0.record(stream)
        for s in range(args.steps):
            res = step(s)
        e1.record(stream)
        barrier()
    ms = e0.elapsed_time(e1)
    pipe.sync()
    stage, calls = pipe.stage_ms()
    pipe.set_profiling(False)
    launches = pipe.kernel_launches - l0
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    fps = world * B * args.steps / (ms * 1e-3)

    # sanity inside the bench: every frame of the batch found the whole array, and the records of the tiled batch
    # repeat with the period of the unique frames (the unique frames themselves are checked against the oracle,
    # stage by stage, in tests/test_gpu_parity.py::test_full_1080p_batch256_properties - same seeds)
    hres = res.to_host()
    found = int(hres.n_markers.min()), int(hres.n_markers.max())
    U = len(frames_u)
    periodic = {k: periodic_ok(outs[0][k][:, :found[0]] if k.startswith("marker_") else outs[0][k], U, start=1 if k in ("pos3d", "pos_flags") else 0)
                for k in ("n_markers", "marker_xy", "marker_axes", "row_det", "row_cxy", "plane", "pos3d", "pos_flags")}

    line = None
    if rank == 0:
        peak, peak_src = peaks()
        blur_ms = stage["blur_dog_area"] / max(calls, 1)
        ach = B * b_alg / (blur_ms * 1e-3) / 1e9
        whole = B * b_alg / (ms / args.steps * 1e-3) / 1e9
        traffic = traffic_tc = None
        tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if os.path.exists(tp):
            t = json.load(open(tp)).get(args.workload)
            if t and t.get("batch") == B:
                traffic, traffic_tc = t.get("blur_area_cs_kernel"), t.get("blur_area_tc_kernel")
        line = {"metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8 (integer blur/labels) + f32/f64 (NCC, geometry)", "data": "synthetic", "config": config,
                "roofline": {"bound": "hbm", "kernel": "blur_area_cs_kernel<39,101> (gray -> DoG -> area mask)", "achieved": ach, "peak": peak,
                             "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": B * b_alg, "kernel_ms": blur_ms,
                             "whole_path_achieved": whole, "whole_path_frac": whole / peak,
                             "note": "ALU-bound path (integer dot products + FMA chains), see DESIGN.md; HBM fraction reported as specified"},
                "stage_ms_per_step": {k: v / max(calls, 1) for k, v in stage.items()},
                "gpu_launches": int(launches), "clocks": clk.summary(), "markers_per_frame": found,
                "checks": {"periodic_records": periodic, "ok": bool(all(periodic.values()) and found == (n_markers, n_markers))}}
        # the binding roofline of the dominant kernel: instruction issue.  blur_area_cs_kernel spreads
        # its work over the fma pipe (integer dot products) and the ALU pipe (column-sum adds); ncu: 4672 warp instructions
        # per 8-row step of a 128-px strip, 148 steps per 135 output steps of a whole-height CTA at 1080p
        csum = line["clocks"]
        if H > 480 and csum.get("sm_mhz"):
            winstr_per_px = 4672.0 / 1024.0 * (H / 8.0 + 13.0) / (H / 8.0)
            sms = torch.cuda.get_device_properties(local).multi_processor_count
            rate = winstr_per_px * B * H * W / (blur_ms * 1e-3 * sms * csum["sm_mhz"] * 1e6)
            line["roofline"]["alu"] = {"pipe": "instruction issue (IDP.4A/IDP.2A on the fma pipe, three-input adds on the ALU pipe)", "achieved": rate,
                                       "peak": 4.0, "unit": "warp-instr/clk/SM", "frac": rate / 4.0,
                                       "peak_source": "4 schedulers x 1 instruction per clock; ncu profiles/r02_ncu_full_blur_cs_batch64.txt: issue active 73 %, "
                                                      "fma pipe 72 % and ALU pipe 82 % of their half-rate peaks",
                                       "instr_per_pixel": winstr_per_px}

    # ---- opt-in arm: the same step with the two blurs on the tensor cores (vbs_set_blur_tc).  The default
    #      arm above is the headline (north_star: no tensor cores); this one is reported beside it.
    if not args.no_tc:
        keep = {k: outs[0][k].clone() for k in ("n_markers", "marker_xy", "marker_axes", "row_det", "row_cxy", "pos3d", "pos_flags", "plane")}
        pipe.set_blur_tc(True)
        t0c = pipe.tc_launches
        for s in range(max(args.warmup, 2)):
            step(s)
        pipe.sync()
        pipe.set_profiling(True)
        stage_before, calls_before = pipe.stage_ms()       # the per-stage accumulators run on from the default arm: difference them
        barrier()
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record(stream)
        for s in range(args.steps):
            step(s)
        t1e.record(stream)
        barrier()
        ms_tc = t0e.elapsed_time(t1e)
        pipe.sync()
        stage_tc, calls_tc = pipe.stage_ms()
        stage_tc = {k: v - stage_before[k] for k, v in stage_tc.items()}
        calls_tc -= calls_before
        pipe.set_profiling(False)
        used = pipe.tc_launches - t0c
        same = all(bool(torch.equal(keep[k].contiguous().view(torch.uint8), outs[0][k].contiguous().view(torch.uint8))) for k in keep)
        pipe.set_blur_tc(False)
        if world > 1:
            t = torch.tensor([ms_tc], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_tc = float(t.item())
        if rank == 0:
            bms = stage_tc["blur_dog_area"] / max(calls_tc, 1)
            ach_tc = B * b_alg / (bms * 1e-3) / 1e9
            line["tc_blur"] = {"value": world * B * args.steps / (ms_tc * 1e-3), "unit": UNIT, "ms_per_step": ms_tc / args.steps,
                               "kernel": "blur_area_tc_kernel (tcgen05.mma kind::i8 banded-Toeplitz GEMMs, TMEM accumulators, TMA operands)",
                               "kernel_ms": bms, "tc_launches": int(used), "records_equal_default_arm": bool(same),
                               "stage_ms_per_step": {k: v / max(calls_tc, 1) for k, v in stage_tc.items()},
                               "roofline": {"bound": "hbm", "achieved": ach_tc, "peak": peak, "unit": "GB/s", "frac": ach_tc / peak, "traffic": traffic_tc,
                                            "algorithmic_bytes_per_launch": B * b_alg, "kernel_ms": bms},
                               "note": "opt-in (VBS_BLUR_TC=1 / vbs_set_blur_tc); not the default because the north_star rules tensor cores out"}

    # ---- e2e: host frames -> records on the host (rank 0's host for N > 1) through the public API, all copies and the
    #      NCCL gather in the timed region
    if not args.no_e2e:
        pin = torch.from_numpy(host_frames).pin_memory()
        on_dev = world > 1                                  # N > 1: records land in device memory, are gathered over NCCL, then one D2H on rank 0
        houts = [pipe.alloc_outputs(B, on_dev, compact=True) for _ in range(2)]
        if on_dev and rank == 0:
            hgath = {k: torch.empty((world * v.shape[0],) + tuple(v.shape[1:]), dtype=v.dtype).pin_memory() for k, v in houts[0][0].items() if k in REC}
        pipe.reset_sequence()

        def land(s):
            """batch s is complete on this rank: gather it to rank 0 and bring it to rank 0's host"""
            if not on_dev:
                return
            g = gather(houts[s & 1][0])              # stream-ordered: the host does not wait, the H2D copies of the next batches keep flowing
            if rank == 0:
                for k, v in g.items():
                    hgath[k].copy_(v, non_blocking=True)

        # streaming use of the public host API: submit batch s+1 (its H2D copy starts at once) before
        # waiting for batch s; every step still moves its own frames host->device and results device->host
        def run(nsteps, s0):
            pipe.submit_host_ptr(pin.data_ptr(), B, H * W, W, s0 * B, houts[s0 & 1])
            for s in range(s0 + 1, s0 + nsteps):
                pipe.submit_host_ptr(pin.data_ptr(), B, H * W, W, s * B, houts[s & 1])
                pipe.wait_host()
                land(s - 1)
            pipe.wait_host()
            land(s0 + nsteps - 1)
            torch.cuda.current_stream().synchronize()          # the last records are on rank 0's host

        def timed(chunk):
            pipe.set_host_chunk(chunk)
            run(2, 0)
            barrier()
            t0 = time.perf_counter()
            run(args.steps, 2)
            torch.cuda.synchronize()
            d = time.perf_counter() - t0
            if world > 1:
                tt = torch.tensor([d], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                d = float(tt.item())
            return world * B * args.steps / d

        chunk = args.host_chunk or 64
        modes = {"whole_batch": timed(0), f"chunked_{chunk}": timed(chunk)}
        pipe.set_host_chunk(args.host_chunk)

        # N > 1: the host feeds the GPUs at different rates when every rank copies at once (8 ranks on this pool's boxes:
        # 23 .. 36 GB/s per link, tools/h2d_concurrent.py), and with an equal split the slowest link sets the step time.
        # Balanced split: the same global batch (world x B frames per step) is cut in proportion to the link rates
        # measured here, each rank submits its share through the same API, the records are gathered (padded to the
        # largest share) and land on rank 0's host as before.
        balance = None
        if world > 1 and not args.no_balance:
            cap = (B * 3 // 2 + 7) // 8 * 8
            probe = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            probe.copy_(pin, non_blocking=True)
            barrier()
            ev0.record(stream)
            for _ in range(3):
                probe.copy_(pin, non_blocking=True)
            ev1.record(stream)
            torch.cuda.synchronize()
            bw = torch.tensor([3 * pin.numel() / (ev0.elapsed_time(ev1) * 1e-3) / 1e9], device=dev, dtype=torch.float64)
            bws = [torch.empty_like(bw) for _ in range(world)]
            dist.all_gather(bws, bw)
            bws = [float(x.item()) for x in bws]
            del probe
            share = [min(cap, max(8, int(round(world * B * x / sum(bws) / 8)) * 8)) for x in bws]
            i = 0
            while sum(share) != world * B and i < 64 * world:        # hand the rounding remainder out in steps of 8, fastest links first
                order = sorted(range(world), key=lambda r: -bws[r])
                r = order[i % world]
                d = 8 if sum(share) < world * B else -8
                if 8 <= share[r] + d <= cap:
                    share[r] += d
                i += 1
            if sum(share) == world * B:
                nb = share[rank]
                cap = max(share)                              # rows every rank contributes to the gather (NCCL wants equal blocks)
                reps_b = (cap + len(frames_u) - 1) // len(frames_u)
                pin_b = torch.from_numpy(np.ascontiguousarray(np.tile(frames_u, (reps_b, 1, 1))[:cap])).pin_memory()
                pipe_b, _, _, _ = setup_pipe(args, frames_u, H, W, rows, cols, local, cap)
                bouts = [pipe_b.alloc_outputs(cap, True, compact=True) for _ in range(2)]
                if rank == 0:
                    hgb = {k: torch.empty((world * v.shape[0],) + tuple(v.shape[1:]), dtype=v.dtype).pin_memory() for k, v in bouts[0][0].items() if k in REC}

                def land_b(s):
                    g = gather(bouts[s & 1][0])
                    if rank == 0:
                        for k, v in g.items():
                            hgb[k].copy_(v, non_blocking=True)

                def run_b(nsteps, s0):
                    pipe_b.submit_host_ptr(pin_b.data_ptr(), nb, H * W, W, s0 * nb, bouts[s0 & 1])
                    for s in range(s0 + 1, s0 + nsteps):
                        pipe_b.submit_host_ptr(pin_b.data_ptr(), nb, H * W, W, s * nb, bouts[s & 1])
                        pipe_b.wait_host()
                        land_b(s - 1)
                    pipe_b.wait_host()
                    land_b(s0 + nsteps - 1)
                    torch.cuda.current_stream().synchronize()

                pipe_b.reset_sequence()
                pipe_b.set_host_chunk(chunk)
                run_b(2, 0)
                barrier()
                t0 = time.perf_counter()
                run_b(args.steps, 2)
                torch.cuda.synchronize()
                d = time.perf_counter() - t0
                tt = torch.tensor([d], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                modes[f"balanced_chunked_{chunk}"] = world * B * args.steps / float(tt.item())
                balance = {"h2d_gbs_per_rank": [round(x, 1) for x in bws], "frames_per_rank": share, "padded_to": cap}
                pipe_b.close()
        # the synchronous call (one batch in, results out, nothing in flight afterwards) for comparison
        souts = pipe.alloc_outputs(B, False, compact=True)
        for s in range(2):
            pipe.process_host_ptr(pin.data_ptr(), B, H * W, W, s * B, souts)
        barrier()
        t1 = time.perf_counter()
        for s in range(min(args.steps, 10)):
            pipe.process_host_ptr(pin.data_ptr(), B, H * W, W, s * B, souts)
        dts = (time.perf_counter() - t1) / min(args.steps, 10)
        if world > 1:
            t = torch.tensor([dts], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dts = float(t[0].item())
        if rank == 0:
            best = max(modes, key=modes.get)
            d2h = sum(int(a.numel() * a.element_size()) if hasattr(a, "numel") else a.nbytes for a in houts[0][0].values())
            line["e2e"] = {"value": modes[best], "unit": UNIT, "h2d_bytes_per_step": int(B * H * W), "d2h_bytes_per_step": int(d2h),
                           "mode": best, "modes": modes, "balance": balance,
                           "api": "MarkerPipeline.submit_host_ptr / wait_host -> vbs_submit_host / vbs_wait_host (pinned host frames in, compact record "
                                  "block out, two batches in flight); the copy schedule (whole batch vs chunks, vbs_set_host_chunk) is picked by "
                                  "measuring both" + ("; records gathered to rank 0 over NCCL and copied to its pinned host memory every step; `balanced_*`: the "
                                                      "global batch of world x B frames is split across the ranks in proportion to their measured "
                                                      "host-to-device rates instead of equally" if on_dev else ""),
                           "synchronous_call_value": world * B / dts,
                           "synchronous_api": "MarkerPipeline.process_host_ptr -> vbs_process_host (chunked copy/compute overlap inside one call, per-rank host results)"}

    # ---- cpu_baseline: oracle port on a bounded sample, rank 0, N=1 only
    if rank == 0 and world == 1 and not args.no_cpu:
        fps_cpu, n, threads = cpu_single_process(frames_u, keys, xy, cam_params, plane_params, args.cpu_seconds)
        line["cpu_baseline"] = {"value": fps_cpu, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{n} frames of the same workload, one process, OpenCV threads={threads} (SciPy parts single-threaded), "
                                          f"host has {os.cpu_count()} cores"}
    if rank == 0:
        print(json.dumps(line))
    pipe.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
