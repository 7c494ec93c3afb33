"""Worker of tests/test_gpu_multi.py: one process per GPU under torchrun (NCCL).  Streams a short synthetic
sequence through vbs_b200.streaming (contiguous shards, last-seen exchange, NCCL gather of records + plane
tilt) and, on rank 0, compares the gathered records BYTE FOR BYTE with one sequential run over the whole
sequence on one GPU.  Prints one JSON line on rank 0; exit code 0 only if everything is identical."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build_sequence(n_frames):
    from vbs_b200 import synth
    h, w, rows, cols = 560, 640, 6, 8
    centres = synth.grid_layout(h, w, rows, cols, 60.0)
    seq = synth.compression_sequence(h, w, centres, 11.0, n_frames, tilt=0.4, depth=1.0, seed0=500)
    half = n_frames // 2
    for f in (half - 2, half - 1, half, half + 1):            # dropouts that straddle the shard boundary
        x, y = centres[13].astype(int); seq[f, y - 20:y + 20, x - 20:x + 20] = 170
    for f in range(half - 5, half + 3):
        x, y = centres[30].astype(int); seq[f, y - 20:y + 20, x - 20:x + 20] = 170
    return seq, (h, w, rows, cols)


def main():
    n_frames, batch = int(sys.argv[1]), int(sys.argv[2])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    import torch
    import torch.distributed as dist
    import vbs_b200  # noqa: F401
    from vbs_b200 import pipeline, reference_state, streaming, synth

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    seq, (h, w, rows, cols) = build_sequence(n_frames)
    frames = torch.from_numpy(seq).to(dev)
    K, D, R, T = synth.synthetic_camera()
    K = K.copy(); K[0, 2] = w / 2 + 3.1; K[1, 2] = h / 2 - 2.3

    def make(max_batch):
        p = pipeline.MarkerPipeline(h, w, 1, max_batch=max_batch, max_markers=256, max_refs=rows * cols)
        assert p.device == local                      # helper contexts follow torch's current device
        r0 = p.process(frames[:1], 0); p.sync(); h0 = r0.to_host()
        keys, xy = reference_state.grid_ids(h0.marker_xy[0, : int(h0.n_markers[0])], cols)
        p.set_reference([k[0] for k in keys], [k[1] for k in keys], xy[:, 0], xy[:, 1], 20.0)
        p.set_camera(K, D, R, T, 2.0, 5.0, 50.0, warmup_frames=2)
        r0 = p.process(frames[:1], 0); p.sync()
        start = np.nan_to_num(r0.to_host().pos3d[0, :, :3])
        ref_xyz = np.stack([(xy[:, 0] - w / 2) / 11.0, (xy[:, 1] - h / 2) / 11.0, np.zeros(len(xy))], 1)
        p.set_plane(ref_xyz, start, None)
        return p

    pipe = make(batch)
    sink = streaming.make_sink(pipe, n_frames, rank, world)     # rank 0: preallocated landing area (ragged shards travel padded)
    got, rec = streaming.run_stream(pipe, lambda lo, hi: frames[lo:hi], n_frames, batch, rank, world, sink=sink)
    if n_frames % 2 == 0:                                         # a second pass through the same sink must give the same bytes
        got, rec = streaming.run_stream(pipe, lambda lo, hi: frames[lo:hi], n_frames, batch, rank, world, sink=sink)
    torch.cuda.synchronize()
    ok, report = True, {}
    if rank == 0:
        one = make(n_frames)
        want, _ = streaming.run_stream(one, lambda lo, hi: frames[lo:hi], n_frames, n_frames, 0, 1)
        torch.cuda.synchronize()
        for k in streaming.RECORD_KEYS:
            a, b = got[k].contiguous().view(torch.uint8), want[k].contiguous().view(torch.uint8)
            same = a.shape == b.shape and bool(torch.equal(a, b))
            report[k] = same
            ok = ok and same
        flags = got["pos_flags"].cpu().numpy()
        report["displacement_rows"] = int(((flags & 4) != 0).sum())
        report["tilt_finite_frames"] = int(torch.isfinite(got["plane"][:, 3]).sum().item())
        report["frames"] = int(got["pos3d"].shape[0])
        ok = ok and report["displacement_rows"] > 0 and report["tilt_finite_frames"] > n_frames // 2 and report["frames"] == n_frames
        one.close()
        print(json.dumps({"ok": ok, "world": world, "report": report}))
    pipe.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
