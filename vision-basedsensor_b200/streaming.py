"""Streaming form of the path (BASELINE.json config 5): a long frame sequence cut into contiguous shards,
one per GPU (one process per GPU), each shard processed batch by batch; the per-frame records stay on the
device until the shard is done, the last-seen state is patched across shard boundaries and the records are
gathered to rank 0 over NCCL.

    MD:434-458  the frame loop of ``MarkerTracker.process`` - here one ``vbs_process_device`` per batch
    R3:277,314  "previous = last seen observation of the same key" - the only cross-frame state; every
                shard starts from an empty table and ``sharding.finish_shard`` emits the one missing
                displacement row per marker once the tables of the preceding shards are known
    FD:141-159  the plane tilt of every frame travels with the records

The result on rank 0 is byte-identical to one sequential run over the whole sequence (tested under NCCL at
world size 2 in tests/test_gpu_multi.py and under gloo on the host logic).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

from . import capi, sharding

RECORD_KEYS = ("row_det", "row_cxy", "row_axes", "pos3d", "pos_flags", "plane", "plane_n", "n_markers")


@dataclass
class ShardRecords:
    """Per-frame records of one shard (device tensors, first axis = frame)."""
    lo: int
    hi: int
    row_det: object
    row_cxy: object
    row_axes: object
    pos3d: object
    pos_flags: object
    plane: object
    plane_n: object
    n_markers: object
    works: object = ()                   # NCCL work handles of the per-batch gathers in flight

    def tensors(self) -> dict:
        return {k: getattr(self, k) for k in RECORD_KEYS if getattr(self, k) is not None}


def process_shard(pipe, frames_of: Callable[[int, int], object], n_frames: int, batch: int, rank: int, world: int,
                  first_frame: int = 0, sink=None) -> ShardRecords:
    """Run this rank's contiguous shard of an ``n_frames`` sequence through ``pipe`` batch by batch.

    ``frames_of(lo, hi)`` returns the device tensor ``[hi-lo, H, W(,3)]`` uint8 of global frames lo..hi-1
    (``hi - lo <= batch``).  The reference array, camera and (optionally) plane inputs of ``pipe`` must be set.
    Records accumulate in device memory (96 B per reference entry + 32 B per frame).  With a ``sink``
    (equal-length shards) the records of every batch start travelling to rank 0 as soon as the batch is queued:
    the NCCL gather of batch k runs beside the kernels of batch k + 1."""
    import torch
    lo, hi = sharding.shard_bounds(n_frames, rank, world)
    n, R = hi - lo, pipe.R
    dev = torch.device("cuda", pipe.device)
    f64, i32, u8 = torch.float64, torch.int32, torch.uint8
    rec = ShardRecords(lo, hi,
                       torch.empty((n, R), dtype=i32, device=dev), torch.empty((n, R, 2), dtype=f64, device=dev),
                       torch.empty((n, R, 3), dtype=f64, device=dev), torch.empty((n, R, 7), dtype=f64, device=dev),
                       torch.empty((n, R), dtype=u8, device=dev),
                       torch.empty((n, 4), dtype=f64, device=dev) if pipe.have_plane else None,
                       torch.empty((n,), dtype=i32, device=dev) if pipe.have_plane else None,
                       torch.empty((n,), dtype=i32, device=dev))
    pipe.reset_sequence()
    pipe.set_first_frame(first_frame)            # the warm-up window (R3:255-256) counts from the GLOBAL first frame
    works = []
    for s in range(0, n, batch):                 # every batch writes straight into its slice of the shard's record block
        e = min(n, s + batch)
        arrays, o = {}, capi.VbsOutputs()
        for k in RECORD_KEYS:
            t = getattr(rec, k)
            if t is not None:
                arrays[k] = t[s:e]
                setattr(o, k, arrays[k].data_ptr())
        pipe.process(frames_of(lo + s, lo + e), lo + s, out=(arrays, o))
        if sink is not None and world > 1:
            works += sharding.gather_rows_async(rec.tensors(), s, e, rank, world, sink)
    rec.works = works
    return rec


def finish_and_gather(pipe, rec: ShardRecords, n_frames: int, rank: int, world: int, dst: int = 0, sink=None) -> Optional[dict]:
    """Patch the shard's first observation of every marker with the last-seen table arriving from the
    preceding shards, and deliver the records to ``dst`` (concatenated in global frame order).

    Without a sink: patch, then one gather of the whole shard.  With a sink the records are already on their way
    (``process_shard``): only the R x 4 tail tables are exchanged, and ``dst`` applies every shard's patch to the
    gathered block itself (``vbs_fix_displacement`` needs nothing but the block, the camera and the incoming table)."""
    import torch
    dev = torch.device("cuda", pipe.device)
    pipe.sync()
    if sink is None or world == 1:
        sharding.finish_shard(pipe, rec, rank, world, dev)
        counts = [hi - lo for lo, hi in (sharding.shard_bounds(n_frames, r, world) for r in range(world))]
        return sharding.gather_records(rec.tensors(), rank, world, dst, counts)
    tails = sharding.all_tail_tables(pipe.get_last_seen(), rank, world, dev)
    for w in getattr(rec, "works", ()):
        w.wait()                                 # the current stream now waits for the gathers
    if rank != dst:
        return None
    for r in range(1, world):
        lo, hi = sink.bounds[r]
        inc = sharding.incoming_last_seen(tails, r)
        capi.check(pipe._ctx, capi.lib.vbs_fix_displacement(pipe._ctx, sink.data["pos3d"][lo:hi].data_ptr(), sink.data["pos_flags"][lo:hi].data_ptr(),
                                                            hi - lo, inc.ctypes.data))
    return sink.data


def make_sink(pipe, n_frames: int, rank: int, world: int, dst: int = 0):
    """Preallocated destination of the gathered records on ``dst`` (None elsewhere, and None when the shards are not all
    the same length: the collective then pads instead, see ``sharding.gather_records``)."""
    import torch
    counts = {hi - lo for lo, hi in (sharding.shard_bounds(n_frames, r, world) for r in range(world))}
    if world == 1 or len(counts) != 1:
        return None
    dev = torch.device("cuda", pipe.device)
    R = pipe.R
    like = {"row_det": torch.empty((0, R), dtype=torch.int32), "row_cxy": torch.empty((0, R, 2), dtype=torch.float64),
            "row_axes": torch.empty((0, R, 3), dtype=torch.float64), "pos3d": torch.empty((0, R, 7), dtype=torch.float64),
            "pos_flags": torch.empty((0, R), dtype=torch.uint8), "n_markers": torch.empty((0,), dtype=torch.int32)}
    if pipe.have_plane:
        like.update({"plane": torch.empty((0, 4), dtype=torch.float64), "plane_n": torch.empty((0,), dtype=torch.int32)})
    return sharding.RecordSink(like, n_frames if rank == dst else 0, world, dev) if rank == dst else _RemoteSink()


class _RemoteSink:
    """Placeholder on the ranks that only send."""
    def parts(self, name, s, e):
        return None


def run_stream(pipe, frames_of, n_frames: int, batch: int, rank: int, world: int, first_frame: int = 0, sink=None):
    """process_shard + finish_and_gather; returns (records dict on rank 0 / None elsewhere, ShardRecords).
    ``sink`` = ``make_sink(...)`` (reusable across calls) overlaps the gather with the kernels."""
    rec = process_shard(pipe, frames_of, n_frames, batch, rank, world, first_frame, sink)
    return finish_and_gather(pipe, rec, n_frames, rank, world, sink=sink), rec
