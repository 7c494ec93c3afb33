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

import numpy as np

from . import capi, sharding

RECORD_KEYS = ("row_det", "row_cxy", "row_axes", "pos3d", "pos_flags", "plane", "plane_n", "n_markers")


def record_layout(n: int, R: int, have_plane: bool):
    """Sections of one shard's record blob: [(key, shape, torch dtype name, byte offset)], total bytes.  Every rank's
    records live in ONE contiguous allocation so that the whole shard travels to rank 0 in one collective."""
    spec = [("row_det", (n, R), "int32", 4), ("row_cxy", (n, R, 2), "float64", 8), ("row_axes", (n, R, 3), "float64", 8),
            ("pos3d", (n, R, 7), "float64", 8), ("pos_flags", (n, R), "uint8", 1)]
    if have_plane:
        spec += [("plane", (n, 4), "float64", 8), ("plane_n", (n,), "int32", 4)]
    spec += [("n_markers", (n,), "int32", 4)]
    out, off = [], 0
    for key, shape, dt, isz in spec:
        nbytes = isz * int(np.prod(shape))
        out.append((key, shape, dt, off))
        off = (off + nbytes + 255) // 256 * 256
    return out, off


def blob_views(blob, n: int, R: int, have_plane: bool) -> dict:
    """Typed views of the sections of a record blob (no copies)."""
    import torch
    views = {}
    for key, shape, dt, off in record_layout(n, R, have_plane)[0]:
        dtype = getattr(torch, dt)
        nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        views[key] = blob[off: off + nbytes].view(dtype).view(shape)
    return views


@dataclass
class ShardRecords:
    """Per-frame records of one shard (device tensors, first axis = frame), all views of ``blob``."""
    lo: int
    hi: int
    blob: object
    views: dict

    def __getattr__(self, key):
        views = object.__getattribute__(self, "views")
        if key in RECORD_KEYS:
            return views.get(key)
        raise AttributeError(key)

    def tensors(self) -> dict:
        return dict(self.views)


class RecordSink:
    """Rank 0's landing area: ``[world, blob_bytes]`` (allocated once, reused across streams) plus the assembled
    ``[n_frames, ...]`` arrays in global frame order."""

    def __init__(self, n_frames: int, R: int, have_plane: bool, world: int, device):
        import torch
        self.n_frames, self.R, self.have_plane, self.world = n_frames, R, have_plane, world
        self.bounds = [sharding.shard_bounds(n_frames, r, world) for r in range(world)]
        self.blob_bytes = max(record_layout(hi - lo, R, have_plane)[1] for lo, hi in self.bounds)
        self.blobs = torch.empty((world, self.blob_bytes), dtype=torch.uint8, device=device)
        self.data = {}
        for key, shape, dt, _ in record_layout(n_frames, R, have_plane)[0]:
            self.data[key] = torch.empty(shape, dtype=getattr(torch, dt), device=device)

    def shard_views(self, r: int) -> dict:
        lo, hi = self.bounds[r]
        return blob_views(self.blobs[r], hi - lo, self.R, self.have_plane)

    def assemble(self) -> dict:
        for r, (lo, hi) in enumerate(self.bounds):
            for key, v in self.shard_views(r).items():
                self.data[key][lo:hi].copy_(v)
        return self.data


def blob_bytes_for(n_frames: int, R: int, have_plane: bool, world: int) -> int:
    return max(record_layout(hi - lo, R, have_plane)[1] for lo, hi in (sharding.shard_bounds(n_frames, r, world) for r in range(world)))


def process_shard(pipe, frames_of: Callable[[int, int], object], n_frames: int, batch: int, rank: int, world: int,
                  first_frame: int = 0) -> ShardRecords:
    """Run this rank's contiguous shard of an ``n_frames`` sequence through ``pipe`` batch by batch.

    ``frames_of(lo, hi)`` returns the device tensor ``[hi-lo, H, W(,3)]`` uint8 of global frames lo..hi-1
    (``hi - lo <= batch``).  The reference array, camera and (optionally) plane inputs of ``pipe`` must be set.
    Records accumulate in one device blob (96 B per reference entry + 32 B per frame); every batch writes
    straight into its rows of it."""
    import torch
    lo, hi = sharding.shard_bounds(n_frames, rank, world)
    n, R = hi - lo, pipe.R
    dev = torch.device("cuda", pipe.device)
    key = (n_frames, world, R, pipe.have_plane)
    if getattr(pipe, "_stream_blob_key", None) != key:           # one allocation per stream geometry, reused across streams
        pipe._stream_blob = torch.empty((blob_bytes_for(n_frames, R, pipe.have_plane, world),), dtype=torch.uint8, device=dev)
        pipe._stream_blob_key = key
    rec = ShardRecords(lo, hi, pipe._stream_blob, blob_views(pipe._stream_blob, n, R, pipe.have_plane))
    pipe.reset_sequence()
    pipe.set_first_frame(first_frame)            # the warm-up window (R3:255-256) counts from the GLOBAL first frame
    for s in range(0, n, batch):
        e = min(n, s + batch)
        arrays, o = {}, capi.VbsOutputs()
        for k, t in rec.views.items():
            arrays[k] = t[s:e]
            setattr(o, k, arrays[k].data_ptr())
        pipe.process(frames_of(lo + s, lo + e), lo + s, out=(arrays, o))
    return rec


def finish_and_gather(pipe, rec: ShardRecords, n_frames: int, rank: int, world: int, dst: int = 0, sink=None) -> Optional[dict]:
    """Deliver the records to ``dst`` in global frame order with the shard boundaries patched: ONE gather of the
    record blobs, one small all_gather of the R x 4 last-seen tables, and on ``dst`` one ``vbs_fix_displacement``
    per later shard (it needs nothing but the gathered block, the camera and the table arriving from the
    preceding shards) - results byte-identical to one sequential run."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", pipe.device)
    pipe.sync()
    if world == 1:
        return rec.tensors()
    tails = sharding.all_tail_tables(pipe.get_last_seen(), rank, world, dev)
    if rank == dst and sink is None:
        sink = RecordSink(n_frames, pipe.R, pipe.have_plane, world, dev)
    dist.gather(rec.blob, [sink.blobs[r] for r in range(world)] if rank == dst else None, dst=dst)
    if rank != dst:
        return None
    for r in range(1, world):
        lo, hi = sink.bounds[r]
        v = sink.shard_views(r)
        inc = sharding.incoming_last_seen(tails, r)
        capi.check(pipe._ctx, capi.lib.vbs_fix_displacement(pipe._ctx, v["pos3d"].data_ptr(), v["pos_flags"].data_ptr(), hi - lo, inc.ctypes.data))
    return sink.assemble()


def make_sink(pipe, n_frames: int, rank: int, world: int, dst: int = 0):
    """Preallocated landing area on ``dst`` (None elsewhere / for one rank): pass it to ``run_stream`` so that no
    allocation happens inside a timed region."""
    import torch
    if world == 1 or rank != dst:
        return None
    return RecordSink(n_frames, pipe.R, pipe.have_plane, world, torch.device("cuda", pipe.device))


def run_stream(pipe, frames_of, n_frames: int, batch: int, rank: int, world: int, first_frame: int = 0, sink=None):
    """process_shard + finish_and_gather; returns (records dict on rank 0 / None elsewhere, ShardRecords)."""
    rec = process_shard(pipe, frames_of, n_frames, batch, rank, world, first_frame)
    return finish_and_gather(pipe, rec, n_frames, rank, world, sink=sink), rec
