"""Frame sharding across the GPUs of one box (one process per GPU, ``torch.distributed``).

Frames are independent units, so the data path needs NO collective: rank r takes the contiguous
block ``shard_bounds(n, r, world)`` and runs the whole pipeline on it.  Two small exchanges remain:

  * the last-seen table (R x 4 doubles) - the only cross-frame state of the path (R3:277,314
    "previous = last seen observation of the same key").  Every rank processes its shard from an
    empty table, then one ``all_gather`` of the tail tables lets each rank patch the single missing
    displacement row per marker (``vbs_fix_displacement``); results are then byte-identical to a
    one-GPU run.
  * the gather of the per-frame records (3D field, IDs, plane) to rank 0 - NCCL over NVLink on a
    GPU box, ``gloo`` in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of rank ``rank``; blocks differ by at most one frame."""
    base, rem = divmod(n_frames, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def incoming_last_seen(tables: np.ndarray, rank: int) -> np.ndarray:
    """Last-seen table entering shard ``rank``: per reference entry the most recent observation
    (largest frame number) among the tail tables of ranks < rank; frame = -1 where none exists."""
    tables = np.asarray(tables, dtype=np.float64)
    world, R, _ = tables.shape
    out = np.zeros((R, 4), dtype=np.float64)
    out[:, 3] = -1.0
    for r in range(rank):
        newer = tables[r, :, 3] > out[:, 3]
        out[newer] = tables[r, newer]
    return out


def exchange_last_seen(table: np.ndarray, rank: int, world: int, device=None) -> np.ndarray:
    """all_gather the tail tables and return the table entering this rank's shard."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return incoming_last_seen(table[None], 0)
    t = torch.from_numpy(np.ascontiguousarray(table, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    return incoming_last_seen(np.stack([p.cpu().numpy() for p in parts]), rank)


def gather_records(tensors: dict, rank: int, world: int, dst: int = 0, counts=None):
    """Gather per-frame record tensors to ``dst``; returns a dict of tensors concatenated in rank order
    (= global frame order for contiguous shards) on ``dst``, else None.  ``counts[r]`` = frames of rank r when
    the shards are not all the same length (``shard_bounds`` blocks differ by at most one frame): the shorter
    blocks are padded for the collective and trimmed on ``dst``."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return dict(tensors)
    ragged = counts is not None and len(set(counts)) > 1
    nmax = max(counts) if ragged else None
    out = {}
    for name, t in tensors.items():
        if ragged and t.shape[0] < nmax:
            pad = torch.zeros((nmax - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            t = torch.cat([t, pad], dim=0)
        t = t.contiguous()
        parts = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
        dist.gather(t, parts, dst=dst)
        if rank == dst:
            if ragged:
                parts = [p[:c] for p, c in zip(parts, counts)]
            out[name] = torch.cat(parts, dim=0)
    return out if rank == dst else None


def finish_shard(pipe, result, rank: int, world: int, device=None):
    """After ``pipe`` processed this rank's shard from an empty table: exchange tails and patch the
    missing displacement rows in ``result`` (device tensors pos3d / pos_flags) in place."""
    from . import capi
    incoming = exchange_last_seen(pipe.get_last_seen(), rank, world, device)
    n = result.pos3d.shape[0]
    capi.check(pipe._ctx, capi.lib.vbs_fix_displacement(pipe._ctx, result.pos3d.data_ptr(), result.pos_flags.data_ptr(), n,
                                                        incoming.ctypes.data))
    return incoming


def all_tail_tables(table: np.ndarray, rank: int, world: int, device=None) -> np.ndarray:
    """[world, R, 4]: the last-seen tables every shard ended with (one small all_gather)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(table, dtype=np.float64))
    if world == 1:
        return t.numpy()[None]
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    return np.stack([p.cpu().numpy() for p in parts])
