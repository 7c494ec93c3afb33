"""world_size-2 gloo tests of the multi-GPU host logic (frame sharding, last-seen exchange,
record gather) - no GPU needed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import vbs_b200  # noqa: F401
from vbs_b200 import sharding


def test_shard_bounds_partition_every_frame_once():
    for n in (0, 1, 7, 256, 65536, 65537):
        for world in (1, 2, 4, 8):
            spans = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_incoming_last_seen_picks_latest_earlier_shard():
    t = np.zeros((3, 4, 4)); t[:, :, 3] = -1
    t[0, 0] = (1, 2, 3, 10); t[1, 0] = (4, 5, 6, 50)          # seen in shard 0 and 1
    t[0, 1] = (7, 8, 9, 20)                                   # only in shard 0 (dropout in shard 1)
    t[2, 2] = (1, 1, 1, 90)                                   # only in shard 2
    assert (sharding.incoming_last_seen(t, 0)[:, 3] == -1).all()
    a = sharding.incoming_last_seen(t, 1)
    assert a[0].tolist() == [1, 2, 3, 10] and a[1].tolist() == [7, 8, 9, 20] and a[2, 3] == -1
    b = sharding.incoming_last_seen(t, 2)
    assert b[0].tolist() == [4, 5, 6, 50] and b[1].tolist() == [7, 8, 9, 20] and b[2, 3] == -1 and b[3, 3] == -1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        R, n = 5, 11
        lo, hi = sharding.shard_bounds(n, rank, world)
        # tail table of this shard: marker m was last seen at frame hi-1-m (if that lies in the shard)
        tail = np.zeros((R, 4)); tail[:, 3] = -1
        for m in range(R):
            f = hi - 1 - m
            if f >= lo:
                tail[m] = (100 + f, 200 + f, 10 + m, f)
        inc = sharding.exchange_last_seen(tail, rank, world)
        # records: one row per frame holding its global frame number
        rec = {"frameno": torch.arange(lo, hi, dtype=torch.float64).reshape(-1, 1).repeat(1, 3)[: (n // world)]}
        got = sharding.gather_records(rec, rank, world, dst=0)
        # ragged shards (6 + 5 frames): padded for the collective, trimmed on the destination
        full = {"frameno": torch.arange(lo, hi, dtype=torch.float64).reshape(-1, 1), "flag": torch.full((hi - lo, 2), rank, dtype=torch.uint8)}
        rag = sharding.gather_records(full, rank, world, dst=0, counts=[b - a for a, b in (sharding.shard_bounds(n, r, world) for r in range(world))])
        q.put((rank, inc, None if got is None else got["frameno"].numpy(), None if rag is None else (rag["frameno"].numpy(), rag["flag"].numpy())))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_exchange_and_gather():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(world):
        r, inc, rec, rag = q.get(timeout=120)
        out[r] = (inc, rec, rag)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert (out[0][0][:, 3] == -1).all()                       # nothing precedes shard 0
    lo0, hi0 = sharding.shard_bounds(11, 0, 2)
    for m in range(5):                                         # shard 1 receives shard 0's tail
        assert out[1][0][m, 3] == hi0 - 1 - m and out[1][0][m, 0] == 100 + hi0 - 1 - m
    assert out[1][1] is None
    frames = out[0][1][:, 0]
    assert frames.tolist() == list(range(0, 5)) + list(range(6, 11))      # rank-major = frame order
    assert out[1][2] is None
    rf, rflag = out[0][2]
    assert rf[:, 0].tolist() == list(range(11)) and rflag[:, 0].tolist() == [0] * 6 + [1] * 5
