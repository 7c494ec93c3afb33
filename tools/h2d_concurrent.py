#!/usr/bin/env python
"""Aggregate pinned H2D bandwidth with every rank copying at once (torchrun): the ceiling of the end-to-end
leg when several GPUs are fed from one host."""
import os, time, torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")
n = 256 * 1080 * 1920
src = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(2)]
for s in src:
    s.fill_(7)
dst = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(3):
    dst.copy_(src[0], non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
reps = 20
for i in range(reps):
    dst.copy_(src[i & 1], non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
gbs = reps * n / dt / 1e9
if world > 1:
    t = torch.tensor([gbs]); parts = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(parts, t)
    if rank == 0:
        v = [float(p) for p in parts]
        print(f"{world} ranks copying at once: per rank {min(v):.1f} .. {max(v):.1f} GB/s, total {sum(v):.1f} GB/s = {sum(v) * 1e9 / (1080 * 1920):.0f} 1080p frames/s")
else:
    print(f"1 rank: {gbs:.1f} GB/s")
