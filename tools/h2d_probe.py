import torch, time
n = 531 * 1000 * 1000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for chunks in (1, 4, 8):
    step = n // chunks
    torch.cuda.synchronize()
    for rep in range(2):
        t0 = time.perf_counter()
        for c in range(chunks):
            d[c * step:(c + 1) * step].copy_(h[c * step:(c + 1) * step], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"H2D pinned {n/1e6:.0f} MB in {chunks} chunk(s): {dt*1e3:.2f} ms = {n/dt/1e9:.1f} GB/s")
o = torch.empty(22 * 1000 * 1000, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter(); o.copy_(d[:o.numel()], non_blocking=True); torch.cuda.synchronize()
print(f"D2H 22 MB: {(time.perf_counter()-t0)*1e3:.2f} ms")
