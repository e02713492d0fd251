"""Aggregate pinned H2D bandwidth with 1, 2, 4 concurrent streams (development aid)."""
import time, torch
n = 8 * 1024**3 // 4
h = torch.empty(n, dtype=torch.float32, pin_memory=True)
d = torch.empty(n, dtype=torch.float32, device="cuda")
for k in (1, 2, 4):
    streams = [torch.cuda.Stream() for _ in range(k)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    chunk = n // 64
    for c in range(64):
        with torch.cuda.stream(streams[c % k]):
            d[c * chunk:(c + 1) * chunk].copy_(h[c * chunk:(c + 1) * chunk], non_blocking=True)
    torch.cuda.synchronize()
    print(f"{k} stream(s): {n*4/(time.perf_counter()-t0)/1e9:.1f} GB/s")
# D2H concurrently with H2D
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n // 2, dtype=torch.float32, pin_memory=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
with torch.cuda.stream(s2): h2.copy_(d[: n // 2], non_blocking=True)
torch.cuda.synchronize(); el = time.perf_counter() - t0
print(f"H2D 8 GiB + D2H 4 GiB concurrently: {el*1e3:.0f} ms -> H2D-equivalent {n*4/el/1e9:.1f} GB/s")
