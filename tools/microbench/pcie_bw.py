"""PCIe bandwidth of the box with page-locked buffers of one 4K picture (what bounds bench.py's e2e figure)."""
import time
import torch
n = 31 * 1024 * 1024
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h2.copy_(d2, non_blocking=True))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    print(name, "%.1f GB/s" % (20 * n / (time.perf_counter() - t0) / 1e9))
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
print("both directions concurrently: %.1f GB/s each" % (20 * n / (time.perf_counter() - t0) / 1e9))
