"""How fast can this box copy pinned host memory to the GPU?  (the ceiling of bench.py's end-to-end leg: 1 byte of text per base)"""
import time, torch
n = 2_530_000_000
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for label, chunk in (("one copy", n), ("5 MB pieces", 5_060_000), ("253 MB pieces", 253_000_000)):
    torch.cuda.synchronize()
    best = 0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for o in range(0, n, chunk):
            d[o:o + chunk].copy_(h[o:o + chunk], non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        best = max(best, n / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    print(f"{label}: {best:.1f} GB/s")
