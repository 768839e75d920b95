#!/usr/bin/env python
"""Raw pinned host->device copy rate of the box (the wall behind bench.py's e2e number)."""
import torch
n = 1536 << 20
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); h.fill_(65)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for chunk in (n, 64 << 20, 32 << 20):
    best = 0.0
    for rep in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for o in range(0, n, chunk):
            d[o:o + chunk].copy_(h[o:o + chunk], non_blocking=True)
        b.record(); torch.cuda.synchronize()
        best = max(best, n / (a.elapsed_time(b) * 1e-3) / 1e9)
    print(f"H2D pinned, {n >> 20} MiB in {chunk >> 20} MiB copies on one stream: {best:.1f} GB/s", flush=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
best = 0.0
for rep in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    s1.wait_event(a); s2.wait_event(a)
    for i, o in enumerate(range(0, n, 32 << 20)):
        with torch.cuda.stream(s1 if i % 2 == 0 else s2):
            d[o:o + (32 << 20)].copy_(h[o:o + (32 << 20)], non_blocking=True)
    e1, e2 = torch.cuda.Event(), torch.cuda.Event()
    e1.record(s1); e2.record(s2)
    torch.cuda.current_stream().wait_event(e1); torch.cuda.current_stream().wait_event(e2)
    b.record(); torch.cuda.synchronize()
    best = max(best, n / (a.elapsed_time(b) * 1e-3) / 1e9)
print(f"H2D pinned, 32 MiB copies alternating over two streams: {best:.1f} GB/s", flush=True)
