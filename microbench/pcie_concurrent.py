#!/usr/bin/env python
"""Host->device ingest ceiling of the box: pinned copies to the first N GPUs AT THE SAME TIME, N = 1, 2, 4, 8.
Per-GPU and aggregate GB/s, 1 GiB per GPU in 64 MiB copies (what ka_annotate issues), best of 4 rounds.
Also each GPU alone, to show asymmetries between host paths.  One process, one stream per device."""
import json, sys, time
import torch
n_dev = torch.cuda.device_count()
n = 1 << 30
chunk = 64 << 20
host = [torch.empty(n, dtype=torch.uint8, pin_memory=True).fill_(65) for _ in range(n_dev)]
dev = [torch.empty(n, dtype=torch.uint8, device=f"cuda:{d}") for d in range(n_dev)]
streams = [torch.cuda.Stream(device=d) for d in range(n_dev)]


def run(devs):
    best = None
    for rep in range(4):
        evs = []
        for d in devs:
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        for d in devs:
            with torch.cuda.device(d), torch.cuda.stream(streams[d]):
                a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                a.record()
                for o in range(0, n, chunk):
                    dev[d][o:o + chunk].copy_(host[d][o:o + chunk], non_blocking=True)
                b.record()
                evs.append((d, a, b))
        for d in devs:
            torch.cuda.synchronize(d)
        wall = time.perf_counter() - t0
        per = {d: n / (a.elapsed_time(b) * 1e-3) / 1e9 for d, a, b in evs}
        agg = len(devs) * n / wall / 1e9
        if best is None or agg > best[0]:
            best = (agg, per)
    return best


for d in range(n_dev):
    agg, per = run([d])
    print(json.dumps({"gpus": [d], "aggregate_GBps": round(agg, 1)}), flush=True)
for k in (2, 4, 8):
    if k > n_dev:
        break
    agg, per = run(list(range(k)))
    print(json.dumps({"gpus": list(range(k)), "aggregate_GBps": round(agg, 1), "per_gpu_GBps": {str(d): round(v, 1) for d, v in per.items()}}), flush=True)
if n_dev >= 8:
    for devs in ([4, 5, 6, 7], [0, 2, 4, 6], [0, 1, 4, 5]):
        agg, per = run(devs)
        print(json.dumps({"gpus": devs, "aggregate_GBps": round(agg, 1), "per_gpu_GBps": {str(d): round(v, 1) for d, v in per.items()}}), flush=True)
