#!/usr/bin/env python
"""Throughput of the mid-sequence tile launch and of big_kernel on batches of long sequences."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kmers_anno_b200 as ka
from kmers_anno_b200 import synth
fam = synth.Families(30000)
kmers, roles = fam.table(int(1e8), K=8)
res, off, _ = fam.batch(0, 40, n_prot=4500)
eng = ka.Engine([0]); eng.db_load(kmers, roles, 8)
total = int(off[-1])
for L in (1500, 4000, 8000, 20000, 100000):
    n = total // L
    o = (np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
    b = eng.upload(res[: n * L], o)
    for _ in range(3): eng.annotate_resident(b, 5)
    st = eng.stats(); b.free()
    print(f"L={L:6d}: {n} sequences, kernels {st['kernel_ms']:.3f} ms, {st['probes']/st['kernel_ms']/1e6:.1f} G probes/s", flush=True)
