#!/usr/bin/env python
"""Single-proteome (config 2) latency through ka_annotate vs chunk size."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kmers_anno_b200 as ka
from kmers_anno_b200 import synth
from kmers_anno_b200.engine import pinned_array
fam = synth.Families(30000)
kmers, roles = fam.table(int(1e8), K=8)
res, off, _ = fam.batch(0, 1, n_prot=4500, alloc=pinned_array)
n = off.shape[0] - 1
out = (pinned_array(n, np.int32), pinned_array(n, np.int32), pinned_array(n, np.uint8))
eng = ka.Engine([0]); eng.db_load(kmers, roles, 8)
for chunk in (32 << 20, 800000, 500000, 400000, 300000, 200000):
    eng.set_option("chunk_residues", chunk)
    for _ in range(10): eng.annotate(res, off, 5, out=out)
    t = time.perf_counter()
    for _ in range(200): eng.annotate(res, off, 5, out=out)
    dt = (time.perf_counter() - t) / 200 * 1e6
    st = eng.stats()
    print(f"chunk {chunk:9d}: e2e {dt:7.1f} us  kernel_sum {st['kernel_ms']*1e3:6.1f} us launches {st['kernel_launches']}", flush=True)
