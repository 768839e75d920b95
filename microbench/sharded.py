#!/usr/bin/env python
"""Replicated vs sharded table on all visible GPUs of one process (ka_annotate, pinned host buffers)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kmers_anno_b200 as ka
from kmers_anno_b200 import synth
from kmers_anno_b200.engine import pinned_array
n_dev = int(sys.argv[1]); genomes = int(sys.argv[2]) if len(sys.argv) > 2 else 400
fam = synth.Families(30000)
kmers, roles = fam.table(int(1e8), K=8)
res, off, _ = fam.batch(0, genomes, n_prot=4500, alloc=pinned_array)
n = off.shape[0] - 1
out = (pinned_array(n, np.int32), pinned_array(n, np.int32), pinned_array(n, np.uint8))
ref = None
for mode, chunk in ((0, 32 << 20), (1, 32 << 20), (2, 32 << 20), (2, 8 << 20)):
    eng = ka.Engine(list(range(n_dev)))
    eng.set_option("table_mode", mode)
    eng.set_option("chunk_residues", chunk)
    t = time.time(); eng.db_load(kmers, roles, 8); tl = time.time() - t
    info = eng.db_info()
    best = 1e9
    for r in range(4):
        t = time.perf_counter(); eng.annotate(res, off, 5, out=out); dt = (time.perf_counter() - t) * 1e3
        if r: best = min(best, dt)
    st = eng.stats()
    sig = hash(out[0].tobytes()) ^ hash(out[1].tobytes())
    if ref is None: ref = sig
    print(f"chunk {chunk>>20:2d} Mi mode {mode} ({('replicated', 'sharded / peer loads', 'sharded / NCCL routed')[mode]}) on {n_dev} GPUs: db load {tl:.1f}s table/GPU {info['table_bytes']/1e6/(n_dev if mode else 1):.0f} MB; "
          f"e2e {best:.2f} ms {st['probes']/best/1e6:.1f} G probes/s  kernel max {st['kernel_ms']:.2f} ms  same={sig==ref}", flush=True)
    eng.close()
