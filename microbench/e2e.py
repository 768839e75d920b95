#!/usr/bin/env python
"""End-to-end ka_annotate timing (pinned host buffers) vs chunk size, plus raw H2D rate."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kmers_anno_b200 as ka
from kmers_anno_b200 import synth
from kmers_anno_b200.engine import pinned_array
genomes = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
opts = dict(kv.split("=") for kv in sys.argv[2:])
chunks = [int(c) << 20 for c in opts.pop("chunks", "16,32,48,64,96,128,256").split(",")]
fam = synth.Families(30000)
kmers, roles = fam.table(int(1e8), K=8)
res, off, _ = fam.batch(0, genomes, n_prot=4500, alloc=pinned_array)
n = off.shape[0] - 1
out = (pinned_array(n, np.int32), pinned_array(n, np.int32), pinned_array(n, np.uint8))
eng = ka.Engine([0])
for k, v in opts.items(): eng.set_option(k, float(v))
eng.db_load(kmers, roles, 8)
t = time.perf_counter(); b = eng.upload(res, off); dt = time.perf_counter() - t
print(f"upload (sync cudaMemcpy from pinned) {len(res)/dt/1e9:.1f} GB/s", flush=True)
eng.annotate_resident(b, 5); eng.annotate_resident(b, 5); print("resident kernel ms", eng.stats()["kernel_ms"], flush=True); b.free()
codes, off32 = eng.pack(res, off, alloc=pinned_array)
for chunk in chunks:
    eng.set_option("chunk_residues", chunk)
    for name, call in (("bytes ", lambda: eng.annotate(res, off, 5, out=out)), ("packed", lambda: eng.annotate_packed(codes, off32, 5, out=out))):
        best = 1e9
        for r in range(4):
            t = time.perf_counter(); call(); dt = (time.perf_counter() - t) * 1e3
            if r: best = min(best, dt)
        st = eng.stats()
        print(f"chunk {chunk>>20:4d} Mi {name}: e2e {best:.2f} ms  ({n/best/1e3:.1f} M seq/s)  kernel_sum {st['kernel_ms']:.2f} ms  tile_sum {st['tile_kernel_ms']:.2f}  launches {st['kernel_launches']}  wall_inside {st['wall_ms']:.2f}", flush=True)
