#!/usr/bin/env python
"""Hit-rate sensitivity of the tile kernel: the C3 batch (~30 % of windows are in the table) vs a
batch drawn from unrelated families (~0.4 % hits), with and without presence signatures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kmers_anno_b200 as ka
from kmers_anno_b200 import synth
fam = synth.Families(30000)
kmers, roles = fam.table(int(1e8), K=8)
other = synth.Families(30000, seed=777)
batches = {"c3": fam.batch(0, 60, n_prot=4500)[:2], "unrelated": other.batch(0, 60, n_prot=4500)[:2]}
import json
configs = [json.loads(a) for a in sys.argv[1:]] or [{}, {"filter": 1}]
for opts in configs:
    eng = ka.Engine([0])
    for k, v in opts.items(): eng.set_option(k, float(v))
    eng.db_load(kmers, roles, 8)
    for name, (res, off) in batches.items():
        b = eng.upload(res, off)
        for _ in range(3): eng.annotate_resident(b, 5)
        role, hits, flag = eng.download(b)
        st = eng.stats()
        print(opts, name, "tile ms %.3f" % st["tile_kernel_ms"], "G probes/s %.1f" % (st["probes"] / st["tile_kernel_ms"] / 1e6),
              "hits/probe %.4f" % (hits.sum() / st["probes"]), flush=True)
        b.free()
    eng.close()
