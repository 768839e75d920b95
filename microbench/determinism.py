#!/usr/bin/env python
"""Build the full-size table twice and check annotate results against each other and the oracle."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kmers_anno_b200 as ka
from kmers_anno_b200 import synth
import oracle
n = float(sys.argv[1]) if len(sys.argv) > 1 else 1e8
fam = synth.Families(30000)
kmers, roles = fam.table(int(n), K=8)
km = kmers.reshape(-1, 8)
print("distinct kmers in DB:", len(np.unique(km.view(np.uint64))), "of", len(roles), flush=True)
res, off, _ = fam.batch(0, 20, n_prot=4500)
outs = []
for rep in range(3):
    with ka.Engine([0]) as eng:
        if rep == 2: eng.set_option("slot_bits", 64)
        eng.db_load(kmers, roles, 8)
        print(eng.db_info(), flush=True)
        outs.append(eng.annotate(res, off, 5))
want = oracle.OracleDb(kmers, roles, 8, threads=16).apply(res, off, 5, threads=16)
for i, o in enumerate(outs):
    d = [(int((x != y).sum())) for x, y in zip(o, want)]
    print("build", i, "vs oracle diffs (role,hits,flag):", d, flush=True)
    if any(d):
        bad = np.nonzero(o[1] != want[1])[0][:10]
        print("  first bad seqs", bad, "got hits", o[1][bad], "want", want[1][bad], "flags", o[2][bad], want[2][bad])
