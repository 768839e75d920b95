#!/usr/bin/env python
"""GPU `build` (ka_build) vs the oracle's Java-shaped restatement on the config-1 shape:
20 synthetic genomes x 4,500 pegs, 500 good roles (SURVEY.md §8d "Small DB (C1)")."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kmers_anno_b200 as ka
from kmers_anno_b200 import synth
from oracle import binding

genomes = int(sys.argv[1]) if len(sys.argv) > 1 else 20
good = 500
fam = synth.Families(30000)
res, off, true_role = fam.batch(5000, genomes, n_prot=4500)
# classification the Java side would do (Feature.getUsefulRoles ∩ goodRoles): a family protein of one
# of the `good` most frequent roles has one good role, everything else none
n_roles = ((true_role >= 0) & (true_role < good)).astype(np.int32)
peg_role = np.where(n_roles == 1, true_role, -1).astype(np.int32)
print(f"{genomes} genomes, {len(n_roles)} pegs, {len(res)} aa, {int(n_roles.sum())} single-role pegs", flush=True)
with ka.Engine([0]) as eng:
    eng.build(res[:100000], off[:300], n_roles[:299], peg_role[:299], 8)   # warm-up
    best = 1e9
    for _ in range(3):
        t = time.perf_counter(); k, r = eng.build(res, off, n_roles, peg_role, 8); best = min(best, time.perf_counter() - t)
print(f"GPU ka_build (host buffers in, k-mers out): {best*1e3:.1f} ms, {len(r)} k-mers, {len(res)/best/1e6:.0f} M residues/s", flush=True)
t = time.perf_counter(); wk, wr, stats = binding.build_db(res, off, n_roles, peg_role, 8, good); dt = time.perf_counter() - t
print(f"oracle build (1 thread, Java-shaped): {dt:.2f} s, {len(wr)} k-mers, {stats}; speed-up {dt/best:.0f}x", flush=True)
a = {(k[i*8:(i+1)*8].tobytes(), int(r[i])) for i in range(len(r))}
b = {(wk[i*8:(i+1)*8].tobytes(), int(wr[i])) for i in range(len(wr))}
print("same set of lines:", a == b, flush=True)
