#!/usr/bin/env python
"""One resident annotate of N proteomes against the 1e8 table with the given options (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kmers_anno_b200 as ka
from kmers_anno_b200 import synth
genomes = int(sys.argv[1]); opts = dict(kv.split("=") for kv in sys.argv[2:])
fam = synth.Families(30000)
kmers, roles = fam.table(int(1e8), K=8)
bseed = int(opts.pop("batch_seed", 0))   # non-zero: proteins from unrelated families (nearly all misses)
res, off, _ = (synth.Families(30000, seed=bseed) if bseed else fam).batch(0, genomes, n_prot=4500)
eng = ka.Engine([0])
for k, v in opts.items(): eng.set_option(k, float(v))
eng.db_load(kmers, roles, 8)
b = eng.upload(res, off)
for _ in range(3):
    eng.annotate_resident(b, 5)
st = eng.stats(); print(opts, "tile ms", st["tile_kernel_ms"], "G probes/s", st["probes"] / st["tile_kernel_ms"] / 1e6)
