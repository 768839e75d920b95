#!/usr/bin/env python
"""Free device memory across repeated ka_db_load / annotate cycles of one engine, per table layout and input form."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import kmers_anno_b200 as ka
from cases import csr, ragged_case

def free_mb():
    torch.cuda.synchronize()
    return torch.cuda.mem_get_info(0)[0] / 1e6

torch.zeros(1, device="cuda")
eng = ka.Engine([0])
cases = {K: ragged_case(5 + K, n_seq=40, K=K) for K in (3, 8, 12)}
print("start", round(free_mb()))
for label, K, opts, packed in (("auto K=8 bytes", 8, {}, False), ("auto K=8 packed", 8, {}, True), ("line K=8", 8, {"slot_bits": 16}, True),
                               ("sector32 K=8", 8, {"slot_bits": 32}, False), ("sector64 K=12", 12, {"slot_bits": 64}, True),
                               ("wide K=12", 12, {"wide": 1, "slot_bits": 64}, False), ("line K=3", 3, {"slot_bits": 16}, False),
                               ("failing line K=12", 12, {"slot_bits": 16}, False), ("sector32 K=12", 12, {"slot_bits": 32}, False)):
    seqs, kmers, roles = cases[K]
    res, off = csr(seqs)
    f0 = free_mb()
    for it in range(30):
        eng.set_option("wide", 0); eng.set_option("slot_bits", 0)
        for k, v in opts.items(): eng.set_option(k, v)
        try:
            eng.db_load(kmers, roles, K)
        except ka.KmerAnnoError as err:
            if it == 0: print("   ", label, "->", err)
            continue
        if packed:
            codes, off32 = eng.pack(res, off, threads=1)
            eng.annotate_packed(codes, off32, 2)
        else:
            eng.annotate(res, off, 2)
    print(f"{label:22s} free before {f0:9.0f} MB after 30 cycles {free_mb():9.0f} MB", flush=True)
eng.close()
print("closed", round(free_mb()))
