#!/usr/bin/env python
"""Option sweep on one table + one resident batch: prints tile-kernel probes/s per config."""
import argparse, itertools, os, sys, time, zlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kmers_anno_b200 as ka
from kmers_anno_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--table-kmers", type=float, default=1e8)
ap.add_argument("--roles", type=int, default=30000)
ap.add_argument("--genomes", type=int, default=200)
ap.add_argument("--K", type=int, default=8)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--configs", default="")   # "k=v,k=v;k=v"
a = ap.parse_args()

fam = synth.Families(a.roles)
t = time.time(); kmers, roles = fam.table(int(a.table_kmers), K=a.K); print(f"table {len(roles)} in {time.time()-t:.1f}s", flush=True)
res, off, _ = fam.batch(0, a.genomes, n_prot=4500, mode=a.mode, K=a.K)
configs = [dict(kv.split("=") for kv in c.split(",") if kv) for c in a.configs.split(";")] if a.configs else [{}]
ref = None
last_db = None
eng = None
for cfg in configs:
    DBOPTS = ("load_factor", "filter", "slot_bits")
    db_key = tuple(cfg.get(k, "") for k in DBOPTS)
    if eng is None or db_key != last_db:
        if eng: eng.close()
        eng = ka.Engine([0])
        for k in DBOPTS:
            if k in cfg: eng.set_option(k, float(cfg[k]))
        t = time.time(); eng.db_load(kmers, roles, a.K); info = eng.db_info(); t_load = time.time() - t
        last_db = db_key
    for k, v in cfg.items():
        if k not in DBOPTS: eng.set_option(k, float(v))
    b = eng.upload(res, off)
    probes = eng.stats()["probes"]
    best = 1e9
    for r in range(a.reps + 1):
        eng.annotate_resident(b, 5)
        st = eng.stats()
        if r: best = min(best, st["tile_kernel_ms"])
    out = eng.download(b); b.free()
    sig = zlib.crc32(out[0].tobytes()) ^ zlib.crc32(out[1].tobytes()) ^ zlib.crc32(out[2].tobytes())
    if ref is None: ref = sig
    print(f"{str(cfg):70s} cls={info['slot_bits']} table={info['table_bytes']/1e6:.0f}MB maxchain={info['max_probe']} "
          f"tile={best:.3f} ms  {probes/best/1e6:.2f} G probes/s  kernel_total={st['kernel_ms']:.3f} ms same={sig==ref} (db load {t_load:.1f}s)", flush=True)
eng.close()
