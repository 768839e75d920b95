#!/usr/bin/env python
"""File-to-report throughput of the C++ `apply` command (host mirror + ingest pipeline):
N synthetic proteome FASTA files + a kmerdb.tbl on disk -> APPLY report, vs parser threads."""
import os, subprocess, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmers_anno_b200 import synth
n_genomes = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n_kmers = int(float(sys.argv[2])) if len(sys.argv) > 2 else 5_000_000
fam = synth.Families(3000)
kmers, roles = fam.table(n_kmers, K=8)
root = tempfile.mkdtemp(prefix="ka_cli_")
gdir = os.path.join(root, "genomes"); os.mkdir(gdir)
t = time.time()
with open(os.path.join(root, "kmerdb.tbl"), "wb") as fh:
    km = kmers.reshape(-1, 8)
    lines = [km[i].tobytes() + b"\tRole%05d\n" % roles[i] for i in range(len(roles))]
    fh.write(b"".join(lines))
with open(os.path.join(root, "roles.in.use"), "w") as fh:
    for r in range(3000): fh.write(f"Role{r:05d}\trole number {r}\n")
total = 0
for g in range(n_genomes):
    res, off, _ = fam.batch(g, 1, n_prot=4500)
    total += len(res)
    with open(os.path.join(gdir, f"{1000 + g}.1.faa"), "wb") as fh:
        buf = []
        for i in range(4500):
            buf.append(b">fig|%d.1.peg.%d hypothetical protein\n" % (1000 + g, i + 1))
            buf.append(res[int(off[i]):int(off[i + 1])].tobytes()); buf.append(b"\n")
        fh.write(b"".join(buf))
print(f"wrote {n_genomes} FASTA files ({total/1e6:.0f} M aa) and a {len(roles)}-line kmerdb.tbl in {time.time()-t:.1f}s", flush=True)
cli = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kmers.anno_b200", "bin", "kmers-anno")
outs = []
batch = sys.argv[3] if len(sys.argv) > 3 else "100"
os.environ["KA_CLI_TRACE"] = "1"
for threads, batch in ((1, batch), (4, batch), (16, batch), (16, "25"), (16, "50")):
    t = time.time()
    r = subprocess.run([cli, "apply", "--threads", str(threads), "--batch", batch, os.path.join(root, "kmerdb.tbl"),
                        os.path.join(root, "roles.in.use"), gdir], capture_output=True)
    dt = time.time() - t
    assert r.returncode == 0, r.stderr[-500:]
    outs.append(r.stdout)
    marks = [l for l in r.stderr.decode().splitlines() if "t=" in l]
    print("   ", " | ".join(m[-40:] for m in marks[:-1]))
    print("   ", marks[-1])
    print("   ", " | ".join(l for l in r.stderr.decode().splitlines() if l.startswith("[ingest]"))[-600:])
    print(f"--threads {threads:2d} --batch {batch}: {dt:.2f} s wall for the whole command ({n_genomes*4500/dt/1e3:.0f} k proteins/s incl. DB load), report {len(r.stdout)} bytes", flush=True)
print("reports identical:", all(o == outs[0] for o in outs))
