"""GPU: the C++ `apply` command (host mirror of ApplyKmerProcessor + reporters over the C ABI)
must print the golden VERIFY / APPLY reports (oracle-derived, tests/golden/make_golden.py)."""
import os
import shutil
import subprocess

import pytest

from test_host import BIN, GOLD, load_small, write_gto

pytestmark = pytest.mark.gpu
CLI = os.path.join(BIN, "kmers-anno")
DB = os.path.join(GOLD, "small.kmerdb.tbl")
ROLES = os.path.join(GOLD, "small.roles.in.use")


def run(args):
    r = subprocess.run([CLI, "apply"] + args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout, r.stderr


def test_apply_verify_and_default_reports(tmp_path):
    gid, pegs = load_small()
    write_gto(str(tmp_path / f"{gid}.gto"), gid, pegs)
    out, err = run(["--format", "VERIFY", DB, ROLES, str(tmp_path)])
    assert out == open(os.path.join(GOLD, "small.verify.tsv")).read()
    assert "Kmer size is 8." in err and "1 genomes found" in err
    out, _ = run([DB, ROLES, str(tmp_path)])                   # default format = APPLY (:78)
    assert out == open(os.path.join(GOLD, "small.apply.tsv")).read()
    # --min raises the threshold: fewer rows, all with hits >= 300
    out, _ = run(["--format", "VERIFY", "-m", "300", DB, ROLES, str(tmp_path)])
    rows = out.splitlines()[1:]
    assert 0 < len(rows) < 81 and all(int(r.split("\t")[3]) >= 300 for r in rows)


def test_apply_many_genomes_order_and_batching(tmp_path):
    gid, pegs = load_small()
    ids = ["100.1", "100.10", "100.2", "99.5"]                # sorted by file name, as GenomeDirectory does
    for g in ids:
        renamed = [(fid.replace(gid, g), fun, prot) for fid, fun, prot in pegs]
        if g.endswith("2"):
            with open(tmp_path / f"{g}.faa", "w") as fh:       # FASTA input path
                for fid, fun, prot in renamed:
                    fh.write(f">{fid} {fun}\n{prot}\n")
        else:
            write_gto(str(tmp_path / f"{g}.gto"), g, renamed)
    want_rows = open(os.path.join(GOLD, "small.verify.tsv")).read().splitlines()[1:]
    outs = []
    for batch in ("1", "3", "64"):
        out, _ = run(["--format", "VERIFY", "--batch", batch, DB, ROLES, str(tmp_path)])
        outs.append(out)
        lines = out.splitlines()
        assert lines[0] == "genome_id\tpeg_id\trole\thits\tfunction"
        body = lines[1:]
        assert len(body) == len(want_rows) * len(ids)
        for k, g in enumerate(sorted(ids)):
            block = body[k * len(want_rows):(k + 1) * len(want_rows)]
            assert block == [r.replace(gid, g) for r in want_rows]
    assert outs[0] == outs[1] == outs[2]
    out, _ = run(["--batch", "2", DB, ROLES, str(tmp_path)])
    vec = open(os.path.join(GOLD, "small.apply.tsv")).read().split("\t", 1)[1]
    assert out == "".join(f"{g}\t{vec}" for g in sorted(ids))


def test_build_verb_then_apply(tmp_path):
    """`build` over the reference's small genome reproduces the golden kmerdb.tbl (as a set of
    lines; the reference's order is HashMap order) and `apply` on its output the golden report."""
    gid, pegs = load_small()
    gdir = tmp_path / "genomes"
    gdir.mkdir()
    write_gto(str(gdir / f"{gid}.gto"), gid, pegs)
    r = subprocess.run([CLI, "build", ROLES, ROLES, str(gdir)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    got = sorted(r.stdout.splitlines())
    want = sorted(open(DB).read().splitlines())
    assert len(got) == 23796 and got == want
    assert "discriminating kmers remaining" in r.stderr
    db = tmp_path / "kmerdb.tbl"
    db.write_text(r.stdout)
    out, _ = run(["--format", "VERIFY", str(db), ROLES, str(gdir)])
    assert out == open(os.path.join(GOLD, "small.verify.tsv")).read()
    # -K changes the k-mer length of the DB and apply follows it
    r = subprocess.run([CLI, "build", "-K", "10", ROLES, ROLES, str(gdir)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and all(len(l.split("\t")[0]) == 10 for l in r.stdout.splitlines())
    db.write_text(r.stdout)
    _, err = run([str(db), ROLES, str(gdir)])
    assert "Kmer size is 10." in err


def test_genes_copies_aliases_of_the_closest_feature(tmp_path):
    """`genes` end to end (GeneCopyProcessor.java:110-166): the target pegs get the aliases of the
    closest same-function source peg within maxDist; expectation from the oracle's distances."""
    import json
    import numpy as np
    import oracle
    from cases import csr
    rng = np.random.default_rng(12)
    gid, pegs = load_small()
    pegs = pegs[:120]

    def mutate(p, rate):
        a = bytearray(p.encode())
        for i in range(len(a)):
            if rng.random() < rate:
                a[i] = b"ACDEFGHIKLMNPQRSTVWY"[int(rng.integers(20))]
        return a.decode()

    src_feats, tgt_feats = [], []
    for i, (fid, fun, prot) in enumerate(pegs):
        fun = fun if i % 5 else "hypothetical protein"              # a function shared by many pegs
        f = {"id": fid, "type": "CDS", "function": fun + (" # src note" if i % 3 == 0 else ""), "protein_translation": prot}
        if i % 4 != 3:
            f["alias_pairs"] = [["gene_name", f"gen{i}"], ["locus_tag", f"b{i:04d}"], ["gene_name", f"alt{i}"]]
        src_feats.append(f)
        rate = [0.0, 0.02, 0.08, 0.5][i % 4]
        t = {"id": fid.replace(gid, "9.9"), "type": "CDS", "function": fun.upper() if i % 2 else fun,
             "protein_translation": mutate(prot, rate)}
        if i % 10 == 0:
            t["alias_pairs"] = [["locus_tag", f"b{i:04d}"]]         # already has one of the aliases
        tgt_feats.append(t)
    src_feats.append({"id": f"fig|{gid}.rna.1", "type": "rna", "function": "tRNA", "alias_pairs": [["gene_name", "rrn"]]})
    json.dump({"id": gid, "scientific_name": "source", "features": src_feats}, open(tmp_path / "s.gto", "w"))
    open(tmp_path / "t.gto", "w").write(json.dumps({"id": "9.9", "scientific_name": "target", "features": tgt_feats,
                                                     "extra": {"keep": [1, 2.5]}}).replace("2.5]", "2.50]"))

    def norm(s):
        s = s.split(" #")[0]
        return " ".join("".join(c.lower() if c.isalnum() else " " for c in s).split())

    for K, max_dist in ((8, 0.5), (10, 0.2)):
        out_file = tmp_path / f"o{K}.gto"
        r = subprocess.run([CLI, "genes", "-K", str(K), "-m", str(max_dist), str(tmp_path / "s.gto"), str(tmp_path / "t.gto"),
                            str(out_file)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        out = json.load(open(out_file))
        assert out["extra"] == {"keep": [1, 2.5]} and "2.50" in open(out_file).read()
        # expectation: same normalised function, closest within max_dist, later candidate wins ties
        by_fun = {}
        for f in src_feats:
            if ".peg." in f["id"] and f.get("alias_pairs"):
                by_fun.setdefault(norm(f["function"]), []).append(f)
        updates = 0
        for t_in, t_out in zip(tgt_feats, out["features"]):
            cands = by_fun.get(norm(t_in["function"]), [])
            want = [list(p) for p in t_in.get("alias_pairs", [])]
            if cands:
                seqs = [t_in["protein_translation"].encode()] + [c["protein_translation"].encode() for c in cands]
                res, off = csr(seqs)
                _, _, _, dist = oracle.kmer_distance_pairs(res, off, np.zeros(len(cands), np.uint32),
                                                           np.arange(1, len(cands) + 1, dtype=np.uint32), K)
                found, fdist = None, max_dist
                for c, d in zip(cands, dist):
                    if d <= fdist:
                        fdist, found = d, c
                if found is not None:
                    updates += 1
                    amap = {}
                    for ty, al in found["alias_pairs"]:
                        amap.setdefault(ty, set()).add(al)
                    for ty in sorted(amap):
                        for al in sorted(amap[ty]):
                            if [ty, al] not in want:
                                want.append([ty, al])
            assert t_out.get("alias_pairs", []) == want, t_in["id"]
        assert f"with {updates} updates" in r.stderr and updates > 20


@pytest.mark.parametrize("mode", ["1", "2", "3"])
def test_apply_with_a_sharded_table(tmp_path, mode):
    """`apply --devices 0,1 --table-mode 1|2`: the table sharded over two GPUs gives the same reports."""
    import kmers_anno_b200 as ka
    try:
        ka.Engine([0, 1]).close()
    except ka.KmerAnnoError:
        pytest.skip("needs at least 2 GPUs")
    gid, pegs = load_small()
    write_gto(str(tmp_path / f"{gid}.gto"), gid, pegs)
    out, err = run(["--format", "VERIFY", "--devices", "0,1", "--table-mode", mode, DB, ROLES, str(tmp_path)])
    want = open(os.path.join(GOLD, "small.verify.tsv")).read()
    assert out == want, [l for l in out.splitlines() if l not in want.splitlines()][:5]
    assert "loaded on 2" in err
