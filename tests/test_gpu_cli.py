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
