"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol,
the C++ mirror of the reference's genome readers / reporters behaves like the Java it
mirrors, the CLI fails loudly without a GPU, and the multi-rank sharding is loss-free."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
BIN = os.path.join(ROOT, "kmers.anno_b200", "bin")
ROLES_FILE = os.path.join(GOLD, "small.roles.in.use")


def test_abi_library_exports_every_declared_symbol():
    import ctypes
    from kmers_anno_b200.engine import ABI_SYMBOLS, LIB_PATH
    header = open(os.path.join(ROOT, "include", "kmeranno.h")).read()
    declared = set(re.findall(r"\b(ka_[a-z_]+)\s*\(", header))
    assert declared == set(ABI_SYMBOLS), declared ^ set(ABI_SYMBOLS)
    lib = ctypes.CDLL(LIB_PATH)
    for name in declared:
        assert getattr(lib, name) is not None
    lib.ka_abi_version.restype = ctypes.c_int
    assert lib.ka_abi_version() == 2      # no compute call: fine without a GPU


def test_jni_shim_compiles_against_a_stub_jni_header():
    """No JDK in this image: the JNI shim is at least syntax- and type-checked (gcc -fsyntax-only -Wall -Werror)
    against tests/stub_jni/jni.h, and every ka_* function it calls is a declared entry point."""
    src = os.path.join(ROOT, "kmers.anno_b200", "java", "jni", "kmerengine_jni.c")
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "tests", "stub_jni"),
                        "-I", os.path.join(ROOT, "include"), src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    from kmers_anno_b200.engine import ABI_SYMBOLS
    used = set(re.findall(r"\b(ka_[a-z_]+)\s*\(", open(src).read()))
    assert used and used <= set(ABI_SYMBOLS), used - set(ABI_SYMBOLS)
    # the Java class declares one native method per shim function
    java = open(os.path.join(ROOT, "kmers.anno_b200", "java", "org", "theseed", "proteins", "kmers", "gpu", "KmerEngine.java")).read()
    natives = set(re.findall(r"native\s+[\w\[\]]+\s+(\w+)\(", java))
    shims = set(re.findall(r"Java_org_theseed_proteins_kmers_gpu_KmerEngine_(\w+)\(", open(src).read()))
    assert natives == shims, natives ^ shims


def test_mixed_kmer_lengths_are_rejected_before_any_gpu_work(tmp_path):
    """HashMap<String,String> takes k-mers of any length (ApplyKmerProcessor.java:106); one packed table has one K:
    the C++ `apply` refuses such a DB loudly while parsing it (documented deviation, docs/SEMANTICS.md #2)."""
    db = tmp_path / "kmerdb.tbl"
    db.write_text("ACDEFGHI\tRoleA\nACDEFGHIK\tRoleB\n")
    roles = tmp_path / "roles.in.use"
    roles.write_text("RoleA\tsome role\nRoleB\tother role\n")
    gdir = tmp_path / "genomes"
    gdir.mkdir()
    r = subprocess.run([os.path.join(BIN, "kmers-anno"), "apply", str(db), str(roles), str(gdir)], capture_output=True, text=True)
    assert r.returncode == 1 and "mixed k-mer lengths (8 and 9)" in r.stderr and r.stdout == ""


def test_no_gpu_means_error_not_fallback():
    from conftest import has_gpu
    if has_gpu():
        pytest.skip("a GPU is present")
    import kmers_anno_b200 as ka
    with pytest.raises(ka.KmerAnnoError) as e:
        ka.Engine([0])
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    r = subprocess.run([os.path.join(BIN, "kmers-anno"), "apply", os.path.join(GOLD, "small.kmerdb.tbl"),
                        os.path.join(GOLD, "small.roles.in.use"), GOLD], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


def test_cli_argument_errors_follow_the_reference():
    cli = os.path.join(BIN, "kmers-anno")
    # validateParms order (ApplyKmerProcessor.java:85-96): directory, db file, min hits, roles file
    r = subprocess.run([cli, "apply", "nodb", "noroles", "nodir"], capture_output=True, text=True)
    assert r.returncode == 1 and "Input directory nodir not found or invalid." in r.stderr
    r = subprocess.run([cli, "apply", "nodb", "noroles", GOLD], capture_output=True, text=True)
    assert r.returncode == 1 and "Kmer database file nodb not found or unreadable." in r.stderr
    db = os.path.join(GOLD, "small.kmerdb.tbl")
    r = subprocess.run([cli, "apply", "-m", "0", db, "noroles", GOLD], capture_output=True, text=True)
    assert r.returncode == 1 and "Min-hits must be positive." in r.stderr
    r = subprocess.run([cli, "apply", db, "noroles", GOLD], capture_output=True, text=True)
    assert r.returncode == 1 and "Roles-to-use file noroles not found or unreadable." in r.stderr
    r = subprocess.run([cli, "apply", "--format", "TRAIN", db, "x", GOLD], capture_output=True, text=True)
    assert r.returncode == 1 and "--format" in r.stderr
    r = subprocess.run([cli, "apply", db], capture_output=True, text=True)
    assert r.returncode == 1 and "is required" in r.stderr
    r = subprocess.run([cli, "build", ROLES_FILE, "noroles", GOLD], capture_output=True, text=True)
    assert r.returncode == 1 and "Good-role file noroles not found or unreadable." in r.stderr
    r = subprocess.run([cli, "build", ROLES_FILE, ROLES_FILE, "nodir"], capture_output=True, text=True)
    assert r.returncode == 1 and "Genome directory nodir not found or invalid." in r.stderr
    r = subprocess.run([cli, "nosuchverb"], capture_output=True, text=True)
    assert r.returncode == 1 and "Invalid command nosuchverb." in r.stderr


def load_small():
    pegs = []
    with open(os.path.join(GOLD, "small_proteins.tsv")) as fh:
        gid = fh.readline().rstrip("\n").split("\t")[1]
        for line in fh:
            pegs.append(line.rstrip("\n").split("\t"))
    return gid, pegs


def write_gto(path, gid, pegs):
    """A GTO-shaped JSON with the nesting, non-peg features and escapes the reader must survive."""
    feats = [{"id": f"fig|{gid}.rna.1", "type": "rna", "function": "16S rRNA", "location": [["c", "1", "+", 9]]}]
    for fid, fun, prot in pegs:
        feats.append({"type": "CDS", "annotations": [["Add \"feature\"", "PATRIC", 1.5e9, ""]], "aliases": [],
                      "function": fun, "location": [[f"{gid}.con.0001", "1159", "-", 549]],
                      "protein_translation": prot, "family_assignments": [], "id": fid})
    doc = {"domain": "Bacteria", "taxonomy": ["a", "b"], "features": feats, "id": gid,
           "scientific_name": "Test é genome", "contigs": [{"id": "c", "dna": "acgt"}], "genetic_code": 11,
           "close_genomes": [], "nested": {"a": [1, 2, {"b": None}], "t": True}}
    json.dump(doc, open(path, "w"), indent=3)


def test_genome_readers_and_reporters(tmp_path):
    gid, pegs = load_small()
    gto = tmp_path / f"{gid}.gto"
    write_gto(str(gto), gid, pegs)
    faa = tmp_path / f"{gid}.faa"
    with open(faa, "w") as fh:
        for fid, fun, prot in pegs:
            fh.write(f">{fid} {fun}\n")
            for i in range(0, len(prot), 60):
                fh.write(prot[i:i + 60] + "\n")
    # the calls of the golden VERIFY run, replayed through the reporters
    ids = [l.split("\t")[0] for l in open(os.path.join(GOLD, "small.roles.in.use"))]
    exp = np.load(os.path.join(GOLD, "small.expected.npz"))
    calls = tmp_path / "calls.tsv"
    with open(calls, "w") as fh:
        for i in np.nonzero(exp["flag"] == 1)[0]:
            fh.write(f"{i}\t{ids[exp['role'][i]]}\t{exp['hits'][i]}\n")
    st = os.path.join(BIN, "kmers-anno-selftest")
    roles = os.path.join(GOLD, "small.roles.in.use")
    for genome_file in (gto, faa):
        r = subprocess.run([st, str(genome_file), roles, "VERIFY", str(calls)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert f"pegs {len(pegs)}" in r.stderr
        assert r.stdout == open(os.path.join(GOLD, "small.verify.tsv")).read()
        dump = [l.rstrip("\n").split("\t") for l in open(str(genome_file) + ".dump")]
        assert dump == pegs                      # ids, functions and proteins read back exactly, in file order
        r = subprocess.run([st, str(genome_file), roles, "APPLY", str(calls)], capture_output=True, text=True)
        assert r.stdout == open(os.path.join(GOLD, "small.apply.tsv")).read()
    # a role that is not in roles.in.use is dropped by APPLY (getRoleIdx -> 0) and kept by VERIFY
    with open(calls, "w") as fh:
        fh.write("0\tNoSuchRole\t9\n")
    r = subprocess.run([st, str(gto), roles, "APPLY", str(calls)], capture_output=True, text=True)
    assert r.stdout.split("\t")[1:] == ["0"] * (len(ids) - 1) + ["0\n"]
    r = subprocess.run([st, str(gto), roles, "VERIFY", str(calls)], capture_output=True, text=True)
    assert r.stdout.splitlines()[1].split("\t")[2:4] == ["NoSuchRole", "9"]


def test_residue_balanced_cuts():
    from kmers_anno_b200.sharding import residue_balanced_cuts, shard
    rng = np.random.default_rng(3)
    lens = rng.integers(0, 900, size=5000)
    off = np.zeros(len(lens) + 1, np.uint64); off[1:] = np.cumsum(lens)
    res = rng.integers(65, 90, size=int(off[-1]), dtype=np.uint8)
    for parts in (1, 2, 3, 8):
        cuts = residue_balanced_cuts(off, parts)
        assert cuts[0] == 0 and cuts[-1] == len(lens) and (np.diff(cuts) >= 0).all()
        sizes = [int(off[cuts[i + 1]] - off[cuts[i]]) for i in range(parts)]
        assert max(sizes) - min(sizes) <= 2 * 900
        back = [shard(res, off, r, parts) for r in range(parts)]
        assert np.array_equal(np.concatenate([b[0] for b in back]), res)
        assert [b[2] for b in back] == list(cuts[:-1])
    # degenerate: more parts than sequences, empty batch
    assert list(residue_balanced_cuts(np.asarray([0, 5], np.uint64), 4)) in ([0, 0, 0, 0, 1], [0, 0, 0, 1, 1], [0, 1, 1, 1, 1], [0, 0, 1, 1, 1])
    assert list(residue_balanced_cuts(np.asarray([7], np.uint64), 3)) == [0, 0, 0, 0]


WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
from cases import csr, ragged_case
from kmers_anno_b200.sharding import shard
import oracle
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
seqs, kmers, roles = ragged_case(77, n_seq=600, K=8)
res, off = csr(seqs)
# every rank annotates its residue-balanced shard (the oracle stands in for the engine: no GPU here)
my_res, my_off, first = shard(res, off, rank, world)
mine = oracle.OracleDb(kmers, roles, 8).apply(my_res, my_off, 3)
gathered = [None] * world
dist.all_gather_object(gathered, (first, [x.tolist() for x in mine]))
if rank == 0:
    gathered.sort(key=lambda g: g[0])
    whole = oracle.OracleDb(kmers, roles, 8).apply(res, off, 3)
    for j in range(3):
        got = np.concatenate([np.asarray(g[1][j]) for g in gathered])
        assert np.array_equal(got, whole[j]), j
    print("SHARDED_OK", world, len(seqs))
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_gloo_sharding(tmp_path):
    """world_size 2 over gloo: shard -> annotate per rank -> gather == single-process result."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARDED_OK 2 600" in r.stdout


def test_json_document_round_trip(tmp_path):
    """JsonDoc (the GTO document the `genes` command rewrites): parse + serialise keeps every value,
    member order and number text."""
    gid, pegs = load_small()
    src = tmp_path / "in.json"
    write_gto(str(src), gid, pegs[:40])
    doc = json.load(open(src))
    doc["odd"] = {"esc": "tab\t quote\" back\\ nl\n unié中 \U0001F600 ctl\x01", "nums": [0, -1, 1.5e9, 2.5E-3, 1e+2],
                  "empty": [{}, [], ""], "null": None, "bools": [True, False]}
    json.dump(doc, open(src, "w"), indent=2, ensure_ascii=True)      # \uXXXX escapes incl. a surrogate pair
    st = os.path.join(BIN, "kmers-anno-selftest")
    out = tmp_path / "out.json"
    r = subprocess.run([st, "--json", str(src), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    back = json.load(open(out, encoding="utf-8"))
    assert back == doc and list(back.keys()) == list(doc.keys())
    assert "1500000000.0" in open(out, encoding="utf-8").read()      # number text is kept as written
    bad = tmp_path / "bad.json"
    bad.write_text('{"a": [1, 2}')
    r = subprocess.run([st, "--json", str(bad), str(out)], capture_output=True, text=True)
    assert r.returncode == 1 and "JSON error" in r.stderr


def test_genes_cli_validation_and_no_candidates(tmp_path):
    """`genes` (GeneCopyProcessor.java:88-108): option checks and messages; a source genome without
    aliases gives no candidates, no engine call, and the target comes back unchanged."""
    cli = os.path.join(BIN, "kmers-anno")
    st = os.path.join(BIN, "kmers-anno-selftest")
    r = subprocess.run([cli, "genes", "-m", "1.5", "a", "b", "c"], capture_output=True, text=True)
    assert r.returncode == 1 and "Distance must be between 0 and 1." in r.stderr
    r = subprocess.run([cli, "genes", "-K", "1", "a", "b", "c"], capture_output=True, text=True)
    assert r.returncode == 1 and "Kmer size must be at least 2." in r.stderr
    r = subprocess.run([cli, "genes", "a", "b", "c"], capture_output=True, text=True)
    assert r.returncode == 1 and "Input genome file a not found or unreadable." in r.stderr
    r = subprocess.run([cli, "genes", "a", "b"], capture_output=True, text=True)
    assert r.returncode == 1 and "Three arguments are required" in r.stderr
    gid, pegs = load_small()
    write_gto(str(tmp_path / "s.gto"), gid, pegs[:30])
    write_gto(str(tmp_path / "t.gto"), "9.9", [(f.replace(gid, "9.9"), fun, p) for f, fun, p in pegs[:30]])
    r = subprocess.run([cli, "genes", str(tmp_path / "s.gto"), str(tmp_path / "t.gto"), str(tmp_path / "o.gto")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "0 features with aliases, 0 functions found." in r.stderr and "with 0 updates" in r.stderr
    assert json.load(open(tmp_path / "o.gto")) == json.load(open(tmp_path / "t.gto"))
    norm = lambda s: subprocess.run([st, "--norm", s], capture_output=True, text=True).stdout.rstrip("\n")
    assert norm("DNA polymerase III  alpha subunit (EC 2.7.7.7) # frameshift") == "dna polymerase iii alpha subunit ec 2 7 7 7"
    assert norm("Hypothetical protein ! truncated") == norm("hypothetical  PROTEIN")


def test_committed_ncu_exports_feed_the_bench_roofline():
    """bench.py takes roofline.traffic from the committed `ncu --page raw --csv` exports (and refuses to run without
    them): both exports parse, name the kernels of their layout and give plausible DRAM bytes per probe."""
    sys.path.insert(0, ROOT)
    import bench
    per_probe, note, parts = bench.ncu_dram_bytes_per_probe(16)
    assert [p["kernel"].split("<")[0].split()[-1] for p in parts] == ["line_filter_kernel", "line_probe_kernel", "line_tally_kernel"]
    assert 40.0 < per_probe < 70.0 and abs(sum(p["dram_bytes_per_probe"] for p in parts) - per_probe) < 1e-6
    assert max(parts, key=lambda p: p["ncu_ms"])["kernel"].startswith("line_probe_kernel")      # the HBM-bound pass dominates
    assert "profiles/r02_line_passes_raw.csv" in note
    per_probe32, note32, parts32 = bench.ncu_dram_bytes_per_probe(32)
    assert len(parts32) == 1 and "tile_kernel" in parts32[0]["kernel"] and 85.0 < per_probe32 < 105.0
    assert per_probe < 0.6 * per_probe32              # the point of the line table


def test_line_table_key_arithmetic_on_the_host(tmp_path):
    """The key arithmetic of the line table (csrc/ka_line.cuh: Feistel mixer, c-way split of the top bits,
    line_locate, filter hash) is host + device code: tests/native/line_geom_check.cu checks exhaustively on small
    key spaces that the mixer and the split are bijections and that (line, tag) identifies the key."""
    exe = tmp_path / "line_geom_check"
    src = os.path.join(ROOT, "tests", "native", "line_geom_check.cu")
    subprocess.run(["nvcc", "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a",
                    "-I" + os.path.join(ROOT, "kmers.anno_b200", "csrc"), "-o", str(exe), src], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert out.strip() == "ok", out
