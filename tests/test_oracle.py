"""CPU tests of the oracle (no GPU): the reference-held pins it CAN be checked against
(RoleTests.java:15-36; the substring property of AppTest.java:145-161), public Java facts
about String.hashCode / HashMap, agreement of its three independent statements (Java-shaped
C, packed-integer C, pure Python), and the committed golden fixtures."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from oracle import binding
from cases import csr, py_apply, ragged_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_role_counter_reference_vector():
    """RoleTests.java:15-36, assertion by assertion."""
    L = oracle.lib()
    c = binding.RoleCounter()
    A, B = 1, 2
    L.orc_role_counter_init(C.byref(c), A)
    assert (c.role_id, c.good, c.bad, L.orc_role_counter_is_good(C.byref(c))) == (A, 0, 0, 1)
    L.orc_role_counter_count(C.byref(c), A)
    assert (c.role_id, c.good, c.bad, L.orc_role_counter_is_good(C.byref(c))) == (A, 1, 0, 1)
    L.orc_role_counter_count(C.byref(c), A)
    assert (c.role_id, c.good, c.bad, L.orc_role_counter_is_good(C.byref(c))) == (A, 2, 0, 1)
    L.orc_role_counter_count(C.byref(c), B)
    assert (c.role_id, c.good, c.bad, L.orc_role_counter_is_good(C.byref(c))) == (A, 2, 1, 0)


def load_small_proteins():
    pegs = []
    with open(os.path.join(GOLD, "small_proteins.tsv")) as fh:
        genome_id = fh.readline().rstrip("\n").split("\t")[1]
        for line in fh:
            fid, fun, prot = line.rstrip("\n").split("\t")
            pegs.append((fid, fun, prot))
    return genome_id, pegs


def test_count_peg_kmers_property_on_small_gto():
    """AppTest.java:145-161 (testKmerPegCounts): every k-mer reported at 1-based `left` equals
    prot.substring(left-1, left-1+8) and holds no 'X' / '*'; the loop `i < L-K`
    (KmerReference.java:134-137) drops the last window."""
    L = oracle.lib()
    _, pegs = load_small_proteins()
    assert len(pegs) == 712 and sum(len(p) for _, _, p in pegs) == 221060  # SURVEY §4 fixture facts
    total = 0
    for _, _, prot in pegs:
        b = prot.encode("latin-1")
        pos = np.zeros(max(len(b), 1), np.uint32)
        n = L.orc_count_peg_kmers_positions(b, len(b), 8, pos.ctypes.data)
        for left in pos[:n]:
            kmer = prot[left - 1: left - 1 + 8]
            assert len(kmer) == 8 and "X" not in kmer and "*" not in kmer
        want = [i + 1 for i in range(max(len(prot) - 8, 0)) if "X" not in prot[i:i + 8]]
        assert list(pos[:n]) == want
        total += n
    # ProteinKmers keeps the last window (recalled semantics, the oracle default): one more
    # window per protein than countPegKmers
    offs = np.zeros(len(pegs) + 1, np.uint64)
    offs[1:] = np.cumsum([len(p) for _, _, p in pegs])
    assert oracle.count_probes(offs, 8) == 216076            # 216,076 8-mers
    assert oracle.count_probes(offs, 8) - total == len(pegs)  # no X in the fixture


def jhash(s):
    h = 0
    for ch in s.encode("latin-1"):
        h = (31 * h + ch) & 0xFFFFFFFF
    return h


def test_java_hashmap_restatement():
    L = oracle.lib()
    # public Java facts: "Aa".hashCode() == "BB".hashCode() == 2112, "hello" -> 99162322
    assert jhash("Aa") == jhash("BB") == 2112 and jhash("hello") == 99162322
    m = L.orc_map_new(-1)  # new HashMap<>()
    keys = [b"BB", b"Aa", b"b", b"a", b"hello", b"c"]
    for i, k in enumerate(keys):
        assert L.orc_map_put(m, k, len(k), i) == 1
    assert L.orc_map_put(m, b"Aa", 2, 77) == 0           # existing key: value replaced
    assert L.orc_map_capacity(m) == 16 and L.orc_map_size(m) == 6
    v = C.c_int32()
    assert L.orc_map_get(m, b"Aa", 2, C.byref(v)) == 1 and v.value == 77
    assert L.orc_map_get(m, b"zz", 2, C.byref(v)) == 0
    out = np.zeros(64, np.uint8); kl = np.zeros(8, np.uint32); vals = np.zeros(8, np.int32)
    n = L.orc_map_dump(m, out.ctypes.data, kl.ctypes.data, vals.ctypes.data)
    got, o = [], 0
    for i in range(n):
        got.append(out[o:o + kl[i]].tobytes()); o += kl[i]
    # iteration = bin index of (h ^ h>>>16) & 15, collisions in insertion order
    def bin_of(k):
        h = jhash(k.decode()); return (h ^ (h >> 16)) & 15
    want = sorted(keys, key=lambda k: (bin_of(k), keys.index(k)))
    assert got == want and got.index(b"BB") + 1 == got.index(b"Aa")
    # growth: 13th entry doubles the table (threshold 12) and keeps every mapping
    for i in range(20):
        k = f"key{i}".encode()
        L.orc_map_put(m, k, len(k), 100 + i)
    assert L.orc_map_capacity(m) == 64 and L.orc_map_size(m) == 26
    for i in range(20):
        k = f"key{i}".encode()
        assert L.orc_map_get(m, k, len(k), C.byref(v)) == 1 and v.value == 100 + i
    assert L.orc_map_remove(m, b"hello", 5) == 1 and L.orc_map_remove(m, b"hello", 5) == 0
    assert L.orc_map_size(m) == 25
    L.orc_map_free(m)
    m = L.orc_map_new(100)                                # tableSizeFor(100) = 128
    L.orc_map_put(m, b"x", 1, 0)
    assert L.orc_map_capacity(m) == 128
    L.orc_map_free(m)


@pytest.mark.parametrize("K", [1, 3, 8, 12, 15])
def test_three_statements_agree(K):
    seqs, kmers, roles = ragged_case(900 + K, n_seq=250, K=K, max_len=400)
    res, off = csr(seqs)
    for min_hits in (1, 4):
        a = oracle.OracleDb(kmers, roles, K).apply(res, off, min_hits)
        b = py_apply(seqs, kmers, roles, K, min_hits)
        # the packed port needs (n_sym+1)^K < 2^64; the Java-shaped oracle takes any K
        c = oracle.FastDb(kmers, roles, K).apply(res, off, min_hits, threads=3) if K <= 12 else a
        for x, y, z in zip(a, b, c):
            assert np.array_equal(x, y) and np.array_equal(x, z)


def test_semantic_switches():
    """The recalled points of ProteinKmers each have a switch."""
    prot = b"ACDEFGHIK" * 3                      # 27 aa, 20 windows, 9 distinct 8-mers
    kmers = [prot[i:i + 8] for i in range(9)]
    roles = np.zeros(9, np.int32)
    res, off = csr([prot])
    db = oracle.OracleDb(kmers, roles, 8)
    assert db.apply(res, off, 1)[1][0] == 9                          # distinct windows (default)
    assert db.apply(res, off, 1, distinct=False)[1][0] == 20         # positional count
    assert db.apply(res, off, 1, include_last=False)[1][0] == 9      # the last window is a repeat here
    res2, off2 = csr([b"ACDEFGHIKL"])                                # 3 windows
    db2 = oracle.OracleDb([b"ACDEFGHI", b"CDEFGHIK", b"DEFGHIKL"], np.zeros(3, np.int32), 8)
    assert db2.apply(res2, off2, 1)[1][0] == 3
    assert db2.apply(res2, off2, 1, include_last=False)[1][0] == 2   # countPegKmers-style loop


def test_threads_do_not_change_results():
    seqs, kmers, roles = ragged_case(5, n_seq=400, K=8)
    res, off = csr(seqs)
    db = oracle.OracleDb(kmers, roles, 8)
    a = db.apply(res, off, 3, threads=1)
    b = db.apply(res, off, 3, threads=7)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_build_rule():
    """BuildKmerProcessor.java:138-223 on a hand-made case."""
    K = 4
    pegs = [(b"AAAACCCC", 1, 0),      # role 0
            (b"CCCCDDDD", 1, 1),      # role 1 shares CCCC with role 0 -> non-unique, pruned
            (b"EEEEFFFF", 0, -1),     # no good role: buffered, its k-mers are deleted in pass 2
            (b"GGGGEEEF", 1, 0),      # role 0; EEEF also occurs in the buffered peg
            (b"HHHHIIII", 2, -1)]     # two good roles: ignored entirely
    res, off = csr([p[0] for p in pegs])
    out = binding.build_db(res, off, [p[1] for p in pegs], [p[2] for p in pegs], K, 2)
    kmers, roles, stats = out
    got = {kmers[i * K:(i + 1) * K].tobytes(): int(roles[i]) for i in range(len(roles))}
    want = {}
    for w in (b"AAAA", b"AAAC", b"AACC", b"ACCC", b"GGGG", b"GGGE", b"GGEE", b"GEEE"):
        want[w] = 0
    for w in (b"CCCD", b"CCDD", b"CDDD", b"DDDD"):
        want[w] = 1
    assert got == want
    assert stats["buffered"] == 1 and stats["non_unique"] == 1 and stats["deleted_pass2"] == 1
    # Java's `goodRoles.size() * 700000` is int arithmetic: 3068 roles overflow to a negative
    # capacity and HashMap throws; the restatement reports that as None
    assert binding.build_db(res, off, [p[1] for p in pegs], [p[2] for p in pegs], K, 3068) is None


def test_golden_fixtures_regression():
    """oracle-derived goldens (tests/golden/make_golden.py): kmerdb.tbl + proteins -> VERIFY."""
    genome_id, pegs = load_small_proteins()
    ids = [l.split("\t")[0] for l in open(os.path.join(GOLD, "small.roles.in.use"))]
    idx = {r: i for i, r in enumerate(ids)}
    kmers, roles = [], []
    for line in open(os.path.join(GOLD, "small.kmerdb.tbl")):
        k, r = line.rstrip("\n").split("\t")
        kmers.append(k.encode()); roles.append(idx[r])
    res, off = csr([p.encode() for _, _, p in pegs])
    role, hits, flag = oracle.OracleDb(kmers, np.asarray(roles, np.int32), 8).apply(res, off, 5)
    exp = np.load(os.path.join(GOLD, "small.expected.npz"))
    assert np.array_equal(role, exp["role"]) and np.array_equal(hits, exp["hits"]) and np.array_equal(flag, exp["flag"])
    lines = ["genome_id\tpeg_id\trole\thits\tfunction"]
    for i, (fid, fun, _) in enumerate(pegs):
        if flag[i] == 1:
            lines.append(f"{genome_id}\t{fid}\t{ids[role[i]]}\t{hits[i]}\t{fun}")
    assert "\n".join(lines) + "\n" == open(os.path.join(GOLD, "small.verify.tsv")).read()


def test_kmer_distance_restatement_against_python_sets():
    """oracle.kmer_distance_pairs (GeneCopyProcessor.java:137-142, recalled ProteinKmers.distance)
    against Python sets on the proteins of small.gto."""
    prots = [p.encode() for _, _, p in load_small_proteins()[1]][:60] + [b"", b"ACD", b"ACDEFGHI"]
    res, off = csr(prots)
    rng = np.random.default_rng(4)
    qa = rng.integers(0, len(prots), 400).astype(np.uint32)
    qb = rng.integers(0, len(prots), 400).astype(np.uint32)
    qb[:40] = qa[:40]
    for K in (2, 8, 10):
        sa, sb, co, dist = oracle.kmer_distance_pairs(res, off, qa, qb, K)
        sets = [{p[i:i + K] for i in range(len(p) - K + 1)} for p in prots]
        for m in range(400):
            a, b = sets[qa[m]], sets[qb[m]]
            sim = len(a & b)
            assert (sa[m], sb[m], co[m]) == (len(a), len(b), sim)
            assert dist[m] == (1.0 if sim == 0 else 1.0 - sim / ((len(a) + len(b)) - sim))
