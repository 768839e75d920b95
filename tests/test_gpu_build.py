"""GPU `build` (ka_build, BuildKmerProcessor.java:138-223) against the oracle's restatement:
the same SET of (k-mer, role) lines (the reference's line order is HashMap iteration order,
i.e. unspecified for consumers), and build -> apply end to end."""
import os

import numpy as np
import pytest

from cases import csr, random_seq
from test_gpu_parity import assert_same

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ka():
    import kmers_anno_b200 as ka
    return ka


def as_set(kmers, roles, K):
    kmers = np.asarray(kmers, np.uint8).reshape(-1, K)
    return {(kmers[i].tobytes(), int(roles[i])) for i in range(len(roles))}


def training_set(seed, n_pegs=400, n_roles=9, K=8):
    """Pegs that share segments so that every rule fires: k-mers under two roles, k-mers that
    also occur in role-less pegs, pegs with two good roles, short pegs."""
    rng = np.random.default_rng(seed)
    motifs = [random_seq(rng, int(rng.integers(K, 40))) for _ in range(60)]
    seqs, n_r, role = [], [], []
    for i in range(n_pegs):
        parts = []
        for _ in range(int(rng.integers(0, 6))):
            parts.append(motifs[int(rng.integers(0, len(motifs)))] if rng.random() < 0.5
                         else random_seq(rng, int(rng.integers(1, 60))))
        s = b"".join(parts)
        if rng.random() < 0.1:
            s = s[: int(rng.integers(0, K))]
        seqs.append(s)
        u = rng.random()
        if u < 0.55:
            n_r.append(1); role.append(int(rng.integers(0, n_roles)))
        elif u < 0.9:
            n_r.append(0); role.append(-1)
        else:
            n_r.append(int(rng.integers(2, 4))); role.append(-1)
    return seqs, np.asarray(n_r, np.int32), np.asarray(role, np.int32)


@pytest.mark.parametrize("K", [3, 8, 12])
def test_build_matches_oracle(ka, K):
    from oracle import binding
    seqs, n_r, role = training_set(40 + K, K=K)
    res, off = csr(seqs)
    want_k, want_r, stats = binding.build_db(res, off, n_r, role, K, 9)
    with ka.Engine([0]) as eng:
        got_k, got_r = eng.build(res, off, n_r, role, K)
    assert stats["non_unique"] > 0 and stats["deleted_pass2"] > 0 and stats["remaining"] > 0
    assert len(got_r) == len(want_r)
    assert as_set(got_k, got_r, K) == as_set(want_k, want_r, K)


def test_build_small_gto_golden(ka):
    """The golden kmerdb.tbl of tests/golden (oracle build over the pegs of the reference's
    small.gto, tests/golden/make_golden.py): the GPU build must produce the same lines."""
    import re
    from test_oracle import load_small_proteins
    _, pegs = load_small_proteins()
    names = [l.rstrip("\n").split("\t")[1] for l in open(os.path.join(GOLD, "small.roles.in.use"))]
    ids = [l.split("\t")[0] for l in open(os.path.join(GOLD, "small.roles.in.use"))]
    idx = {n: i for i, n in enumerate(names)}

    def roles_of_function(fun):   # same recalled rule as make_golden.py
        fun = re.split(r"\s*[#!]", fun, maxsplit=1)[0]
        return [r.strip() for r in re.split(r"\s+/\s+|\s+@\s+|;\s+", fun) if r.strip()]

    n_r = np.zeros(len(pegs), np.int32); role = np.full(len(pegs), -1, np.int32)
    for i, (_, fun, _) in enumerate(pegs):
        good = [r for r in roles_of_function(fun) if r in idx]
        n_r[i] = len(good)
        if len(good) == 1:
            role[i] = idx[good[0]]
    res, off = csr([p.encode() for _, _, p in pegs])
    with ka.Engine([0]) as eng:
        got_k, got_r = eng.build(res, off, n_r, role, 8)
    want = set()
    for line in open(os.path.join(GOLD, "small.kmerdb.tbl")):
        k, r = line.rstrip("\n").split("\t")
        want.add((k.encode(), ids.index(r)))
    assert len(got_r) == 23796 and as_set(got_k, got_r, 8) == want


def test_build_then_apply_chain(ka):
    """build with load_as_db, then annotate the training pegs: equals oracle build + oracle apply."""
    import oracle
    from oracle import binding
    seqs, n_r, role = training_set(7, n_pegs=600)
    res, off = csr(seqs)
    want_k, want_r, _ = binding.build_db(res, off, n_r, role, 8, 9)
    with ka.Engine([0]) as eng:
        got_k, got_r = eng.build(res, off, n_r, role, 8, load_as_db=True)
        assert eng.db_info()["n_keys"] == len(want_r)
        got = eng.annotate(res, off, 2)
    want = oracle.OracleDb(want_k, want_r, 8).apply(res, off, 2)
    assert_same(got, want, "build -> apply")
    assert (got[2] == 1).sum() > 20


def test_build_errors(ka):
    res, off = csr([b"ACDEFGHIKL"])
    with ka.Engine([0]) as eng:
        with pytest.raises(ka.KmerAnnoError) as e:
            eng.build(res, off, [1], [-3], 8)
        assert e.value.code == -8
        with pytest.raises(ka.KmerAnnoError) as e:
            eng.build(res, off, [1], [0], 13)
        assert e.value.code == -5
        k, r = eng.build(res, off, [2], [0], 8)            # a two-role peg contributes nothing
        assert len(r) == 0
        k, r = eng.build(res, off, [1], [5], 8)
        assert as_set(k, r, 8) == {(b"ACDEFGHI", 5), (b"CDEFGHIK", 5), (b"DEFGHIKL", 5)}
