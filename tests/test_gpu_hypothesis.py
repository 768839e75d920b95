"""Property-based GPU parity (hypothesis): arbitrary byte alphabets, K, ragged sequences and DBs
with duplicate / conflicting lines — the engine must equal the pure-Python statement of
ApplyKmerProcessor.java:122-148 on every generated case."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from cases import csr, py_apply

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    import kmers_anno_b200 as ka
    eng = ka.Engine([0])
    yield eng
    eng.close()


@st.composite
def case(draw):
    K = draw(st.integers(1, 12))
    n_sym = draw(st.integers(1, 31))
    alphabet = draw(st.lists(st.integers(0, 255), min_size=n_sym, max_size=n_sym, unique=True))
    extra = draw(st.lists(st.integers(0, 255), min_size=0, max_size=4))      # bytes the DB may never use
    sym = st.sampled_from(alphabet)
    # a few motifs reused across sequences so that hits, duplicates and conflicts happen
    motifs = draw(st.lists(st.lists(sym, min_size=K, max_size=K + 6).map(bytes), min_size=1, max_size=6))
    piece = st.one_of(st.sampled_from(motifs), st.lists(st.sampled_from(alphabet + extra), max_size=20).map(bytes))
    seqs = draw(st.lists(st.lists(piece, max_size=6).map(b"".join), min_size=1, max_size=25))
    windows = sorted({m[i:i + K] for m in motifs for i in range(len(m) - K + 1)})
    kmers = draw(st.lists(st.sampled_from(windows), min_size=1, max_size=40))
    roles = draw(st.lists(st.integers(0, 5), min_size=len(kmers), max_size=len(kmers)))
    min_hits = draw(st.integers(1, 4))
    wide = draw(st.integers(0, 1))          # narrow / wide-table kernels
    layout = draw(st.sampled_from([0, 16, 16, 32, 64]))   # 16 = the 128-byte-line table, when the keys fit it
    packed = draw(st.booleans())            # ka_annotate_packed (5-bit stream) instead of ka_annotate
    return K, seqs, kmers, roles, min_hits, wide, layout, packed


@settings(max_examples=int(os.environ.get("KA_HYP_EXAMPLES", 150)), deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(case())
def test_engine_equals_python_statement(engine, c):
    import kmers_anno_b200 as ka
    K, seqs, kmers, roles, min_hits, wide, layout, packed = c
    res, off = csr(seqs)
    if layout == 16:
        wide = 0
    engine.set_option("wide", wide)
    engine.set_option("slot_bits", layout)
    try:
        engine.db_load(kmers, np.asarray(roles, np.int32), K)
    except ka.KmerAnnoError as err:
        # keys of more than 39 bits (or a one-letter alphabet) do not fit the line table (KA_ERR_TOO_BIG), and a forced
        # narrow slot under 12-mers needs a table of ~100 GB + its build scratch, which the HBM left by the other
        # fixtures of the session may not hold (KA_ERR_OOM): a loud error either way, then the default layout
        assert (layout == 16 and err.code == -10) or (layout in (32, 64) and K >= 11 and err.code == -7), err
        engine.set_option("slot_bits", 0)
        engine.db_load(kmers, np.asarray(roles, np.int32), K)
    else:
        assert layout == 0 or engine.db_info()["slot_bits"] == layout
    if packed:
        codes, off32 = engine.pack(res, off, threads=1)
        got = engine.annotate_packed(codes, off32, min_hits)
    else:
        got = engine.annotate(res, off, min_hits)
    want = py_apply(seqs, kmers, roles, K, min_hits)
    for g, w in zip(got, want):
        assert np.array_equal(g, w), (K, seqs, kmers, roles, min_hits, layout, packed, g.tolist(), w.tolist())


@st.composite
def distance_case(draw):
    K = draw(st.integers(1, 12))
    n_sym = draw(st.integers(1, 31))
    alphabet = draw(st.lists(st.integers(0, 255), min_size=n_sym, max_size=n_sym, unique=True))
    sym = st.sampled_from(alphabet)
    motifs = draw(st.lists(st.lists(sym, min_size=K, max_size=K + 8).map(bytes), min_size=1, max_size=5))
    piece = st.one_of(st.sampled_from(motifs), st.lists(sym, max_size=15).map(bytes))
    seqs = draw(st.lists(st.lists(piece, max_size=8).map(b"".join), min_size=1, max_size=12))
    n = len(seqs)
    groups = draw(st.lists(st.tuples(st.integers(0, n - 1), st.lists(st.integers(0, n - 1), max_size=6)), max_size=8))
    return K, seqs, groups


@settings(max_examples=int(os.environ.get("KA_HYP_EXAMPLES", 100)), deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(distance_case())
def test_distance_equals_python_sets(engine, c):
    """ka_kmer_distance against Python sets (GeneCopyProcessor.java:137-142, recalled ProteinKmers.distance)."""
    K, seqs, groups = c
    res, off = csr(seqs)
    q = np.asarray([g[0] for g in groups], np.uint32)
    cs = np.asarray([x for g in groups for x in g[1]], np.uint32)
    go = np.concatenate([[0], np.cumsum([len(g[1]) for g in groups])]).astype(np.uint64)
    size, common, dist = engine.kmer_distance(res, off, K, q, go, cs)
    sets = [{s[i:i + K] for i in range(len(s) - K + 1)} for s in seqs]
    assert [int(x) for x in size] == [len(t) for t in sets]
    m = 0
    for g in groups:
        for x in g[1]:
            a, b = sets[g[0]], sets[x]
            sim = len(a & b)
            assert common[m] == sim and dist[m] == (1.0 if sim == 0 else 1.0 - sim / ((len(a) + len(b)) - sim)), (K, seqs, groups)
            m += 1
