"""Property-based GPU parity (hypothesis): arbitrary byte alphabets, K, ragged sequences and DBs
with duplicate / conflicting lines — the engine must equal the pure-Python statement of
ApplyKmerProcessor.java:122-148 on every generated case."""
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
    return K, seqs, kmers, roles, min_hits


@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(case())
def test_engine_equals_python_statement(engine, c):
    K, seqs, kmers, roles, min_hits = c
    res, off = csr(seqs)
    engine.db_load(kmers, np.asarray(roles, np.int32), K)
    got = engine.annotate(res, off, min_hits)
    want = py_apply(seqs, kmers, roles, K, min_hits)
    for g, w in zip(got, want):
        assert np.array_equal(g, w), (K, seqs, kmers, roles, min_hits, g.tolist(), w.tolist())
