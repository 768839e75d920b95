"""GPU parity of the pairwise k-mer distance (GeneCopyProcessor.java:137-142) against the oracle:
set sizes and similarities exact, the double distance bit for bit (tolerance 0: it is one IEEE
division and one subtraction of exactly representable integers)."""
import numpy as np
import pytest

from cases import csr, random_seq

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ka():
    import kmers_anno_b200 as ka
    return ka


@pytest.fixture(scope="module")
def oracle():
    import oracle
    return oracle


def mutate(rng, s, rate):
    a = bytearray(s)
    for i in range(len(a)):
        if rng.random() < rate:
            a[i] = b"ACDEFGHIKLMNPQRSTVWY"[int(rng.integers(20))]
    return bytes(a)


def family_case(seed, n_fam=40, per_fam=6, max_len=900, odd=True):
    """Protein families (ancestor + mutated members, a few truncated / extended) plus strays."""
    rng = np.random.default_rng(seed)
    seqs, fam = [], []
    for f in range(n_fam):
        anc = random_seq(rng, int(rng.integers(20, max_len)))
        for m in range(per_fam):
            s = mutate(rng, anc, float(rng.choice([0.0, 0.02, 0.1, 0.3])))
            if m % 3 == 1:
                s = s[: max(1, len(s) // 2)]
            if m % 3 == 2:
                s = s + random_seq(rng, 40)
            seqs.append(s); fam.append(f)
    seqs += [b"", b"ACD", b"ACDEFGHI", b"ACDEFGHIACDEFGHIACDEFGHI", b"MKV" * 100]
    fam += [-1] * 5
    if odd:
        seqs += [b"acdea" * 10, b"ACDEFXXXGHIKL*", b"ACDEFGHIKLMNPQRSTVWYBZ"]
        fam += [-1] * 3
    return seqs, np.asarray(fam)


def groups_for(fam, rng, extra=3):
    """Every sequence is a query; candidates = its family members (incl. itself) + random others."""
    n = len(fam)
    q, go, cs = [], [0], []
    for i in range(n):
        cand = [j for j in range(n) if fam[j] == fam[i] and fam[i] >= 0] + [int(x) for x in rng.integers(0, n, extra)]
        if i % 11 == 0:
            cand = []                                  # a query without candidates
        q.append(i); cs += cand; go.append(len(cs))
    return np.asarray(q, np.uint32), np.asarray(go, np.uint64), np.asarray(cs, np.uint32)


def check(ka, oracle, seqs, K, q, go, cs, devices=(0,)):
    res, off = csr(seqs)
    with ka.Engine(list(devices)) as eng:
        size, common, dist = eng.kmer_distance(res, off, K, q, go, cs)
    qa = np.repeat(q, np.diff(go).astype(np.int64))
    sa, sb, co, dd = oracle.kmer_distance_pairs(res, off, qa, cs, K)
    assert np.array_equal(common, co), np.nonzero(common != co)[0][:8]
    assert np.array_equal(size[qa], sa) and np.array_equal(size[cs], sb)
    assert np.array_equal(dist.view(np.uint64), dd.view(np.uint64))      # bit-exact doubles
    return size, common, dist


@pytest.mark.parametrize("K", [1, 2, 5, 8, 10, 12])
def test_distance_all_k(ka, oracle, K):
    seqs, fam = family_case(K)
    q, go, cs = groups_for(fam, np.random.default_rng(K))
    size, common, dist = check(ka, oracle, seqs, K, q, go, cs)
    assert dist.min() == 0.0 and dist.max() == 1.0       # identical and disjoint pairs both occur


def test_distance_against_pure_python(ka, oracle):
    seqs, fam = family_case(77, n_fam=10, per_fam=4, max_len=200)
    q, go, cs = groups_for(fam, np.random.default_rng(1))
    size, common, dist = check(ka, oracle, seqs, 8, q, go, cs)
    sets = [{s[i:i + 8] for i in range(len(s) - 7)} for s in seqs]
    qa = np.repeat(q, np.diff(go).astype(np.int64))
    for m in range(len(cs)):
        a, b = sets[qa[m]], sets[cs[m]]
        sim = len(a & b)
        want = 1.0 if sim == 0 else 1.0 - sim / ((len(a) + len(b)) - sim)
        assert common[m] == sim and dist[m] == want
    assert [int(x) for x in size] == [len(s) for s in sets]


def test_distance_long_sequences_use_global_sets(ka, oracle):
    """Sequences beyond 4096 windows do not fit the shared-memory set: global scratch path."""
    rng = np.random.default_rng(9)
    big = random_seq(rng, 30000)
    seqs = [big, mutate(rng, big, 0.05), big[:12000], random_seq(rng, 5000), random_seq(rng, 300), big[100:400], b"MK" * 6000]
    n = len(seqs)
    q = np.arange(n, dtype=np.uint32)
    cs = np.tile(np.arange(n, dtype=np.uint32), n)
    go = np.arange(n + 1, dtype=np.uint64) * n
    size, common, dist = check(ka, oracle, seqs, 8, q, go, cs)
    assert size[6] == 2 and common[0] == size[0]


def test_distance_closest_feature_rule(ka, oracle):
    """The reference's selection loop (GeneCopyProcessor.java:139-146) on top of the distances:
    `f2Dist <= fDist` lets a LATER candidate win a tie."""
    base = b"MKVLAAGIVALLLAGCSSAPKEDTSWVRLHNQ" * 4
    seqs = [base, base, b"W" * 50, base]
    res, off = csr(seqs)
    with ka.Engine([0]) as eng:
        _, _, dist = eng.kmer_distance(res, off, 8, [0], [0, 3], [1, 2, 3])
    found, fdist = None, 0.5
    for j, d in enumerate(dist):
        if d <= fdist:
            fdist, found = d, j
    assert found == 2 and fdist == 0.0 and dist[1] == 1.0


def test_distance_multi_device(ka, oracle):
    n_dev = 0
    for n in (4, 2):
        try:
            ka.Engine(list(range(n))).close(); n_dev = n; break
        except ka.KmerAnnoError:
            continue
    if n_dev < 2:
        pytest.skip("needs at least 2 GPUs")
    seqs, fam = family_case(5, n_fam=60)
    q, go, cs = groups_for(fam, np.random.default_rng(2))
    check(ka, oracle, seqs, 8, q, go, cs, devices=range(n_dev))


def test_distance_errors(ka):
    res, off = csr([b"ACDEFGHIKL", b"ACDEFGHIKL"])
    with ka.Engine([0]) as eng:
        with pytest.raises(ka.KmerAnnoError):
            eng.kmer_distance(res, off, 13, [0], [0, 1], [1])           # K beyond the 5-bit packing
        with pytest.raises(ka.KmerAnnoError):
            eng.kmer_distance(res, off, 8, [5], [0, 1], [1])            # query outside the batch
        with pytest.raises(ka.KmerAnnoError):
            eng.kmer_distance(res, off, 8, [0], [0, 1], [7])            # candidate outside the batch
        many = bytes(range(40, 80))
        r2, o2 = csr([many, many])
        with pytest.raises(ka.KmerAnnoError):
            eng.kmer_distance(r2, o2, 8, [0], [0, 1], [1])              # 40 distinct bytes
        size, common, dist = eng.kmer_distance(res, off, 8, [0], [0, 1], [1])
        assert list(size) == [3, 3] and common[0] == 3 and dist[0] == 0.0
