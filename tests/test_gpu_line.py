"""GPU parity of the 128-byte-line table (slot class 16: 16-bit tags + roles, spill inside the line,
overflow table, L2-resident presence filter) and of the packed 5-bit input form
(ka_annotate_packed), against the CPU oracle through the C ABI.  Integer work: bit-exact."""
import numpy as np
import pytest

from cases import csr, py_apply, ragged_case, random_seq

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ka():
    import kmers_anno_b200 as ka
    return ka


@pytest.fixture(scope="module")
def oracle():
    import oracle
    return oracle


def assert_same(got, want, what=""):
    for name, g, w in zip(("role", "hits", "flag"), got, want):
        if not np.array_equal(g, w):
            bad = np.nonzero(g != w)[0]
            raise AssertionError(f"{what}: {name} differs on {bad.size}/{g.size} sequences; first "
                                 f"{bad[:8]} got {g[bad[:8]]} want {w[bad[:8]]}")


def run_line(ka, oracle, seqs, kmers, roles, K, min_hits=3, options=None, form="bytes", threads=1):
    res, off = csr(seqs)
    with ka.Engine([0]) as eng:
        eng.set_option("slot_bits", 16)
        for k, v in (options or {}).items():
            eng.set_option(k, v)
        eng.db_load(kmers, roles, K)
        info = eng.db_info()
        assert info["slot_bits"] == 16
        if form == "packed":
            codes, off32 = eng.pack(res, off, threads=threads)
            got = eng.annotate_packed(codes, off32, min_hits)
        elif form == "resident":
            b = eng.upload(res, off)
            eng.annotate_resident(b, min_hits)
            got = eng.download(b)
            b.free()
        else:
            got = eng.annotate(res, off, min_hits)
    odb = oracle.OracleDb(kmers, roles, K)
    assert info["n_keys"] == odb.size()
    assert_same(got, odb.apply(res, off, min_hits), f"line table K={K} form={form} opts={options}")
    return got, info


@pytest.mark.parametrize("K", [2, 3, 5, 7, 8, 9])
@pytest.mark.parametrize("form", ["bytes", "packed", "resident"])
def test_line_table_ragged(ka, oracle, K, form):
    seqs, kmers, roles = ragged_case(200 + K, n_seq=400, K=K)
    got, _ = run_line(ka, oracle, seqs, kmers, roles, K, form=form)
    if K >= 5:
        assert set(np.unique(got[2])) == {0, 1, 2, 3}


def test_line_table_rejects_what_it_cannot_hold(ka):
    seqs, kmers, roles = ragged_case(3, n_seq=50, K=12)
    with ka.Engine([0]) as eng:
        eng.set_option("slot_bits", 16)
        with pytest.raises(ka.KmerAnnoError) as ei:
            eng.db_load(kmers, roles, 12)                      # 20^12: halves of 26 + 26 bits
        assert ei.value.code == -10
        with pytest.raises(ka.KmerAnnoError) as ei:
            eng.db_load([b"A", b"C"], np.asarray([1, 2], np.int32), 1)     # K = 1 has no two halves
        assert ei.value.code == -10
        seqs, kmers, roles = ragged_case(3, n_seq=50, K=8)
        big = roles.copy(); big[0] = 70000
        with pytest.raises(ka.KmerAnnoError) as ei:
            eng.db_load(kmers, big, 8)                         # role ids beyond 16 bits
        assert ei.value.code == -1


def test_line_table_against_pure_python(ka, oracle):
    seqs, kmers, roles = ragged_case(12, n_seq=120, K=8, max_len=300)
    got, _ = run_line(ka, oracle, seqs, kmers, roles, 8, min_hits=2, form="packed")
    assert_same(got, py_apply(seqs, kmers, roles, 8, 2), "pure python")


def test_line_table_duplicates_last_line_and_odd_bytes(ka, oracle):
    unit = b"ACDEFGHIK"
    prot = unit * 40
    kmers = [prot[i:i + 8] for i in range(9)] * 2
    roles = np.asarray([4] * 9 + [65535] * 9, np.int32)       # last line wins; the largest 16-bit role id
    seqs = [prot, prot[:30], unit, b"acdefghik" * 3, b"ACDEFGHIXK", b"", b"ACDEFGH", b"\x00\xff" * 8 + prot[:12]]
    for form in ("bytes", "packed"):
        got, _ = run_line(ka, oracle, seqs, kmers, roles, 8, min_hits=5, form=form)
        assert got[1][0] == 9 and got[0][0] == 65535 and got[2][0] == 1


def test_line_table_spill_and_overflow(ka, oracle):
    """7-mers at load factor 0.8: ~20 % of the sectors spill into their line, ~9 % of the lines
    overflow into the overflow table; every distinct key must survive and resolve exactly."""
    from kmers_anno_b200 import synth
    fam = synth.Families(2000)
    K = 7
    kmers, roles = fam.table(3_000_000, K=K)
    res, off, _ = fam.batch(5, 2, n_prot=3000, K=K)
    want = oracle.OracleDb(kmers, roles, K, threads=8).apply(res, off, 5, threads=8)
    for filt in (1, 0):
        with ka.Engine([0]) as eng:
            eng.set_option("slot_bits", 16)
            eng.set_option("load_factor", 0.8)
            eng.set_option("filter", filt)
            eng.db_load(kmers, roles, K)
            info = eng.db_info()
            got = eng.annotate(res, off, 5)
        assert info["n_keys"] == oracle.OracleDb(kmers, roles, K, threads=8).size()
        assert info["n_spilled"] > info["n_keys"] // 50 and info["n_overflow"] > 1000, info
        assert_same(got, want, f"spill / overflow heavy line table, filter={filt}")


def test_line_table_tiny_sequences_and_long_ones(ka, oracle):
    rng = np.random.default_rng(6)
    seqs = []
    for i in range(5000):                                      # far more than 64 sequences per tile
        r = rng.random()
        seqs.append(b"" if r < 0.5 else (b"ACDEFGHI" if r < 0.75 else random_seq(rng, int(rng.integers(1, 12)))))
    got, _ = run_line(ka, oracle, seqs, [b"ACDEFGHI"], np.asarray([7], np.int32), 8, min_hits=1, form="packed")
    assert (got[0] == 7).sum() >= 1100
    lengths = [20000, 300, 7000, 1537, 1536, 64, 0, 12000, 8192, 8193, 1535, 5000]
    seqs, kmers, roles = ragged_case(22, n_seq=len(lengths), K=8, lengths=lengths, db_frac=0.5)
    for form in ("bytes", "packed", "resident"):
        run_line(ka, oracle, seqs, kmers, roles, 8, form=form)
    run_line(ka, oracle, seqs, kmers, roles, 8, options={"tile_span": 256, "long_seq": 256, "mid_seq": 512}, form="packed")


def test_line_table_segments_of_one_sequence_share_one_kmer_set(ka, oracle):
    """A sequence with tiles of its own is cut into segments that are filtered and probed by different warps and
    tallied together: a k-mer that occurs in several segments counts once, and windows that straddle a segment
    boundary are looked up like any other (here every window of the periodic sequences is in the DB)."""
    rng = np.random.default_rng(11)
    K = 8
    motifs = [random_seq(rng, 50), random_seq(rng, 37), random_seq(rng, 64)]
    seqs = [motifs[0] * 120, random_seq(rng, 400), motifs[1] * 100 + random_seq(rng, 900) + motifs[1] * 60,
            motifs[2] * 128, random_seq(rng, 3000) + motifs[0] * 20]           # 6000, 400, 6820, 8192, 4000 residues
    kmers, roles = [], []
    for r, m in enumerate(motifs):
        mm = m + m[: K - 1]
        for i in range(len(m)):
            kmers.append(mm[i:i + K]); roles.append(3 + r)
    roles = np.asarray(roles, np.int32)
    for form in ("bytes", "packed", "resident"):
        got, _ = run_line(ka, oracle, seqs, kmers, roles, K, min_hits=5, form=form)
        assert got[1][0] == 50 and got[0][0] == 3 and got[2][0] == 1          # 50 distinct k-mers, however often they recur
        assert got[1][3] == 64 and got[0][3] == 5
        assert got[2][4] == 1 and got[0][4] == 3
    run_line(ka, oracle, seqs, kmers, roles, K, min_hits=5, options={"tile_span": 256, "long_seq": 512, "mid_seq": 8192}, form="packed")


def test_packed_form_on_every_layout(ka, oracle):
    """ka_annotate_packed == ka_annotate on sector-class tables too (device unpack), with a batch whose
    first offset is not zero and with several chunks; ka_pack_residues from several threads."""
    seqs, kmers, roles = ragged_case(56, n_seq=900, K=8, max_len=500)
    res, off = csr(seqs)
    pad = 37
    res2 = np.concatenate([np.full(pad, ord("A"), np.uint8), res])
    off2 = off + np.uint64(pad)
    want = oracle.OracleDb(kmers, roles, 8).apply(res, off, 3)
    for slot_bits in (16, 32, 64, 128):
        with ka.Engine([0]) as eng:
            eng.set_option("slot_bits", slot_bits)
            eng.set_option("chunk_residues", 20000)
            eng.db_load(kmers, roles, 8)
            lut = eng.alphabet()
            assert sorted(set(lut.tolist()) - {31}) == list(range(20)) and lut[ord("A")] == 0 and lut[ord("X")] == 31
            for threads in (1, 5):
                codes, off32 = eng.pack(res2, off2, threads=threads)
                assert_same(eng.annotate_packed(codes, off32, 3), want, f"packed, slot_bits={slot_bits}")
            assert_same(eng.annotate(res2, off2, 3), want, f"bytes, slot_bits={slot_bits}")


def test_options_are_validated_atomically(ka):
    """A rejected ka_set_option changes nothing; table options only act at the next load; a resident
    batch uploaded under other tiling options is refused instead of being mis-tiled."""
    seqs, kmers, roles = ragged_case(9, n_seq=200, K=8)
    res, off = csr(seqs)
    with ka.Engine([0]) as eng:
        eng.db_load(kmers, roles, 8)
        base = eng.annotate(res, off, 3)
        for name, bad in (("tile_span", 100000), ("tile_span", 65536), ("mid_seq", 10), ("long_seq", 1e9),
                          ("slot_bits", 48), ("table_mode", 5), ("nonsense", 1)):
            with pytest.raises(ka.KmerAnnoError):
                eng.set_option(name, bad)
        assert_same(eng.annotate(res, off, 3), base, "after rejected options")
        eng.set_option("table_mode", 2)                        # acts at the next ka_db_load only: no NCCL call now
        assert_same(eng.annotate(res, off, 3), base, "table_mode set after the load")
        eng.set_option("table_mode", 0)
        b = eng.upload(res, off)
        eng.set_option("long_seq", 4096)
        with pytest.raises(ka.KmerAnnoError):
            eng.annotate_resident(b, 3)
        eng.set_option("long_seq", 2048)
        eng.annotate_resident(b, 3)
        assert_same(eng.download(b), base, "resident after restoring the tiling options")
        eng.db_load(kmers, roles, 8)
        with pytest.raises(ka.KmerAnnoError):
            eng.annotate_resident(b, 3)                        # the DB was reloaded
        b.free()
