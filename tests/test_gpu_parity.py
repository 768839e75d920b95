"""GPU parity: the CUDA path through the C ABI (libkmeranno.so) must equal the CPU oracle
bit for bit — role call, hit count and flag of every sequence (integer work: exact)."""
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


def run_case(ka, oracle, seqs, kmers, roles, K, min_hits=5, options=None, resident=False):
    res, off = csr(seqs)
    with ka.Engine([0]) as eng:
        for k, v in (options or {}).items():
            eng.set_option(k, v)
        eng.db_load(kmers, roles, K)
        if resident:
            b = eng.upload(res, off)
            eng.annotate_resident(b, min_hits)
            got = eng.download(b)
            b.free()
        else:
            got = eng.annotate(res, off, min_hits)
        info = eng.db_info()
    want = oracle.OracleDb(kmers, roles, K).apply(res, off, min_hits)
    assert info["n_keys"] == oracle.OracleDb(kmers, roles, K).size()
    assert_same(got, want, f"K={K} min_hits={min_hits} opts={options}")
    return got


@pytest.mark.parametrize("K", [1, 2, 5, 8, 10, 12])
def test_ragged_all_k(ka, oracle, K):
    seqs, kmers, roles = ragged_case(100 + K, n_seq=400, K=K)
    got = run_case(ka, oracle, seqs, kmers, roles, K, min_hits=3)
    # every outcome class must actually occur for the case to mean anything
    if K >= 5:
        assert set(np.unique(got[2])) == {0, 1, 2, 3}


@pytest.mark.parametrize("min_hits", [1, 2, 5, 50])
def test_min_hits(ka, oracle, min_hits):
    seqs, kmers, roles = ragged_case(7, n_seq=300, K=8)
    run_case(ka, oracle, seqs, kmers, roles, 8, min_hits=min_hits)


def test_against_pure_python(ka, oracle):
    seqs, kmers, roles = ragged_case(11, n_seq=120, K=8, max_len=300)
    got = run_case(ka, oracle, seqs, kmers, roles, 8, min_hits=2)
    want = py_apply(seqs, kmers, roles, 8, 2)
    assert_same(got, want, "pure python")


def test_duplicate_kmers_count_once(ka, oracle):
    # a protein that is one 9-mer unit repeated: 9 distinct 8-mers however long it is
    unit = b"ACDEFGHIK"
    prot = unit * 40
    kmers = [prot[i:i + 8] for i in range(9)]
    roles = np.zeros(9, np.int32) + 4
    got = run_case(ka, oracle, [prot, prot[:30], unit], kmers, roles, 8, min_hits=5)
    assert got[1][0] == 9 and got[0][0] == 4 and got[2][0] == 1


def test_last_db_line_wins(ka, oracle):
    prot = b"MKVLAAGIVALLLAGCSSAPKE"
    kmers = [prot[i:i + 8] for i in range(6)] * 2
    roles = np.asarray([1] * 6 + [9] * 6, np.int32)
    got = run_case(ka, oracle, [prot], kmers, roles, 8, min_hits=5)
    assert got[0][0] == 9 and got[1][0] == 6


def test_empty_and_short(ka, oracle):
    seqs = [b"", b"", b"ACD", b"ACDEFGH", b"ACDEFGHI", b"", b"ACDEFGHIK", b""]
    kmers = [b"ACDEFGHI", b"CDEFGHIK"]
    got = run_case(ka, oracle, seqs, kmers, np.asarray([3, 3], np.int32), 8, min_hits=1)
    assert list(got[1]) == [0, 0, 0, 0, 1, 0, 2, 0]
    assert list(got[0]) == [-1, -1, -1, -1, 3, -1, 3, -1]


def test_bytes_outside_db_alphabet(ka, oracle):
    # case-sensitive, no filtering: 'a' != 'A', X / * never match, windows over them miss
    kmers = [b"ACDEFGHI", b"CDEFGHIK", b"DEFGHIKL"]
    seqs = [b"ACDEFGHIKL", b"acdefghikl", b"ACDEFGHIXKL", b"ACDEFGHI*", b"\x00\xff" * 8 + b"ACDEFGHI"]
    got = run_case(ka, oracle, seqs, kmers, np.asarray([2, 2, 2], np.int32), 8, min_hits=1)
    assert list(got[1]) == [3, 0, 1, 1, 1]


def test_many_tiny_sequences_sub_batches(ka, oracle):
    # thousands of sequences starting inside one residue tile: > MAX_TILE_SEQ per tile
    rng = np.random.default_rng(5)
    seqs = []
    for i in range(6000):
        r = rng.random()
        seqs.append(b"" if r < 0.5 else (b"ACDEFGHI" if r < 0.75 else random_seq(rng, int(rng.integers(1, 12)))))
    kmers = [b"ACDEFGHI"]
    got = run_case(ka, oracle, seqs, kmers, np.asarray([7], np.int32), 8, min_hits=1)
    assert (got[0] == 7).sum() >= 1400


def test_long_sequences_use_big_kernel(ka, oracle):
    lengths = [20000, 300, 7000, 2049, 2048, 64, 0, 12000, 8192, 8193, 2047, 5000]
    seqs, kmers, roles = ragged_case(21, n_seq=len(lengths), K=8, lengths=lengths, db_frac=0.5)
    run_case(ka, oracle, seqs, kmers, roles, 8, min_hits=3)
    run_case(ka, oracle, seqs, kmers, roles, 8, min_hits=3, options={"wide": 1})
    # the same inputs with a small long_seq so that most sequences take the long path
    run_case(ka, oracle, seqs, kmers, roles, 8, min_hits=3, options={"tile_span": 256, "long_seq": 256})


@pytest.mark.parametrize("opts", [
    {"tile_span": 256, "long_seq": 1024},
    {"tile_span": 4096, "long_seq": 8192, "mid_seq": 8192},
    {"tile_span": 512, "long_seq": 512, "mid_seq": 600},
    {"mid_seq": 2048},
    {"long_seq": 2048},
    {"chunk_residues": 4096},
    {"load_factor": 0.9},
    {"load_factor": 0.05},
    {"l2_persist": 0},
    {"slot_bits": 32},
    {"slot_bits": 64},
    {"slot_bits": 128},
    {"slot_bits": 64, "load_factor": 0.9},
    {"slot_bits": 32, "load_factor": 0.9},
    {"wide": 1},
    {"wide": 1, "slot_bits": 64, "load_factor": 0.9},
    {"wide": 1, "slot_bits": 32, "load_factor": 0.9, "tile_span": 256, "long_seq": 300, "mid_seq": 700},
    {"wide": 1, "slot_bits": 128},      # 128-bit slots have no wide form: the option is ignored
    # the 128-byte-line table (16-bit tags and roles, spill inside the line, presence filter)
    {"slot_bits": 16},
    {"slot_bits": 16, "filter": 0},
    {"slot_bits": 16, "tile_span": 256, "long_seq": 300, "mid_seq": 700},
    {"slot_bits": 16, "tile_span": 4096, "long_seq": 8192, "mid_seq": 8192},
    {"slot_bits": 16, "chunk_residues": 4096},
    {"slot_bits": 16, "load_factor": 0.8, "filter": 0},
])
def test_options_do_not_change_results(ka, oracle, opts):
    seqs, kmers, roles = ragged_case(33, n_seq=500, K=8, max_len=900)
    run_case(ka, oracle, seqs, kmers, roles, 8, min_hits=3, options=opts)


@pytest.mark.parametrize("K,max_role,want_bits", [(5, 400, (16, 32)), (8, 400, (16, 64)), (8, 70000, (64, 128)), (12, 30000, (64,)),
                                                   (12, 2**31 - 2, (128,)), (3, 10, (16, 32))])
def test_slot_class_selection(ka, oracle, K, max_role, want_bits):
    """A replicated table with K <= 10 and role ids below 65536 gets the line table (slot_bits 16) when its layout
    fits the key space without padding; otherwise the engine picks the narrowest sector slot that holds remainder +
    role (ka_common.cuh)."""
    seqs, kmers, roles = ragged_case(70 + K, n_seq=300, K=K, n_roles=12)
    roles = (roles.astype(np.int64) * (max_role // 11)).astype(np.int32)   # spread ids up to max_role
    res, off = csr(seqs)
    with ka.Engine([0]) as eng:
        eng.db_load(kmers, roles, K)
        assert eng.db_info()["slot_bits"] in want_bits
        got = eng.annotate(res, off, 3)
    assert_same(got, oracle.OracleDb(kmers, roles, K).apply(res, off, 3), f"slot class K={K}")


@pytest.mark.parametrize("slot_bits,lf,filt", [(32, 0.9, 0), (64, 0.9, 0), (128, 0.9, 0), (32, 0.4, 0), (32, 0.9, -2), (64, 0.9, -2)])
def test_overflow_heavy_table(ka, oracle, slot_bits, lf, filt):
    """Millions of keys at a high load factor: ~13 % of the keys leave their home sector.
    Quotiented slots keep only a remainder, so those keys must live in the overflow table
    under their whole mixed value — every distinct key must survive and resolve exactly."""
    from kmers_anno_b200 import synth
    fam = synth.Families(2000)
    kmers, roles = fam.table(3_000_000, K=8)
    res, off, _ = fam.batch(5, 2, n_prot=4500)
    with ka.Engine([0]) as eng:
        eng.set_option("slot_bits", slot_bits)
        eng.set_option("load_factor", lf)
        if filt == -2:
            eng.set_option("wide", 1)         # wide-table kernels on the same table
        eng.db_load(kmers, roles, 8)
        info = eng.db_info()
        got = eng.annotate(res, off, 5)
    assert info["n_keys"] == 3_000_000 and info["slot_bits"] == slot_bits
    want = oracle.OracleDb(kmers, roles, 8, threads=8).apply(res, off, 5, threads=8)
    assert_same(got, want, f"overflow-heavy table slot_bits={slot_bits} lf={lf}")


def test_resident_path(ka, oracle):
    seqs, kmers, roles = ragged_case(44, n_seq=500, K=10, max_len=900)
    run_case(ka, oracle, seqs, kmers, roles, 10, min_hits=3, resident=True)


def test_offsets_with_nonzero_base(ka, oracle):
    seqs, kmers, roles = ragged_case(55, n_seq=200, K=8)
    res, off = csr(seqs)
    pad = 37
    res2 = np.concatenate([np.full(pad, ord("A"), np.uint8), res])
    off2 = off + np.uint64(pad)
    with ka.Engine([0]) as eng:
        eng.db_load(kmers, roles, 8)
        got = eng.annotate(res2, off2, 3)
    want = oracle.OracleDb(kmers, roles, 8).apply(res, off, 3)
    assert_same(got, want, "nonzero base")


def test_synthetic_proteome_c1_shape(ka, oracle):
    """SURVEY config 1 shape: one 4,500-protein proteome vs a small family DB; the fast
    oracle cross-checks the Java-shaped one."""
    from kmers_anno_b200 import synth
    fam = synth.Families(500)
    kmers, roles = fam.table(400000, K=8)
    res, off, true_role = fam.batch(3, 1, n_prot=4500)
    with ka.Engine([0]) as eng:
        eng.db_load(kmers, roles, 8)
        got = eng.annotate(res, off, 5)
        st = eng.stats()
    want = oracle.OracleDb(kmers, roles, 8).apply(res, off, 5, threads=4)
    fast = oracle.FastDb(kmers, roles, 8).apply(res, off, 5, threads=4)
    assert_same(got, want, "C1 vs Java-shaped oracle")
    assert_same(got, fast, "C1 vs packed oracle")
    assert st["probes"] == oracle.count_probes(off, 8)
    called = got[2] == 1
    assert called.sum() > 1000 and (got[0][called] == true_role[called]).all()


@pytest.mark.parametrize("mode,K", [(1, 8), (2, 10), (1, 12)])
def test_skewed_lengths_c4_shape(ka, oracle, mode, K):
    """SURVEY config 4: K sweep with log-uniform / bimodal 50..5000 aa lengths."""
    from kmers_anno_b200 import synth
    fam = synth.Families(300)
    kmers, roles = fam.table(300000, K=K)
    res, off, _ = fam.batch(1, 1, n_prot=1500, mode=mode, K=K)
    with ka.Engine([0]) as eng:
        eng.db_load(kmers, roles, K)
        got = eng.annotate(res, off, 5)
    want = oracle.OracleDb(kmers, roles, K).apply(res, off, 5, threads=4)
    assert_same(got, want, f"C4 mode={mode} K={K}")
    assert (got[2] == 2).sum() > 0 and (got[2] == 1).sum() > 0


def test_multi_device_engine(ka, oracle):
    """One engine over every visible GPU: the batch is cut into residue-balanced ranges, one
    per device, with a replicated table; the gathered result must equal the oracle's."""
    n_dev = 0
    for n in (8, 4, 2):
        try:
            ka.Engine(list(range(n))).close()
            n_dev = n
            break
        except ka.KmerAnnoError:
            continue
    if n_dev < 2:
        pytest.skip("needs at least 2 GPUs")
    from kmers_anno_b200 import synth
    fam = synth.Families(300)
    kmers, roles = fam.table(300000, K=8)
    res, off, _ = fam.batch(2, 3, n_prot=2000)
    with ka.Engine(list(range(n_dev))) as eng:
        eng.set_option("chunk_residues", 200000)      # several chunks per device
        eng.db_load(kmers, roles, 8)
        got = eng.annotate(res, off, 5)
        st = eng.stats()
    want = oracle.OracleDb(kmers, roles, 8).apply(res, off, 5, threads=4)
    assert_same(got, want, f"{n_dev}-device engine")
    assert st["sequences"] == off.shape[0] - 1 and st["kernel_launches"] >= 2 * n_dev


@pytest.mark.parametrize("slot_bits,lf,wide", [(0, 0.4, 0), (32, 0.9, 0), (64, 0.9, 0), (0, 0.4, 1), (64, 0.9, 1)])
def test_sharded_table_peer_loads(ka, oracle, slot_bits, lf, wide):
    """table_mode=1: the table is split by sector range over the engine's GPUs and probes read
    remote sectors through NVLink peer memory (the config-5 shape, at test size)."""
    n_dev = 0
    for n in (8, 4, 2):
        try:
            ka.Engine(list(range(n))).close()
            n_dev = n
            break
        except ka.KmerAnnoError:
            continue
    if n_dev < 2:
        pytest.skip("needs at least 2 GPUs")
    from kmers_anno_b200 import synth
    fam = synth.Families(1000)
    kmers, roles = fam.table(2_000_000, K=8)
    res, off, _ = fam.batch(9, 2, n_prot=4500)
    with ka.Engine(list(range(n_dev))) as eng:
        eng.set_option("table_mode", 1)
        eng.set_option("wide", wide)
        eng.set_option("slot_bits", slot_bits)
        eng.set_option("load_factor", lf)
        eng.db_load(kmers, roles, 8)
        info = eng.db_info()
        got = eng.annotate(res, off, 5)
    assert info["n_keys"] == 2_000_000
    want = oracle.OracleDb(kmers, roles, 8, threads=8).apply(res, off, 5, threads=8)
    assert_same(got, want, f"sharded table over {n_dev} GPUs slot_bits={slot_bits}")


@pytest.mark.parametrize("mode", [2, 3])
@pytest.mark.parametrize("slot_bits,lf,chunk,wide", [(0, 0.4, 32 << 20, 0), (32, 0.9, 300000, 0), (64, 0.9, 1 << 20, 0),
                                                     (0, 0.4, 32 << 20, 1), (32, 0.9, 300000, 1)])
def test_routed_table_nccl_all_to_all(ka, oracle, slot_bits, lf, chunk, wide, mode):
    """table_mode=2: same sharding, but the keys are routed to the owning GPU with NCCL send/recv
    (all-to-all), probed there and the answers come back in request order.  table_mode=3: the same routing
    with the exchanges fused into the scatter / lookup kernels as NVLink peer stores."""
    n_dev = 0
    for n in (8, 4, 2):
        try:
            ka.Engine(list(range(n))).close()
            n_dev = n
            break
        except ka.KmerAnnoError:
            continue
    if n_dev < 2:
        pytest.skip("needs at least 2 GPUs")
    from kmers_anno_b200 import synth
    fam = synth.Families(1000)
    kmers, roles = fam.table(2_000_000, K=8)
    res, off, _ = fam.batch(9, 2, n_prot=4500)
    # uneven ranges: a long tail of empty sequences gives the last device fewer rounds than the first
    off = np.concatenate([off, np.full(5000, off[-1], np.uint64)])
    with ka.Engine(list(range(n_dev))) as eng:
        eng.set_option("table_mode", mode)
        eng.set_option("wide", wide)
        eng.set_option("slot_bits", slot_bits)
        eng.set_option("load_factor", lf)
        eng.set_option("chunk_residues", chunk)
        eng.db_load(kmers, roles, 8)
        got = eng.annotate(res, off, 5)
        got2 = eng.annotate(res, off, 1)
        # sequences longer than mid_seq (the reference annotates proteins of any length): their keys are not routed, the
        # long-sequence kernel reads the shards through peer loads
        lres = np.concatenate([res[: int(off[40])], res[: int(off[120])], res[: int(off[3])]])
        loff = np.asarray([0, int(off[40]), int(off[40]) + int(off[120]), len(lres)], np.uint64)
        assert int(off[120]) > 8192
        glong = eng.annotate(lres, loff, 5)
        again = eng.annotate(res, off, 5)               # the engine stays usable
        # the packed input form: the extract and tally kernels stage the 5-bit stream themselves
        codes, off32 = eng.pack(res, off)
        gpacked = eng.annotate_packed(codes, off32, 5)
        lcodes, loff32 = eng.pack(lres, loff)
        glong_packed = eng.annotate_packed(lcodes, loff32, 5)
    want = oracle.OracleDb(kmers, roles, 8, threads=8).apply(res, off, 5, threads=8)
    assert_same(got, want, f"routed table over {n_dev} GPUs slot_bits={slot_bits}")
    assert_same(again, want, "routed table, second call")
    assert_same(glong, oracle.OracleDb(kmers, roles, 8, threads=8).apply(lres, loff, 5, threads=8), "routed table, long sequences")
    assert_same(gpacked, want, "routed table, packed input")
    assert_same(glong_packed, glong, "routed table, long sequences, packed input")
    assert_same(got2, oracle.OracleDb(kmers, roles, 8, threads=8).apply(res, off, 1, threads=8), "routed, min_hits 1")


def planted_proteins(rng, kmers, roles, n_prot, K):
    """Proteins made of DB k-mers of one role (some of two roles) joined by random residues."""
    by_role = {}
    for i, r in enumerate(roles):
        by_role.setdefault(int(r), []).append(i)
    role_ids = sorted(by_role)
    seqs = []
    for p in range(n_prot):
        r = role_ids[int(rng.integers(len(role_ids)))]
        picks = [by_role[r][int(j)] for j in rng.integers(0, len(by_role[r]), int(rng.integers(1, 12)))]
        if p % 7 == 0:                                   # a second role: ambiguous
            r2 = role_ids[int(rng.integers(len(role_ids)))]
            picks.append(by_role[r2][0])
        parts = []
        for i in picks:
            parts.append(kmers[i].tobytes())
            parts.append(random_seq(rng, int(rng.integers(0, 30))))
        seqs.append(b"".join(parts))
    return seqs


@pytest.mark.parametrize("K,n,n_roles,opts", [(12, 200_000, 300, {}), (12, 200_000, 300, {"wide": 1}),
                                              (8, 50_000, 40, {"load_factor": 0.9}), (5, 3_000_000, 17, {})])
def test_synthetic_db_matches_host_lines(ka, oracle, K, n, n_roles, opts):
    """ka_db_load_synthetic generates the lines on the device; synth.synthetic_db_lines regenerates
    them on the host for the oracle (K=5: 3e6 draws from 3.2e6 possible k-mers: duplicates, last wins)."""
    from kmers_anno_b200 import synth
    seed = 20261018 + K
    kmers, roles = synth.synthetic_db_lines(np.arange(n, dtype=np.uint64), K, n_roles, seed)
    rng = np.random.default_rng(K)
    seqs = planted_proteins(rng, kmers, roles, 1500, K) + [random_seq(rng, 400) for _ in range(200)]
    res, off = csr(seqs)
    with ka.Engine([0]) as eng:
        for k, v in opts.items():
            eng.set_option(k, v)
        eng.db_load_synthetic(n, K, n_roles, seed)
        info = eng.db_info()
        got = eng.annotate(res, off, 3)
    db = oracle.OracleDb(kmers.reshape(-1), roles, K, threads=8)
    assert info["n_keys"] == db.size() and info["n_lines"] == n
    assert_same(got, db.apply(res, off, 3, threads=8), f"synthetic DB K={K}")
    if K >= 8:
        assert set(np.unique(got[2])) >= {1, 2}


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_synthetic_db_sharded(ka, oracle, mode):
    """The oversized-table path at test size: device-generated lines, sharded wide table."""
    n_dev = 0
    for n in (8, 4, 2):
        try:
            ka.Engine(list(range(n))).close()
            n_dev = n
            break
        except ka.KmerAnnoError:
            continue
    if n_dev < 2:
        pytest.skip("needs at least 2 GPUs")
    from kmers_anno_b200 import synth
    K, n, n_roles, seed = 12, 4_000_000, 3000, 99
    kmers, roles = synth.synthetic_db_lines(np.arange(n, dtype=np.uint64), K, n_roles, seed)
    rng = np.random.default_rng(3)
    seqs = planted_proteins(rng, kmers, roles, 4000, K) + [random_seq(rng, 400) for _ in range(500)]
    res, off = csr(seqs)
    with ka.Engine(list(range(n_dev))) as eng:
        eng.set_option("table_mode", mode)
        eng.set_option("wide", 1)
        eng.db_load_synthetic(n, K, n_roles, seed)
        info = eng.db_info()
        got = eng.annotate(res, off, 3)
    db = oracle.OracleDb(kmers.reshape(-1), roles, K, threads=8)
    assert info["n_keys"] == db.size()
    assert_same(got, db.apply(res, off, 3, threads=8), f"synthetic sharded DB mode {mode}")


def test_error_paths(ka):
    with ka.Engine([0]) as eng:
        with pytest.raises(ka.KmerAnnoError) as e:
            eng.annotate(np.zeros(8, np.uint8), np.asarray([0, 8], np.uint64), 5)
        assert e.value.code == -6                     # KA_ERR_NO_DB
        with pytest.raises(ka.KmerAnnoError) as e:
            eng.db_load([b"A" * 13], np.zeros(1, np.int32), 13)
        assert e.value.code == -5                     # KA_ERR_K
        many = [bytes([65 + (i % 26), 97 + (i % 26)] * 4) for i in range(26)]
        with pytest.raises(ka.KmerAnnoError) as e:
            eng.db_load(many, np.zeros(26, np.int32), 8)
        assert e.value.code == -4                     # KA_ERR_ALPHABET: 52 distinct bytes
        with pytest.raises(ka.KmerAnnoError) as e:
            eng.db_load([b"ACDEFGHI"], np.asarray([-1], np.int32), 8)
        assert e.value.code == -8                     # KA_ERR_ROLE
        eng.db_load([b"ACDEFGHI"], np.asarray([0], np.int32), 8)
        with pytest.raises(ka.KmerAnnoError) as e:
            eng.annotate(np.zeros(8, np.uint8), np.asarray([0, 8], np.uint64), 0)
        assert e.value.code == -1                     # min_hits must be positive (:91-92)
        with pytest.raises(ka.KmerAnnoError) as e:
            eng.annotate(np.zeros(8, np.uint8), np.asarray([8, 0], np.uint64), 1)
        assert e.value.code == -9                     # KA_ERR_OFFSETS
    with pytest.raises(ka.KmerAnnoError) as e:
        ka.Engine([99])
    assert e.value.code == -2                         # KA_ERR_NO_DEVICE
