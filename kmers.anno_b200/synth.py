"""ctypes binding of libkasynth.so (host/ka_synth.cpp): seeded synthetic role families,
proteomes and signature tables of the shapes SURVEY.md §8(d) defines.  Test/bench support."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkasynth.so")
SEED = 20261018
_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        lib.kas_families_new.argtypes = [C.c_uint64, C.c_uint32]
        lib.kas_families_new.restype = vp
        lib.kas_families_free.argtypes = [vp]
        lib.kas_families_free.restype = None
        lib.kas_family_len.argtypes = [vp, C.c_uint32]
        lib.kas_family_len.restype = C.c_uint64
        lib.kas_batch_size.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_int]
        lib.kas_batch_size.restype = C.c_uint64
        lib.kas_batch_fill.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_int,
                                       C.c_int, vp, vp, vp]
        lib.kas_batch_fill.restype = None
        lib.kas_table.argtypes = [vp, C.c_uint64, C.c_int, C.c_uint32, C.c_uint64, C.c_int, vp, vp]
        lib.kas_table.restype = C.c_uint64
        _lib = lib
    return _lib


class Families:
    def __init__(self, n_roles, seed=SEED):
        self.lib = _load()
        self.seed = seed
        self.n_roles = n_roles
        self.h = self.lib.kas_families_new(seed, n_roles)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.kas_families_free(self.h)
            self.h = None

    def batch(self, g0, n_genomes, n_prot=4500, mode=0, K=8, threads=None, alloc=None):
        """CSR batch of genomes [g0, g0+n_genomes): (residues u8, offsets u64, true_role i32).
        `alloc(shape, dtype)` lets the caller provide pinned arrays."""
        threads = threads or min(os.cpu_count() or 1, 32)
        n = n_genomes * n_prot
        alloc = alloc or (lambda shape, dtype: np.empty(shape, dtype))
        offsets = alloc(n + 1, np.uint64)
        true_role = np.empty(n, np.int32)
        # sizes first (single pass inside fill computes them again; sizing is cheap)
        total = self.lib.kas_batch_size(self.h, self.seed, g0, n_genomes, n_prot, mode, K)
        residues = alloc(max(int(total), 1), np.uint8)
        self.lib.kas_batch_fill(self.h, self.seed, g0, n_genomes, n_prot, mode, K, threads,
                                residues.ctypes.data, offsets.ctypes.data, true_role.ctypes.data)
        return residues[:int(total)], offsets, true_role

    def table(self, target, K=8, members_per_role=8, threads=None):
        """(kmers u8[n*K], roles i32[n]) of up to `target` discriminating k-mers."""
        threads = threads or min(os.cpu_count() or 1, 32)
        kmers = np.empty(target * K, np.uint8)
        roles = np.empty(target, np.int32)
        n = self.lib.kas_table(self.h, self.seed, K, members_per_role, target, threads,
                               kmers.ctypes.data, roles.ctypes.data)
        return kmers[: n * K], roles[:n]


_AA20 = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)


def synthetic_db_lines(index, K, n_roles, seed):
    """Lines `index` (uint64 array) of the device-generated DB of ka_db_load_synthetic
    (include/kmeranno.h): (kmers u8[len(index), K], roles i32[len(index)])."""
    index = np.asarray(index, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = np.uint64(seed) + index * np.uint64(0x9E3779B97F4A7C15)
        for _ in range(2):
            x ^= x >> np.uint64(32)
            x *= np.uint64(0xD6E8FEB86659FD93)
        x ^= x >> np.uint64(32)
    kmers = np.empty((index.shape[0], K), np.uint8)
    for j in range(K):
        kmers[:, j] = _AA20[((x >> np.uint64(5 * j)) & np.uint64(31)) % np.uint64(20)]
    return kmers, (index % np.uint64(n_roles)).astype(np.int32)


def planted_batch(n_keys, n_prot, K, n_roles, seed, alloc=None, rng_seed=5, min_hits=5):
    """Queries for a device-generated DB (ka_db_load_synthetic) that is too large to regenerate on the host:
    PLANTED proteins made of DB k-mers of one role (every 7th protein: plus one k-mer of a second role) joined
    by random residues, so that the expected call of every protein is known by construction, up to chance
    hits of the random windows.  Returns (residues u8, offsets u64, expected role i32, expected hits i32,
    ambiguous bool, probes)."""
    alloc = alloc or (lambda shape, dtype: np.empty(shape, dtype))
    rng = np.random.default_rng(rng_seed)
    h = rng.integers(1, 12, n_prot)                      # planted k-mers per protein
    role = rng.integers(0, n_roles, n_prot)
    ambiguous = (np.arange(n_prot) % 7) == 0
    n_seg = h + ambiguous                                # + one k-mer of another role
    seg_prot = np.repeat(np.arange(n_prot), n_seg)
    seg_first = np.concatenate([[0], np.cumsum(n_seg)])[:-1]
    is_extra = np.zeros(seg_prot.shape[0], bool)
    is_extra[(seg_first + n_seg - 1)[ambiguous]] = True
    seg_role = role[seg_prot].copy()
    seg_role[is_extra] = (seg_role[is_extra] + 1 + rng.integers(0, n_roles - 1, int(is_extra.sum()))) % n_roles
    lines_per_role = n_keys // n_roles
    seg_line = (seg_role + n_roles * rng.integers(0, lines_per_role, seg_prot.shape[0])).astype(np.uint64)
    seg_kmers, seg_roles_chk = synthetic_db_lines(seg_line, K, n_roles, seed)
    assert np.array_equal(seg_roles_chk, seg_role.astype(np.int32))
    spacer = rng.integers(0, 60, seg_prot.shape[0])
    seg_start = np.concatenate([[0], np.cumsum(K + spacer)])
    total = int(seg_start[-1])
    res = alloc(total, np.uint8)
    res[:] = _AA20[rng.integers(0, 20, total)]
    pos = (seg_start[:-1, None] + np.arange(K)[None, :]).reshape(-1)
    res[pos] = seg_kmers.reshape(-1)
    off = alloc(n_prot + 1, np.uint64)
    off[:] = np.concatenate([[0], seg_start[1:][np.cumsum(n_seg) - 1]])
    # expectation: distinct planted k-mers of the protein's role (duplicate picks count once)
    key = np.zeros(seg_prot.shape[0], np.uint64)
    for j in range(K):
        key = key * np.uint64(32) + seg_kmers[:, j].astype(np.uint64)
    order = np.lexsort((key, seg_prot))
    sp, sk = seg_prot[order], key[order]
    first = np.ones(sp.shape[0], bool)
    first[1:] = (sp[1:] != sp[:-1]) | (sk[1:] != sk[:-1])
    distinct = np.bincount(sp[first], minlength=n_prot)
    exp_hits = np.where(ambiguous, 0, distinct).astype(np.int32)
    exp_role = np.where(ambiguous | (distinct < min_hits), -1, role).astype(np.int32)
    probes = int(np.maximum((off[1:] - off[:-1]).astype(np.int64) - K + 1, 0).sum())
    return res, off, exp_role, exp_hits, ambiguous, probes
