"""ctypes binding of oracle/libkaoracle.so (ka_oracle.c + ka_oracle_fast.c)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkaoracle.so")
_lib = None


class RoleCounter(C.Structure):
    _fields_ = [("role_id", C.c_int32), ("good", C.c_int32), ("bad", C.c_int32)]


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.orc_map_new.argtypes = [C.c_int64]; L.orc_map_new.restype = vp
        L.orc_map_free.argtypes = [vp]; L.orc_map_free.restype = None
        L.orc_map_size.argtypes = [vp]; L.orc_map_size.restype = C.c_uint64
        L.orc_map_capacity.argtypes = [vp]; L.orc_map_capacity.restype = C.c_uint32
        L.orc_map_put.argtypes = [vp, C.c_char_p, C.c_uint32, C.c_int32]
        L.orc_map_get.argtypes = [vp, C.c_char_p, C.c_uint32, C.POINTER(C.c_int32)]
        L.orc_map_remove.argtypes = [vp, C.c_char_p, C.c_uint32]
        L.orc_map_dump.argtypes = [vp, vp, vp, vp]; L.orc_map_dump.restype = C.c_uint64
        L.orc_db_load.argtypes = [vp, vp, C.c_uint64, C.c_int, C.c_int64]; L.orc_db_load.restype = vp
        L.orc_db_load_mt.argtypes = [vp, vp, C.c_uint64, C.c_int, C.c_int64, C.c_int]; L.orc_db_load_mt.restype = vp
        L.orc_apply.argtypes = [vp, vp, vp, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp]
        L.orc_build.argtypes = [vp, vp, C.c_uint64, vp, vp, C.c_int, C.c_int32, C.c_int, C.c_int, vp]
        L.orc_build.restype = vp
        L.orc_role_counter_init.argtypes = [C.POINTER(RoleCounter), C.c_int32]
        L.orc_role_counter_init.restype = None
        L.orc_role_counter_count.argtypes = [C.POINTER(RoleCounter), C.c_int32]
        L.orc_role_counter_is_good.argtypes = [C.POINTER(RoleCounter)]
        L.orc_count_peg_kmers_positions.argtypes = [C.c_char_p, C.c_uint64, C.c_int, vp]
        L.orc_count_peg_kmers_positions.restype = C.c_uint64
        L.orc_count_probes.argtypes = [vp, C.c_uint64, C.c_int]; L.orc_count_probes.restype = C.c_uint64
        L.orc_kmer_distance_pairs.argtypes = [vp, vp, vp, vp, C.c_uint64, C.c_int, vp, vp, vp, vp]
        L.orc_kmer_distance_pairs.restype = None
        L.orf_db_load.argtypes = [vp, vp, C.c_uint64, C.c_int]; L.orf_db_load.restype = vp
        L.orf_db_free.argtypes = [vp]; L.orf_db_free.restype = None
        L.orf_db_size.argtypes = [vp]; L.orf_db_size.restype = C.c_uint64
        L.orf_apply.argtypes = [vp, vp, vp, C.c_uint64, C.c_int, C.c_int, C.c_int, vp, vp, vp]
        _lib = L
    return _lib


def _kmers_array(kmers):
    if not isinstance(kmers, np.ndarray):
        kmers = np.frombuffer(b"".join(k if isinstance(k, bytes) else k.encode("latin-1") for k in kmers),
                              dtype=np.uint8)
    return np.ascontiguousarray(kmers, dtype=np.uint8).reshape(-1)


def count_probes(offsets, K):
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    return lib().orc_count_probes(offsets.ctypes.data, offsets.shape[0] - 1, K)


class OracleDb:
    """HashMap<String,String> kmerRoleMap of ApplyKmerProcessor.java:53,99-110, Java-shaped."""

    def __init__(self, kmers, role_ids, K, file_len_bytes=None, threads=1):
        self.L = lib()
        kmers = _kmers_array(kmers)
        role_ids = np.ascontiguousarray(role_ids, dtype=np.int32)
        n = role_ids.shape[0]
        assert kmers.shape[0] == n * K
        if file_len_bytes is None:
            file_len_bytes = n * (K + 8)
        self.K = K
        self.h = self.L.orc_db_load_mt(kmers.ctypes.data, role_ids.ctypes.data, n, K, file_len_bytes, threads)
        if not self.h:
            raise MemoryError("orc_db_load failed")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_map_free(self.h)
            self.h = None

    def size(self):
        return self.L.orc_map_size(self.h)

    def get(self, kmer):
        v = C.c_int32()
        b = kmer if isinstance(kmer, bytes) else kmer.encode("latin-1")
        return v.value if self.L.orc_map_get(self.h, b, len(b), C.byref(v)) else None

    def apply(self, residues, offsets, min_hits=5, distinct=True, include_last=True, threads=1, K=None):
        residues = np.ascontiguousarray(residues, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = offsets.shape[0] - 1
        role = np.empty(n, np.int32); hits = np.empty(n, np.int32); flag = np.empty(n, np.uint8)
        rc = self.L.orc_apply(self.h, residues.ctypes.data, offsets.ctypes.data, n, K or self.K, min_hits,
                              int(distinct), int(include_last), threads,
                              role.ctypes.data, hits.ctypes.data, flag.ctypes.data)
        if rc:
            raise ValueError("orc_apply rejected its arguments")
        return role, hits, flag


class FastDb:
    """Packed-integer CPU port (ka_oracle_fast.c)."""

    def __init__(self, kmers, role_ids, K):
        self.L = lib()
        kmers = _kmers_array(kmers)
        role_ids = np.ascontiguousarray(role_ids, dtype=np.int32)
        self.h = self.L.orf_db_load(kmers.ctypes.data, role_ids.ctypes.data, role_ids.shape[0], K)
        if not self.h:
            raise MemoryError("orf_db_load failed")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orf_db_free(self.h)
            self.h = None

    def size(self):
        return self.L.orf_db_size(self.h)

    def apply(self, residues, offsets, min_hits=5, distinct=True, threads=1):
        residues = np.ascontiguousarray(residues, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = offsets.shape[0] - 1
        role = np.empty(n, np.int32); hits = np.empty(n, np.int32); flag = np.empty(n, np.uint8)
        rc = self.L.orf_apply(self.h, residues.ctypes.data, offsets.ctypes.data, n, min_hits, int(distinct),
                              threads, role.ctypes.data, hits.ctypes.data, flag.ctypes.data)
        if rc:
            raise ValueError("orf_apply rejected its arguments")
        return role, hits, flag


def build_db(residues, offsets, n_roles, peg_role, K, n_good_roles, distinct=True, include_last=True):
    """BuildKmerProcessor.java:138-223.  Returns (kmers list[bytes], roles int32[], stats) in
    HashMap iteration order, or None when Java's capacity expression goes negative."""
    L = lib()
    residues = np.ascontiguousarray(residues, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n_roles = np.ascontiguousarray(n_roles, dtype=np.int32)
    peg_role = np.ascontiguousarray(peg_role, dtype=np.int32)
    stats = np.zeros(4, np.uint64)
    h = L.orc_build(residues.ctypes.data, offsets.ctypes.data, offsets.shape[0] - 1, n_roles.ctypes.data,
                    peg_role.ctypes.data, K, n_good_roles, int(distinct), int(include_last), stats.ctypes.data)
    if not h:
        return None
    n = L.orc_map_size(h)
    keys = np.empty(max(n * K, 1), np.uint8)
    vals = np.empty(max(n, 1), np.int32)
    L.orc_map_dump(h, keys.ctypes.data, None, vals.ctypes.data)
    L.orc_map_free(h)
    return keys[: n * K], vals[:n], {"buffered": int(stats[0]), "non_unique": int(stats[1]),
                                    "deleted_pass2": int(stats[2]), "remaining": int(stats[3])}


def kmer_distance_pairs(residues, offsets, qa, qb, K):
    """ProteinKmers.distance of the pairs (qa[m], qb[m]) of a CSR batch (GeneCopyProcessor.java:137-142):
    (size_a i32[M], size_b i32[M], common i32[M], distance f64[M])."""
    residues = np.ascontiguousarray(residues, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    qa = np.ascontiguousarray(qa, dtype=np.uint32)
    qb = np.ascontiguousarray(qb, dtype=np.uint32)
    m = qa.shape[0]
    sa, sb, co = np.empty(m, np.int32), np.empty(m, np.int32), np.empty(m, np.int32)
    dist = np.empty(m, np.float64)
    lib().orc_kmer_distance_pairs(residues.ctypes.data, offsets.ctypes.data, qa.ctypes.data, qb.ctypes.data,
                                  m, K, sa.ctypes.data, sb.ctypes.data, co.ctypes.data, dist.ctypes.data)
    return sa, sb, co, dist
