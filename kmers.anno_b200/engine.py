"""ctypes binding of libkmeranno.so (include/kmeranno.h).

Mirrors what the Java `KmerEngine` JNI class of INTEGRATION.md does: load the DB
(ApplyKmerProcessor.java:99-110), annotate CSR batches (:122-148).  No compute here and no
fallback — every method either calls the CUDA library or raises KmerAnnoError.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# KMERANNO_LIB selects another build of the same library (e.g. the bounds-checking `make debuglib`)
LIB_PATH = os.environ.get("KMERANNO_LIB") or os.path.join(_HERE, "libkmeranno.so")

FLAG_NONE, FLAG_CALLED, FLAG_AMBIGUOUS, FLAG_BELOW_MIN = 0, 1, 2, 3

# every symbol include/kmeranno.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "ka_create", "ka_destroy", "ka_last_error", "ka_set_option", "ka_db_load", "ka_db_load_synthetic", "ka_db_get_info",
    "ka_annotate", "ka_annotate_packed", "ka_db_get_alphabet", "ka_pack_residues", "ka_build", "ka_kmer_distance", "ka_batch_upload", "ka_annotate_resident", "ka_batch_download", "ka_batch_free",
    "ka_host_alloc", "ka_host_free", "ka_get_stats", "ka_probe_roofline", "ka_abi_version",
]


class KmerAnnoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"kmeranno error {code}: {msg}")
        self.code = code


class DbInfo(C.Structure):
    _fields_ = [("K", C.c_int32), ("n_symbols", C.c_int32), ("n_lines", C.c_uint64),
                ("n_keys", C.c_uint64), ("n_buckets", C.c_uint64), ("table_bytes", C.c_uint64),
                ("max_probe", C.c_uint32), ("slot_bits", C.c_uint32), ("filter_bytes", C.c_uint64),
                ("n_spilled", C.c_uint64), ("n_overflow", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [("sequences", C.c_uint64), ("residues", C.c_uint64), ("probes", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("kernel_ms", C.c_double), ("tile_kernel_ms", C.c_double), ("wall_ms", C.c_double)]


_lib = None


def load_library():
    """dlopen libkmeranno.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KmerAnnoError(-2, f"{LIB_PATH} is missing: run `make` or __graft_entry__.build(); "
                                "there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, u8p, i32p, u64p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
    lib.ka_abi_version.restype = C.c_int
    lib.ka_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    lib.ka_destroy.argtypes = [vp]
    lib.ka_destroy.restype = None
    lib.ka_last_error.argtypes = [vp]
    lib.ka_last_error.restype = C.c_char_p
    lib.ka_set_option.argtypes = [vp, C.c_char_p, C.c_double]
    lib.ka_db_load.argtypes = [vp, u8p, i32p, C.c_uint64, C.c_int]
    lib.ka_db_load_synthetic.argtypes = [vp, C.c_uint64, C.c_int, C.c_int32, C.c_uint64]
    lib.ka_db_get_info.argtypes = [vp, C.POINTER(DbInfo)]
    lib.ka_annotate.argtypes = [vp, u8p, u64p, C.c_uint64, C.c_int32, i32p, i32p, u8p]
    lib.ka_annotate_packed.argtypes = [vp, u8p, vp, C.c_uint64, C.c_int32, i32p, i32p, u8p]
    lib.ka_db_get_alphabet.argtypes = [vp, u8p]
    lib.ka_pack_residues.argtypes = [vp, u8p, C.c_uint64, C.c_uint64, u8p]
    lib.ka_build.argtypes = [vp, u8p, u64p, C.c_uint64, i32p, i32p, C.c_int, C.c_uint64, u8p, i32p,
                             C.POINTER(C.c_uint64), C.c_int]
    lib.ka_kmer_distance.argtypes = [vp, u8p, u64p, C.c_uint64, C.c_int, vp, u64p, C.c_uint64, vp, i32p, i32p, vp]
    lib.ka_batch_upload.argtypes = [vp, C.c_int, u8p, u64p, C.c_uint64, C.POINTER(vp)]
    lib.ka_annotate_resident.argtypes = [vp, vp, C.c_int32]
    lib.ka_batch_download.argtypes = [vp, vp, i32p, i32p, u8p]
    lib.ka_batch_free.argtypes = [vp, vp]
    lib.ka_batch_free.restype = None
    lib.ka_host_alloc.argtypes = [C.c_size_t]
    lib.ka_host_alloc.restype = C.c_void_p
    lib.ka_host_free.argtypes = [C.c_void_p]
    lib.ka_host_free.restype = None
    lib.ka_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.ka_probe_roofline.argtypes = [vp, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_int,
                                      C.POINTER(C.c_double)]
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data


def pinned_array(shape, dtype):
    """numpy array over cudaHostAlloc'd memory (ka_host_alloc); keeps itself alive."""
    lib = load_library()
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = lib.ka_host_alloc(max(n, 1))
    if not p:
        raise KmerAnnoError(-7, "ka_host_alloc failed")
    buf = (C.c_uint8 * max(n, 1)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[p] = buf
    return arr


_PINNED = {}


class Batch:
    def __init__(self, engine, handle, n):
        self.engine, self.handle, self.n = engine, handle, n

    def free(self):
        if self.handle:
            self.engine._lib.ka_batch_free(self.engine._h, self.handle)
            self.handle = None


class Engine:
    """One engine = the `kmerRoleMap` + the peg loop of ApplyKmerProcessor on the GPU(s)."""

    def __init__(self, device_ids=None):
        self._lib = load_library()
        h = C.c_void_p()
        if device_ids is None:
            rc = self._lib.ka_create(None, 0, C.byref(h))
        else:
            ids = (C.c_int * len(device_ids))(*device_ids)
            rc = self._lib.ka_create(ids, len(device_ids), C.byref(h))
        if rc != 0:
            raise KmerAnnoError(rc, (self._lib.ka_last_error(None) or b"").decode())
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ka_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise KmerAnnoError(rc, (self._lib.ka_last_error(self._h) or b"").decode())

    def set_option(self, name, value):
        self._check(self._lib.ka_set_option(self._h, name.encode(), float(value)))

    def db_load(self, kmers, role_ids, K):
        """kmers: uint8 array of n*K residue bytes (or list of str/bytes); role_ids: int32[n]."""
        if not isinstance(kmers, np.ndarray):
            kmers = np.frombuffer(b"".join(k if isinstance(k, bytes) else k.encode("latin-1")
                                           for k in kmers), dtype=np.uint8)
        kmers = np.ascontiguousarray(kmers, dtype=np.uint8).reshape(-1)
        role_ids = np.ascontiguousarray(role_ids, dtype=np.int32)
        n = role_ids.shape[0]
        if kmers.shape[0] != n * K:
            raise KmerAnnoError(-1, f"kmers holds {kmers.shape[0]} bytes, expected n*K = {n * K}")
        self._check(self._lib.ka_db_load(self._h, _ptr(kmers), _ptr(role_ids), n, K))

    def db_load_synthetic(self, n, K, n_roles, seed):
        """Device-generated DB of n lines (oversized-table configuration); synth.synthetic_db_lines
        regenerates any line on the host."""
        self._check(self._lib.ka_db_load_synthetic(self._h, int(n), int(K), int(n_roles), int(seed)))

    def db_info(self):
        info = DbInfo()
        self._check(self._lib.ka_db_get_info(self._h, C.byref(info)))
        return {f: getattr(info, f) for f, _ in DbInfo._fields_}

    def annotate(self, residues, offsets, min_hits=5, out=None):
        """residues uint8[R], offsets uint64[N+1] (host).  Returns (role, hits, flag)."""
        residues = np.ascontiguousarray(residues, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = max(offsets.shape[0] - 1, 0)
        if out is None:
            out = (np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.uint8))
        role, hits, flag = out
        self._check(self._lib.ka_annotate(self._h, _ptr(residues), _ptr(offsets), n, int(min_hits),
                                          _ptr(role), _ptr(hits), _ptr(flag)))
        return role, hits, flag

    def alphabet(self):
        """code_of_byte[256] of the loaded DB: digit 0..n-1 in byte order, 31 = byte not in the DB."""
        lut = np.empty(256, np.uint8)
        self._check(self._lib.ka_db_get_alphabet(self._h, _ptr(lut)))
        return lut

    def pack(self, residues, offsets, alloc=None, threads=None, out=None):
        """The packed form of a CSR batch for annotate_packed: (codes u8 stream, offsets u32).  The
        stream is indexed like `residues` (residue r at bits [5r, 5r+5)); ka_pack_residues on
        8-aligned slices, one per thread.  `out` = (codes, offsets32) of an earlier call to write into."""
        residues = np.ascontiguousarray(residues, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        total = int(offsets[-1]) if offsets.shape[0] else 0
        if total >= 1 << 32:
            raise KmerAnnoError(-10, "annotate_packed takes at most 2^32 - 1 residues per call")
        alloc = alloc or (lambda shape, dtype: np.empty(shape, dtype))
        if out is not None:
            codes, off32 = out
            if codes.shape[0] < (total * 5 + 7) // 8 + 16 or off32.shape[0] != offsets.shape[0]:
                raise ValueError("pack: `out` does not fit this batch")
        else:
            codes = alloc((total * 5 + 7) // 8 + 16, np.uint8)
            off32 = alloc(offsets.shape[0], np.uint32)
        off32[:] = offsets
        threads = threads or min(os.cpu_count() or 1, 32)
        step = max(1 << 20, -(-total // threads) + 7 & ~7)
        cuts = list(range(0, total, step)) + [total]

        def work(i):
            a, b = cuts[i], cuts[i + 1]
            rc = self._lib.ka_pack_residues(self._h, residues.ctypes.data + a, b - a, a, _ptr(codes))
            if rc != 0:
                raise KmerAnnoError(rc, (self._lib.ka_last_error(self._h) or b"").decode())

        if len(cuts) > 2:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(threads) as ex:
                list(ex.map(work, range(len(cuts) - 1)))
        elif total:
            work(0)
        return codes, off32

    def annotate_packed(self, codes, offsets32, min_hits=5, out=None):
        """codes: 5-bit stream (see pack), offsets32 uint32[N+1] (host).  Returns (role, hits, flag)."""
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        offsets32 = np.ascontiguousarray(offsets32, dtype=np.uint32)
        n = max(offsets32.shape[0] - 1, 0)
        if out is None:
            out = (np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.uint8))
        role, hits, flag = out
        self._check(self._lib.ka_annotate_packed(self._h, _ptr(codes), _ptr(offsets32), n, int(min_hits),
                                                 _ptr(role), _ptr(hits), _ptr(flag)))
        return role, hits, flag

    def build(self, residues, offsets, n_roles, peg_role, K, load_as_db=False):
        """GPU `build` (BuildKmerProcessor.java:138-223): returns (kmers u8[n*K], roles i32[n]), unordered."""
        residues = np.ascontiguousarray(residues, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n_roles = np.ascontiguousarray(n_roles, dtype=np.int32)
        peg_role = np.ascontiguousarray(peg_role, dtype=np.int32)
        n = offsets.shape[0] - 1
        lens = (offsets[1:] - offsets[:-1]).astype(np.int64)
        cap = int(np.maximum(lens[n_roles == 1] - K + 1, 0).sum())
        kmers = np.empty(max(cap * K, 1), np.uint8)
        roles = np.empty(max(cap, 1), np.int32)
        found = C.c_uint64()
        self._check(self._lib.ka_build(self._h, _ptr(residues), _ptr(offsets), n, _ptr(n_roles), _ptr(peg_role),
                                       K, cap, _ptr(kmers), _ptr(roles), C.byref(found), int(load_as_db)))
        return kmers[: found.value * K], roles[: found.value]

    def kmer_distance(self, residues, offsets, K, query_seq, group_offsets, cand_seq):
        """ProteinKmers.distance of every query against its candidates (GeneCopyProcessor.java:137-142).
        Returns (set_size i32[N], common i32[M], distance f64[M])."""
        residues = np.ascontiguousarray(residues, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        query_seq = np.ascontiguousarray(query_seq, dtype=np.uint32)
        group_offsets = np.ascontiguousarray(group_offsets, dtype=np.uint64)
        cand_seq = np.ascontiguousarray(cand_seq, dtype=np.uint32)
        n, q, m = offsets.shape[0] - 1, query_seq.shape[0], cand_seq.shape[0]
        if group_offsets.shape[0] != q + 1 or (q and int(group_offsets[-1]) != m):
            raise KmerAnnoError(-1, "group_offsets must have Q+1 entries ending at len(cand_seq)")
        size = np.empty(max(n, 1), np.int32)
        common = np.empty(max(m, 1), np.int32)
        dist = np.empty(max(m, 1), np.float64)
        self._check(self._lib.ka_kmer_distance(self._h, _ptr(residues), _ptr(offsets), n, int(K), _ptr(query_seq),
                                               _ptr(group_offsets), q, _ptr(cand_seq), _ptr(size), _ptr(common), _ptr(dist)))
        return size[:n], common[:m], dist[:m]

    def upload(self, residues, offsets, dev_index=0):
        residues = np.ascontiguousarray(residues, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = offsets.shape[0] - 1
        h = C.c_void_p()
        self._check(self._lib.ka_batch_upload(self._h, dev_index, _ptr(residues), _ptr(offsets), n,
                                              C.byref(h)))
        return Batch(self, h, n)

    def annotate_resident(self, batch, min_hits=5):
        self._check(self._lib.ka_annotate_resident(self._h, batch.handle, int(min_hits)))

    def download(self, batch):
        role = np.empty(batch.n, np.int32)
        hits = np.empty(batch.n, np.int32)
        flag = np.empty(batch.n, np.uint8)
        self._check(self._lib.ka_batch_download(self._h, batch.handle, _ptr(role), _ptr(hits), _ptr(flag)))
        return role, hits, flag

    def stats(self):
        s = Stats()
        self._check(self._lib.ka_get_stats(self._h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in Stats._fields_}

    def probe_roofline(self, table_bytes, n_probes, slot_bytes=32, reps=5, dev_index=0):
        v = C.c_double()
        self._check(self._lib.ka_probe_roofline(self._h, dev_index, int(table_bytes), int(n_probes),
                                                int(slot_bytes), int(reps), C.byref(v)))
        return v.value
