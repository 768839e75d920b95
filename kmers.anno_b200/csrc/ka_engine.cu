// ka_engine.cu — C ABI (include/kmeranno.h) over the sm_100a kernels: device table build,
// multi-device sharding, pipelined H2D -> plan -> tile -> big -> D2H chunks.
//
// Replaces /root/reference/src/main/java/org/theseed/proteins/kmers/anno/
// ApplyKmerProcessor.java:99-110 (DB load) and :122-148 (peg loop).  No CPU fallback: every
// path below either runs the CUDA kernels or returns an error code.
#include "ka_engine_internal.cuh"

using namespace ka;
using namespace kai;

namespace {
thread_local std::string g_create_error;
}

namespace kai {

int fail(ka_engine* e, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (e) e->err = buf; else g_create_error = buf;
    return code;
}

int dev_fail(Device& d, int code, const char* what, cudaError_t ce) {
    char buf[512];
    snprintf(buf, sizeof buf, "device %d: %s: %s", d.id, what, cudaGetErrorString(ce));
    d.err = code; d.errmsg = buf;
    return code;
}

int pipe_init(Device& d, Pipe& p) {
    DCK(d, cudaStreamCreateWithFlags(&p.st, cudaStreamNonBlocking));
    DCK(d, cudaEventCreate(&p.ev_k0));
    DCK(d, cudaEventCreate(&p.ev_t0));
    DCK(d, cudaEventCreate(&p.ev_t1));
    DCK(d, cudaEventCreate(&p.ev_k1));
    DCK(d, cudaEventCreateWithFlags(&p.done, cudaEventDisableTiming));
    DCK(d, cudaEventCreateWithFlags(&p.ev_in, cudaEventDisableTiming));
    DCK(d, cudaEventCreateWithFlags(&p.ev_out, cudaEventDisableTiming));
    DCK(d, cudaMalloc((void**)&p.ctr, 32));
    DCK(d, cudaMemset(p.ctr, 0, 32));
    return KA_OK;
}

void pipe_free(Pipe& p) {
    if (p.res) cudaFree(p.res);
    if (p.pk) cudaFree(p.pk);
    if (p.off) cudaFree(p.off);
    if (p.off32_in) cudaFree(p.off32_in);
    if (p.off32) cudaFree(p.off32);
    if (p.first) cudaFree(p.first);
    if (p.surv) cudaFree(p.surv);
    if (p.surv_cnt) cudaFree(p.surv_cnt);
    if (p.role) cudaFree(p.role);
    if (p.hits) cudaFree(p.hits);
    if (p.flag) cudaFree(p.flag);
    if (p.ctr) cudaFree(p.ctr);
    if (p.big) cudaFree(p.big);
    if (p.mid) cudaFree(p.mid);
    if (p.scratch) cudaFree(p.scratch);
    if (p.ev_k0) cudaEventDestroy(p.ev_k0);
    if (p.ev_t0) cudaEventDestroy(p.ev_t0);
    if (p.ev_t1) cudaEventDestroy(p.ev_t1);
    if (p.ev_k1) cudaEventDestroy(p.ev_k1);
    if (p.done) cudaEventDestroy(p.done);
    if (p.st) cudaStreamDestroy(p.st);
    if (p.ev_in) cudaEventDestroy(p.ev_in);
    if (p.ev_out) cudaEventDestroy(p.ev_out);
    if (p.via_buf || p.via_st || p.via_ev) {
        // (allocated on the via device; cudaFree / destroy work from any current device)
        if (p.via_buf) cudaFree(p.via_buf);
        if (p.via_st) cudaStreamDestroy(p.via_st);
        if (p.via_ev) cudaEventDestroy(p.via_ev);
    }
    p = Pipe();
}

// size the per-chunk device buffers
int pipe_reserve(Device& d, Pipe& p, uint64_t n_res, uint64_t n_seq, uint64_t n_tiles,
                 uint64_t n_long, uint64_t long_res, uint64_t n_mid, bool wide, bool need_bytes, bool need_codes,
                 bool need_surv) {
    int rc;
    if (need_bytes && (rc = ensure(d, p.res, p.res_cap, n_res + 64, "residues"))) return rc;
    // 5-bit codes of up to 127 lead residues + the chunk, in groups of 32 residues = 5 words, + over-read slack
    if (need_codes && (rc = ensure(d, p.pk, p.pk_cap, ((n_res + 127 + 31) / 32) * 5 + 32, "packed residues"))) return rc;
    size_t want_seq = n_seq + 1;
    if (want_seq > p.seq_cap || !p.off) {
        if (p.off) cudaFree(p.off);
        if (p.off32_in) cudaFree(p.off32_in);
        if (p.off32) cudaFree(p.off32);
        p.off32_in = p.off32 = nullptr;
        if (p.role) cudaFree(p.role);
        if (p.hits) cudaFree(p.hits);
        if (p.flag) cudaFree(p.flag);
        p.off = nullptr; p.role = p.hits = nullptr; p.flag = nullptr; p.seq_cap = 0;
        size_t n = want_seq + want_seq / 8 + 64;
        cudaError_t ce;
        if ((ce = cudaMalloc((void**)&p.off, n * 8)) != cudaSuccess ||
            (ce = cudaMalloc((void**)&p.off32_in, n * 4)) != cudaSuccess ||
            (ce = cudaMalloc((void**)&p.off32, n * 4)) != cudaSuccess ||
            (ce = cudaMalloc((void**)&p.role, n * 4)) != cudaSuccess ||
            (ce = cudaMalloc((void**)&p.hits, n * 4)) != cudaSuccess ||
            (ce = cudaMalloc((void**)&p.flag, n)) != cudaSuccess)
            return dev_fail(d, KA_ERR_OOM, "sequence buffers", ce);
        p.seq_cap = n;
    }
    if ((rc = ensure(d, p.first, p.first_cap, n_tiles + 2, "tile index"))) return rc;
    if (need_surv && ((rc = ensure(d, p.surv, p.surv_cap, n_res + 128 + 64, "survivor list")) ||
                      (rc = ensure(d, p.surv_cnt, p.surv_cnt_cap, 2 * (n_seq + 2 + n_mid), "survivor counts")))) return rc;
    if ((rc = ensure(d, p.big, p.big_cap, n_long + 1, "long-sequence list"))) return rc;
    if ((rc = ensure(d, p.mid, p.mid_cap, n_mid + 1, "mid-sequence tiles"))) return rc;
    if ((rc = ensure(d, p.scratch, p.scratch_cap, (wide ? 2 : 1) * (2 * long_res + 4), "long-sequence tokens")))
        return rc;
    return KA_OK;
}

// validate offsets of [cs, ce) and collect shape numbers; false = offsets not monotone
template <typename OffT>
static bool scan_offsets_t(const OffT* off, uint64_t cs, uint64_t ce, uint32_t long_seq, uint32_t mid_seq, int K,
                           ChunkShape& s) {
    s = ChunkShape();
    for (uint64_t i = cs; i < ce; i++) {
        if (off[i + 1] < off[i]) return false;
        uint64_t L = (uint64_t)off[i + 1] - (uint64_t)off[i];
        if (L > mid_seq) { s.n_long++; s.long_res += L; }
        else if (L > long_seq) { s.n_mid++; s.n_mid_seg += (L + LINE_MID_SEG - 1) / LINE_MID_SEG; }
        if (L >= (uint64_t)K) s.probes += L - K + 1;
    }
    s.n_res = (uint64_t)off[ce] - (uint64_t)off[cs];
    return true;
}

bool scan_offsets(const BatchIn& in, uint64_t cs, uint64_t ce, uint32_t long_seq, uint32_t mid_seq, int K,
                  ChunkShape& s) {
    return in.off64 ? scan_offsets_t(in.off64, cs, ce, long_seq, mid_seq, K, s)
                    : scan_offsets_t(in.off32, cs, ce, long_seq, mid_seq, K, s);
}

void fill_params(ka_engine* e, Device& d, Pipe& p, uint64_t base, uint64_t n_res, uint64_t n_seq,
                 int32_t min_hits, AnnotParams& ap) {
    ap.pk = nullptr;
    ap.pk_lead = 0;
    ap.res = p.res;
    ap.off = p.off;
    ap.base = base;
    ap.n_seq = (uint32_t)n_seq;
    ap.tile_span = e->tile_span;
    ap.long_seq = e->long_seq;
    ap.mid_seq = std::max(e->mid_seq, e->long_seq);
    ap.mid_desc = p.mid;
    ap.mid_count = p.ctr + 1;
    ap.ext_max = e->tile_span + e->long_seq;
    ap.n_tiles = (uint32_t)(n_res / e->tile_span + 1);
    tile_smem_bytes(ap.ext_max, &ap.res_bytes, false);
    ap.first = p.first;
    ap.tab = e->geom;
    ap.tab.sectors = d.table;
    ap.tab.ovf = d.ovf;
    ap.tab.shard_sectors = d.shard_sectors;
    ap.tab.shard_ovf = d.shard_ovf;
    ap.lut = d.lut;
    ap.min_hits = min_hits;
    ap.out_role = p.role;
    ap.out_hits = p.hits;
    ap.out_flag = p.flag;
    ap.big_count = p.ctr;
    ap.tok_cursor = reinterpret_cast<unsigned long long*>(p.ctr + 2);
    ap.big_list = p.big;
    ap.scratch = p.scratch;
    ap.dbg = p.ctr + 4;
}

// every tile-kernel instantiation may use up to the opt-in shared-memory limit of the device
int ensure_tile_smem(Device& d) {
    if (d.smem_set) return KA_OK;
    int optin = 0;
    DCK(d, cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, d.id));
    for (int cls : {32, 64, 128})
        for (int v = 0; v < N_VARIANTS; v++) DCK(d, tile_kernel_set_smem(cls, v, (size_t)optin - 2048));  // minus the static part
    DCK(d, tile_kernel_mode_set_smem((size_t)optin - 2048));
    DCK(d, line_tile_set_smem((size_t)optin - 2048));
    d.smem_set = (size_t)optin - 2048;
    return KA_OK;
}

// enqueue plan + tile + big on the pipe's stream, bracketed by timing events
int enqueue_kernels(ka_engine* e, Device& d, Pipe& p, const AnnotParams& ap, uint64_t n_long, uint64_t n_mid) {
    const bool wide = ap.tab.wide != 0;
    size_t smem = tile_smem_bytes(ap.ext_max, nullptr, wide);
    { int rc = ensure_tile_smem(d); if (rc) return rc; }
    if (smem > d.smem_set) return dev_fail(d, KA_ERR_INVALID, "tile shared memory exceeds the device limit", cudaErrorInvalidValue);
    DCK(d, cudaMemsetAsync(p.ctr, 0, 16, p.st));
    DCK(d, cudaEventRecord(p.ev_k0, p.st));
    DCK(d, launch_plan(ap, p.st));
    DCK(d, cudaEventRecord(p.ev_t0, p.st));
    DCK(d, launch_tiles(ap, 0, smem, p.st));
    d.launches += 2;
    if (n_mid) {
        // sequences of long_seq < L <= mid_seq: one tile each, same kernel, larger shared-memory shape
        AnnotParams am = ap;
        am.first = p.mid;
        am.n_tiles = (uint32_t)n_mid;
        am.ext_max = ap.mid_seq;
        size_t smem_mid = tile_smem_bytes(am.ext_max, &am.res_bytes, wide);
        if (smem_mid > d.smem_set) return dev_fail(d, KA_ERR_INVALID, "mid tile shared memory exceeds the device limit", cudaErrorInvalidValue);
        {
            cudaError_t ce = launch_tiles(am, 1, smem_mid, p.st);
            if (ce != cudaSuccess) {
                char buf[256];
                snprintf(buf, sizeof buf, "mid tile launch (tiles %u, smem %zu, cls %d)", am.n_tiles, smem_mid, am.tab.cls);
                return dev_fail(d, KA_ERR_CUDA, buf, ce);
            }
        }
        d.launches += 1;
    }
    DCK(d, cudaEventRecord(p.ev_t1, p.st));
    if (n_long) {
        int grid = (int)std::min<uint64_t>(n_long, (uint64_t)d.sm_count * 4);
        DCK(d, launch_big(ap, grid, p.st));
        d.launches += 1;
    }
    DCK(d, cudaEventRecord(p.ev_k1, p.st));
    return KA_OK;
}

int collect_times(Device& d, Pipe& p) {
#ifdef KA_DEBUG
    uint32_t dbg = 0;
    DCK(d, cudaMemcpy(&dbg, p.ctr + 4, 4, cudaMemcpyDeviceToHost));
    if (dbg) {
        char buf[96];
        snprintf(buf, sizeof buf, "KA_DEBUG bounds check failed in a kernel (codes 0x%x)", dbg);
        d.err = KA_ERR_CUDA; d.errmsg = buf;
        cudaMemset(p.ctr + 4, 0, 4);
        return d.err;
    }
#endif
    float a = 0, b = 0;
    DCK(d, cudaEventElapsedTime(&a, p.ev_k0, p.ev_k1));
    DCK(d, cudaEventElapsedTime(&b, p.ev_t0, p.ev_t1));
    d.kernel_ms += a;
    d.tile_ms += b;
    return KA_OK;
}

void set_l2_window(ka_engine* e, Device& d, cudaStream_t st) {
    // The line table keeps its filter in L2 through cache hints on the loads themselves (evict_last on
    // the filter words, evict_first on the table lines); a persisting carve-out measured slower
    // (microbench/filter_probe.cu: 93 -> 86 G probes/s), so no window is set for it.
    if (!e->l2_persist || !d.table || e->line) {
        cudaStreamAttrValue v;
        memset(&v, 0, sizeof v);
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v);
        // give back the carve-out an earlier table of this process may have set: it is taken from the L2 that
        // ordinary data (here: the filter words) can use — measured 8.4 -> 9.5 ms on the line table when left behind
        size_t cur = 0;
        if (cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize) == cudaSuccess && cur != 0) {
            cudaCtxResetPersistingL2Cache();
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
        }
        cudaGetLastError();
        return;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, d.id) != cudaSuccess) return;
    if (prop.persistingL2CacheMaxSize <= 0 || prop.accessPolicyMaxWindowSize <= 0) return;
    size_t persist = (size_t)prop.persistingL2CacheMaxSize;
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist);
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof v);
    // pin as much of the table as the persisting carve-out holds
    size_t span = (size_t)e->info.table_bytes;
    size_t win = std::min<size_t>(span, (size_t)prop.accessPolicyMaxWindowSize);
    v.accessPolicyWindow.base_ptr = (void*)d.table;
    v.accessPolicyWindow.num_bytes = win;
    v.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)persist / (double)win);
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v);
    cudaGetLastError();  // the window is an optimisation; never fail the call on it
}

// ---- line table (slot class 16): parameters and launches of one chunk ----
void fill_line_params(ka_engine* e, Device& d, Pipe& p, uint64_t n_res, uint64_t n_seq, int32_t min_hits, LineParams& lp) {
    lp.pk = p.pk;
    lp.off = p.off32;
    lp.n_seq = (uint32_t)n_seq;
    lp.tile_span = line_span(e, d, n_res);
    lp.n_tiles = (uint32_t)(n_res / lp.tile_span + 1);
    lp.long_seq = e->long_seq;
    lp.mid_seq = std::max(e->mid_seq, e->long_seq);
    lp.ext_max = lp.tile_span + e->long_seq;
    lp.first = p.first;
    lp.surv = p.surv;
    lp.surv_cnt = p.surv_cnt;
    lp.hit_cnt = p.surv_cnt + (p.surv_cnt_cap / 2);
    lp.mid_desc = p.mid;
    lp.mid_count = p.ctr + 1;
    lp.big_count = p.ctr;
    lp.tok_cursor = reinterpret_cast<unsigned long long*>(p.ctr + 2);
    lp.big_list = p.big;
    lp.scratch = p.scratch;
    lp.tab = e->lgeom;
    lp.tab.lines = d.table;
    lp.tab.ovf = d.ovf;
    lp.tab.filt = e->filter ? d.filt : nullptr;
    lp.min_hits = min_hits;
    lp.out_role = p.role;
    lp.out_hits = p.hits;
    lp.out_flag = p.flag;
    lp.dbg = p.ctr + 4;
}

// plan + tiles (+ single-sequence tiles, + long sequences) of the line table.  The kernels of ALL chunks of a device
// run on one compute stream, in chunk order: the copies of the pipes still overlap them, but the passes of two
// chunks never run side by side — the probe pass of one would evict the filter words the filter pass of the other
// needs from L2 (measured: 29.9 -> 32.7-38 ms end to end when every pipe launched on its own stream).
int enqueue_line_kernels(ka_engine* e, Device& d, Pipe& p, const LineParams& lp, bool off_is_64, uint64_t origin,
                         uint64_t n_long, uint64_t n_mid, bool solo) {
    const size_t smem = line_tile_smem_bytes(lp.ext_max);
    { int rc = ensure_tile_smem(d); if (rc) return rc; }
    if (smem > d.smem_set) return dev_fail(d, KA_ERR_INVALID, "tile shared memory exceeds the device limit", cudaErrorInvalidValue);
    // (a call that is a single chunk on this device has nothing to be kept apart from: it stays on the pipe's stream)
    if (!solo && !d.line_st) DCK(d, cudaStreamCreateWithFlags(&d.line_st, cudaStreamNonBlocking));
    cudaStream_t st = solo ? p.st : d.line_st;
    DCK(d, cudaMemsetAsync(p.ctr, 0, 16, p.st));
    if (!solo) {
        DCK(d, cudaEventRecord(p.ev_in, p.st));
        DCK(d, cudaStreamWaitEvent(st, p.ev_in, 0));
    }
    DCK(d, cudaEventRecord(p.ev_k0, st));
    DCK(d, launch_line_plan(lp, off_is_64 ? p.off : nullptr, p.off32_in, origin, st));
    d.launches += 1;
    DCK(d, cudaEventRecord(p.ev_t0, st));
    // filter and probe passes over the single-sequence tiles and the ordinary tiles in one launch each, then the
    // tally pass, whose shared memory depends on the tile size, once per kind.  One CTA per tile: the hardware's CTA
    // scheduler balances tiles of very different cost (a grid of resident warps striding over the tiles measured
    // 0.95 against 0.75 ms for the probe pass of 60 proteomes).
    const unsigned all = 0x7fffffffu;
    static const bool trace = getenv("KA_LINE_TRACE") != nullptr;     // debugging aid: per-pass times on stderr (synchronises)
    cudaEvent_t tv[5] = {};
    if (trace) for (auto& ev : tv) { cudaEventCreate(&ev); }
    if (trace) cudaEventRecord(tv[0], st);
    LineParams q = lp;
    q.n_mid_tiles = (uint32_t)n_mid; q.tally_mid = 0;   // (n_mid counts SEGMENTS of mid sequences here)
    q.tile0 = 0; q.tile1 = (uint32_t)n_mid + lp.n_tiles;
    DCK(d, launch_line_filter(q, all, st));
    if (trace) cudaEventRecord(tv[1], st);
    DCK(d, launch_line_probe(q, all, st));
    if (trace) cudaEventRecord(tv[2], st);
    q.n_mid_tiles = 0; q.tally_mid = 0; q.tile1 = lp.n_tiles;
    DCK(d, launch_line_tally(q, all, st));
    if (trace) cudaEventRecord(tv[3], st);
    d.launches += 3;
    if (n_mid) {
        LineParams lm = lp;
        lm.first = p.mid;
        lm.n_mid_tiles = 0; lm.tally_mid = 1; lm.tile0 = 0; lm.tile1 = (uint32_t)n_mid;
        lm.ext_max = lp.mid_seq;
        const size_t smem_mid = line_tile_smem_bytes(lm.ext_max);
        if (smem_mid > d.smem_set) return dev_fail(d, KA_ERR_INVALID, "mid tile shared memory exceeds the device limit", cudaErrorInvalidValue);
        DCK(d, launch_line_tally(lm, all, st));
        d.launches += 1;
    }
    if (trace) {
        cudaEventRecord(tv[4], st);
        cudaEventSynchronize(tv[4]);
        float f = 0, pr = 0, ta = 0, tm = 0;
        cudaEventElapsedTime(&f, tv[0], tv[1]); cudaEventElapsedTime(&pr, tv[1], tv[2]);
        cudaEventElapsedTime(&ta, tv[2], tv[3]); cudaEventElapsedTime(&tm, tv[3], tv[4]);
        fprintf(stderr, "[line trace dev%d] %u tiles of %u + %llu mid segments: filter %.1f us, probe %.1f us, tally %.1f us, mid tally %.1f us\n",
                d.id, lp.n_tiles, lp.tile_span, (unsigned long long)n_mid, f * 1e3, pr * 1e3, ta * 1e3, tm * 1e3);
        for (auto& ev : tv) cudaEventDestroy(ev);
    }
    DCK(d, cudaEventRecord(p.ev_t1, st));
    if (n_long) {
        int grid = (int)std::min<uint64_t>(n_long, (uint64_t)d.sm_count * 4);
        DCK(d, launch_line_big(lp, grid, st));
        d.launches += 1;
    }
    DCK(d, cudaEventRecord(p.ev_k1, st));
    if (!solo) {
        DCK(d, cudaEventRecord(p.ev_out, st));
        DCK(d, cudaStreamWaitEvent(p.st, p.ev_out, 0));
    }
    return KA_OK;
}

// Host -> device copy of a chunk, directly or through the via device: H2D into a staging buffer on the via GPU
// (its PCIe path), then a peer copy over NVLink on the pipe's stream.
static int h2d_chunk(ka_engine* e, Device& d, Pipe& p, void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return KA_OK;
    if (e->ingest_via < 0 || e->ingest_via == d.id) {
        DCK(d, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, p.st));
        return KA_OK;
    }
    const int via = e->ingest_via;
    if (!p.via_st || bytes > p.via_cap) {
        DCK(d, cudaSetDevice(via));
        if (!p.via_st) { DCK(d, cudaStreamCreateWithFlags(&p.via_st, cudaStreamNonBlocking)); DCK(d, cudaEventCreateWithFlags(&p.via_ev, cudaEventDisableTiming)); }
        if (bytes > p.via_cap) {
            if (p.via_buf) cudaFree(p.via_buf);
            p.via_buf = nullptr; p.via_cap = 0;
            const size_t want = bytes + bytes / 8 + 4096;
            cudaError_t ce = cudaMalloc((void**)&p.via_buf, want);
            if (ce != cudaSuccess) { cudaSetDevice(d.id); return dev_fail(d, KA_ERR_OOM, "ingest_via staging buffer", ce); }
            p.via_cap = want;
        }
        DCK(d, cudaSetDevice(d.id));
    }
    // (the staging buffer of this pipe is free: the pipe's previous chunk has completed, see p.busy)
    DCK(d, cudaMemcpyAsync(p.via_buf, src, bytes, cudaMemcpyHostToDevice, p.via_st));
    DCK(d, cudaEventRecord(p.via_ev, p.via_st));
    DCK(d, cudaStreamWaitEvent(p.st, p.via_ev, 0));
    DCK(d, cudaMemcpyPeerAsync(dst, d.id, p.via_buf, via, bytes, p.st));
    return KA_OK;
}

template <typename OffT>
static uint64_t chunk_end(const OffT* off, uint64_t cs, uint64_t s_end, uint64_t chunk_residues) {
    const uint64_t lim = (uint64_t)off[cs] + chunk_residues;
    // last sequence boundary at or below the limit (at least one sequence per chunk)
    uint64_t lo = cs + 1, hi = s_end + 1;
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        if ((uint64_t)off[mid] <= lim) lo = mid + 1; else hi = mid;
    }
    return lo - 1;
}

// Annotate sequences [s_begin, s_end) of the host batch on device d.  Four combinations:
//   line table   + packed input : H2D codes                 -> line kernels        (fast path)
//   line table   + byte input   : H2D bytes -> pack kernel  -> line kernels
//   sector table + byte input   : H2D bytes                 -> sector kernels
//   sector table + packed input : H2D codes -> unpack kernel -> sector kernels
int annotate_range(ka_engine* e, Device& d, const BatchIn& in, uint64_t s_begin, uint64_t s_end, int32_t min_hits,
                   int32_t* out_role, int32_t* out_hits, uint8_t* out_flag) {
    DCK(d, cudaSetDevice(d.id));
    d.kernel_ms = d.tile_ms = 0; d.launches = 0; d.h2d = d.d2h = 0; d.probes = 0;
    const bool line = e->line, packed = in.packed();
    int slot = 0;
    unsigned n_chunk = 0;
    uint64_t cs = s_begin;
    while (cs < s_end) {
        // chunk = as many whole sequences as fit in chunk_residues (at least one)
        // Line table and packed input, automatic chunk size: 16 Mi residues first, doubling up to 256 Mi, and never more than half of
        // what is left — the first copy and the last launch group, which nothing can hide, stay short, and the groups
        // in between are few and long (every group ends in the tail of three kernels: 23 chunks of 64 Mi cost
        // 26.2 ms of kernels against 22.8 for one launch).
        uint64_t chunk = chunk_of(e, false);
        if (line && packed && e->chunk_residues == 0) {     // (the byte form is copy-bound on any box: 64 Mi measured best)
            const uint64_t left = in.off(s_end) - in.off(cs);
            chunk = std::min<uint64_t>(256ull << 20, (16ull << 20) << std::min(n_chunk, 8u));
            chunk = std::min(chunk, std::max<uint64_t>(16ull << 20, left / 2));
        }
        n_chunk++;
        uint64_t ce = in.off64 ? chunk_end(in.off64, cs, s_end, chunk) : chunk_end(in.off32, cs, s_end, chunk);
        if (ce <= cs) ce = cs + 1;
        if (ce - cs > 0xfffffff0ull) ce = cs + 0xfffffff0ull;
        ChunkShape sh;
        if (!scan_offsets(in, cs, ce, e->long_seq, e->mid_seq, e->info.K, sh)) {
            d.err = KA_ERR_OFFSETS; d.errmsg = "offsets are not monotone";
            return d.err;
        }
        if (sh.n_res > 0x7fffff00ull) {
            d.err = KA_ERR_TOO_BIG; d.errmsg = "a single sequence exceeds 2^31 residues";
            return d.err;
        }
        d.probes += sh.probes;
        Pipe& p = d.pipe[slot];
        if (p.busy) {
            DCK(d, cudaEventSynchronize(p.done));
            int rc = collect_times(d, p);
            if (rc) return rc;
            p.busy = false;
        }
        const uint64_t n = ce - cs;
        const uint64_t n_tiles = tiles_of(e, d, sh.n_res);
        int rc = pipe_reserve(d, p, sh.n_res, n, n_tiles, sh.n_long, sh.long_res, line ? sh.n_mid_seg : sh.n_mid, e->geom.wide != 0,
                              !packed || !line, packed || line, line);
        if (rc) return rc;
        const uint64_t r_begin = in.off(cs), r_end = in.off(ce);
        const uint64_t origin = r_begin & ~127ull;          // chunk-relative residue 0 (5 * 128 bits = 80 bytes: byte aligned)
        const uint32_t lead = (uint32_t)(r_begin - origin);
        if (packed) {
            const uint64_t byte0 = origin * 5 / 8, byte1 = (r_end * 5 + 7) / 8;
            if ((rc = h2d_chunk(e, d, p, p.pk, in.codes + byte0, byte1 - byte0))) return rc;
            DCK(d, cudaMemcpyAsync(p.off32_in, in.off32 + cs, (n + 1) * 4, cudaMemcpyHostToDevice, p.st));
            d.h2d += (byte1 - byte0) + (n + 1) * 4;
        } else {
            if ((rc = h2d_chunk(e, d, p, p.res, in.residues + r_begin, sh.n_res))) return rc;
            DCK(d, cudaMemcpyAsync(p.off, in.off64 + cs, (n + 1) * 8, cudaMemcpyHostToDevice, p.st));
            d.h2d += sh.n_res + (n + 1) * 8;
        }
        if (line) {
            if (!packed) { DCK(d, launch_pack(p.res, lead, sh.n_res, d.lut5, p.pk, p.st)); d.launches += 1; }
            LineParams lp;
            fill_line_params(e, d, p, sh.n_res, n, min_hits, lp);
            rc = enqueue_line_kernels(e, d, p, lp, !packed, origin, sh.n_long, sh.n_mid_seg, cs == s_begin && ce == s_end);
        } else {
            // packed input: the narrow probe kernels stage the code stream themselves; the wide / sharded forms and
            // the long-sequence kernel read residue bytes, so those chunks are expanded first
            const bool stage_packed = packed && !e->geom.wide && e->geom.n_shards <= 1;
            if (packed) {
                if (!stage_packed || sh.n_long) { DCK(d, launch_unpack(p.pk, lead, sh.n_res, d.inv32, p.res, p.st)); d.launches += 1; }
                DCK(d, launch_widen_offsets(p.off32_in, n + 1, p.off, p.st));
                d.launches += 1;
            }
            AnnotParams ap;
            fill_params(e, d, p, r_begin, sh.n_res, n, min_hits, ap);
            if (stage_packed) { ap.pk = p.pk; ap.pk_lead = lead; }
            rc = enqueue_kernels(e, d, p, ap, sh.n_long, sh.n_mid);
        }
        if (rc) return rc;
        DCK(d, cudaMemcpyAsync(out_role + cs, p.role, n * 4, cudaMemcpyDeviceToHost, p.st));
        DCK(d, cudaMemcpyAsync(out_hits + cs, p.hits, n * 4, cudaMemcpyDeviceToHost, p.st));
        d.d2h += n * 8;
        if (out_flag) {
            DCK(d, cudaMemcpyAsync(out_flag + cs, p.flag, n, cudaMemcpyDeviceToHost, p.st));
            d.d2h += n;
        }
        DCK(d, cudaEventRecord(p.done, p.st));
        p.busy = true;
        slot = (slot + 1) % NPIPE;
        cs = ce;
    }
    for (int i = 0; i < NPIPE; i++) {
        Pipe& p = d.pipe[i];
        if (p.busy) {
            DCK(d, cudaEventSynchronize(p.done));
            int rc = collect_times(d, p);
            if (rc) return rc;
            p.busy = false;
        }
    }
    return KA_OK;
}

}  // namespace kai

// ======================================================================================
// C ABI
// ======================================================================================
extern "C" {

int ka_abi_version(void) { return KA_ABI_VERSION; }

int ka_create(const int* device_ids, int n_devices, ka_engine** out) {
    if (!out) return fail(nullptr, KA_ERR_INVALID, "ka_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t ce = cudaGetDeviceCount(&count);
    if (ce != cudaSuccess || count == 0)
        return fail(nullptr, KA_ERR_NO_DEVICE,
                    "ka_create: no CUDA device (%s); this engine has no CPU fallback",
                    ce == cudaSuccess ? "device count is 0" : cudaGetErrorString(ce));
    int dflt = 0;
    if (!device_ids) { device_ids = &dflt; n_devices = 1; }
    if (n_devices < 1) return fail(nullptr, KA_ERR_INVALID, "ka_create: n_devices < 1");
    ka_engine* e = new ka_engine();
    e->devs.resize(n_devices);
    for (int i = 0; i < n_devices; i++) {
        Device& d = e->devs[i];
        d.id = device_ids[i];
        if (d.id < 0 || d.id >= count) {
            int bad = d.id;
            e->devs.resize(i);
            ka_destroy(e);
            return fail(nullptr, KA_ERR_NO_DEVICE, "ka_create: device id %d out of range (0..%d)", bad, count - 1);
        }
        cudaDeviceProp prop;
        if (cudaSetDevice(d.id) != cudaSuccess || cudaGetDeviceProperties(&prop, d.id) != cudaSuccess) {
            e->devs.resize(i);
            ka_destroy(e);
            return fail(nullptr, KA_ERR_NO_DEVICE, "ka_create: cannot open device %d", device_ids[i]);
        }
        if (prop.major < 10) {
            int bad = d.id;
            e->devs.resize(i);
            ka_destroy(e);
            return fail(nullptr, KA_ERR_NO_DEVICE,
                        "ka_create: device %d is sm_%d%d; this build is sm_100a only", bad, prop.major, prop.minor);
        }
        d.sm_count = prop.multiProcessorCount;
        int rc = KA_OK;
        for (int k = 0; k < NPIPE && rc == KA_OK; k++) rc = pipe_init(d, d.pipe[k]);
        if (rc == KA_OK && (cudaMalloc((void**)&d.lut, 256) != cudaSuccess || cudaMalloc((void**)&d.lut5, 256) != cudaSuccess ||
                            cudaMalloc((void**)&d.inv32, 32) != cudaSuccess)) rc = KA_ERR_OOM;
        if (rc) {
            std::string m = d.errmsg.empty() ? "device allocation failed" : d.errmsg;
            e->devs.resize(i + 1);
            ka_destroy(e);
            return fail(nullptr, rc, "ka_create: %s", m.c_str());
        }
    }
    *out = e;
    return KA_OK;
}

void ka_destroy(ka_engine* e) {
    if (!e) return;
    for (Device& d : e->devs) {
        cudaSetDevice(d.id);
        for (int k = 0; k < NPIPE; k++) pipe_free(d.pipe[k]);
        if (d.table) cudaFree(d.table);
        if (d.ovf) cudaFree(d.ovf);
        if (d.filt) cudaFree(d.filt);
        if (d.line_st) cudaStreamDestroy(d.line_st);
        if (d.lut5) cudaFree(d.lut5);
        if (d.inv32) cudaFree(d.inv32);
        route_destroy_comm(d);
        for (auto& ln : d.lane) {
            for (void* q : {(void*)ln.r_keys, (void*)ln.r_send, (void*)ln.r_recv, (void*)ln.r_ans_recv, (void*)ln.r_ans_sorted,
                            (void*)ln.r_small, (void*)ln.r_pos}) if (q) cudaFree(q);
            if (ln.h_cnt) cudaFreeHost(ln.h_cnt);
            if (ln.ev_counts) cudaEventDestroy(ln.ev_counts);
            for (cudaEvent_t ev : {ln.ev_scatter, ln.ev_lookup, ln.ev_tally}) if (ev) cudaEventDestroy(ev);
        }
        if (d.ev_route0) cudaEventDestroy(d.ev_route0);
        if (d.ev_route1) cudaEventDestroy(d.ev_route1);
        if (d.shard_sectors) cudaFree((void*)d.shard_sectors);
        if (d.shard_ovf) cudaFree((void*)d.shard_ovf);
        if (d.lut) cudaFree(d.lut);
    }
    delete e;
}

const char* ka_last_error(const ka_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int ka_set_option(ka_engine* e, const char* name, double v) {
    if (!e || !name) return KA_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    const std::string n(name);
    // every value is validated into a copy; the engine changes only when the whole call succeeds
    double load_factor = e->load_factor;
    uint32_t tile_span = e->tile_span, long_seq = e->long_seq, mid_seq = e->mid_seq;
    uint64_t chunk_residues = e->chunk_residues;
    int l2_persist = e->l2_persist, table_mode = e->table_mode, wide = e->wide, filter = e->filter, slot_bits = e->slot_bits;
    int resident_packed = e->resident_packed, ingest_via = e->ingest_via;
    if (n == "load_factor") {
        if (!(v > 0.0 && v <= 0.9)) return fail(e, KA_ERR_INVALID, "load_factor must be in (0, 0.9]");
        load_factor = v;
    } else if (n == "tile_span") {
        if (!(v >= 256 && v <= 65536)) return fail(e, KA_ERR_INVALID, "tile_span must be in [256, 65536]");
        tile_span = (uint32_t)v & ~15u;
        if (long_seq < tile_span) long_seq = tile_span;
    } else if (n == "long_seq") {
        if (!(v >= 256 && v <= (1 << 20))) return fail(e, KA_ERR_INVALID, "long_seq must be in [256, 2^20]");
        long_seq = (uint32_t)v;
        if (long_seq < tile_span) long_seq = tile_span;
    } else if (n == "mid_seq") {
        if (!(v >= 256 && v <= 49152)) return fail(e, KA_ERR_INVALID, "mid_seq must be in [256, 49152]");
        mid_seq = (uint32_t)v;
    } else if (n == "chunk_residues") {
        if (!(v == 0 || (v >= 4096 && v <= (double)(1ull << 30)))) return fail(e, KA_ERR_INVALID, "chunk_residues must be 0 (automatic) or in [4096, 2^30]");
        chunk_residues = (uint64_t)v;
    } else if (n == "l2_persist") {
        l2_persist = v != 0;
    } else if (n == "table_mode") {
        if (v != 0 && v != 1 && v != 2 && v != 3)
            return fail(e, KA_ERR_INVALID, "table_mode must be 0 (replicated), 1 (sharded, peer loads), 2 (sharded, NCCL routed) or 3 (sharded, routed by peer stores)");
        table_mode = (int)v;
    } else if (n == "wide") {
        wide = v != 0;
    } else if (n == "filter") {
        filter = v != 0;
    } else if (n == "resident_packed") {
        resident_packed = v != 0;
    } else if (n == "ingest_via") {
        int count = 0;
        cudaGetDeviceCount(&count);
        if (v != -1 && !(v >= 0 && v < count && v == (int)v)) return fail(e, KA_ERR_INVALID, "ingest_via must be -1 or a CUDA device id");
        if (v >= 0 && e->devs.size() != 1) return fail(e, KA_ERR_INVALID, "ingest_via applies to a single-device engine");
        if (v >= 0 && (int)v != e->devs[0].id) {
            int can = 0;
            cudaDeviceCanAccessPeer(&can, e->devs[0].id, (int)v);
            if (!can) return fail(e, KA_ERR_NO_DEVICE, "ingest_via: device %d cannot reach device %d over NVLink / P2P", e->devs[0].id, (int)v);
            cudaSetDevice(e->devs[0].id);
            cudaError_t pe = cudaDeviceEnablePeerAccess((int)v, 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) return fail(e, KA_ERR_CUDA, "ingest_via: cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(pe));
            cudaGetLastError();
        }
        ingest_via = (int)v;
    } else if (n == "slot_bits") {
        if (v != 0 && v != 16 && v != 32 && v != 64 && v != 128) return fail(e, KA_ERR_INVALID, "slot_bits must be 0, 16, 32, 64 or 128");
        slot_bits = (int)v;
    } else {
        return fail(e, KA_ERR_INVALID, "unknown option '%s'", name);
    }
    const uint32_t mid_eff = std::max(mid_seq, long_seq);
    if (tile_smem_bytes(tile_span + long_seq, nullptr, true) > 225 * 1024 || tile_smem_bytes(mid_eff, nullptr, true) > 225 * 1024 ||
        line_tile_smem_bytes(tile_span + long_seq) > 225 * 1024 || line_tile_smem_bytes(mid_eff) > 225 * 1024)
        return fail(e, KA_ERR_INVALID, "tile_span + long_seq (or mid_seq) needs more than 227 KB of shared memory");
    e->load_factor = load_factor; e->tile_span = tile_span; e->long_seq = long_seq; e->mid_seq = mid_seq;
    e->chunk_residues = chunk_residues; e->l2_persist = l2_persist; e->table_mode = table_mode; e->wide = wide;
    e->filter = filter; e->slot_bits = slot_bits; e->resident_packed = resident_packed; e->ingest_via = ingest_via;
    return KA_OK;
}

int ka_db_get_info(ka_engine* e, ka_db_info* out) {
    if (!e || !out) return KA_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    if (!e->have_db) return fail(e, KA_ERR_NO_DB, "no k-mer database loaded");
    *out = e->info;
    return KA_OK;
}

static int annotate_impl(ka_engine* e, const BatchIn& in, uint64_t N, int32_t min_hits, int32_t* out_role,
                         int32_t* out_hits, uint8_t* out_flag, const char* who) {
    std::lock_guard<std::mutex> lk(e->mu);
    if (!e->have_db) return fail(e, KA_ERR_NO_DB, "%s: no k-mer database loaded", who);
    if (min_hits < 1) return fail(e, KA_ERR_INVALID, "%s: min_hits must be positive", who);  // ApplyKmerProcessor.java:91-92
    if (N && ((!in.off64 && !in.off32) || !out_role || !out_hits)) return fail(e, KA_ERR_INVALID, "%s: NULL argument", who);
    if (N && in.off(N) > in.off(0) && !in.residues && !in.codes) return fail(e, KA_ERR_INVALID, "%s: residues is NULL", who);
    auto t0 = std::chrono::steady_clock::now();
    e->stats = ka_stats{};
    if (N == 0) return KA_OK;
    if (in.off(N) < in.off(0)) return fail(e, KA_ERR_OFFSETS, "%s: offsets are not monotone", who);

    // residue-balanced contiguous ranges, one per device
    size_t nd = e->devs.size();
    std::vector<uint64_t> cut(nd + 1, 0);
    cut[nd] = N;
    uint64_t total = in.off(N) - in.off(0);
    for (size_t i = 1; i < nd; i++) {
        const uint64_t target = in.off(0) + total / nd * i;
        uint64_t lo = 0, hi = N + 1;                         // first index with off >= target
        while (lo < hi) { const uint64_t mid = lo + ((hi - lo) >> 1); if (in.off(mid) < target) lo = mid + 1; else hi = mid; }
        cut[i] = std::min<uint64_t>(std::max<uint64_t>(lo, cut[i - 1]), N);
    }
    int rc;
    if (e->db_table_mode >= 2) {
        if ((e->db_table_mode == 2 && !e->nccl_ready) || e->geom.n_shards <= 1) return fail(e, KA_ERR_INVALID, "%s: the loaded table is not a routed table", who);
        // routed sharded table: every device must walk every round, even with an empty range
        RouteShared shared((int)nd);
        rc = for_each_device(e, [&](Device& d, int i) {
            return annotate_routed_range(e, d, i, shared, in, cut[i], cut[i + 1], min_hits, out_role, out_hits, out_flag);
        });
    } else {
        rc = for_each_device(e, [&](Device& d, int i) {
            if (cut[i] == cut[i + 1]) { d.kernel_ms = d.tile_ms = 0; d.launches = d.h2d = d.d2h = d.probes = 0; return (int)KA_OK; }
            return annotate_range(e, d, in, cut[i], cut[i + 1], min_hits, out_role, out_hits, out_flag);
        });
    }
    if (rc) {
        for (Device& d : e->devs) { cudaSetDevice(d.id); cudaDeviceSynchronize(); for (auto& p : d.pipe) p.busy = false; }
        return rc;
    }
    ka_stats& s = e->stats;
    s.sequences = N;
    s.residues = total;
    for (Device& d : e->devs) {
        s.kernel_launches += d.launches;
        s.probes += d.probes;
        s.h2d_bytes += d.h2d;
        s.d2h_bytes += d.d2h;
        s.kernel_ms = std::max(s.kernel_ms, d.kernel_ms);
        s.tile_kernel_ms = std::max(s.tile_kernel_ms, d.tile_ms);
    }
    s.wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return KA_OK;
}

int ka_annotate(ka_engine* e, const uint8_t* residues, const uint64_t* offsets, uint64_t N,
                int32_t min_hits, int32_t* out_role, int32_t* out_hits, uint8_t* out_flag) {
    if (!e) return KA_ERR_INVALID;
    BatchIn in;
    in.residues = residues; in.off64 = offsets;
    return annotate_impl(e, in, N, min_hits, out_role, out_hits, out_flag, "ka_annotate");
}

int ka_annotate_packed(ka_engine* e, const uint8_t* codes, const uint32_t* offsets, uint64_t N,
                       int32_t min_hits, int32_t* out_role, int32_t* out_hits, uint8_t* out_flag) {
    if (!e) return KA_ERR_INVALID;
    BatchIn in;
    in.codes = codes; in.off32 = offsets;
    return annotate_impl(e, in, N, min_hits, out_role, out_hits, out_flag, "ka_annotate_packed");
}

int ka_db_get_alphabet(ka_engine* e, uint8_t* code_of_byte) {
    if (!e || !code_of_byte) return KA_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    if (!e->have_db) return fail(e, KA_ERR_NO_DB, "ka_db_get_alphabet: no k-mer database loaded");
    memcpy(code_of_byte, e->lut5, 256);
    return KA_OK;
}

// 8 residues -> 40 bits -> 5 bytes; the tail (n % 8 residues) is written bit by bit into zeroed bytes
int ka_pack_residues(ka_engine* e, const uint8_t* residues, uint64_t n, uint64_t first_index, uint8_t* codes) {
    if (!e || (n && (!residues || !codes))) return KA_ERR_INVALID;
    uint8_t lut[256];
    {
        std::lock_guard<std::mutex> lk(e->mu);
        if (!e->have_db) return fail(e, KA_ERR_NO_DB, "ka_pack_residues: no k-mer database loaded");
        if (first_index % 8) return fail(e, KA_ERR_INVALID, "ka_pack_residues: first_index must be a multiple of 8 (byte boundary of the stream)");
        memcpy(lut, e->lut5, 256);
    }
    uint8_t* o = codes + first_index / 8 * 5;
    uint64_t i = 0;
    for (; i + 8 <= n; i += 8, o += 5) {
        const uint8_t* r = residues + i;
        const uint64_t v = (uint64_t)lut[r[0]] | (uint64_t)lut[r[1]] << 5 | (uint64_t)lut[r[2]] << 10 | (uint64_t)lut[r[3]] << 15 |
                           (uint64_t)lut[r[4]] << 20 | (uint64_t)lut[r[5]] << 25 | (uint64_t)lut[r[6]] << 30 | (uint64_t)lut[r[7]] << 35;
        o[0] = (uint8_t)v; o[1] = (uint8_t)(v >> 8); o[2] = (uint8_t)(v >> 16); o[3] = (uint8_t)(v >> 24); o[4] = (uint8_t)(v >> 32);
    }
    if (i < n) {
        uint64_t v = 0;
        const uint64_t rest = n - i;
        for (uint64_t k = 0; k < rest; k++) v |= (uint64_t)lut[residues[i + k]] << (5 * k);
        const uint64_t nbytes = (rest * 5 + 7) / 8;
        for (uint64_t k = 0; k < nbytes; k++) o[k] = (uint8_t)(v >> (8 * k));
    }
    return KA_OK;
}

int ka_batch_upload(ka_engine* e, int dev_index, const uint8_t* residues, const uint64_t* offsets,
                    uint64_t N, ka_batch** out) {
    if (!e || !out) return KA_ERR_INVALID;
    *out = nullptr;
    std::lock_guard<std::mutex> lk(e->mu);
    if (!e->have_db) return fail(e, KA_ERR_NO_DB, "ka_batch_upload: load the k-mer database first");
    if (dev_index < 0 || dev_index >= (int)e->devs.size()) return fail(e, KA_ERR_INVALID, "ka_batch_upload: bad device index");
    if (e->db_table_mode >= 2) return fail(e, KA_ERR_INVALID, "ka_batch_upload: resident batches are not available with the routed table (table_mode 2 / 3)");
    if (N == 0 || !offsets) return fail(e, KA_ERR_INVALID, "ka_batch_upload: empty batch");
    if (N > 0xfffffff0ull) return fail(e, KA_ERR_TOO_BIG, "ka_batch_upload: too many sequences");
    Device& d = e->devs[dev_index];
    cudaSetDevice(d.id);
    BatchIn in;
    in.residues = residues; in.off64 = offsets;
    ChunkShape sh;
    if (!scan_offsets(in, 0, N, e->long_seq, e->mid_seq, e->info.K, sh)) return fail(e, KA_ERR_OFFSETS, "ka_batch_upload: offsets are not monotone");
    if (e->line && sh.n_res > 0x7fffff00ull) return fail(e, KA_ERR_TOO_BIG, "ka_batch_upload: more than 2^31 residues in one resident batch");
    ka_batch* b = new ka_batch();
    b->dev_index = dev_index; b->n_seq = N; b->n_res = sh.n_res; b->base = offsets[0];
    b->long_res = sh.long_res; b->n_long = sh.n_long; b->n_mid = e->line ? sh.n_mid_seg : sh.n_mid;
    b->origin = offsets[0] & ~127ull;
    b->tile_span = e->tile_span; b->long_seq = e->long_seq; b->mid_seq = e->mid_seq; b->db_serial = e->db_serial;
    int rc = pipe_init(d, b->p);
    // resident form: the 5-bit stream wherever the tile kernels can stage it (line table; narrow unsharded sector tables)
    const bool keep_codes = e->line || (!e->geom.wide && e->geom.n_shards <= 1 && e->resident_packed);
    if (rc == KA_OK) rc = pipe_reserve(d, b->p, sh.n_res, N, tiles_of(e, d, sh.n_res), sh.n_long, sh.long_res, e->line ? sh.n_mid_seg : sh.n_mid, e->geom.wide != 0,
                                       true, keep_codes, e->line);
    cudaError_t ce = cudaSuccess;
    if (rc == KA_OK && sh.n_res) ce = cudaMemcpy(b->p.res, residues + offsets[0], sh.n_res, cudaMemcpyHostToDevice);
    if (rc == KA_OK && ce == cudaSuccess) ce = cudaMemcpy(b->p.off, offsets, (N + 1) * 8, cudaMemcpyHostToDevice);
    if (rc == KA_OK && ce == cudaSuccess && keep_codes) {
        // (what ka_annotate_packed ships; the long-sequence kernel of the sector tables still reads bytes)
        ce = launch_pack(b->p.res, (uint32_t)(offsets[0] - b->origin), sh.n_res, d.lut5, b->p.pk, b->p.st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(b->p.st);
        if (ce == cudaSuccess && (e->line || sh.n_long == 0)) { cudaFree(b->p.res); b->p.res = nullptr; b->p.res_cap = 0; }
    }
    if (rc || ce != cudaSuccess) {
        std::string m = rc ? d.errmsg : std::string(cudaGetErrorString(ce));
        pipe_free(b->p);
        delete b;
        return fail(e, rc ? rc : KA_ERR_CUDA, "ka_batch_upload: %s", m.c_str());
    }
    set_l2_window(e, d, b->p.st);
    e->stats = ka_stats{};
    e->stats.sequences = N; e->stats.residues = sh.n_res; e->stats.probes = sh.probes;
    *out = b;
    return KA_OK;
}

int ka_annotate_resident(ka_engine* e, ka_batch* b, int32_t min_hits) {
    if (!e || !b) return KA_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    if (!e->have_db) return fail(e, KA_ERR_NO_DB, "ka_annotate_resident: no k-mer database loaded");
    if (min_hits < 1) return fail(e, KA_ERR_INVALID, "ka_annotate_resident: min_hits must be positive");
    if (b->db_serial != e->db_serial)
        return fail(e, KA_ERR_INVALID, "ka_annotate_resident: the k-mer database was reloaded after this batch was uploaded; upload it again");
    if (b->tile_span != e->tile_span || b->long_seq != e->long_seq || b->mid_seq != e->mid_seq)
        return fail(e, KA_ERR_INVALID, "ka_annotate_resident: tile_span / long_seq / mid_seq changed after this batch was uploaded; upload it again");
    Device& d = e->devs[b->dev_index];
    cudaSetDevice(d.id);
    auto t0 = std::chrono::steady_clock::now();
    d.kernel_ms = d.tile_ms = 0; d.launches = 0;
    int rc;
    if (e->line) {
        LineParams lp;
        fill_line_params(e, d, b->p, b->n_res, b->n_seq, min_hits, lp);
        rc = enqueue_line_kernels(e, d, b->p, lp, true, b->origin, b->n_long, b->n_mid, true);
    } else {
        AnnotParams ap;
        fill_params(e, d, b->p, b->base, b->n_res, b->n_seq, min_hits, ap);
        if (b->p.pk) { ap.pk = b->p.pk; ap.pk_lead = (uint32_t)(b->base - b->origin); }
        rc = enqueue_kernels(e, d, b->p, ap, b->n_long, b->n_mid);
    }
    if (rc == KA_OK) {
        cudaError_t ce = cudaStreamSynchronize(b->p.st);
        if (ce != cudaSuccess) rc = dev_fail(d, KA_ERR_CUDA, "annotate kernels", ce);
    }
    if (rc == KA_OK) rc = collect_times(d, b->p);
    if (rc) { e->err = d.errmsg; return rc; }
    e->stats.kernel_launches = d.launches;
    e->stats.kernel_ms = d.kernel_ms;
    e->stats.tile_kernel_ms = d.tile_ms;
    e->stats.h2d_bytes = e->stats.d2h_bytes = 0;
    e->stats.wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return KA_OK;
}

int ka_batch_download(ka_engine* e, ka_batch* b, int32_t* out_role, int32_t* out_hits, uint8_t* out_flag) {
    if (!e || !b) return KA_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    Device& d = e->devs[b->dev_index];
    cudaSetDevice(d.id);
    cudaError_t ce = cudaSuccess;
    if (out_role) ce = cudaMemcpy(out_role, b->p.role, b->n_seq * 4, cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess && out_hits) ce = cudaMemcpy(out_hits, b->p.hits, b->n_seq * 4, cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess && out_flag) ce = cudaMemcpy(out_flag, b->p.flag, b->n_seq, cudaMemcpyDeviceToHost);
    if (ce != cudaSuccess) return fail(e, KA_ERR_CUDA, "ka_batch_download: %s", cudaGetErrorString(ce));
    return KA_OK;
}

void ka_batch_free(ka_engine* e, ka_batch* b) {
    if (!e || !b) return;
    std::lock_guard<std::mutex> lk(e->mu);
    cudaSetDevice(e->devs[b->dev_index].id);
    pipe_free(b->p);
    delete b;
}

void* ka_host_alloc(size_t bytes) {
    void* p = nullptr;
    // KA_PINNED_WC=1: write-combined pinned memory (no CPU cache snooping on the DMA reads; the
    // host must then only WRITE these buffers) — an experiment knob for multi-GPU ingest
    const char* wc = getenv("KA_PINNED_WC");
    unsigned flags = cudaHostAllocPortable | ((wc && wc[0] == '1') ? cudaHostAllocWriteCombined : 0u);
    if (cudaHostAlloc(&p, bytes ? bytes : 1, flags) != cudaSuccess) return nullptr;
    return p;
}

void ka_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int ka_get_stats(ka_engine* e, ka_stats* out) {
    if (!e || !out) return KA_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    *out = e->stats;
    return KA_OK;
}

int ka_probe_roofline(ka_engine* e, int dev_index, uint64_t table_bytes, uint64_t n_probes,
                      int slot_bytes, int reps, double* probes_per_s) {
    if (!e || !probes_per_s) return KA_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    if (dev_index < 0 || dev_index >= (int)e->devs.size()) return fail(e, KA_ERR_INVALID, "bad device index");
    if (slot_bytes != 16 && slot_bytes != 32) return fail(e, KA_ERR_INVALID, "slot_bytes must be 16 or 32");
    if (table_bytes < 4096 || n_probes == 0 || reps < 1) return fail(e, KA_ERR_INVALID, "bad roofline arguments");
    Device& d = e->devs[dev_index];
    cudaSetDevice(d.id);
    uint4* buf = nullptr; unsigned long long* sink = nullptr;
    uint64_t n16 = table_bytes / 16;
    if (cudaMalloc((void**)&buf, n16 * 16) != cudaSuccess) return fail(e, KA_ERR_OOM, "roofline buffer allocation failed");
    if (cudaMalloc((void**)&sink, 8) != cudaSuccess) { cudaFree(buf); return fail(e, KA_ERR_OOM, "roofline sink allocation failed"); }
    cudaStream_t st = d.pipe[0].st;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaError_t ce = cudaMemsetAsync(sink, 0, 8, st);
    if (ce == cudaSuccess) ce = launch_fill_random(buf, n16, st);
    uint64_t n_slots = slot_bytes == 32 ? n16 / 2 : n16;
    float best = 1e30f;
    for (int r = 0; r < reps + 1 && ce == cudaSuccess; r++) {  // first launch is warm-up
        cudaEventRecord(a, st);
        ce = launch_random_probe(buf, n_slots, slot_bytes, n_probes, 0x5151ull * (r + 1), sink, st);
        cudaEventRecord(b, st);
        if (ce == cudaSuccess) ce = cudaEventSynchronize(b);
        float ms = 0;
        if (ce == cudaSuccess) cudaEventElapsedTime(&ms, a, b);
        if (r > 0 && ms > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(buf); cudaFree(sink);
    if (ce != cudaSuccess) return fail(e, KA_ERR_CUDA, "roofline kernel: %s", cudaGetErrorString(ce));
    *probes_per_s = (double)n_probes / ((double)best * 1e-3);
    return KA_OK;
}

}  // extern "C"
