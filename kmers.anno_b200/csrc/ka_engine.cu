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
    DCK(d, cudaMalloc((void**)&p.ctr, 32));
    DCK(d, cudaMemset(p.ctr, 0, 32));
    return KA_OK;
}

void pipe_free(Pipe& p) {
    if (p.res) cudaFree(p.res);
    if (p.off) cudaFree(p.off);
    if (p.first) cudaFree(p.first);
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
    p = Pipe();
}

// size the per-chunk device buffers
int pipe_reserve(Device& d, Pipe& p, uint64_t n_res, uint64_t n_seq, uint64_t n_tiles,
                 uint64_t n_long, uint64_t long_res, uint64_t n_mid, bool wide) {
    int rc;
    if ((rc = ensure(d, p.res, p.res_cap, n_res + 64, "residues"))) return rc;
    size_t want_seq = n_seq + 1;
    if (want_seq > p.seq_cap || !p.off) {
        if (p.off) cudaFree(p.off);
        if (p.role) cudaFree(p.role);
        if (p.hits) cudaFree(p.hits);
        if (p.flag) cudaFree(p.flag);
        p.off = nullptr; p.role = p.hits = nullptr; p.flag = nullptr; p.seq_cap = 0;
        size_t n = want_seq + want_seq / 8 + 64;
        cudaError_t ce;
        if ((ce = cudaMalloc((void**)&p.off, n * 8)) != cudaSuccess ||
            (ce = cudaMalloc((void**)&p.role, n * 4)) != cudaSuccess ||
            (ce = cudaMalloc((void**)&p.hits, n * 4)) != cudaSuccess ||
            (ce = cudaMalloc((void**)&p.flag, n)) != cudaSuccess)
            return dev_fail(d, KA_ERR_OOM, "sequence buffers", ce);
        p.seq_cap = n;
    }
    if ((rc = ensure(d, p.first, p.first_cap, n_tiles + 2, "tile index"))) return rc;
    if ((rc = ensure(d, p.big, p.big_cap, n_long + 1, "long-sequence list"))) return rc;
    if ((rc = ensure(d, p.mid, p.mid_cap, n_mid + 1, "mid-sequence tiles"))) return rc;
    if ((rc = ensure(d, p.scratch, p.scratch_cap, (wide ? 2 : 1) * (2 * long_res + 4), "long-sequence tokens")))
        return rc;
    return KA_OK;
}

// validate offsets of [cs, ce) and collect shape numbers; false = offsets not monotone
bool scan_offsets(const uint64_t* off, uint64_t cs, uint64_t ce, uint32_t long_seq, uint32_t mid_seq, int K,
                  ChunkShape& s) {
    s = ChunkShape();
    for (uint64_t i = cs; i < ce; i++) {
        if (off[i + 1] < off[i]) return false;
        uint64_t L = off[i + 1] - off[i];
        if (L > mid_seq) { s.n_long++; s.long_res += L; }
        else if (L > long_seq) s.n_mid++;
        if (L >= (uint64_t)K) s.probes += L - K + 1;
    }
    s.n_res = off[ce] - off[cs];
    return true;
}

void fill_params(ka_engine* e, Device& d, Pipe& p, uint64_t base, uint64_t n_res, uint64_t n_seq,
                 int32_t min_hits, AnnotParams& ap) {
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
    ap.tab.sig = d.sig;
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
    for (int cls : {32, 64, 128}) DCK(d, tile_kernel_filt_set_smem(cls, (size_t)optin - 2048));
    DCK(d, tile_kernel_mode_set_smem((size_t)optin - 2048));
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
    if (ap.tab.sig && e->two_phase && !wide) {
        size_t smem_f = tile_smem_bytes_filt(ap.ext_max, nullptr);
        if (smem_f > d.smem_set) return dev_fail(d, KA_ERR_INVALID, "tile shared memory exceeds the device limit", cudaErrorInvalidValue);
        DCK(d, launch_tiles_filt(ap, smem_f, p.st));
    } else {
        DCK(d, launch_tiles(ap, e->variant, smem, p.st));
    }
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
            cudaError_t ce = launch_tiles(am, e->mid_variant, smem_mid, p.st);
            if (ce != cudaSuccess) {
                char buf[256];
                snprintf(buf, sizeof buf, "mid tile launch (tiles %u, smem %zu, variant %d, cls %d)", am.n_tiles, smem_mid, e->mid_variant, am.tab.cls);
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
    if (!e->l2_persist || !d.table) return;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, d.id) != cudaSuccess) return;
    if (prop.persistingL2CacheMaxSize <= 0 || prop.accessPolicyMaxWindowSize <= 0) return;
    size_t persist = (size_t)prop.persistingL2CacheMaxSize;
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist);
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof v);
    // with the presence filter the signatures are the hot, reusable data: pin them; otherwise
    // pin as much of the table as the persisting carve-out holds
    const bool sig = d.sig != nullptr;
    size_t span = sig ? ((size_t)2 << e->geom.bbits) : (size_t)e->info.table_bytes;
    size_t win = std::min<size_t>(span, (size_t)prop.accessPolicyMaxWindowSize);
    v.accessPolicyWindow.base_ptr = sig ? (void*)d.sig : (void*)d.table;
    v.accessPolicyWindow.num_bytes = win;
    v.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)persist / (double)win);
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v);
    cudaGetLastError();  // the window is an optimisation; never fail the call on it
}

// Annotate sequences [s_begin, s_end) of the host batch on device d.
int annotate_range(ka_engine* e, Device& d, const uint8_t* residues, const uint64_t* offsets,
                   uint64_t s_begin, uint64_t s_end, int32_t min_hits, int32_t* out_role,
                   int32_t* out_hits, uint8_t* out_flag) {
    DCK(d, cudaSetDevice(d.id));
    d.kernel_ms = d.tile_ms = 0; d.launches = 0; d.h2d = d.d2h = 0; d.probes = 0;
    int slot = 0;
    uint64_t cs = s_begin;
    while (cs < s_end) {
        // chunk = as many whole sequences as fit in chunk_residues (at least one)
        uint64_t lim = offsets[cs] + e->chunk_residues;
        uint64_t ce = std::upper_bound(offsets + cs + 1, offsets + s_end + 1, lim) - offsets - 1;
        if (ce <= cs) ce = cs + 1;
        if (ce - cs > 0xfffffff0ull) ce = cs + 0xfffffff0ull;
        ChunkShape sh;
        if (!scan_offsets(offsets, cs, ce, e->long_seq, e->mid_seq, e->info.K, sh)) {
            d.err = KA_ERR_OFFSETS; d.errmsg = "offsets are not monotone";
            return d.err;
        }
        if (sh.n_res > 0x7fffffffull) {
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
        uint64_t n = ce - cs;
        uint64_t n_tiles = sh.n_res / e->tile_span + 1;
        int rc = pipe_reserve(d, p, sh.n_res, n, n_tiles, sh.n_long, sh.long_res, sh.n_mid, e->geom.wide != 0);
        if (rc) return rc;
        if (sh.n_res)
            DCK(d, cudaMemcpyAsync(p.res, residues + offsets[cs], sh.n_res, cudaMemcpyHostToDevice, p.st));
        DCK(d, cudaMemcpyAsync(p.off, offsets + cs, (n + 1) * 8, cudaMemcpyHostToDevice, p.st));
        d.h2d += sh.n_res + (n + 1) * 8;
        AnnotParams ap;
        fill_params(e, d, p, offsets[cs], sh.n_res, n, min_hits, ap);
        if (!out_flag) ap.out_flag = p.flag;  // kernel always writes flags; host may skip them
        rc = enqueue_kernels(e, d, p, ap, sh.n_long, sh.n_mid);
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
        if (rc == KA_OK && cudaMalloc((void**)&d.lut, 256) != cudaSuccess) rc = KA_ERR_OOM;
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
        if (d.sig) cudaFree(d.sig);
        route_destroy_comm(d);
        for (auto& ln : d.lane) {
            for (void* q : {(void*)ln.r_keys, (void*)ln.r_send, (void*)ln.r_recv, (void*)ln.r_ans_recv, (void*)ln.r_ans_sorted,
                            (void*)ln.r_small, (void*)ln.r_pos}) if (q) cudaFree(q);
            if (ln.h_cnt) cudaFreeHost(ln.h_cnt);
            if (ln.ev_counts) cudaEventDestroy(ln.ev_counts);
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
    std::string n(name);
    if (n == "load_factor") {
        if (!(v > 0.0 && v <= 0.9)) return fail(e, KA_ERR_INVALID, "load_factor must be in (0, 0.9]");
        e->load_factor = v;
    } else if (n == "tile_span") {
        if (v < 256 || v > 65536) return fail(e, KA_ERR_INVALID, "tile_span must be in [256, 65536]");
        e->tile_span = (uint32_t)v & ~15u;
        if (e->long_seq < e->tile_span) e->long_seq = e->tile_span;
    } else if (n == "long_seq") {
        if (v < 256 || v > (1 << 20)) return fail(e, KA_ERR_INVALID, "long_seq must be in [256, 2^20]");
        e->long_seq = (uint32_t)v;
        if (e->long_seq < e->tile_span) e->long_seq = e->tile_span;
    } else if (n == "mid_seq") {
        if (v < 256 || v > 49152) return fail(e, KA_ERR_INVALID, "mid_seq must be in [256, 49152]");
        e->mid_seq = (uint32_t)v;
    } else if (n == "mid_variant") {
        if (v < 0 || v >= N_VARIANTS) return fail(e, KA_ERR_INVALID, "mid_variant must be 0..%d", N_VARIANTS - 1);
        e->mid_variant = (int)v;
    } else if (n == "chunk_residues") {
        if (v < 4096 || v > (double)(1ull << 30)) return fail(e, KA_ERR_INVALID, "chunk_residues must be in [4096, 2^30]");
        e->chunk_residues = (uint64_t)v;
    } else if (n == "l2_persist") {
        e->l2_persist = v != 0;
    } else if (n == "two_phase") {
        e->two_phase = v != 0;
    } else if (n == "table_mode") {
        if (v != 0 && v != 1 && v != 2) return fail(e, KA_ERR_INVALID, "table_mode must be 0 (replicated), 1 (sharded, peer loads) or 2 (sharded, routed)");
        e->table_mode = (int)v;
    } else if (n == "wide") {
        e->wide = v != 0;
    } else if (n == "filter") {
        e->filter = v < 0 ? -1 : (v != 0);
    } else if (n == "slot_bits") {
        if (v != 0 && v != 32 && v != 64 && v != 128) return fail(e, KA_ERR_INVALID, "slot_bits must be 0, 32, 64 or 128");
        e->slot_bits = (int)v;
    } else if (n == "variant") {
        if (v < 0 || v >= N_VARIANTS) return fail(e, KA_ERR_INVALID, "variant must be 0..%d", N_VARIANTS - 1);
        e->variant = (int)v;
    } else {
        return fail(e, KA_ERR_INVALID, "unknown option '%s'", name);
    }
    if (tile_smem_bytes(e->tile_span + e->long_seq, nullptr, e->wide != 0) > 225 * 1024 ||
        tile_smem_bytes(std::max(e->mid_seq, e->long_seq), nullptr, e->wide != 0) > 225 * 1024)
        return fail(e, KA_ERR_INVALID, "tile_span + long_seq (or mid_seq) needs more than 227 KB of shared memory");
    return KA_OK;
}

int ka_db_get_info(ka_engine* e, ka_db_info* out) {
    if (!e || !out) return KA_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    if (!e->have_db) return fail(e, KA_ERR_NO_DB, "no k-mer database loaded");
    *out = e->info;
    return KA_OK;
}

int ka_annotate(ka_engine* e, const uint8_t* residues, const uint64_t* offsets, uint64_t N,
                int32_t min_hits, int32_t* out_role, int32_t* out_hits, uint8_t* out_flag) {
    if (!e) return KA_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    if (!e->have_db) return fail(e, KA_ERR_NO_DB, "ka_annotate: no k-mer database loaded");
    if (min_hits < 1) return fail(e, KA_ERR_INVALID, "ka_annotate: min_hits must be positive");  // ApplyKmerProcessor.java:91-92
    if (N && (!offsets || !out_role || !out_hits)) return fail(e, KA_ERR_INVALID, "ka_annotate: NULL argument");
    if (N && offsets[N] > offsets[0] && !residues) return fail(e, KA_ERR_INVALID, "ka_annotate: residues is NULL");
    auto t0 = std::chrono::steady_clock::now();
    e->stats = ka_stats{};
    if (N == 0) return KA_OK;
    if (offsets[N] < offsets[0]) return fail(e, KA_ERR_OFFSETS, "ka_annotate: offsets are not monotone");

    // residue-balanced contiguous ranges, one per device
    size_t nd = e->devs.size();
    std::vector<uint64_t> cut(nd + 1, 0);
    cut[nd] = N;
    uint64_t total = offsets[N] - offsets[0];
    for (size_t i = 1; i < nd; i++) {
        uint64_t target = offsets[0] + total / nd * i;
        uint64_t c = std::lower_bound(offsets, offsets + N + 1, target) - offsets;
        cut[i] = std::min<uint64_t>(std::max<uint64_t>(c, cut[i - 1]), N);
    }
    int rc;
    if (e->table_mode == 2) {
        // routed sharded table: every device must walk every round, even with an empty range
        RouteShared shared((int)nd);
        rc = for_each_device(e, [&](Device& d, int i) {
            return annotate_routed_range(e, d, i, shared, residues, offsets, cut[i], cut[i + 1], min_hits, out_role, out_hits, out_flag);
        });
    } else {
        rc = for_each_device(e, [&](Device& d, int i) {
            if (cut[i] == cut[i + 1]) { d.kernel_ms = d.tile_ms = 0; d.launches = d.h2d = d.d2h = d.probes = 0; return (int)KA_OK; }
            return annotate_range(e, d, residues, offsets, cut[i], cut[i + 1], min_hits, out_role, out_hits, out_flag);
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

int ka_batch_upload(ka_engine* e, int dev_index, const uint8_t* residues, const uint64_t* offsets,
                    uint64_t N, ka_batch** out) {
    if (!e || !out) return KA_ERR_INVALID;
    *out = nullptr;
    std::lock_guard<std::mutex> lk(e->mu);
    if (!e->have_db) return fail(e, KA_ERR_NO_DB, "ka_batch_upload: load the k-mer database first");
    if (dev_index < 0 || dev_index >= (int)e->devs.size()) return fail(e, KA_ERR_INVALID, "ka_batch_upload: bad device index");
    if (e->table_mode == 2) return fail(e, KA_ERR_INVALID, "ka_batch_upload: resident batches are not available with the routed table (table_mode 2)");
    if (N == 0 || !offsets) return fail(e, KA_ERR_INVALID, "ka_batch_upload: empty batch");
    if (N > 0xfffffff0ull) return fail(e, KA_ERR_TOO_BIG, "ka_batch_upload: too many sequences");
    Device& d = e->devs[dev_index];
    cudaSetDevice(d.id);
    ChunkShape sh;
    if (!scan_offsets(offsets, 0, N, e->long_seq, e->mid_seq, e->info.K, sh)) return fail(e, KA_ERR_OFFSETS, "ka_batch_upload: offsets are not monotone");
    ka_batch* b = new ka_batch();
    b->dev_index = dev_index; b->n_seq = N; b->n_res = sh.n_res; b->base = offsets[0];
    b->long_res = sh.long_res; b->n_long = sh.n_long; b->n_mid = sh.n_mid;
    int rc = pipe_init(d, b->p);
    if (rc == KA_OK) rc = pipe_reserve(d, b->p, sh.n_res, N, sh.n_res / e->tile_span + 1, sh.n_long, sh.long_res, sh.n_mid, e->geom.wide != 0);
    cudaError_t ce = cudaSuccess;
    if (rc == KA_OK && sh.n_res) ce = cudaMemcpy(b->p.res, residues + offsets[0], sh.n_res, cudaMemcpyHostToDevice);
    if (rc == KA_OK && ce == cudaSuccess) ce = cudaMemcpy(b->p.off, offsets, (N + 1) * 8, cudaMemcpyHostToDevice);
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
    Device& d = e->devs[b->dev_index];
    cudaSetDevice(d.id);
    auto t0 = std::chrono::steady_clock::now();
    d.kernel_ms = d.tile_ms = 0; d.launches = 0;
    AnnotParams ap;
    fill_params(e, d, b->p, b->base, b->n_res, b->n_seq, min_hits, ap);
    int rc = enqueue_kernels(e, d, b->p, ap, b->n_long, b->n_mid);
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
