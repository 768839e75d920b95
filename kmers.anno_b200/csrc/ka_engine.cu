// ka_engine.cu — C ABI (include/kmeranno.h) over the sm_100a kernels: device table build,
// multi-device sharding, pipelined H2D -> plan -> tile -> big -> D2H chunks.
//
// Replaces /root/reference/src/main/java/org/theseed/proteins/kmers/anno/
// ApplyKmerProcessor.java:99-110 (DB load) and :122-148 (peg loop).  No CPU fallback: every
// path below either runs the CUDA kernels or returns an error code.
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <dlfcn.h>
#include <nccl.h>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/kmeranno.h"
#include "ka_kernels.cuh"

using namespace ka;

namespace {

constexpr int NPIPE = 4;  // chunks in flight per device

thread_local std::string g_create_error;

struct Pipe {
    cudaStream_t st = nullptr;
    uint8_t* res = nullptr; size_t res_cap = 0;
    unsigned long long* off = nullptr; size_t seq_cap = 0;
    uint4* first = nullptr; size_t first_cap = 0;
    int32_t* role = nullptr; int32_t* hits = nullptr; uint8_t* flag = nullptr;
    uint32_t* ctr = nullptr;  // 16 bytes: [0] big_count, [2..3] token cursor (u64)
    BigItem* big = nullptr; size_t big_cap = 0;
    uint4* mid = nullptr; size_t mid_cap = 0;
    uint32_t* scratch = nullptr; size_t scratch_cap = 0;
    cudaEvent_t ev_k0 = nullptr, ev_t0 = nullptr, ev_t1 = nullptr, ev_k1 = nullptr, done = nullptr;
    bool busy = false;
};

struct Device {
    int id = 0;
    int sm_count = 148;
    uint4* table = nullptr;
    uint4* ovf = nullptr;     // overflow table (cls 32/64)
    uint16_t* sig = nullptr;  // per-sector presence signatures
    const uint4** shard_sectors = nullptr;  // sharded mode: device array of peer pointers, one per shard
    const uint4** shard_ovf = nullptr;
    // routed mode (table_mode 2)
    ncclComm_t comm = nullptr;
    struct RouteLane {   // buffers of one round in flight (two lanes alternate, see annotate_routed_range)
        unsigned long long *r_keys = nullptr, *r_send = nullptr, *r_recv = nullptr, *r_ans_recv = nullptr,
                           *r_ans_sorted = nullptr, *r_small = nullptr;  // r_small: 8 counts, 8 offsets, 8 cursors
        uint32_t* r_pos = nullptr;             // send slot of every residue position
        size_t r_cap_pos = 0, r_cap_recv = 0;
        unsigned long long* h_cnt = nullptr;   // pinned: per-owner key counts of the round
        cudaEvent_t ev_counts = nullptr;
    } lane[2];
    cudaEvent_t ev_route0 = nullptr, ev_route1 = nullptr;
    uint8_t* lut = nullptr;
    Pipe pipe[NPIPE];
    size_t smem_set = 0;   // opt-in shared-memory limit once the tile kernels are configured
    bool route_smem_set = false;
    // per-call accounting
    double kernel_ms = 0, tile_ms = 0;
    uint64_t launches = 0, h2d = 0, d2h = 0, probes = 0;
    int err = KA_OK;
    std::string errmsg;
};

}  // namespace

struct ka_batch {
    int dev_index = 0;
    uint64_t n_seq = 0, n_res = 0, base = 0, long_res = 0, n_long = 0, n_mid = 0;
    Pipe p;  // owns device buffers of the resident batch
};

struct ka_engine {
    std::vector<Device> devs;
    std::mutex mu;
    std::string err;
    // options
    double load_factor = 0.4;
    uint32_t tile_span = 1536;
    uint32_t long_seq = 2048;
    uint32_t mid_seq = 8192;
    int mid_variant = 1;
    uint64_t chunk_residues = 48ull << 20;
    int l2_persist = 1;
    int variant = 0;
    int slot_bits = 0;  // 0 = choose automatically
    int filter = 0;     // 1 = per-sector presence signatures in front of the table (measured slower
                        // in the fused kernel: 40 vs 46 G probes/s, profiles/r01_summary.md), -1 = auto
    bool have_sig = false;
    int two_phase = 0;  // with signatures: 1 = two-phase tile kernel (measured slower, kept as an experiment),
                        // 0 = signature test inside the fused kernel
    int table_mode = 0; // 0 = table replicated on every device, 1 = sharded by sector range (peer loads)
    int wide = 0;       // 1 = force the wide-table kernels (64-bit sector indices and tokens) on any table
    bool peers_enabled = false;
    bool nccl_ready = false;
    // db
    bool have_db = false;
    ka_db_info info{};
    TableView geom{};   // geometry of the loaded table (sectors pointer filled per device)
    uint8_t lut[256];
    ka_stats stats{};
};

namespace {

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

#define DCK(d, call)                                                         \
    do {                                                                     \
        cudaError_t _ce = (call);                                            \
        if (_ce != cudaSuccess) return dev_fail((d), KA_ERR_CUDA, #call, _ce); \
    } while (0)

template <typename T>
int ensure(Device& d, T*& ptr, size_t& cap, size_t want, const char* what) {
    if (want <= cap && ptr) return KA_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr; cap = 0;
    size_t n = want + want / 8 + 64;
    cudaError_t ce = cudaMalloc((void**)&ptr, n * sizeof(T));
    if (ce != cudaSuccess) { ptr = nullptr; return dev_fail(d, KA_ERR_OOM, what, ce); }
    cap = n;
    return KA_OK;
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

struct ChunkShape {
    uint64_t n_res = 0, n_long = 0, long_res = 0, probes = 0, n_mid = 0;
};

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

// enqueue plan + tile + big on the pipe's stream, bracketed by timing events
int enqueue_kernels(ka_engine* e, Device& d, Pipe& p, const AnnotParams& ap, uint64_t n_long, uint64_t n_mid) {
    const bool wide = ap.tab.wide != 0;
    size_t smem = tile_smem_bytes(ap.ext_max, nullptr, wide);
    if (!d.smem_set) {
        // every tile-kernel instantiation may use up to the opt-in shared-memory limit of the device
        int optin = 0;
        DCK(d, cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, d.id));
        for (int cls : {32, 64, 128})
            for (int v = 0; v < N_VARIANTS; v++) DCK(d, tile_kernel_set_smem(cls, v, (size_t)optin - 2048));  // minus the static part
        for (int cls : {32, 64, 128}) DCK(d, tile_kernel_filt_set_smem(cls, (size_t)optin - 2048));
        DCK(d, tile_kernel_mode_set_smem((size_t)optin - 2048));
        d.smem_set = (size_t)optin - 2048;
    }
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

// Source of the DB lines: host arrays, or the synthetic generator (kmers == NULL).
struct DbSource {
    const uint8_t* kmers = nullptr;
    const int32_t* roles = nullptr;
    uint64_t n = 0;
    uint64_t seed = 0;        // synthetic only
    uint32_t n_roles = 0;     // synthetic only
    uint32_t role_bits = 1;   // bits of the largest role id
    bool synthetic = false;
};

// Build the table replica (or shard) of one device from the DB lines.
int build_table(ka_engine* e, Device& d, const TableView& geom, const DbSource& src, uint64_t* n_keys, uint32_t* max_probe) {
    const uint8_t* kmers = src.kmers;
    const int32_t* roles = src.roles;
    const uint64_t n = src.n;
    DCK(d, cudaSetDevice(d.id));
    if (d.table) { cudaFree(d.table); d.table = nullptr; }
    if (d.ovf) { cudaFree(d.ovf); d.ovf = nullptr; }
    if (d.sig) { cudaFree(d.sig); d.sig = nullptr; }
    const int K = geom.K;
    const size_t n_sectors = (size_t)1 << (geom.n_shards > 1 ? geom.shard_shift : geom.bbits);  // of this device
    const size_t bytes = n_sectors * 32;
    const size_t n_slots = n_sectors * (geom.cls == 32 ? 8 : (geom.cls == 64 ? 4 : 2));
    cudaError_t ce = cudaMalloc((void**)&d.table, bytes);
    if (ce != cudaSuccess) { d.table = nullptr; return dev_fail(d, KA_ERR_OOM, "table", ce); }
    const size_t ovf_bytes = geom.cls == 128 ? 0 : ((size_t)64 << geom.ovf_bbits);
    if (ovf_bytes) {
        ce = cudaMalloc((void**)&d.ovf, ovf_bytes);
        if (ce != cudaSuccess) { d.ovf = nullptr; return dev_fail(d, KA_ERR_OOM, "overflow table", ce); }
    }
    const bool use_sig = geom.n_shards <= 1 && !geom.wide && (e->filter == 1 || (e->filter < 0 && geom.bbits >= 20));
    const size_t sig_bytes = use_sig ? (n_sectors * 2 + 4) : 0;
    if (sig_bytes) {
        ce = cudaMalloc((void**)&d.sig, sig_bytes);
        if (ce != cudaSuccess) { d.sig = nullptr; return dev_fail(d, KA_ERR_OOM, "signature array", ce); }
    }
    cudaStream_t st = d.pipe[0].st;
    TableView tab = geom;
    tab.sectors = d.table;
    tab.ovf = d.ovf;
    tab.sig = d.sig;
    const uint64_t CH = (src.synthetic ? 64ull : 16ull) << 20;  // k-mers per upload / per generator launch
    uint8_t* dk = nullptr; int32_t* dr = nullptr; unsigned long long* best = nullptr;
    unsigned long long* dc = nullptr; uint32_t* de = nullptr;
    uint64_t ch = std::min<uint64_t>(CH, n ? n : 1);
    const bool packed = geom.cls != 128;
    // cls 32/64: 8 bytes per primary slot hold the winning (line, role) until db_finalize
    if ((ce = cudaMalloc((void**)&dk, ch * K)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&dr, ch * 4)) != cudaSuccess ||
        (packed && (ce = cudaMalloc((void**)&best, n_slots * 8)) != cudaSuccess) ||
        (ce = cudaMalloc((void**)&dc, 16)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&de, 16)) != cudaSuccess) {
        if (dk) cudaFree(dk);
        if (dr) cudaFree(dr);
        if (best) cudaFree(best);
        if (dc) cudaFree(dc);
        return dev_fail(d, KA_ERR_OOM, "DB staging", ce);
    }
    int rc = KA_OK;
    auto step = [&](cudaError_t c, const char* what) {
        if (c != cudaSuccess && rc == KA_OK) rc = dev_fail(d, KA_ERR_CUDA, what, c);
    };
    step(cudaMemsetAsync(d.table, 0, bytes, st), "memset table");
    if (ovf_bytes) step(cudaMemsetAsync(d.ovf, 0, ovf_bytes, st), "memset overflow table");
    if (sig_bytes) step(cudaMemsetAsync(d.sig, 0, sig_bytes, st), "memset signatures");
    if (packed) step(cudaMemsetAsync(best, 0, n_slots * 8, st), "memset best");
    step(cudaMemcpyAsync(d.lut, e->lut, 256, cudaMemcpyHostToDevice, st), "H2D lut");
    step(cudaMemsetAsync(dc, 0, 16, st), "memset counters");
    step(cudaMemsetAsync(de, 0, 16, st), "memset errs");
    for (uint64_t i = 0; i < n && rc == KA_OK; i += ch) {
        uint64_t m = std::min(ch, n - i);
        if (!src.synthetic) {
            step(cudaMemcpyAsync(dk, kmers + i * K, m * K, cudaMemcpyHostToDevice, st), "H2D kmers");
            step(cudaMemcpyAsync(dr, roles + i, m * 4, cudaMemcpyHostToDevice, st), "H2D roles");
        } else {
            step(launch_db_generate(i, m, K, src.seed, src.n_roles, dk, dr, st), "db_generate");
        }
        step(launch_db_insert(tab, dk, dr, m, i, d.lut, best, src.role_bits, dc, de, st), "db_insert");
        step(cudaStreamSynchronize(st), "db_insert sync");
    }
    if (rc == KA_OK) step(launch_db_finalize(tab, best, src.role_bits, st), "db_finalize");
    step(cudaStreamSynchronize(st), "db_finalize sync");
    unsigned long long hc[2] = {0, 0};
    uint32_t he[4] = {0, 0, 0, 0};
    step(cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost), "D2H counters");
    step(cudaMemcpy(he, de, 16, cudaMemcpyDeviceToHost), "D2H errs");
    cudaFree(dk); cudaFree(dr); cudaFree(dc); cudaFree(de);
    if (best) cudaFree(best);
    if (rc) return rc;
    if (he[0]) { d.err = KA_ERR_ALPHABET; d.errmsg = "k-mer byte outside the DB alphabet (internal)"; return d.err; }
    if (he[1]) { d.err = KA_ERR_ROLE; d.errmsg = "negative role id in the DB"; return d.err; }
    if (he[2]) { d.err = KA_ERR_TOO_BIG; d.errmsg = "overflow table full"; return d.err; }  // caller retries larger
    *n_keys = hc[0];
    *max_probe = (uint32_t)hc[1];
    return KA_OK;
}

uint32_t ceil_log2(double x) {
    uint32_t b = 0;
    while ((double)(1ull << b) < x && b < 62) b++;
    return b;
}

// Pick slot class and sector count: the smallest table that holds n keys at the requested
// load factor with remainder + role fitting the slot (see ka_common.cuh).
bool choose_geometry(uint64_t n, int K, int32_t max_role, double lf, int force_cls, uint32_t n_shards, bool force_wide, TableView& g) {
    const uint32_t w = 5u * (uint32_t)K;
    uint32_t role_bits = 1;
    while (((uint64_t)max_role + 1) >> role_bits) role_bits++;
    bool found = false;
    uint64_t best_bytes = 0;
    for (int cls : {32, 64, 128}) {
        if (force_cls && cls != force_cls) continue;
        if (n_shards > 1 && cls == 128) continue;   // chaining across shards is not supported: quotiented classes only
        const int S = 256 / cls;
        uint32_t b = ceil_log2((double)(n ? n : 1) / ((double)S * lf));
        if (b < 6) b = 6;
        uint32_t shard_log = 0;
        while ((1u << shard_log) < n_shards) shard_log++;
        if (b < 6 + shard_log) b = 6 + shard_log;
        uint32_t rem_bits = 0;
        if (cls != 128) {
            if ((int)(w + role_bits) - cls > (int)b) b = w + role_bits - (uint32_t)cls;
            if (b > w) b = w;
            if (b < shard_log) continue;
            rem_bits = w - b;
            if (rem_bits + role_bits > (uint32_t)cls) continue;
        }
        uint32_t slot_log = cls == 32 ? 3 : (cls == 64 ? 2 : 1);
        // narrow tables: slot index + 1 must fit the 32-bit de-dup token; beyond that the quotiented
        // classes switch to the wide kernels (64-bit sector indices, the key is the token)
        bool wide = force_wide && cls != 128;
        if (b + slot_log > 31) {
            if (cls == 128 || b > 40) continue;
            wide = true;
        }
        uint64_t bytes = 32ull << b;
        if (!found || bytes < best_bytes) {
            found = true; best_bytes = bytes;
            g.cls = cls; g.bbits = b; g.rem_bits = rem_bits; g.wbits = w; g.K = K;
            g.key_mask = (1ull << w) - 1;
            g.rem_mask = rem_bits ? ((1ull << rem_bits) - 1) : 0;
            g.sectors = nullptr;
            g.ovf = nullptr;
            g.sig = nullptr;
            g.n_primary_slots = wide ? 0u : (uint32_t)((uint64_t)S << b);
            g.wide = wide ? 1u : 0u;
            // expected keys beyond S per sector under Poisson(n / sectors) arrivals
            double lam = (double)n / (double)(1ull << b), pk = std::exp(-lam), over = 0;
            for (int k = 1; k < S + 400; k++) {
                pk *= lam / k;
                if (k > S) over += (k - S) * pk;
            }
            double want = 4.0 * over * (double)(1ull << b) + 4096;
            g.ovf_bbits = cls == 128 ? 0 : ceil_log2(want / 2.0 / (n_shards ? n_shards : 1));
            g.n_shards = n_shards;
            g.shard_shift = b - shard_log;
            g.my_shard = 0;
            g.shard_sectors = nullptr;
            g.shard_ovf = nullptr;
        }
    }
    return found;
}

// ---- NCCL, loaded on demand: only the routed sharded table needs it -------------------------
struct NcclApi {
    void* h = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    bool load() {
        if (h) return true;
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!h) return false;
        GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
        CommInitAll = (decltype(CommInitAll))dlsym(h, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
        GroupStart = (decltype(GroupStart))dlsym(h, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(h, "ncclGroupEnd");
        Send = (decltype(Send))dlsym(h, "ncclSend");
        Recv = (decltype(Recv))dlsym(h, "ncclRecv");
        return GetErrorString && CommInitAll && CommDestroy && GroupStart && GroupEnd && Send && Recv;
    }
};
NcclApi g_nccl;

class Barrier {
public:
    explicit Barrier(int n) : n_(n) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m_);
        const int gen = gen_;
        if (++count_ == n_) { count_ = 0; gen_++; cv_.notify_all(); }
        else cv_.wait(lk, [&] { return gen != gen_; });
    }
private:
    std::mutex m_;
    std::condition_variable cv_;
    int n_, count_ = 0, gen_ = 0;
};

struct RouteShared {
    Barrier bar;
    std::vector<std::array<unsigned long long, 8>> counts;   // counts[d][o]: keys device d sends to owner o this round
    std::vector<size_t> n_chunks;
    std::atomic<int> abort{0};                                // a device could not size its receive buffers
    explicit RouteShared(int n) : bar(n), counts(n), n_chunks(n, 0) {}
};

int route_reserve(Device& dev, Device::RouteLane& d, size_t n_pos, size_t n_recv) {
    if (n_pos > d.r_cap_pos) {
        for (void* q : {(void*)d.r_keys, (void*)d.r_send, (void*)d.r_ans_sorted, (void*)d.r_pos}) if (q) cudaFree(q);
        d.r_keys = d.r_send = d.r_ans_sorted = nullptr; d.r_pos = nullptr; d.r_cap_pos = 0;
        size_t n = n_pos + n_pos / 8 + 1024;
        cudaError_t ce;
        if ((ce = cudaMalloc((void**)&d.r_keys, n * 8)) != cudaSuccess || (ce = cudaMalloc((void**)&d.r_send, n * 8)) != cudaSuccess ||
            (ce = cudaMalloc((void**)&d.r_ans_sorted, n * 8)) != cudaSuccess ||
            (ce = cudaMalloc((void**)&d.r_pos, n * 4)) != cudaSuccess)
            return dev_fail(dev, KA_ERR_OOM, "routing buffers", ce);
        d.r_cap_pos = n;
    }
    if (n_recv > d.r_cap_recv) {
        if (d.r_recv) cudaFree(d.r_recv);
        if (d.r_ans_recv) cudaFree(d.r_ans_recv);
        d.r_recv = d.r_ans_recv = nullptr; d.r_cap_recv = 0;
        size_t n = n_recv + n_recv / 8 + 1024;
        cudaError_t ce;
        if ((ce = cudaMalloc((void**)&d.r_recv, n * 8)) != cudaSuccess || (ce = cudaMalloc((void**)&d.r_ans_recv, n * 8)) != cudaSuccess)
            return dev_fail(dev, KA_ERR_OOM, "routing receive buffers", ce);
        d.r_cap_recv = n;
    }
    if (!d.r_small && cudaMalloc((void**)&d.r_small, 24 * 8) != cudaSuccess) return dev_fail(dev, KA_ERR_OOM, "routing counters", cudaErrorMemoryAllocation);
    if (!d.h_cnt && cudaHostAlloc((void**)&d.h_cnt, 64, cudaHostAllocDefault) != cudaSuccess) return dev_fail(dev, KA_ERR_OOM, "routing counters (pinned)", cudaErrorMemoryAllocation);
    if (!d.ev_counts && cudaEventCreateWithFlags(&d.ev_counts, cudaEventDisableTiming) != cudaSuccess) return dev_fail(dev, KA_ERR_CUDA, "routing event", cudaErrorUnknown);
    return KA_OK;
}

#define NCK(d, call)                                                                    \
    do {                                                                                \
        ncclResult_t _nr = (call);                                                      \
        if (_nr != ncclSuccess && (d).err == KA_OK) {                                   \
            (d).err = KA_ERR_CUDA; (d).errmsg = std::string("NCCL: ") + g_nccl.GetErrorString(_nr); \
        }                                                                               \
    } while (0)

// Routed sharded table: every device extracts the keys of its own sequences, the keys travel to
// the GPU that owns their table sector (NCCL send/recv all-to-all over NVLink), the owner probes
// its local shard, the answers travel back in request order and the requester tallies.
//
// Rounds are software-pipelined over TWO lanes (stream + buffers each): the extraction of round
// r+2 is queued behind round r on its lane, and the host issues round r+1's exchange while round
// r's kernels still run, so the NVLink phases of one round overlap the HBM-bound kernels
// (bucket scatter, owner lookup, un-permute, tally) of the other.  NCCL orders the operations of
// one communicator across the two streams itself; every device issues them in the same round order.
// All devices walk the same number of rounds (empty rounds send nothing) so that the
// point-to-point calls always match.  A failure on one device is remembered but the device keeps
// taking part with empty rounds: nobody is left waiting in a collective.
// KA_ROUTE_SERIAL=1 runs one lane with a sync per round (and KA_ROUTE_TRACE=1 then prints the phases).
int annotate_routed_range(ka_engine* e, Device& d, int idx, RouteShared& sh, const uint8_t* residues,
                          const uint64_t* offsets, uint64_t s_begin, uint64_t s_end, int32_t min_hits,
                          int32_t* out_role, int32_t* out_hits, uint8_t* out_flag) {
    const int nd = (int)e->devs.size();
    cudaSetDevice(d.id);
    d.kernel_ms = d.tile_ms = 0; d.launches = 0; d.h2d = d.d2h = 0; d.probes = 0;
    d.err = KA_OK; d.errmsg.clear();
    std::vector<std::pair<uint64_t, uint64_t>> chunks;
    for (uint64_t cs = s_begin; cs < s_end;) {
        uint64_t lim = offsets[cs] + e->chunk_residues;
        uint64_t ce = std::upper_bound(offsets + cs + 1, offsets + s_end + 1, lim) - offsets - 1;
        if (ce <= cs) ce = cs + 1;
        chunks.push_back({cs, ce});
        cs = ce;
    }
    sh.n_chunks[idx] = chunks.size();
    sh.bar.wait();
    size_t rounds = 0;
    for (size_t c : sh.n_chunks) rounds = std::max(rounds, c);
    auto cuda_ok = [&](cudaError_t ce, const char* what) {
        if (ce != cudaSuccess && d.err == KA_OK) dev_fail(d, KA_ERR_CUDA, what, ce);
        return ce == cudaSuccess;
    };
    if (!d.smem_set) {
        int optin = 0;
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, d.id);
        for (int cls : {32, 64, 128})
            for (int v = 0; v < N_VARIANTS; v++) cuda_ok(tile_kernel_set_smem(cls, v, (size_t)optin - 2048), "tile kernel shared memory");
        for (int cls : {32, 64, 128}) cuda_ok(tile_kernel_filt_set_smem(cls, (size_t)optin - 2048), "tile kernel shared memory");
        cuda_ok(tile_kernel_mode_set_smem((size_t)optin - 2048), "routed tile kernel shared memory");
        d.smem_set = (size_t)optin - 2048;
    }
    if (!d.ev_route0) { cuda_ok(cudaEventCreate(&d.ev_route0), "event"); cuda_ok(cudaEventCreate(&d.ev_route1), "event"); }

    const bool serial = getenv("KA_ROUTE_SERIAL") != nullptr;
    const bool trace = serial && idx == 0 && getenv("KA_ROUTE_TRACE") != nullptr;
    const int n_lanes = serial ? 1 : 2;
    std::vector<cudaEvent_t> tev(12, nullptr);
    if (trace) for (auto& ev : tev) cudaEventCreate(&ev);
    double tsum[11] = {0};

    // per-round state kept between the two halves of a round
    struct Round {
        bool has = false;
        uint64_t cs = 0, ce = 0, n = 0;
        ChunkShape shp;
        AnnotParams ap, am;
        size_t smem = 0, smem_mid = 0;
    };
    std::vector<Round> rs(n_lanes);

    // first half of round r on its lane: upload, plan, extract the keys, count them per owner
    auto issue_extract = [&](size_t r) {
        const int L = (int)(r % n_lanes);
        Round& R = rs[L];
        R = Round();
        Device::RouteLane& ln = d.lane[L];
        Pipe& p = d.pipe[L];
        cudaStream_t st = p.st;
        auto mark = [&](int k) { if (trace) cudaEventRecord(tev[k], st); };
        R.has = r < chunks.size() && d.err == KA_OK;
        if (route_reserve(d, ln, 64, 64) != KA_OK) R.has = false;     // counters, pinned counts, event
        if (R.has) {
            R.cs = chunks[r].first; R.ce = chunks[r].second; R.n = R.ce - R.cs;
            if (!scan_offsets(offsets, R.cs, R.ce, e->long_seq, e->mid_seq, e->info.K, R.shp)) { d.err = KA_ERR_OFFSETS; d.errmsg = "offsets are not monotone"; R.has = false; }
            else if (R.shp.n_long) { d.err = KA_ERR_TOO_BIG; d.errmsg = "routed table mode: a sequence is longer than mid_seq (raise the mid_seq option)"; R.has = false; }
            else if (R.shp.n_res > 0x7fffffffull) { d.err = KA_ERR_TOO_BIG; d.errmsg = "chunk exceeds 2^31 residues"; R.has = false; }
        }
        if (R.has) {
            d.probes += R.shp.probes;
            if (pipe_reserve(d, p, R.shp.n_res, R.n, R.shp.n_res / e->tile_span + 1, 0, 0, R.shp.n_mid, e->geom.wide != 0) ||
                route_reserve(d, ln, R.shp.n_res + 64, 0)) R.has = false;
        }
        if (ln.h_cnt) memset(ln.h_cnt, 0, 64);
        if (R.has) {
            if (R.shp.n_res) cuda_ok(cudaMemcpyAsync(p.res, residues + offsets[R.cs], R.shp.n_res, cudaMemcpyHostToDevice, st), "H2D residues");
            cuda_ok(cudaMemcpyAsync(p.off, offsets + R.cs, (R.n + 1) * 8, cudaMemcpyHostToDevice, st), "H2D offsets");
            d.h2d += R.shp.n_res + (R.n + 1) * 8;
            fill_params(e, d, p, offsets[R.cs], R.shp.n_res, R.n, min_hits, R.ap);
            R.ap.route_keys = ln.r_keys;
            R.ap.route_ans = ln.r_ans_sorted;
            R.ap.route_slot = ln.r_pos;
            R.smem = tile_smem_bytes(R.ap.ext_max, nullptr, R.ap.tab.wide != 0);
            R.am = R.ap;
            R.am.first = p.mid; R.am.n_tiles = (uint32_t)R.shp.n_mid; R.am.ext_max = R.ap.mid_seq;
            R.smem_mid = tile_smem_bytes(R.am.ext_max, &R.am.res_bytes, R.ap.tab.wide != 0);
            if (std::max(R.smem, R.smem_mid) > d.smem_set && d.err == KA_OK) { d.err = KA_ERR_INVALID; d.errmsg = "tile shared memory exceeds the device limit"; }
            cuda_ok(cudaMemsetAsync(p.ctr, 0, 16, st), "memset");
            cuda_ok(cudaMemsetAsync(ln.r_keys, 0xff, (R.shp.n_res + 64) * 8, st), "memset keys");
            cuda_ok(cudaMemsetAsync(ln.r_small, 0, 24 * 8, st), "memset counters");
            mark(0);
            cuda_ok(launch_plan(R.ap, st), "plan");
            cuda_ok(launch_tiles_mode(R.ap, 0, 1, R.smem, st), "extract tiles");
            if (R.shp.n_mid) cuda_ok(launch_tiles_mode(R.am, 1, 1, R.smem_mid, st), "extract mid tiles");
            mark(1);
            cuda_ok(launch_route_count(ln.r_keys, R.shp.n_res, R.ap.tab, ln.r_small, st), "route count");
            mark(2);
            cuda_ok(cudaMemcpyAsync(ln.h_cnt, ln.r_small, 64, cudaMemcpyDeviceToHost, st), "D2H counts");
            d.launches += 3 + (R.shp.n_mid ? 1 : 0);
        }
        if (ln.ev_counts) cuda_ok(cudaEventRecord(ln.ev_counts, st), "event");
    };

    cuda_ok(cudaEventRecord(d.ev_route0, d.pipe[0].st), "event");
    for (int L = 0; L < n_lanes && (size_t)L < rounds; L++) issue_extract((size_t)L);
    for (size_t r = 0; r < rounds; r++) {
        const int L = (int)(r % n_lanes);
        Round& R = rs[L];
        Device::RouteLane& ln = d.lane[L];
        Pipe& p = d.pipe[L];
        cudaStream_t st = p.st;
        auto mark = [&](int k) { if (trace) cudaEventRecord(tev[k], st); };
        std::array<unsigned long long, 8> cnt{};
        if (ln.ev_counts) cuda_ok(cudaEventSynchronize(ln.ev_counts), "extract sync");   // the counts of round r are on the host
        if (d.err != KA_OK) R.has = false;
        if (R.has && ln.h_cnt) for (int o = 0; o < 8; o++) cnt[o] = ln.h_cnt[o];
        sh.counts[idx] = cnt;
        sh.bar.wait();                     // every device's counts of this round are visible
        unsigned long long send_off[9] = {0}, recv_cnt[8] = {0}, recv_off[9] = {0};
        for (int o = 0; o < nd; o++) send_off[o + 1] = send_off[o] + sh.counts[idx][o];
        for (int o = 0; o < nd; o++) { recv_cnt[o] = sh.counts[o][idx]; recv_off[o + 1] = recv_off[o] + recv_cnt[o]; }
        const unsigned long long my_cnt[8] = {sh.counts[idx][0], sh.counts[idx][1], sh.counts[idx][2], sh.counts[idx][3],
                                              sh.counts[idx][4], sh.counts[idx][5], sh.counts[idx][6], sh.counts[idx][7]};
        const unsigned long long total_send = send_off[nd], total_recv = recv_off[nd];
        bool recv_ok = route_reserve(d, ln, 0, total_recv + 64) == KA_OK && ln.r_small;
        if (!recv_ok) { sh.abort.store(1); if (d.err == KA_OK) d.err = KA_ERR_OOM; }
        sh.bar.wait();                     // nobody overwrites counts before everyone has read them
        if (sh.abort.load()) {
            // a peer cannot receive: every device skips the exchange of this and all later rounds
            if (d.err == KA_OK) { d.err = KA_ERR_OOM; d.errmsg = "routed table mode: a peer device ran out of memory"; }
            recv_ok = false; R.has = false;
        }
        if (R.has) {
            cuda_ok(cudaMemcpyAsync(ln.r_small + 8, send_off, 64, cudaMemcpyHostToDevice, st), "H2D offsets");
            mark(3);
            cuda_ok(launch_route_scatter(ln.r_keys, R.shp.n_res, R.ap.tab, ln.r_small + 8, ln.r_small + 16, ln.r_send, ln.r_pos, st), "route scatter");
            mark(4);
            d.launches += 1;
        }
        if (recv_ok) {
            // keys to their owners
            NCK(d, g_nccl.GroupStart());
            for (int o = 0; o < nd; o++) {
                if (o == idx) continue;
                if (my_cnt[o]) NCK(d, g_nccl.Send(ln.r_send + send_off[o], my_cnt[o], ncclUint64, o, d.comm, st));
                if (recv_cnt[o]) NCK(d, g_nccl.Recv(ln.r_recv + recv_off[o], recv_cnt[o], ncclUint64, o, d.comm, st));
            }
            NCK(d, g_nccl.GroupEnd());
            if (recv_cnt[idx]) cuda_ok(cudaMemcpyAsync(ln.r_recv + recv_off[idx], ln.r_send + send_off[idx], recv_cnt[idx] * 8, cudaMemcpyDeviceToDevice, st), "self keys");
            // the owner answers from its local shard
            TableView tab = e->geom;
            tab.sectors = d.table; tab.ovf = d.ovf; tab.sig = nullptr; tab.my_shard = (uint32_t)idx;
            mark(5);
            cuda_ok(launch_route_lookup(ln.r_recv, total_recv, tab, ln.r_ans_recv, st), "route lookup");
            mark(6);
            d.launches += 1;
            // answers back to the requesters, in request order
            NCK(d, g_nccl.GroupStart());
            for (int o = 0; o < nd; o++) {
                if (o == idx) continue;
                if (recv_cnt[o]) NCK(d, g_nccl.Send(ln.r_ans_recv + recv_off[o], recv_cnt[o], ncclUint64, o, d.comm, st));
                if (my_cnt[o]) NCK(d, g_nccl.Recv(ln.r_ans_sorted + send_off[o], my_cnt[o], ncclUint64, o, d.comm, st));
            }
            NCK(d, g_nccl.GroupEnd());
            if (recv_cnt[idx]) cuda_ok(cudaMemcpyAsync(ln.r_ans_sorted + send_off[idx], ln.r_ans_recv + recv_off[idx], recv_cnt[idx] * 8, cudaMemcpyDeviceToDevice, st), "self answers");
        }
        if (R.has) {
            mark(7);
            mark(8);
            cuda_ok(launch_tiles_mode(R.ap, 0, 2, R.smem, st), "tally tiles");
            if (R.shp.n_mid) cuda_ok(launch_tiles_mode(R.am, 1, 2, R.smem_mid, st), "tally mid tiles");
            mark(9);
            d.launches += 1 + (R.shp.n_mid ? 1 : 0);
            cuda_ok(cudaMemcpyAsync(out_role + R.cs, p.role, R.n * 4, cudaMemcpyDeviceToHost, st), "D2H role");
            cuda_ok(cudaMemcpyAsync(out_hits + R.cs, p.hits, R.n * 4, cudaMemcpyDeviceToHost, st), "D2H hits");
            d.d2h += R.n * 8;
            if (out_flag) { cuda_ok(cudaMemcpyAsync(out_flag + R.cs, p.flag, R.n, cudaMemcpyDeviceToHost, st), "D2H flag"); d.d2h += R.n; }
        }
        if (serial) {
            cuda_ok(cudaStreamSynchronize(st), "round sync");
            if (trace && R.has) {
                const int pairs[9][2] = {{0, 1}, {1, 2}, {3, 4}, {4, 5}, {5, 6}, {6, 7}, {7, 8}, {8, 9}, {0, 9}};
                for (int k = 0; k < 9; k++) { float ms = 0; if (cudaEventElapsedTime(&ms, tev[pairs[k][0]], tev[pairs[k][1]]) == cudaSuccess) tsum[k] += ms; }
                cudaGetLastError();
            }
        }
        // the extraction of round r + n_lanes queues behind this round on the same lane
        if (r + n_lanes < rounds) issue_extract(r + n_lanes);
    }
    for (int L = 0; L < n_lanes; L++) cuda_ok(cudaStreamSynchronize(d.pipe[L].st), "round sync");
    // device-side span of the whole call (uploads, kernels and exchanges of all rounds)
    if (d.ev_route0 && d.ev_route1) {
        cudaStream_t last = d.pipe[0].st;
        cudaEventRecord(d.ev_route1, last);
        cudaEventSynchronize(d.ev_route1);
        float ms = 0;
        if (cudaEventElapsedTime(&ms, d.ev_route0, d.ev_route1) == cudaSuccess) d.kernel_ms = ms;
        cudaGetLastError();
    }
#ifdef KA_DEBUG
    for (int L = 0; L < n_lanes && d.err == KA_OK; L++) {
        uint32_t dbg = 0;
        if (d.pipe[L].ctr && cudaMemcpy(&dbg, d.pipe[L].ctr + 4, 4, cudaMemcpyDeviceToHost) == cudaSuccess && dbg) {
            char buf[96];
            snprintf(buf, sizeof buf, "KA_DEBUG bounds check failed in a kernel (codes 0x%x)", dbg);
            d.err = KA_ERR_CUDA; d.errmsg = buf;
            cudaMemset(d.pipe[L].ctr + 4, 0, 4);
        }
    }
#endif
    if (trace) {
        fprintf(stderr, "[route trace dev0, %zu rounds] extract %.2f count %.2f scatter %.2f exchange-keys %.2f lookup %.2f exchange-answers %.2f tally %.2f | first-to-last %.2f ms\n",
                rounds, tsum[0], tsum[1], tsum[2], tsum[3], tsum[4], tsum[5], tsum[7], tsum[8]);
        for (auto& ev : tev) cudaEventDestroy(ev);
    }
    return d.err;
}

template <typename F>
int for_each_device(ka_engine* e, F f) {
    if (e->devs.size() == 1) {
        int rc = f(e->devs[0], 0);
        if (rc) e->err = e->devs[0].errmsg;
        return rc;
    }
    std::vector<int> rcs(e->devs.size(), 0);
    std::vector<std::thread> th;
    for (size_t i = 0; i < e->devs.size(); i++)
        th.emplace_back([&, i] { rcs[i] = f(e->devs[i], (int)i); });
    for (auto& t : th) t.join();
    for (size_t i = 0; i < e->devs.size(); i++)
        if (rcs[i]) { e->err = e->devs[i].errmsg; return rcs[i]; }
    return KA_OK;
}

}  // namespace

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
        if (d.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(d.comm);
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

static int db_load_impl(ka_engine* e, const uint8_t* kmers, const int32_t* role_ids, uint64_t n, int K,
                        uint64_t syn_seed = 0, int32_t syn_roles = 0);

int ka_db_load(ka_engine* e, const uint8_t* kmers, const int32_t* role_ids, uint64_t n, int K) {
    if (!e) return KA_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    if (n && (!kmers || !role_ids)) return fail(e, KA_ERR_INVALID, "ka_db_load: NULL input");
    return db_load_impl(e, kmers, role_ids, n, K);
}

int ka_db_load_synthetic(ka_engine* e, uint64_t n, int K, int32_t n_roles, uint64_t seed) {
    if (!e) return KA_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    if (n == 0 || n_roles < 1) return fail(e, KA_ERR_INVALID, "ka_db_load_synthetic: n and n_roles must be positive");
    return db_load_impl(e, nullptr, nullptr, n, K, seed, n_roles);
}

// syn_roles > 0: the lines come from the synthetic generator (syn_seed, syn_roles), kmers/role_ids unused
static int db_load_impl(ka_engine* e, const uint8_t* kmers, const int32_t* role_ids, uint64_t n, int K,
                        uint64_t syn_seed, int32_t syn_roles) {
    if (K < 1 || K > KMAX) return fail(e, KA_ERR_K, "K = %d: this engine packs 5 bits per residue, K must be 1..%d", K, KMAX);
    e->have_db = false;

    // 1. alphabet: the distinct bytes of the DB, scanned on device 0
    Device& d0 = e->devs[0];
    uint32_t bitmap[8] = {0};
    const bool synthetic = syn_roles > 0;
    if (synthetic) {
        for (const char* a = "ACDEFGHIKLMNPQRSTVWY"; *a; a++) bitmap[(uint8_t)*a >> 5] |= 1u << ((uint8_t)*a & 31);
    } else {
        cudaSetDevice(d0.id);
        cudaStream_t st = d0.pipe[0].st;
        uint32_t* dbm = nullptr; uint8_t* dk = nullptr;
        const uint64_t CH = 256ull << 20;
        uint64_t total = n * (uint64_t)K, ch = std::min<uint64_t>(CH, total ? total : 1);
        if (cudaMalloc((void**)&dbm, 32) != cudaSuccess || cudaMalloc((void**)&dk, ch) != cudaSuccess) {
            if (dbm) cudaFree(dbm);
            return fail(e, KA_ERR_OOM, "ka_db_load: alphabet staging allocation failed");
        }
        cudaError_t ce = cudaMemsetAsync(dbm, 0, 32, st);
        for (uint64_t i = 0; i < total && ce == cudaSuccess; i += ch) {
            uint64_t m = std::min(ch, total - i);
            ce = cudaMemcpyAsync(dk, kmers + i, m, cudaMemcpyHostToDevice, st);
            if (ce == cudaSuccess) ce = launch_alphabet_scan(dk, m, dbm, st);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        }
        if (ce == cudaSuccess) ce = cudaMemcpy(bitmap, dbm, 32, cudaMemcpyDeviceToHost);
        cudaFree(dbm); cudaFree(dk);
        if (ce != cudaSuccess) return fail(e, KA_ERR_CUDA, "ka_db_load: alphabet scan: %s", cudaGetErrorString(ce));
    }
    memset(e->lut, 0, 256);
    int nsym = 0;
    for (int b = 0; b < 256; b++)
        if (bitmap[b >> 5] & (1u << (b & 31))) {
            nsym++;
            if (nsym <= 31) e->lut[b] = (uint8_t)nsym;  // codes 1..31 in byte order; 0 = absent
        }
    if (nsym > 31)
        return fail(e, KA_ERR_ALPHABET,
                    "ka_db_load: the DB uses %d distinct residue bytes; at most 31 fit the 5-bit packing", nsym);

    // 2. table geometry (slot class, sector count) from n, K and the largest role id
    int32_t max_role = synthetic ? syn_roles - 1 : 0;
    for (uint64_t i = 0; !synthetic && i < n; i++) {
        if (role_ids[i] < 0) return fail(e, KA_ERR_ROLE, "ka_db_load: negative role id %d at line %llu", role_ids[i], (unsigned long long)i);
        if (role_ids[i] > max_role) max_role = role_ids[i];
    }
    DbSource src;
    src.kmers = synthetic ? nullptr : kmers; src.roles = role_ids; src.n = n; src.seed = syn_seed; src.n_roles = (uint32_t)syn_roles;
    src.synthetic = synthetic;
    while (((uint64_t)max_role + 1) >> src.role_bits) src.role_bits++;
    {
        // a slot keeps (line + 1) << role_bits | role in 64 bits while the DB streams in
        uint32_t line_bits = 1;
        while (line_bits < 64 && ((n + 1) >> line_bits)) line_bits++;
        if (line_bits + src.role_bits > 64)
            return fail(e, KA_ERR_TOO_BIG, "ka_db_load: %llu lines with role ids up to %d exceed the 64-bit (line, role) word",
                        (unsigned long long)n, max_role);
    }
    TableView geom;
    const uint32_t n_shards = e->table_mode >= 1 ? (uint32_t)e->devs.size() : 1u;
    if (e->table_mode >= 1) {
        if (n_shards != 2 && n_shards != 4 && n_shards != 8)
            return fail(e, KA_ERR_INVALID, "ka_db_load: a sharded table needs an engine on 2, 4 or 8 devices (has %u)", n_shards);
        if (!e->peers_enabled) {
            for (Device& a : e->devs) {
                cudaSetDevice(a.id);
                for (Device& b : e->devs) {
                    if (a.id == b.id) continue;
                    int can = 0;
                    cudaDeviceCanAccessPeer(&can, a.id, b.id);
                    if (!can) return fail(e, KA_ERR_NO_DEVICE, "ka_db_load: device %d cannot access device %d's memory (no NVLink/P2P)", a.id, b.id);
                    cudaError_t pe = cudaDeviceEnablePeerAccess(b.id, 0);
                    if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled)
                        return fail(e, KA_ERR_CUDA, "ka_db_load: cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(pe));
                    cudaGetLastError();
                }
            }
            e->peers_enabled = true;
        }
    }
    // communicators first: NCCL sets up its buffers before the table takes most of the HBM
    if (e->table_mode == 2 && !e->nccl_ready) {
        if (!g_nccl.load()) return fail(e, KA_ERR_NO_DEVICE, "ka_db_load: table_mode 2 needs libnccl.so.2 (%s)", dlerror() ? dlerror() : "symbols missing");
        std::vector<ncclComm_t> comms(e->devs.size());
        std::vector<int> ids;
        for (Device& d : e->devs) ids.push_back(d.id);
        auto tn = std::chrono::steady_clock::now();
        ncclResult_t nr = g_nccl.CommInitAll(comms.data(), (int)ids.size(), ids.data());
        if (nr != ncclSuccess) return fail(e, KA_ERR_CUDA, "ka_db_load: ncclCommInitAll: %s", g_nccl.GetErrorString(nr));
        if (getenv("KA_LOAD_TRACE"))
            fprintf(stderr, "[db load] ncclCommInitAll on %zu devices: %.2f s\n", ids.size(),
                    std::chrono::duration<double>(std::chrono::steady_clock::now() - tn).count());
        for (size_t i = 0; i < e->devs.size(); i++) e->devs[i].comm = comms[i];
        e->nccl_ready = true;
    }
    if (!choose_geometry(n, K, max_role, e->load_factor, e->slot_bits, n_shards, e->wide != 0, geom))
        return fail(e, KA_ERR_TOO_BIG, "ka_db_load: %llu k-mers (K=%d, max role %d) do not fit %s",
                    (unsigned long long)n, K, max_role, e->slot_bits ? "the forced slot width" : "any slot class of this build");

    // 3. build one replica per device
    auto tb = std::chrono::steady_clock::now();
    std::vector<uint64_t> nk(e->devs.size(), 0);
    std::vector<uint32_t> mp(e->devs.size(), 0);
    int rc = KA_OK;
    for (int attempt = 0; attempt < 6; attempt++) {
        if (!geom.wide && (uint64_t)geom.n_primary_slots + (uint64_t)n_shards * (2ull << geom.ovf_bbits) >= 0xfffffff0ull) {
            if (geom.cls == 128) return fail(e, KA_ERR_TOO_BIG, "ka_db_load: table exceeds the 32-bit slot index of the 128-bit slot class");
            geom.wide = 1; geom.n_primary_slots = 0;   // overflow entries pushed the token range past 32 bits
        }
        rc = for_each_device(e, [&](Device& d, int i) {
            TableView g = geom;
            g.my_shard = n_shards > 1 ? (uint32_t)i : 0u;
            return build_table(e, d, g, src, &nk[i], &mp[i]);
        });
        if (rc != KA_ERR_TOO_BIG) break;
        geom.ovf_bbits += 2;  // overflow table was too small for this key set: rebuild 4x larger
    }
    if (rc) return rc;
    if (getenv("KA_LOAD_TRACE"))
        fprintf(stderr, "[db load] table build (%llu lines, 2^%u sectors of %d-bit slots%s, %u shard(s)): %.2f s\n",
                (unsigned long long)n, geom.bbits, geom.cls, geom.wide ? ", wide" : "", n_shards,
                std::chrono::duration<double>(std::chrono::steady_clock::now() - tb).count());
    if (n_shards > 1) {
        // every device gets the peer pointers of all shards
        std::vector<const uint4*> ps(8, nullptr), po(8, nullptr);
        for (size_t i = 0; i < e->devs.size(); i++) { ps[i] = e->devs[i].table; po[i] = e->devs[i].ovf; nk[0] += i ? nk[i] : 0; mp[0] = std::max(mp[0], mp[i]); }
        for (Device& d : e->devs) {
            cudaSetDevice(d.id);
            if (!d.shard_sectors && cudaMalloc((void**)&d.shard_sectors, 64) != cudaSuccess) return fail(e, KA_ERR_OOM, "shard pointer table");
            if (!d.shard_ovf && cudaMalloc((void**)&d.shard_ovf, 64) != cudaSuccess) return fail(e, KA_ERR_OOM, "shard pointer table");
            cudaMemcpy((void*)d.shard_sectors, ps.data(), 64, cudaMemcpyHostToDevice);
            cudaMemcpy((void*)d.shard_ovf, po.data(), 64, cudaMemcpyHostToDevice);
        }
    }
    e->geom = geom;
    e->info.K = K;
    e->info.n_symbols = nsym;
    e->info.n_lines = n;
    e->info.n_keys = nk[0];
    e->info.n_buckets = 1ull << geom.bbits;
    e->info.table_bytes = (32ull << geom.bbits) + (geom.cls == 128 ? 0 : (uint64_t)n_shards * (64ull << geom.ovf_bbits));  // all shards
    e->have_sig = e->devs[0].sig != nullptr;
    e->info.max_probe = mp[0];
    e->info.slot_bits = (uint32_t)geom.cls;
    e->have_db = true;
    for (Device& d : e->devs) {
        cudaSetDevice(d.id);
        for (int k = 0; k < NPIPE; k++) set_l2_window(e, d, d.pipe[k].st);
    }
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

int ka_build(ka_engine* e, const uint8_t* residues, const uint64_t* offsets, uint64_t N,
             const int32_t* n_roles, const int32_t* peg_role, int K, uint64_t cap,
             uint8_t* out_kmers, int32_t* out_roles, uint64_t* n_out, int load_as_db) {
    if (!e || !n_out) return KA_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    *n_out = 0;
    if (K < 1 || K > KMAX) return fail(e, KA_ERR_K, "K = %d: this engine packs 5 bits per residue, K must be 1..%d", K, KMAX);
    if (N == 0) return KA_OK;
    if (!offsets || !n_roles || !peg_role || (cap && (!out_kmers || !out_roles))) return fail(e, KA_ERR_INVALID, "ka_build: NULL argument");
    if (N > 0xfffffff0ull) return fail(e, KA_ERR_TOO_BIG, "ka_build: too many pegs");
    uint64_t windows = 0;
    for (uint64_t i = 0; i < N; i++) {
        if (offsets[i + 1] < offsets[i]) return fail(e, KA_ERR_OFFSETS, "ka_build: offsets are not monotone");
        uint64_t L = offsets[i + 1] - offsets[i];
        if (n_roles[i] == 1) {
            if (peg_role[i] < 0) return fail(e, KA_ERR_ROLE, "ka_build: negative role id at peg %llu", (unsigned long long)i);
            if (L >= (uint64_t)K) windows += L - K + 1;
        }
    }
    Device& d = e->devs[0];
    cudaSetDevice(d.id);
    cudaStream_t st = d.pipe[0].st;
    const uint64_t base = offsets[0], n_res = offsets[N] - base;
    uint64_t n_slots = 1024;
    while (n_slots < 2 * windows) n_slots <<= 1;
    uint8_t *d_res = nullptr, *d_lut = nullptr, *d_inv = nullptr, *d_ok = nullptr;
    unsigned long long *d_off = nullptr, *d_cnt = nullptr;
    int32_t *d_nr = nullptr, *d_pr = nullptr, *d_or = nullptr;
    uint32_t* d_bm = nullptr;
    Slot128* d_tab = nullptr;
    auto cleanup = [&] {
        cudaFree(d_res); cudaFree(d_lut); cudaFree(d_inv); cudaFree(d_ok); cudaFree(d_off); cudaFree(d_cnt);
        cudaFree(d_nr); cudaFree(d_pr); cudaFree(d_or); cudaFree(d_bm); cudaFree(d_tab);
    };
    cudaError_t ce;
    if ((ce = cudaMalloc((void**)&d_res, n_res + K + 64)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&d_off, (N + 1) * 8)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&d_nr, N * 4)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&d_pr, N * 4)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&d_lut, 256)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&d_inv, 32)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&d_bm, 32)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&d_cnt, 8)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&d_tab, n_slots * sizeof(Slot128))) != cudaSuccess ||
        (ce = cudaMalloc((void**)&d_ok, std::max<uint64_t>(cap, 1) * K)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&d_or, std::max<uint64_t>(cap, 1) * 4)) != cudaSuccess) {
        cleanup();
        return fail(e, KA_ERR_OOM, "ka_build: device allocation failed: %s", cudaGetErrorString(ce));
    }
    auto bail = [&](const char* what, cudaError_t c) {
        cleanup();
        return fail(e, KA_ERR_CUDA, "ka_build: %s: %s", what, cudaGetErrorString(c));
    };
    if (n_res && (ce = cudaMemcpyAsync(d_res, residues + base, n_res, cudaMemcpyHostToDevice, st)) != cudaSuccess) return bail("H2D residues", ce);
    if ((ce = cudaMemcpyAsync(d_off, offsets, (N + 1) * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) return bail("H2D offsets", ce);
    if ((ce = cudaMemcpyAsync(d_nr, n_roles, N * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) return bail("H2D n_roles", ce);
    if ((ce = cudaMemcpyAsync(d_pr, peg_role, N * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) return bail("H2D peg_role", ce);
    if ((ce = cudaMemsetAsync(d_bm, 0, 32, st)) != cudaSuccess) return bail("memset", ce);
    if ((ce = cudaMemsetAsync(d_cnt, 0, 8, st)) != cudaSuccess) return bail("memset", ce);
    if ((ce = cudaMemsetAsync(d_tab, 0, n_slots * sizeof(Slot128), st)) != cudaSuccess) return bail("memset table", ce);
    // alphabet of the training proteins (same rule as ka_db_load: at most 31 distinct bytes)
    if ((ce = launch_alphabet_scan(d_res, n_res, d_bm, st)) != cudaSuccess) return bail("alphabet scan", ce);
    uint32_t bitmap[8];
    if ((ce = cudaMemcpyAsync(bitmap, d_bm, 32, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return bail("D2H alphabet", ce);
    if ((ce = cudaStreamSynchronize(st)) != cudaSuccess) return bail("alphabet sync", ce);
    uint8_t lut[256], inv[32];
    memset(lut, 0, 256); memset(inv, 0, 32);
    int nsym = 0;
    for (int b = 0; b < 256; b++)
        if (bitmap[b >> 5] & (1u << (b & 31))) {
            nsym++;
            if (nsym <= 31) { lut[b] = (uint8_t)nsym; inv[nsym] = (uint8_t)b; }
        }
    if (nsym > 31) {
        cleanup();
        return fail(e, KA_ERR_ALPHABET, "ka_build: the proteins use %d distinct residue bytes; at most 31 fit the 5-bit packing", nsym);
    }
    if ((ce = cudaMemcpyAsync(d_lut, lut, 256, cudaMemcpyHostToDevice, st)) != cudaSuccess) return bail("H2D lut", ce);
    if ((ce = cudaMemcpyAsync(d_inv, inv, 32, cudaMemcpyHostToDevice, st)) != cudaSuccess) return bail("H2D inv", ce);
    if ((ce = launch_build_pass(1, d_res, d_off, (uint32_t)N, d_nr, d_pr, K, d_lut, d_tab, n_slots, st)) != cudaSuccess) return bail("build pass 1", ce);
    if ((ce = launch_build_pass(2, d_res, d_off, (uint32_t)N, d_nr, d_pr, K, d_lut, d_tab, n_slots, st)) != cudaSuccess) return bail("build pass 2", ce);
    if ((ce = launch_build_emit(d_tab, n_slots, K, d_inv, cap, d_ok, d_or, d_cnt, st)) != cudaSuccess) return bail("build emit", ce);
    unsigned long long found = 0;
    if ((ce = cudaMemcpyAsync(&found, d_cnt, 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return bail("D2H count", ce);
    if ((ce = cudaStreamSynchronize(st)) != cudaSuccess) return bail("build kernels", ce);
    *n_out = found;
    if (found > cap) {
        cleanup();
        return fail(e, KA_ERR_TOO_BIG, "ka_build: %llu k-mers found, output capacity is %llu", found, (unsigned long long)cap);
    }
    if (found) {
        if ((ce = cudaMemcpy(out_kmers, d_ok, found * K, cudaMemcpyDeviceToHost)) != cudaSuccess) return bail("D2H kmers", ce);
        if ((ce = cudaMemcpy(out_roles, d_or, found * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) return bail("D2H roles", ce);
    }
    cleanup();
    if (load_as_db && found) return db_load_impl(e, out_kmers, out_roles, found, K);
    return KA_OK;
}

// Pairwise k-mer distance for query groups (GeneCopyProcessor.java:137-142); see include/kmeranno.h.
int ka_kmer_distance(ka_engine* e, const uint8_t* residues, const uint64_t* offsets, uint64_t N, int K,
                     const uint32_t* query_seq, const uint64_t* group_offsets, uint64_t Q,
                     const uint32_t* cand_seq, int32_t* out_set_size, int32_t* out_common, double* out_distance) {
    if (!e) return KA_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    if (K < 1 || K > KMAX) return fail(e, KA_ERR_K, "K = %d: this engine packs 5 bits per residue, K must be 1..%d", K, KMAX);
    if (N == 0) return Q ? fail(e, KA_ERR_INVALID, "ka_kmer_distance: queries over an empty batch") : (int)KA_OK;
    if (!offsets || (Q && (!query_seq || !group_offsets))) return fail(e, KA_ERR_INVALID, "ka_kmer_distance: NULL argument");
    if (N > 0xfffffff0ull || Q > 0xfffffff0ull) return fail(e, KA_ERR_TOO_BIG, "ka_kmer_distance: too many sequences or queries");
    const uint64_t base = offsets[0];
    for (uint64_t i = 0; i < N; i++) {
        if (offsets[i + 1] < offsets[i]) return fail(e, KA_ERR_OFFSETS, "ka_kmer_distance: offsets are not monotone");
        if (offsets[i + 1] - offsets[i] > 0x3fffffffull) return fail(e, KA_ERR_TOO_BIG, "ka_kmer_distance: a sequence exceeds 2^30 residues");
    }
    const uint64_t n_res = offsets[N] - base;
    if (n_res && !residues) return fail(e, KA_ERR_INVALID, "ka_kmer_distance: residues is NULL");
    const uint64_t M = Q ? group_offsets[Q] : 0;
    if (Q && group_offsets[0] != 0) return fail(e, KA_ERR_OFFSETS, "ka_kmer_distance: group_offsets must start at 0");
    for (uint64_t q = 0; q < Q; q++) {
        if (group_offsets[q + 1] < group_offsets[q]) return fail(e, KA_ERR_OFFSETS, "ka_kmer_distance: group_offsets are not monotone");
        if (group_offsets[q + 1] - group_offsets[q] > 0xfffffff0ull) return fail(e, KA_ERR_TOO_BIG, "ka_kmer_distance: a query has too many candidates");
        if (query_seq[q] >= N) return fail(e, KA_ERR_INVALID, "ka_kmer_distance: query %llu names sequence %u of %llu", (unsigned long long)q, query_seq[q], (unsigned long long)N);
    }
    if (M && (!cand_seq || !out_common || !out_distance)) return fail(e, KA_ERR_INVALID, "ka_kmer_distance: NULL argument");
    for (uint64_t m = 0; m < M; m++)
        if (cand_seq[m] >= N) return fail(e, KA_ERR_INVALID, "ka_kmer_distance: candidate %llu names sequence %u of %llu", (unsigned long long)m, cand_seq[m], (unsigned long long)N);

    // hash-set placement: shared memory for the common lengths, a global slice beyond
    auto windows = [&](uint64_t i) { uint64_t L = offsets[i + 1] - offsets[i]; return (uint32_t)(L >= (uint64_t)K ? L - K + 1 : 0); };
    // shared-memory set size: the smallest power of two that holds the set of 90 % of the sequences
    // (at most 8192 entries = 64 KB); the long tail uses global slices, the common case keeps
    // many CTAs per SM
    uint32_t smem_cap = 64;
    {
        uint64_t by_cap[32] = {0};
        for (uint64_t i = 0; i < N; i++) { uint32_t c = dist_set_cap(windows(i)); int b = 0; while ((1u << b) < c) b++; by_cap[b]++; }
        uint64_t seen = 0;
        for (int b = 6; b <= 13; b++) {
            seen += by_cap[b];
            smem_cap = 1u << b;
            if (seen * 10 >= N * 9) break;
        }
    }
    std::vector<unsigned long long> seq_scratch(N, 0), query_scratch(Q ? Q : 1, 0);
    uint64_t need_seq = 0, need_query = 0;
    for (uint64_t i = 0; i < N; i++) { uint32_t c = dist_set_cap(windows(i)); if (c > smem_cap) { seq_scratch[i] = need_seq; need_seq += c; } }
    for (uint64_t q = 0; q < Q; q++) { uint32_t c = dist_set_cap(windows(query_seq[q])); if (c > smem_cap) { query_scratch[q] = need_query; need_query += c; } }
    const uint64_t n_scratch = std::max<uint64_t>(std::max(need_seq, need_query), 1);

    // work-balanced contiguous query ranges, one per device (work = residues streamed)
    const size_t nd = e->devs.size();
    std::vector<uint64_t> cut(nd + 1, 0);
    {
        std::vector<uint64_t> work(Q + 1, 0);
        for (uint64_t q = 0; q < Q; q++) {
            uint64_t w = offsets[query_seq[q] + 1] - offsets[query_seq[q]] + 64;
            for (uint64_t m = group_offsets[q]; m < group_offsets[q + 1]; m++) w += offsets[cand_seq[m] + 1] - offsets[cand_seq[m]] + 16;
            work[q + 1] = work[q] + w;
        }
        cut[nd] = Q;
        for (size_t i = 1; i < nd; i++)
            cut[i] = std::max<uint64_t>(cut[i - 1], std::lower_bound(work.begin(), work.end(), work[Q] / nd * i) - work.begin());
    }

    auto t0 = std::chrono::steady_clock::now();
    e->stats = ka_stats{};
    int rc = for_each_device(e, [&](Device& d, int i) {
        const uint64_t qa = cut[i], qb = cut[i + 1];
        d.kernel_ms = d.tile_ms = 0; d.launches = 0; d.h2d = d.d2h = 0; d.probes = 0;
        if (i != 0 && qa == qb) return (int)KA_OK;          // device 0 always reports the set sizes
        DCK(d, cudaSetDevice(d.id));
        cudaStream_t st = d.pipe[0].st;
        Pipe& p = d.pipe[0];
        std::vector<void*> owned;
        auto alloc = [&](size_t bytes) -> void* {
            void* q = nullptr;
            if (cudaMalloc(&q, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
            owned.push_back(q);
            return q;
        };
        auto finish = [&](int code, const char* what, cudaError_t ce) {
            for (void* q : owned) cudaFree(q);
            return code == KA_OK ? (int)KA_OK : dev_fail(d, code, what, ce);
        };
        DistParams dp{};
        uint8_t* d_res = (uint8_t*)alloc(n_res + K + 64);
        unsigned long long* d_off = (unsigned long long*)alloc((N + 1) * 8);
        uint32_t* d_qs = (uint32_t*)alloc(Q * 4);
        unsigned long long* d_go = (unsigned long long*)alloc((Q + 1) * 8);
        uint32_t* d_cs = (uint32_t*)alloc(M * 4);
        int32_t* d_size = (int32_t*)alloc(N * 4);
        int32_t* d_common = (int32_t*)alloc(M * 4);
        double* d_dist = (double*)alloc(M * 8);
        unsigned long long* d_sk = (unsigned long long*)alloc(n_scratch * 8);
        uint8_t* d_uniq = (uint8_t*)alloc(n_res + 64);
        unsigned long long* d_ss = (unsigned long long*)alloc(N * 8);
        unsigned long long* d_qsc = (unsigned long long*)alloc((Q ? Q : 1) * 8);
        if (!d_res || !d_off || !d_qs || !d_go || !d_cs || !d_size || !d_common || !d_dist || !d_sk || !d_uniq || !d_ss || !d_qsc)
            return finish(KA_ERR_OOM, "ka_kmer_distance: device allocation", cudaErrorMemoryAllocation);
        cudaError_t ce = cudaSuccess;
        auto step = [&](cudaError_t c) { if (ce == cudaSuccess) ce = c; };
        if (n_res) step(cudaMemcpyAsync(d_res, residues + base, n_res, cudaMemcpyHostToDevice, st));
        step(cudaMemcpyAsync(d_off, offsets, (N + 1) * 8, cudaMemcpyHostToDevice, st));
        if (Q) step(cudaMemcpyAsync(d_qs, query_seq, Q * 4, cudaMemcpyHostToDevice, st));
        if (Q) step(cudaMemcpyAsync(d_go, group_offsets, (Q + 1) * 8, cudaMemcpyHostToDevice, st));
        if (M) step(cudaMemcpyAsync(d_cs, cand_seq, M * 4, cudaMemcpyHostToDevice, st));
        step(cudaMemcpyAsync(d_ss, seq_scratch.data(), N * 8, cudaMemcpyHostToDevice, st));
        if (Q) step(cudaMemcpyAsync(d_qsc, query_scratch.data(), Q * 8, cudaMemcpyHostToDevice, st));
        // alphabet of the batch (at most 31 distinct bytes), scanned from the device copy
        uint8_t* d_lut = (uint8_t*)alloc(256);             // not d.lut: that one belongs to the loaded DB
        uint32_t* d_bm = (uint32_t*)alloc(32);
        if (!d_lut || !d_bm) return finish(KA_ERR_OOM, "ka_kmer_distance: device allocation", cudaErrorMemoryAllocation);
        uint32_t bitmap[8] = {0};
        step(cudaMemsetAsync(d_bm, 0, 32, st));
        step(launch_alphabet_scan(d_res, n_res, d_bm, st));
        step(cudaMemcpyAsync(bitmap, d_bm, 32, cudaMemcpyDeviceToHost, st));
        step(cudaStreamSynchronize(st));
        if (ce != cudaSuccess) return finish(KA_ERR_CUDA, "ka_kmer_distance: alphabet scan", ce);
        uint8_t lut[256];
        memset(lut, 0, 256);
        int nsym = 0;
        for (int b = 0; b < 256; b++)
            if (bitmap[b >> 5] & (1u << (b & 31))) { nsym++; if (nsym <= 31) lut[b] = (uint8_t)nsym; }
        if (nsym > 31) {
            for (void* q : owned) cudaFree(q);
            d.err = KA_ERR_ALPHABET;
            d.errmsg = "ka_kmer_distance: the proteins use " + std::to_string(nsym) + " distinct residue bytes; at most 31 fit the 5-bit packing";
            return (int)KA_ERR_ALPHABET;
        }
        step(cudaMemcpyAsync(d_lut, lut, 256, cudaMemcpyHostToDevice, st));
        d.h2d = n_res + (N + 1) * 8 + Q * 4 + (Q + 1) * 8 + M * 4 + N * 8 + Q * 8;
        step(dist_set_smem(dist_smem_bytes(8192)));
        dp.res = d_res; dp.off = d_off; dp.base = base; dp.n_seq = (uint32_t)N; dp.K = K; dp.key_mask = (1ull << (5 * K)) - 1; dp.lut = d_lut;
        dp.set_size = d_size; dp.query_seq = d_qs; dp.group_off = d_go; dp.q_begin = (uint32_t)qa; dp.q_end = (uint32_t)qb;
        dp.cand_seq = d_cs; dp.common = d_common; dp.dist = d_dist; dp.smem_cap = smem_cap;
        dp.scratch_keys = d_sk; dp.uniq = d_uniq; dp.seq_scratch = d_ss; dp.query_scratch = d_qsc;
        step(cudaEventRecord(p.ev_k0, st));
        step(launch_set_size(dp, d.sm_count, st));
        step(launch_common(dp, d.sm_count, st));
        step(cudaEventRecord(p.ev_k1, st));
        d.launches = 2;
        const uint64_t ma = Q ? group_offsets[qa] : 0, mb = Q ? group_offsets[qb] : 0;
        if (i == 0 && out_set_size) step(cudaMemcpyAsync(out_set_size, d_size, N * 4, cudaMemcpyDeviceToHost, st));
        if (mb > ma) {
            step(cudaMemcpyAsync(out_common + ma, d_common + ma, (mb - ma) * 4, cudaMemcpyDeviceToHost, st));
            step(cudaMemcpyAsync(out_distance + ma, d_dist + ma, (mb - ma) * 8, cudaMemcpyDeviceToHost, st));
        }
        d.d2h = (i == 0 && out_set_size ? N * 4 : 0) + (mb - ma) * 12;
        step(cudaStreamSynchronize(st));
        if (ce != cudaSuccess) return finish(KA_ERR_CUDA, "ka_kmer_distance", ce);
        float ms = 0;
        if (cudaEventElapsedTime(&ms, p.ev_k0, p.ev_k1) == cudaSuccess) d.kernel_ms = ms;
        return finish(KA_OK, "", cudaSuccess);
    });
    if (rc) return rc;
    ka_stats& s = e->stats;
    s.sequences = N; s.residues = n_res;
    for (Device& d : e->devs) {
        s.kernel_launches += d.launches; s.h2d_bytes += d.h2d; s.d2h_bytes += d.d2h;
        s.kernel_ms = std::max(s.kernel_ms, d.kernel_ms);
    }
    s.wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return KA_OK;
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
