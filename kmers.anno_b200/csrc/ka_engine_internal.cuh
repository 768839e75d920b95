// ka_engine_internal.cuh — types and helpers shared by the engine's translation units
// (ka_engine.cu: lifetime, options, annotate pipeline; ka_table.cu: table geometry and build;
// ka_route.cu: NCCL-routed sharded table; ka_build_api.cu: ka_build; ka_distance.cu: ka_kmer_distance).
// Nothing here is part of the C ABI.
#pragma once
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
#include <unistd.h>
#include <vector>

#include "../../include/kmeranno.h"
#include "ka_kernels.cuh"
#include "ka_line.cuh"


namespace kai {

using namespace ka;

constexpr int NPIPE = 4;  // chunks in flight per device

struct Pipe {
    cudaStream_t st = nullptr;
    uint8_t* res = nullptr; size_t res_cap = 0;          // residue bytes of the chunk
    uint32_t* pk = nullptr; size_t pk_cap = 0;           // 5-bit codes of the chunk (words), ka_line.cuh
    unsigned long long* off = nullptr; size_t seq_cap = 0;   // absolute 64-bit offsets
    uint32_t* off32_in = nullptr;                        // absolute 32-bit offsets as given to ka_annotate_packed
    uint32_t* off32 = nullptr;                           // chunk-relative offsets written by line_plan_kernel
    uint4* first = nullptr; size_t first_cap = 0;
    uint2* surv = nullptr; size_t surv_cap = 0;          // line table: survivors of the filter pass (one slot per residue position)
    uint32_t* surv_cnt = nullptr; size_t surv_cnt_cap = 0;
    int32_t* role = nullptr; int32_t* hits = nullptr; uint8_t* flag = nullptr;
    uint32_t* ctr = nullptr;  // 16 bytes: [0] big_count, [2..3] token cursor (u64)
    BigItem* big = nullptr; size_t big_cap = 0;
    uint4* mid = nullptr; size_t mid_cap = 0;
    uint32_t* scratch = nullptr; size_t scratch_cap = 0;
    cudaEvent_t ev_k0 = nullptr, ev_t0 = nullptr, ev_t1 = nullptr, ev_k1 = nullptr, done = nullptr;
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;       // line table: hand-over to and from the device's compute stream
    bool busy = false;
    // option "ingest_via": the chunk lands on another GPU with a faster host path and crosses NVLink from there
    uint8_t* via_buf = nullptr; size_t via_cap = 0;      // staging buffer ON the via device
    cudaStream_t via_st = nullptr;                       // stream on the via device
    cudaEvent_t via_ev = nullptr;
};

struct Device {
    int id = 0;
    int sm_count = 148;
    uint4* table = nullptr;
    uint4* ovf = nullptr;     // overflow table (cls 32/64)
    uint32_t* filt = nullptr; // presence filter of the line table: one word per sector
    cudaStream_t line_st = nullptr;  // line table: the passes of all chunks run on ONE stream (see enqueue_line_kernels)
    uint8_t* lut5 = nullptr;  // 256-byte residue -> radix digit table (31 = not in the DB alphabet)
    uint8_t* inv32 = nullptr; // 32-byte digit -> residue byte table (entry 31 = a byte outside the alphabet)
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
        cudaEvent_t ev_scatter = nullptr, ev_lookup = nullptr, ev_tally = nullptr;   // peer-store transport
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

}  // namespace kai

struct ka_batch {
    int dev_index = 0;
    uint64_t n_seq = 0, n_res = 0, base = 0, long_res = 0, n_long = 0, n_mid = 0;
    uint64_t origin = 0;                       // line table: absolute residue index of chunk-relative residue 0
    uint32_t tile_span = 0, long_seq = 0, mid_seq = 0;   // tiling options the batch was scanned with
    uint64_t db_serial = 0;                    // the DB load this batch was prepared for
    kai::Pipe p;  // owns device buffers of the resident batch
};

struct ka_engine {
    std::vector<kai::Device> devs;
    std::mutex mu;
    std::string err;
    // options
    double load_factor = 0;     // 0 = default of the chosen layout (0.4 sector classes, 0.68 line table)
    uint32_t tile_span = 1536;
    uint32_t long_seq = 2048;
    uint32_t mid_seq = 8192;
    uint64_t chunk_residues = 0;       // 0 = automatic: 64 Mi, 32 Mi on a routed table (see chunk_of)
    int l2_persist = 1;
    int slot_bits = 0;  // 0 = choose automatically; 16 = the 128-byte-line table (ka_line.cuh)
    int filter = 1;     // line table: 1 = L2-resident presence filter in front of it (measurement knob)
    int ingest_via = -1;      // CUDA device id whose PCIe path carries the H2D copies of a single-device engine (-1 = its own)
    int resident_packed = 1;  // resident batches of narrow sector tables are kept as the 5-bit stream (measurement knob)
    int table_mode = 0; // next ka_db_load: 0 = replicated, 1 = sharded by sector range (peer loads), 2 = sharded + NCCL routing
    int wide = 0;       // next ka_db_load: 1 = force the wide-table kernels (64-bit sector indices and tokens)
    // state of the LOADED database (options above only take effect at the next ka_db_load)
    int db_table_mode = 0;
    bool line = false;          // the loaded table is a line table (slot class 16)
    ka::LineTable lgeom{};      // its geometry (pointers filled per device)
    uint64_t db_serial = 0;     // bumped by every successful load
    bool peers_enabled = false;
    bool nccl_ready = false;
    // db
    bool have_db = false;
    ka_db_info info{};
    ka::TableView geom{};   // geometry of the loaded table (sectors pointer filled per device)
    uint8_t lut[256];           // residue byte -> 5-bit field 1..31 (0 = absent): sector classes 32/64/128
    uint8_t lut5[256];          // residue byte -> radix digit 0..n-1 (31 = absent): line table and packed streams
    uint8_t inv32[32];          // digit -> residue byte; inv32[31] = some byte outside the alphabet
    ka_stats stats{};
};


namespace kai {

int fail(ka_engine* e, int code, const char* fmt, ...);
int dev_fail(Device& d, int code, const char* what, cudaError_t ce);

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


// A host batch in one of the two input forms of the ABI.
struct BatchIn {
    const uint8_t* residues = nullptr; const uint64_t* off64 = nullptr;   // ka_annotate: bytes + absolute 64-bit offsets
    const uint8_t* codes = nullptr; const uint32_t* off32 = nullptr;      // ka_annotate_packed: 5-bit stream + 32-bit offsets
    uint64_t off(uint64_t i) const { return off64 ? off64[i] : (uint64_t)off32[i]; }
    bool packed() const { return codes != nullptr; }
};

struct ChunkShape {
    uint64_t n_res = 0, n_long = 0, long_res = 0, probes = 0, n_mid = 0;
    uint64_t n_mid_seg = 0;    // line table: segments of the mid sequences (ka_line.cuh LINE_MID_SEG)
};


int pipe_init(Device& d, Pipe& p);
void pipe_free(Pipe& p);
int pipe_reserve(Device& d, Pipe& p, uint64_t n_res, uint64_t n_seq, uint64_t n_tiles,
                 uint64_t n_long, uint64_t long_res, uint64_t n_mid, bool wide, bool need_bytes, bool need_codes,
                 bool need_surv = false);
bool scan_offsets(const BatchIn& in, uint64_t cs, uint64_t ce, uint32_t long_seq, uint32_t mid_seq, int K,
                  ChunkShape& s);
void fill_params(ka_engine* e, Device& d, Pipe& p, uint64_t base, uint64_t n_res, uint64_t n_seq,
                 int32_t min_hits, AnnotParams& ap);
int ensure_tile_smem(Device& d);
int enqueue_kernels(ka_engine* e, Device& d, Pipe& p, const AnnotParams& ap, uint64_t n_long, uint64_t n_mid);
int collect_times(Device& d, Pipe& p);
void set_l2_window(ka_engine* e, Device& d, cudaStream_t st);
// residues per pipelined chunk: the option, or the measured best of the mode (profiles/r02_summary.md D, G)
static inline uint64_t chunk_of(const ka_engine* e, bool routed) {
    return e->chunk_residues ? e->chunk_residues : (routed ? 32ull << 20 : 64ull << 20);
}

// Tile span of the line passes for a chunk of n_res residues: one tile is one warp of the filter and probe passes,
// so a chunk with fewer tiles than the GPU holds warps (a single proteome: config 2) is cut into smaller tiles.
static inline uint32_t line_span(const ka_engine* e, const Device& d, uint64_t n_res) {
    const uint64_t want = (n_res / ((uint64_t)d.sm_count * 32) + 127) & ~127ull;
    return (uint32_t)std::min<uint64_t>(e->tile_span, std::max<uint64_t>(256, want));
}
static inline uint64_t tiles_of(const ka_engine* e, const Device& d, uint64_t n_res) {
    return n_res / (e->line ? line_span(e, d, n_res) : e->tile_span) + 1;
}

int annotate_range(ka_engine* e, Device& d, const BatchIn& in, uint64_t s_begin, uint64_t s_end, int32_t min_hits,
                   int32_t* out_role, int32_t* out_hits, uint8_t* out_flag);
void fill_line_params(ka_engine* e, Device& d, Pipe& p, uint64_t n_res, uint64_t n_seq, int32_t min_hits, LineParams& lp);
int enqueue_line_kernels(ka_engine* e, Device& d, Pipe& p, const LineParams& lp, bool off_is_64, uint64_t origin,
                         uint64_t n_long, uint64_t n_mid, bool solo);

// ---- ka_table.cu ----
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

int build_table(ka_engine* e, Device& d, const TableView& geom, const DbSource& src, uint64_t* n_keys, uint32_t* max_probe);
// line table (slot class 16): geometry for n keys of K digits over `nsym` symbols; false = not representable
bool choose_line_geometry(uint64_t n, int K, int nsym, double lf, bool forced, LineTable& g);
int build_line_table(ka_engine* e, Device& d, const LineTable& geom, const DbSource& src, uint64_t* counts3);
bool choose_geometry(uint64_t n, int K, int32_t max_role, double lf, int force_cls, uint32_t n_shards, bool force_wide, TableView& g);
// syn_roles > 0: the lines come from the synthetic generator (syn_seed, syn_roles), kmers/role_ids unused.
// on_device: kmers / role_ids are DEVICE pointers (the output of ka_build, never copied to the host) and
// max_role_hint bounds the role ids.
int db_load_impl(ka_engine* e, const uint8_t* kmers, const int32_t* role_ids, uint64_t n, int K,
                 uint64_t syn_seed = 0, int32_t syn_roles = 0, bool on_device = false, int32_t max_role_hint = 0);

// ---- ka_route.cu ----
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
    // peer-store transport (table_mode 3): every device publishes, per lane, its receive and answer buffers and the
    // events other devices order their kernels against
    struct Pub {
        unsigned long long* recv[2] = {nullptr, nullptr};
        unsigned long long* ans[2] = {nullptr, nullptr};
        cudaEvent_t scatter_done[2] = {nullptr, nullptr}, lookup_done[2] = {nullptr, nullptr}, tally_done[2] = {nullptr, nullptr};
    };
    std::vector<Pub> pub;
    explicit RouteShared(int n) : bar(n), counts(n), n_chunks(n, 0), pub(n) {}
};

int route_init_comms(ka_engine* e);      // dlopen libnccl, one communicator per device (idempotent)
void route_destroy_comm(Device& d);
int annotate_routed_range(ka_engine* e, Device& d, int idx, RouteShared& sh, const BatchIn& in,
                          uint64_t s_begin, uint64_t s_end, int32_t min_hits,
                          int32_t* out_role, int32_t* out_hits, uint8_t* out_flag);

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


}  // namespace kai
