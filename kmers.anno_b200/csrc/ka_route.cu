// ka_route.cu — sharded table, routed form (option table_mode = 2): the k-mer keys travel to the GPU
// that owns their table sector with NCCL send/recv (all-to-all over NVLink), see annotate_routed_range.
#include "ka_engine_internal.cuh"

using namespace ka;
using namespace kai;

namespace kai {

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

int route_init_comms(ka_engine* e) {
    if (e->nccl_ready) return KA_OK;
    if (!g_nccl.load()) return fail(e, KA_ERR_NO_DEVICE, "ka_db_load: table_mode 2 needs libnccl.so.2 (%s)", dlerror() ? dlerror() : "symbols missing");
    std::vector<ncclComm_t> comms(e->devs.size());
    std::vector<int> ids;
    for (Device& d : e->devs) ids.push_back(d.id);
    auto tn = std::chrono::steady_clock::now();
    // NCCL writes its log (with NCCL_DEBUG set, at least the version banner) to STDOUT, which belongs to the
    // caller's report (ApplyKmerProcessor.java:94).  The library itself touches neither the environment nor
    // the process's file descriptors; a single-threaded host that wants the banner on stderr (the CLI, bench.py)
    // sets NCCL_DEBUG_FILE itself and opts into the descriptor swap with KA_NCCL_STDOUT_TO_STDERR=1.
    const char* swap = getenv("KA_NCCL_STDOUT_TO_STDERR");
    int saved_stdout = -1;
    if (swap && swap[0] == '1') {
        fflush(stdout);
        saved_stdout = dup(1);
        if (saved_stdout >= 0) dup2(2, 1);
    }
    ncclResult_t nr = g_nccl.CommInitAll(comms.data(), (int)ids.size(), ids.data());
    if (saved_stdout >= 0) { fflush(stdout); dup2(saved_stdout, 1); close(saved_stdout); }
    if (nr != ncclSuccess) return fail(e, KA_ERR_CUDA, "ka_db_load: ncclCommInitAll: %s", g_nccl.GetErrorString(nr));
    if (getenv("KA_LOAD_TRACE"))
        fprintf(stderr, "[db load] ncclCommInitAll on %zu devices: %.2f s\n", ids.size(),
                std::chrono::duration<double>(std::chrono::steady_clock::now() - tn).count());
    for (size_t i = 0; i < e->devs.size(); i++) e->devs[i].comm = comms[i];
    e->nccl_ready = true;
    return KA_OK;
}

void route_destroy_comm(Device& d) {
    if (d.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(d.comm);
    d.comm = nullptr;
}

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
    for (cudaEvent_t* ev : {&d.ev_scatter, &d.ev_lookup, &d.ev_tally})
        if (!*ev && cudaEventCreateWithFlags(ev, cudaEventDisableTiming) != cudaSuccess) return dev_fail(dev, KA_ERR_CUDA, "routing event", cudaErrorUnknown);
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
int annotate_routed_range(ka_engine* e, Device& d, int idx, RouteShared& sh, const BatchIn& bin,
                          uint64_t s_begin, uint64_t s_end, int32_t min_hits,
                          int32_t* out_role, int32_t* out_hits, uint8_t* out_flag) {
    const int nd = (int)e->devs.size();
    cudaSetDevice(d.id);
    d.kernel_ms = d.tile_ms = 0; d.launches = 0; d.h2d = d.d2h = 0; d.probes = 0;
    d.err = KA_OK; d.errmsg.clear();
    const bool packed = bin.packed();
    std::vector<std::pair<uint64_t, uint64_t>> chunks;
    for (uint64_t cs = s_begin; cs < s_end;) {
        const uint64_t lim = bin.off(cs) + chunk_of(e, true);
        uint64_t lo = cs + 1, hi = s_end + 1;            // last sequence boundary at or below the limit
        while (lo < hi) { const uint64_t mid = lo + ((hi - lo) >> 1); if (bin.off(mid) <= lim) lo = mid + 1; else hi = mid; }
        uint64_t ce = lo - 1;
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
    ensure_tile_smem(d);   // a failure is recorded in d.err and turns the rounds of this device into empty ones
    if (!d.ev_route0) { cuda_ok(cudaEventCreate(&d.ev_route0), "event"); cuda_ok(cudaEventCreate(&d.ev_route1), "event"); }

    const bool peer = e->db_table_mode == 3;     // keys and answers travel as NVLink stores of the scatter / lookup kernels
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
            if (!scan_offsets(bin, R.cs, R.ce, e->long_seq, e->mid_seq, e->info.K, R.shp)) { d.err = KA_ERR_OFFSETS; d.errmsg = "offsets are not monotone"; R.has = false; }
            else if (R.shp.n_res > 0x7fffffffull) { d.err = KA_ERR_TOO_BIG; d.errmsg = "chunk exceeds 2^31 residues"; R.has = false; }
        }
        if (R.has) {
            d.probes += R.shp.probes;
            if (pipe_reserve(d, p, R.shp.n_res, R.n, R.shp.n_res / e->tile_span + 1, R.shp.n_long, R.shp.long_res, R.shp.n_mid, e->geom.wide != 0, !packed || R.shp.n_long, packed) ||
                route_reserve(d, ln, R.shp.n_res + 64, 0)) R.has = false;
        }
        if (ln.h_cnt) memset(ln.h_cnt, 0, 64);
        if (R.has) {
            const uint64_t r_begin = bin.off(R.cs), r_end = bin.off(R.ce);
            const uint64_t origin = r_begin & ~127ull;
            if (packed) {
                // the extract and tally kernels stage the 5-bit stream themselves; the long-sequence kernel reads bytes
                const uint64_t byte0 = origin * 5 / 8, byte1 = (r_end * 5 + 7) / 8;
                if (byte1 > byte0) cuda_ok(cudaMemcpyAsync(p.pk, bin.codes + byte0, byte1 - byte0, cudaMemcpyHostToDevice, st), "H2D codes");
                cuda_ok(cudaMemcpyAsync(p.off32_in, bin.off32 + R.cs, (R.n + 1) * 4, cudaMemcpyHostToDevice, st), "H2D offsets");
                cuda_ok(launch_widen_offsets(p.off32_in, R.n + 1, p.off, st), "widen offsets");
                if (R.shp.n_long) cuda_ok(launch_unpack(p.pk, (uint32_t)(r_begin - origin), R.shp.n_res, d.inv32, p.res, st), "unpack");
                d.h2d += (byte1 - byte0) + (R.n + 1) * 4;
            } else {
                if (R.shp.n_res) cuda_ok(cudaMemcpyAsync(p.res, bin.residues + r_begin, R.shp.n_res, cudaMemcpyHostToDevice, st), "H2D residues");
                cuda_ok(cudaMemcpyAsync(p.off, bin.off64 + R.cs, (R.n + 1) * 8, cudaMemcpyHostToDevice, st), "H2D offsets");
                d.h2d += R.shp.n_res + (R.n + 1) * 8;
            }
            fill_params(e, d, p, r_begin, R.shp.n_res, R.n, min_hits, R.ap);
            if (packed) { R.ap.pk = p.pk; R.ap.pk_lead = (uint32_t)(r_begin - origin); }
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
        const unsigned long long total_recv = recv_off[nd];
        bool recv_ok = route_reserve(d, ln, 0, total_recv + 64) == KA_OK && ln.r_small;
        if (!recv_ok) { sh.abort.store(1); if (d.err == KA_OK) d.err = KA_ERR_OOM; }
        if (peer) {
            RouteShared::Pub& pb = sh.pub[idx];
            pb.recv[L] = ln.r_recv; pb.ans[L] = ln.r_ans_sorted;
            pb.scatter_done[L] = ln.ev_scatter; pb.lookup_done[L] = ln.ev_lookup; pb.tally_done[L] = ln.ev_tally;
        }
        // column sums: where my keys start inside every owner's receive buffer, where every sender's keys start in my send order
        unsigned long long recv_base[8] = {0}, their_send_off[8] = {0};
        for (int o = 0; o < nd; o++) {
            for (int s2 = 0; s2 < idx; s2++) recv_base[o] += sh.counts[s2][o];          // my region in owner o's receive buffer
            for (int o2 = 0; o2 < idx; o2++) their_send_off[o] += sh.counts[o][o2];     // send slot of sender o's first key for me
        }
        sh.bar.wait();                     // nobody overwrites counts before everyone has read them
        if (sh.abort.load()) {
            // a peer cannot receive: every device skips the exchange of this and all later rounds
            if (d.err == KA_OK) { d.err = KA_ERR_OOM; d.errmsg = "routed table mode: a peer device ran out of memory"; }
            recv_ok = false; R.has = false;
        }
        if (peer && !sh.abort.load()) {
            // ---- peer-store transport: three kernels per round, ordered across the GPUs by events ----
            // (an event must be RECORDED before another device waits on it, hence the host barriers)
            if (R.has) {
                RouteDst dst;
                for (int o = 0; o < 8; o++) dst.p[o] = nullptr;
                for (int o = 0; o < nd; o++) {
                    // the owner's lookup of the previous round on this lane must have drained its receive buffer
                    if (sh.pub[o].lookup_done[L]) cuda_ok(cudaStreamWaitEvent(st, sh.pub[o].lookup_done[L], 0), "wait lookup");
                    dst.p[o] = sh.pub[o].recv[L] + recv_base[o] - send_off[o];
                }
                cuda_ok(cudaMemcpyAsync(ln.r_small + 8, send_off, 64, cudaMemcpyHostToDevice, st), "H2D offsets");
                mark(3);
                cuda_ok(launch_route_scatter(ln.r_keys, R.shp.n_res, R.ap.tab, ln.r_small + 8, ln.r_small + 16, dst, ln.r_pos, st), "route scatter");
                mark(4);
                d.launches += 1;
            }
            cuda_ok(cudaEventRecord(ln.ev_scatter, st), "event");
            sh.bar.wait();                 // every scatter of this round is enqueued and its event recorded
            {
                RouteAns ra;
                ra.n_regions = (uint32_t)nd;
                for (int s2 = 0; s2 < 8; s2++) { ra.p[s2] = nullptr; ra.first[s2] = ~0ull; }
                for (int s2 = 0; s2 < nd; s2++) {
                    cuda_ok(cudaStreamWaitEvent(st, sh.pub[s2].scatter_done[L], 0), "wait scatter");
                    // the requester's tally of the previous round on this lane must have read its answers
                    if (sh.pub[s2].tally_done[L]) cuda_ok(cudaStreamWaitEvent(st, sh.pub[s2].tally_done[L], 0), "wait tally");
                    ra.first[s2] = recv_off[s2];
                    ra.p[s2] = sh.pub[s2].ans[L] + their_send_off[s2] - recv_off[s2];
                }
                TableView tab = e->geom;
                tab.sectors = d.table; tab.ovf = d.ovf; tab.my_shard = (uint32_t)idx;
                mark(5);
                if (total_recv) { cuda_ok(launch_route_lookup(ln.r_recv, total_recv, tab, nullptr, ra, st), "route lookup"); d.launches += 1; }
                mark(6);
            }
            cuda_ok(cudaEventRecord(ln.ev_lookup, st), "event");
            sh.bar.wait();                 // every lookup of this round is enqueued and its event recorded
            for (int o = 0; o < nd; o++) cuda_ok(cudaStreamWaitEvent(st, sh.pub[o].lookup_done[L], 0), "wait answers");
            recv_ok = false;               // (the NCCL exchange below is not used)
        } else if (R.has) {
            RouteDst dst;
            for (int o = 0; o < 8; o++) dst.p[o] = ln.r_send;
            cuda_ok(cudaMemcpyAsync(ln.r_small + 8, send_off, 64, cudaMemcpyHostToDevice, st), "H2D offsets");
            mark(3);
            cuda_ok(launch_route_scatter(ln.r_keys, R.shp.n_res, R.ap.tab, ln.r_small + 8, ln.r_small + 16, dst, ln.r_pos, st), "route scatter");
            mark(4);
            d.launches += 1;
        }
        if (peer && sh.abort.load()) { sh.bar.wait(); sh.bar.wait(); }   // keep the barrier count of the round
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
            tab.sectors = d.table; tab.ovf = d.ovf; tab.my_shard = (uint32_t)idx;
            mark(5);
            RouteAns none;
            none.n_regions = 0;
            cuda_ok(launch_route_lookup(ln.r_recv, total_recv, tab, ln.r_ans_recv, none, st), "route lookup");
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
            if (R.shp.n_long) {
                // sequences beyond the tile sizes: the long-sequence kernel probes the shards through NVLink peer loads
                cuda_ok(launch_big(R.ap, (int)std::min<uint64_t>(R.shp.n_long, (uint64_t)d.sm_count * 4), st), "long sequences");
                d.launches += 1;
            }
            if (peer) cuda_ok(cudaEventRecord(ln.ev_tally, st), "event");
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

}  // namespace kai
