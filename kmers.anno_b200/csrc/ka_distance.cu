// ka_distance.cu — pairwise protein k-mer distance on the GPU (sm_100a).
//
// Reference being replaced: genome/compare/GeneCopyProcessor.java:137-142 — for every target peg
//   ProteinKmers kmers = new ProteinKmers(feat.getProteinTranslation());            (:137)
//   for (Feature f2 : feats) { ... new ProteinKmers(f2...) ; kmers.distance(f2Kmers) } (:140-142)
// ProteinKmers (org.theseed.sequence, not in the repository; recalled, docs/SEMANTICS.md) is the
// HashSet of the distinct K-windows of a protein; similarity = |A ∩ B|, distance = 1 when the
// sets share nothing, else 1 - |A ∩ B| / (|A| + |B| - |A ∩ B|)  (Jaccard distance, double).
//
// Kernels
//   set_size_kernel   one CTA per sequence: the distinct K-windows go into an open-addressed hash set
//                     (shared memory; an L2-resident global slice for sequences too long for it),
//                     the number of successful inserts is |set|.
//                     It also marks ONE window of every distinct k-mer (the insert that created the
//                     entry), so that later passes can count distinct k-mers without a second set.
//   common_kernel     one CTA per query group: the query's set is built once; every warp then takes
//                     candidates on its own, streams their marked windows through the set and counts
//                     the hits = |A ∩ B|; lane 0 writes it and the distance.
// Integer work in shared memory; the only floating-point operation is the final IEEE double
// division / subtraction, which equals the Java expression bit for bit.
#include "ka_engine_internal.cuh"

namespace ka {

namespace {

constexpr int DTHREADS = 128;

struct SetRef {
    unsigned long long* keys;   // cap entries, 0 = empty (packed keys are never 0)
    uint32_t cap;               // power of two
};

__device__ __forceinline__ SetRef pick_set(const DistParams& p, unsigned char* smem, uint32_t windows,
                                           unsigned long long scratch_first) {
    SetRef s;
    s.cap = dist_set_cap(windows);
    if (s.cap <= p.smem_cap) {
        s.keys = reinterpret_cast<unsigned long long*>(smem);
    } else {
        s.keys = p.scratch_keys + scratch_first;
    }
    return s;
}

// f(w, key) for every K-window w of the W windows starting at residue r0, spread over `nthr`
// cooperating threads (`me` = this thread's index among them): each thread rolls the 5-bit pack
// over ONE contiguous run of windows (K - 1 + run byte loads instead of K per window).
template <typename F>
__device__ __forceinline__ void for_windows(const DistParams& p, const uint8_t* s_lut, const uint8_t* r0,
                                            unsigned long long W, uint32_t me, uint32_t nthr, F f) {
    const unsigned long long run = (W + nthr - 1) / nthr;
    const unsigned long long w0 = (unsigned long long)me * run;
    if (w0 >= W) return;
    const unsigned long long w1 = w0 + run < W ? w0 + run : W;
    const uint8_t* r = r0 + w0;
    unsigned long long key = 0;
    for (int j = 0; j < p.K - 1; j++) key = (key << 5) | s_lut[__ldg(r + j)];
    r += p.K - 1;
    for (unsigned long long w = w0; w < w1; w++, r++) {
        key = ((key << 5) | s_lut[__ldg(r)]) & p.key_mask;
        f(w, key);
    }
}

// insert: true = the key was not in the set yet
__device__ __forceinline__ bool set_insert(const SetRef& s, unsigned long long key) {
    uint32_t j = (uint32_t)mix64(key) & (s.cap - 1);
    for (;;) {
        const unsigned long long old = atomicCAS(s.keys + j, 0ull, key);
        if (old == 0ull) return true;
        if (old == key) return false;
        j = (j + 1) & (s.cap - 1);
    }
}

__device__ __forceinline__ bool set_has(const SetRef& s, unsigned long long key) {
    uint32_t j = (uint32_t)mix64(key) & (s.cap - 1);
    for (;;) {
        const unsigned long long cur = s.keys[j];
        if (cur == key) return true;
        if (cur == 0ull) return false;
        j = (j + 1) & (s.cap - 1);
    }
}

// fill the (cleared) set with the windows of sequence `seq`, the whole CTA cooperating; when
// `uniq` is given, uniq[window position] = 1 for exactly one window of every distinct k-mer
__device__ __forceinline__ int build_set(const DistParams& p, const SetRef& s, const uint8_t* s_lut,
                                         uint32_t seq, uint8_t* uniq) {
    const unsigned long long a = p.off[seq] - p.base, L = p.off[seq + 1] - p.off[seq];
    const unsigned long long W = L >= (unsigned long long)p.K ? L - p.K + 1 : 0;
    for (uint32_t i = threadIdx.x; i < s.cap; i += DTHREADS) s.keys[i] = 0ull;
    __syncthreads();
    int mine = 0;
    for_windows(p, s_lut, p.res + a, W, threadIdx.x, DTHREADS, [&](unsigned long long w, unsigned long long key) {
        const bool fresh = set_insert(s, key);
        mine += fresh ? 1 : 0;
        if (uniq) uniq[a + w] = fresh ? 1 : 0;
    });
    return mine;
}

}  // namespace

// |set| of every sequence, and the first-occurrence marks the pair kernel uses to count every
// distinct k-mer of a candidate once
__global__ void __launch_bounds__(DTHREADS) set_size_kernel(DistParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint8_t s_lut[256];
    __shared__ int s_red[DTHREADS / 32];
    for (int i = threadIdx.x; i < 256; i += DTHREADS) s_lut[i] = p.lut[i];
    __syncthreads();
    for (uint32_t seq = blockIdx.x; seq < p.n_seq; seq += gridDim.x) {
        const unsigned long long L = p.off[seq + 1] - p.off[seq];
        const uint32_t W = L >= (unsigned long long)p.K ? (uint32_t)(L - p.K + 1) : 0u;
        const SetRef s = pick_set(p, smem, W, p.seq_scratch[seq]);
        int n = build_set(p, s, s_lut, seq, p.uniq);
        n = __reduce_add_sync(0xffffffffu, n);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = n;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < DTHREADS / 32; w++) t += s_red[w];
            p.set_size[seq] = t;
        }
        __syncthreads();
    }
}

// One CTA per query: its set is built once in shared memory; then every WARP takes candidates of
// the group on its own — a lane rolls over a contiguous run of the candidate's windows and counts
// the marked (first-occurrence) ones found in the query's set — no barrier per candidate.
__global__ void __launch_bounds__(DTHREADS) common_kernel(DistParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint8_t s_lut[256];
    for (int i = threadIdx.x; i < 256; i += DTHREADS) s_lut[i] = p.lut[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t q = p.q_begin + blockIdx.x; q < p.q_end; q += gridDim.x) {
        const uint32_t qs = p.query_seq[q];
        const unsigned long long LA = p.off[qs + 1] - p.off[qs];
        const uint32_t WA = LA >= (unsigned long long)p.K ? (uint32_t)(LA - p.K + 1) : 0u;
        const SetRef s = pick_set(p, smem, WA, p.query_scratch[q]);
        build_set(p, s, s_lut, qs, nullptr);
        __syncthreads();
        const int size_a = p.set_size[qs];
        for (unsigned long long m = p.group_off[q] + warp; m < p.group_off[q + 1]; m += DTHREADS / 32) {
            const uint32_t cs = p.cand_seq[m];
            const unsigned long long b = p.off[cs] - p.base, LB = p.off[cs + 1] - p.off[cs];
            const unsigned long long WB = LB >= (unsigned long long)p.K ? LB - p.K + 1 : 0;
            int mine = 0;
            for_windows(p, s_lut, p.res + b, WB, lane, 32, [&](unsigned long long w, unsigned long long key) {
                if (p.uniq[b + w] && set_has(s, key)) mine++;
            });
            const int common = __reduce_add_sync(0xffffffffu, mine);
            if (lane == 0) {
                p.common[m] = common;
                double d = 1.0;                                   // nothing shared
                if (common > 0) {
                    const double sim = (double)common;
                    const double uni = (double)(size_a + p.set_size[cs]) - sim;
                    d = 1.0 - sim / uni;
                }
                p.dist[m] = d;
            }
        }
        __syncthreads();
    }
}

size_t dist_smem_bytes(uint32_t smem_cap) { return (size_t)smem_cap * 8; }

cudaError_t dist_set_smem(size_t bytes) {
    cudaError_t ce = cudaFuncSetAttribute(set_size_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(common_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return ce;
}

cudaError_t launch_set_size(const DistParams& p, int sm_count, cudaStream_t st) {
    if (p.n_seq == 0) return cudaSuccess;
    const unsigned grid = (unsigned)(p.n_seq < (uint32_t)sm_count * 16u ? p.n_seq : (uint32_t)sm_count * 16u);
    set_size_kernel<<<grid, DTHREADS, dist_smem_bytes(p.smem_cap), st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_common(const DistParams& p, int sm_count, cudaStream_t st) {
    if (p.q_end <= p.q_begin) return cudaSuccess;
    const uint32_t nq = p.q_end - p.q_begin;
    const unsigned grid = (unsigned)(nq < (uint32_t)sm_count * 16u ? nq : (uint32_t)sm_count * 16u);
    common_kernel<<<grid, DTHREADS, dist_smem_bytes(p.smem_cap), st>>>(p);
    return cudaGetLastError();
}

}  // namespace ka

// ======================================================================================
// C ABI
// ======================================================================================
using namespace ka;
using namespace kai;

extern "C" {

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

}  // extern "C"
