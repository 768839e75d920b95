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
#include "ka_kernels.cuh"

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
