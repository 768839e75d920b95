// ka_kernels.cu — hand-written sm_100a kernels of the k-mer annotation hot path.
//
// Reference being replaced (paths under /root/reference/src/main/java/org/theseed/):
//   proteins/kmers/anno/ApplyKmerProcessor.java:122-148 — per peg: new ProteinKmers(prot)
//   (:123, the set of distinct K-windows), kmerRoleMap.get (:130), unanimous-role tally
//   (:131-144), thresholded call (:146-147); and :99-110 — the DB load (HashMap.put, last
//   line wins).
//
// Kernels
//   plan_kernel    first sequence of every residue tile (binary search on the CSR offsets)
//                  and the list of long sequences.
//   tile_kernel    one CTA per tile of whole sequences: the tile's residues are staged into
//                  shared memory with one TMA bulk copy (cp.async.bulk + mbarrier); each
//                  thread rolls the 5-bit packed key over 8 consecutive window positions,
//                  issues its 8 bucket loads (one 256-bit load = one 32-byte DRAM sector
//                  each) before consuming any, de-duplicates hitting k-mers of a sequence
//                  with a shared-memory token set (HashSet semantics of ProteinKmers), and
//                  reduces (count, min role, max role) per sequence with warp match/redux
//                  and shared-memory atomics; the epilogue applies unanimity + min_hits.
//   big_kernel     same per-position work for sequences too long for a tile's shared
//                  memory: one CTA per sequence, token set in an L2-resident scratch region.
//   db_insert      lock-free insert of the packed DB k-mers (atomicCAS on the key,
//                  atomicMax on (line, role) so the last line wins).
//
// HBM-bound integer work: no tensor cores.  Algorithmic bytes per probe = 32 (one bucket
// sector) + 1 (the residue).
#include "ka_kernels.cuh"

namespace ka {

// ------------------------------------------------------------------------------------
// table lookup
// ------------------------------------------------------------------------------------

// Continue a lookup past a full first bucket (rare at load factor <= 0.5).
__device__ __forceinline__ int lookup_overflow(const TableView& tab, unsigned long long key,
                                            unsigned long long b, uint32_t& slot) {
    for (;;) {
        b = (b + 1 == tab.n_buckets) ? 0 : b + 1;
        uint4 s0, s1;
        load_bucket(tab.buckets + 2 * b, s0, s1);
        unsigned long long k0 = u64_of(s0.x, s0.y), k1 = u64_of(s1.x, s1.y);
        if (k0 == key) { slot = (uint32_t)(2 * b); return (int)s0.z; }
        if (k1 == key) { slot = (uint32_t)(2 * b + 1); return (int)s1.z; }
        if (k1 == 0) return -1;  // slots fill in order: an empty slot ends the chain
    }
}

// Insert a hit token into the sequence's open-addressed de-dup region; true = first time.
template <bool SHARED>
__device__ __forceinline__ bool token_insert(uint32_t* region, uint32_t n, uint32_t token) {
    uint32_t j = (uint32_t)(((unsigned long long)(token * 0x9E3779B1u) * n) >> 32);
    for (;;) {
        uint32_t old = atomicCAS(region + j, TOKEN_EMPTY, token);
        if (old == TOKEN_EMPTY) return true;
        if (old == token) return false;
        j = (j + 1 == n) ? 0 : j + 1;
    }
}

// The reference's decision (ApplyKmerProcessor.java:146-147) from the reduced tally.
__device__ __forceinline__ void emit_call(const AnnotParams& p, uint32_t seq, int cnt, int rmin,
                                          int rmax) {
    int role = -1, hits = 0;
    uint8_t flag = 0;                                   // KA_FLAG_NONE
    if (cnt > 0) {
        if (rmin != rmax) { flag = 2; }                 // badPeg: two roles hit (:140-143)
        else if (cnt >= p.min_hits) { role = rmin; hits = cnt; flag = 1; }  // :146
        else { hits = cnt; flag = 3; }
    }
    p.out_role[seq] = role;
    p.out_hits[seq] = hits;
    if (p.out_flag) p.out_flag[seq] = flag;
}

// ------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------
__global__ void plan_kernel(AnnotParams p) {
    unsigned long long gid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid <= p.n_tiles) {
        // first[t] = first sequence i in [0, n_seq) with off[i] - base >= t * tile_span
        unsigned long long target = p.base + gid * (unsigned long long)p.tile_span;
        uint32_t lo = 0, hi = p.n_seq;
        while (lo < hi) {
            uint32_t mid = lo + ((hi - lo) >> 1);
            if (p.off[mid] < target) lo = mid + 1; else hi = mid;
        }
        p.first[gid] = lo;
    }
    if (gid < p.n_seq) {
        unsigned long long L = p.off[gid + 1] - p.off[gid];
        if (L > p.long_seq) {
            uint32_t idx = atomicAdd(p.big_count, 1u);
            unsigned long long tb = atomicAdd(p.tok_cursor, 2ull * L);
            BigItem it; it.seq = (uint32_t)gid; it.pad = 0; it.tok_base = tb;
            p.big_list[idx] = it;
        }
    }
}

cudaError_t launch_plan(const AnnotParams& p, cudaStream_t st) {
    unsigned long long n = (unsigned long long)p.n_tiles + 1;
    if (p.n_seq > n) n = p.n_seq;
    unsigned blocks = (unsigned)((n + 255) / 256);
    if (blocks == 0) blocks = 1;
    plan_kernel<<<blocks, 256, 0, st>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// tile kernel
// ------------------------------------------------------------------------------------
// dynamic shared memory layout (bytes):
//   [0, res_bytes)                       residue stage (16-byte aligned TMA destination)
//   [+256)                               residue -> code LUT
//   [+4*(MAX_TILE_SEQ+4))                s_off: sequence starts relative to the stage
//   [+3 * 4*MAX_TILE_SEQ)                s_cnt, s_min, s_max
//   [+4*(2*ext_max+4))                   token set
size_t tile_smem_bytes(uint32_t ext_max, uint32_t* res_bytes_out) {
    uint32_t res_bytes = (ext_max + 16 + 32 + 15) & ~15u;  // lead slack + K-1 over-read
    if (res_bytes_out) *res_bytes_out = res_bytes;
    return (size_t)res_bytes + 256 + 4 * (MAX_TILE_SEQ + 4) + 3 * 4 * MAX_TILE_SEQ +
           4 * (2 * (size_t)ext_max + 4);
}

template <int C>
__global__ void __launch_bounds__(TILE_THREADS) tile_kernel(AnnotParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_bar;

    uint8_t* s_res = smem_raw;
    uint8_t* s_lut = s_res + p.res_bytes;
    uint32_t* s_off = (uint32_t*)(s_lut + 256);
    int* s_cnt = (int*)(s_off + MAX_TILE_SEQ + 4);
    int* s_min = s_cnt + MAX_TILE_SEQ;
    int* s_max = s_min + MAX_TILE_SEQ;
    uint32_t* s_tok = (uint32_t*)(s_max + MAX_TILE_SEQ);

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31;
    const uint32_t t = blockIdx.x;
    const uint32_t s0 = p.first[t];
    uint32_t s1 = p.first[t + 1];
    if (s0 >= s1) return;
    // Only the last sequence starting in a tile can be a long one (long_seq >= tile_span
    // pushes the next start past the tile): leave it to big_kernel.
    if (p.off[s1] - p.off[s1 - 1] > p.long_seq) s1--;
    if (s0 >= s1) return;

    s_lut[tid] = p.lut[tid];  // TILE_THREADS == 256
    if (tid == 0) mbar_init(&s_bar, 1);
    __syncthreads();

    const int K = p.tab.K;
    const unsigned long long kmask = p.tab.key_mask;
    uint32_t parity = 0;

    for (uint32_t sb = s0; sb < s1; sb += MAX_TILE_SEQ) {
        const uint32_t ns = min((uint32_t)MAX_TILE_SEQ, s1 - sb);
        const unsigned long long g0 = p.off[sb] - p.base, g1 = p.off[sb + ns] - p.base;
        const unsigned long long g0a = g0 & ~15ull;
        const uint32_t lead = (uint32_t)(g0 - g0a);
        const uint32_t ext = (uint32_t)(g1 - g0a);           // stage-relative end of the residues
        const uint32_t nbytes = (ext + 15u) & ~15u;

        // stage the residues of sequences [sb, sb+ns) with one bulk copy
        if (tid == 0 && nbytes) {
            mbar_expect_tx(&s_bar, nbytes);
            bulk_g2s(s_res, p.res + g0a, nbytes, &s_bar);
        }
        for (uint32_t i = tid; i <= ns; i += TILE_THREADS)
            s_off[i] = (uint32_t)(p.off[sb + i] - p.base - g0a);
        for (uint32_t i = tid; i < ns; i += TILE_THREADS) {
            s_cnt[i] = 0; s_min[i] = 0x7fffffff; s_max[i] = -1;
        }
        {
            const uint32_t ntok = 2u * (ext - lead);
            const uint4 z = make_uint4(0, 0, 0, 0);
            for (uint32_t i = tid * 4; i < ntok; i += TILE_THREADS * 4)
                *reinterpret_cast<uint4*>(s_tok + i) = z;
        }
        __syncthreads();
        if (nbytes) { mbar_wait(&s_bar, parity); parity ^= 1; }

        for (uint32_t pb = 0; pb < ext; pb += TILE_THREADS * C) {
            const uint32_t P0 = pb + tid * C;
            int cur = -1, cnt = 0, mn = 0x7fffffff, mx = -1;
            if (P0 < ext) {
                // sequence containing P0: last i with s_off[i] <= P0 (-1: lead slack)
                int lo = 0, hi = (int)ns + 1;
                while (lo < hi) {
                    int mid = (lo + hi) >> 1;
                    if (s_off[mid] <= P0) lo = mid + 1; else hi = mid;
                }
                int si = lo - 1;

                // rolling 5-bit pack: warm up over K-1 residues, then one key per position
                const uint8_t* r = s_res + P0;
                unsigned long long key = 0;
                int vr = 0;  // consecutive residues inside the DB alphabet
                for (int j = 0; j < K - 1; j++) {
                    uint32_t c = s_lut[r[j]];
                    key = (key << 5) | c;
                    vr = c ? vr + 1 : 0;
                }
                r += K - 1;

                unsigned long long keys[C];
                uint32_t bkt[C];
                int seqi[C];
                uint4 v0[C], v1[C];
                unsigned okmask = 0;
#pragma unroll
                for (int i = 0; i < C; i++) {
                    uint32_t c = s_lut[r[i]];
                    key = (key << 5) | c;
                    vr = c ? vr + 1 : 0;
                    const uint32_t pos = P0 + i;
                    while (si < (int)ns && pos >= s_off[si + 1]) si++;
                    const bool ok = (si >= 0) && (si < (int)ns) && (pos + K <= s_off[si + 1]) &&
                                    (vr >= K);
                    keys[i] = key & kmask;
                    seqi[i] = si;
                    if (ok) {
                        okmask |= 1u << i;
                        const unsigned long long b = bucket_of(keys[i], p.tab.n_buckets);
                        bkt[i] = (uint32_t)b;
                        load_bucket(p.tab.buckets + 2 * b, v0[i], v1[i]);
                    }
                }
#pragma unroll
                for (int i = 0; i < C; i++) {
                    if (okmask & (1u << i)) {
                        const unsigned long long k0 = u64_of(v0[i].x, v0[i].y);
                        const unsigned long long k1 = u64_of(v1[i].x, v1[i].y);
                        int role = -1;
                        uint32_t slot = 0;
                        if (k0 == keys[i]) { role = (int)v0[i].z; slot = 2 * bkt[i]; }
                        else if (k1 == keys[i]) { role = (int)v1[i].z; slot = 2 * bkt[i] + 1; }
                        else if (k1 != 0) role = lookup_overflow(p.tab, keys[i], bkt[i], slot);
                        if (role >= 0) {
                            const int q = seqi[i];
                            const uint32_t a = s_off[q], b = s_off[q + 1];
                            if (token_insert<true>(s_tok + 2 * (a - lead), 2 * (b - a), slot + 1)) {
                                if (q != cur) {
                                    if (cur >= 0 && cnt > 0) {
                                        atomicAdd(&s_cnt[cur], cnt);
                                        atomicMin(&s_min[cur], mn);
                                        atomicMax(&s_max[cur], mx);
                                    }
                                    cur = q; cnt = 0; mn = 0x7fffffff; mx = -1;
                                }
                                cnt++;
                                mn = min(mn, role);
                                mx = max(mx, role);
                            }
                        }
                    }
                }
            }
            // segmented reduction over the lanes that ended on the same sequence
            const unsigned grp = __match_any_sync(0xffffffffu, cur);
            const int tot = __reduce_add_sync(grp, cnt);
            const int gmin = __reduce_min_sync(grp, mn);
            const int gmax = __reduce_max_sync(grp, mx);
            if (cur >= 0 && tot > 0 && lane == (uint32_t)(__ffs(grp) - 1)) {
                atomicAdd(&s_cnt[cur], tot);
                atomicMin(&s_min[cur], gmin);
                atomicMax(&s_max[cur], gmax);
            }
        }
        __syncthreads();
        for (uint32_t i = tid; i < ns; i += TILE_THREADS)
            emit_call(p, sb + i, s_cnt[i], s_min[i], s_max[i]);
        __syncthreads();
    }
}

cudaError_t tile_kernel_set_smem(size_t bytes) {
    return cudaFuncSetAttribute(tile_kernel<POS_PER_THREAD>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

cudaError_t launch_tiles(const AnnotParams& p, size_t smem, cudaStream_t st) {
    if (p.n_tiles == 0) return cudaSuccess;
    tile_kernel<POS_PER_THREAD><<<p.n_tiles, TILE_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// long sequences: one CTA per sequence, token set in global scratch (L2 resident)
// ------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(TILE_THREADS) big_kernel(AnnotParams p) {
    __shared__ uint8_t s_lut[256];
    __shared__ int sh_cnt, sh_min, sh_max;
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    s_lut[tid] = p.lut[tid];
    const uint32_t nbig = *p.big_count;
    const int K = p.tab.K;
    const unsigned long long kmask = p.tab.key_mask;

    for (uint32_t bi = blockIdx.x; bi < nbig; bi += gridDim.x) {
        const BigItem it = p.big_list[bi];
        const unsigned long long a = p.off[it.seq] - p.base;
        const unsigned long long L = p.off[it.seq + 1] - p.off[it.seq];
        const unsigned long long W = L - (unsigned long long)K + 1;  // L > long_seq >= K
        uint32_t* region = p.scratch + it.tok_base;
        const uint32_t nreg = (uint32_t)(2 * L);
        for (uint32_t i = tid; i < nreg; i += TILE_THREADS) region[i] = TOKEN_EMPTY;
        if (tid == 0) { sh_cnt = 0; sh_min = 0x7fffffff; sh_max = -1; }
        __syncthreads();

        int cnt = 0, mn = 0x7fffffff, mx = -1;
        for (unsigned long long pb = 0; pb < W; pb += TILE_THREADS * C) {
            const unsigned long long P0 = pb + (unsigned long long)tid * C;
            if (P0 >= W) continue;
            const uint8_t* r = p.res + a + P0;
            unsigned long long key = 0;
            int vr = 0;
            for (int j = 0; j < K - 1; j++) {
                uint32_t c = s_lut[__ldg(r + j)];
                key = (key << 5) | c;
                vr = c ? vr + 1 : 0;
            }
            r += K - 1;
            unsigned long long keys[C];
            uint32_t bkt[C];
            uint4 v0[C], v1[C];
            unsigned okmask = 0;
#pragma unroll
            for (int i = 0; i < C; i++) {
                if (P0 + i < W) {
                    uint32_t c = s_lut[__ldg(r + i)];
                    key = (key << 5) | c;
                    vr = c ? vr + 1 : 0;
                    keys[i] = key & kmask;
                    if (vr >= K) {
                        okmask |= 1u << i;
                        const unsigned long long b = bucket_of(keys[i], p.tab.n_buckets);
                        bkt[i] = (uint32_t)b;
                        load_bucket(p.tab.buckets + 2 * b, v0[i], v1[i]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < C; i++) {
                if (okmask & (1u << i)) {
                    const unsigned long long k0 = u64_of(v0[i].x, v0[i].y);
                    const unsigned long long k1 = u64_of(v1[i].x, v1[i].y);
                    int role = -1;
                    uint32_t slot = 0;
                    if (k0 == keys[i]) { role = (int)v0[i].z; slot = 2 * bkt[i]; }
                    else if (k1 == keys[i]) { role = (int)v1[i].z; slot = 2 * bkt[i] + 1; }
                    else if (k1 != 0) role = lookup_overflow(p.tab, keys[i], bkt[i], slot);
                    if (role >= 0 && token_insert<false>(region, nreg, slot + 1)) {
                        cnt++;
                        mn = min(mn, role);
                        mx = max(mx, role);
                    }
                }
            }
        }
        const int tot = __reduce_add_sync(0xffffffffu, cnt);
        const int gmin = __reduce_min_sync(0xffffffffu, mn);
        const int gmax = __reduce_max_sync(0xffffffffu, mx);
        if (lane == 0 && tot > 0) {
            atomicAdd(&sh_cnt, tot);
            atomicMin(&sh_min, gmin);
            atomicMax(&sh_max, gmax);
        }
        __syncthreads();
        if (tid == 0) emit_call(p, it.seq, sh_cnt, sh_min, sh_max);
        __syncthreads();
    }
}

cudaError_t launch_big(const AnnotParams& p, int grid, cudaStream_t st) {
    big_kernel<POS_PER_THREAD><<<grid, TILE_THREADS, 0, st>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// DB build
// ------------------------------------------------------------------------------------
__global__ void alphabet_scan_kernel(const uint8_t* __restrict__ bytes, unsigned long long n,
                                     uint32_t* bitmap8) {
    __shared__ uint32_t sb[8];
    if (threadIdx.x < 8) sb[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += stride) {
        const uint32_t b = bytes[i];
        const uint32_t bit = 1u << (b & 31);
        if (!(sb[b >> 5] & bit)) atomicOr(&sb[b >> 5], bit);
    }
    __syncthreads();
    if (threadIdx.x < 8 && sb[threadIdx.x]) atomicOr(&bitmap8[threadIdx.x], sb[threadIdx.x]);
}

cudaError_t launch_alphabet_scan(const uint8_t* bytes, unsigned long long n, uint32_t* bitmap8,
                                 cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    unsigned long long want = (n + 255) / 256;
    unsigned blocks = (unsigned)(want < 148ull * 16 ? want : 148ull * 16);
    alphabet_scan_kernel<<<blocks, 256, 0, st>>>(bytes, n, bitmap8);
    return cudaGetLastError();
}

// HashMap.put for every DB line (ApplyKmerProcessor.java:106): a duplicate k-mer keeps one
// slot, and atomicMax over (line index << 32 | role) leaves the LAST line's role in it.
__global__ void db_insert_kernel(const uint8_t* __restrict__ kmers,
                                 const int32_t* __restrict__ roles, unsigned long long n,
                                 unsigned long long line_base, int K,
                                 const uint8_t* __restrict__ lut, Slot* table,
                                 unsigned long long n_buckets, unsigned long long* counters,
                                 uint32_t* errs) {
    __shared__ uint8_t s_lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = lut[i];
    __syncthreads();
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    uint32_t longest = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += stride) {
        const uint8_t* km = kmers + i * (unsigned long long)K;
        unsigned long long key = 0;
        bool bad = false;
        for (int j = 0; j < K; j++) {
            const uint32_t c = s_lut[km[j]];
            bad |= (c == 0);
            key = (key << 5) | c;
        }
        const int32_t role = roles[i];
        if (bad) { atomicAdd(&errs[0], 1u); continue; }
        if (role < 0) { atomicAdd(&errs[1], 1u); continue; }
        const unsigned long long val = ((line_base + i) << 32) | (uint32_t)role;
        unsigned long long b = bucket_of(key, n_buckets);
        uint32_t chain = 1;
        for (;;) {
            Slot* s = table + 2 * b;
            bool done = false;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if (done) break;
                unsigned long long old = atomicCAS(&s[h].key, 0ull, key);
                if (old == 0ull || old == key) {
                    if (old == 0ull) atomicAdd(&counters[0], 1ull);
                    atomicMax(&s[h].val, val);
                    done = true;
                }
            }
            if (done) break;
            b = (b + 1 == n_buckets) ? 0 : b + 1;
            chain++;
        }
        longest = max(longest, chain);
    }
    longest = __reduce_max_sync(__activemask(), longest);
    if ((threadIdx.x & 31) == 0) atomicMax(&counters[1], (unsigned long long)longest);
}

cudaError_t launch_db_insert(const uint8_t* kmers, const int32_t* roles, unsigned long long n,
                             unsigned long long line_base, int K, const uint8_t* lut, Slot* table,
                             unsigned long long n_buckets, unsigned long long* counters,
                             uint32_t* errs, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    unsigned long long want = (n + 255) / 256;
    unsigned blocks = (unsigned)(want < 148ull * 32 ? want : 148ull * 32);
    db_insert_kernel<<<blocks, 256, 0, st>>>(kmers, roles, n, line_base, K, lut, table, n_buckets,
                                             counters, errs);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// random-probe roofline microbenchmark
// ------------------------------------------------------------------------------------
template <int BYTES>
__global__ void __launch_bounds__(256) random_probe_kernel(const uint4* __restrict__ buf,
                                                           unsigned long long n_slots,
                                                           unsigned long long n_probes,
                                                           unsigned long long seed,
                                                           unsigned long long* sink) {
    constexpr int U = 8;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x * U;
    uint32_t acc = 0;
    for (unsigned long long i0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * U;
         i0 < n_probes; i0 += stride) {
        uint4 a[U], b[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const unsigned long long idx = __umul64hi(mix64(seed + i0 + u + 1), n_slots);
            if (BYTES == 32) load_bucket(buf + 2 * idx, a[u], b[u]);
            else {
                asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                    : "=r"(a[u].x), "=r"(a[u].y), "=r"(a[u].z), "=r"(a[u].w)
                    : "l"(buf + idx));
                b[u] = make_uint4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) acc ^= a[u].x ^ a[u].w ^ b[u].y ^ b[u].z;
    }
    if (acc == 0x9e3779b9u) atomicAdd(sink, 1ull);
}

cudaError_t launch_random_probe(const uint4* buf, unsigned long long n_slots, int slot_bytes,
                                unsigned long long n_probes, unsigned long long seed,
                                unsigned long long* sink, cudaStream_t st) {
    unsigned long long want = (n_probes + 256ull * 8 - 1) / (256ull * 8);
    unsigned blocks = (unsigned)(want < 148ull * 64 ? want : 148ull * 64);
    if (blocks == 0) blocks = 1;
    if (slot_bytes == 32)
        random_probe_kernel<32><<<blocks, 256, 0, st>>>(buf, n_slots, n_probes, seed, sink);
    else
        random_probe_kernel<16><<<blocks, 256, 0, st>>>(buf, n_slots, n_probes, seed, sink);
    return cudaGetLastError();
}

__global__ void fill_random_kernel(uint4* buf, unsigned long long n) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += stride) {
        unsigned long long h = mix64(i + 0x1234567ull);
        buf[i] = make_uint4((uint32_t)h, (uint32_t)(h >> 32), (uint32_t)i, (uint32_t)(i >> 32));
    }
}

cudaError_t launch_fill_random(uint4* buf, unsigned long long n_uint4, cudaStream_t st) {
    fill_random_kernel<<<148 * 16, 256, 0, st>>>(buf, n_uint4);
    return cudaGetLastError();
}

}  // namespace ka
