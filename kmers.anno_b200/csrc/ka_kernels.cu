// ka_kernels.cu — hand-written sm_100a kernels of the k-mer annotation hot path.
//
// Reference being replaced (paths under /root/reference/src/main/java/org/theseed/):
//   proteins/kmers/anno/ApplyKmerProcessor.java:122-148 — per peg: new ProteinKmers(prot)
//   (:123, the set of distinct K-windows), kmerRoleMap.get (:130), unanimous-role tally
//   (:131-144), thresholded call (:146-147); and :99-110 — the DB load (HashMap.put, last
//   line wins).
//
// Kernels
//   plan_kernel    one 16-byte descriptor per residue tile (binary searches on the CSR offsets),
//                  single-sequence tiles for mid-length sequences, the list of long sequences.
//   tile_kernel    one CTA per tile of whole sequences: the tile's residues are staged into
//                  shared memory with one TMA bulk copy (cp.async.bulk + mbarrier); each
//                  thread rolls the 5-bit packed key over a run of consecutive window
//                  positions, issues all its sector loads (one 256-bit load each = one DRAM
//                  line) before consuming any, de-duplicates hitting k-mers of a sequence
//                  with a shared-memory token set (HashSet semantics of ProteinKmers), and
//                  reduces (count, min role, max role) per sequence with warp match/redux
//                  and shared-memory atomics; the epilogue applies unanimity + min_hits.
//   big_kernel     same per-position work for sequences too long for a tile's shared
//                  memory: one CTA per sequence, token set in an L2-resident scratch region.
//   db_insert      lock-free insert of the packed DB k-mers (atomicCAS claims a slot,
//                  atomicMax on the db line index makes the last line win), db_finalize
//                  writes the winning role into every slot.
//
// HBM-bound integer work: no tensor cores.  Algorithmic bytes per probe = 32 (one table
// sector) + 1 (the residue).
#include "ka_kernels.cuh"
#include <type_traits>

// `make DEBUG=1` compiles bounds checks into the kernels (compute-sanitizer is not available on
// the GPU pool): a failed check ORs its code into p.dbg[0] and the engine turns that into an
// error (tests/ run once under this build: no check fires).
#ifdef KA_DEBUG
#define KA_CHECK(cond, code) do { if (!(cond)) atomicOr(p.dbg, (code)); } while (0)
#else
#define KA_CHECK(cond, code) do { } while (0)
#endif

namespace ka {

// ------------------------------------------------------------------------------------
// table lookup: C independent probes per thread, all loads in flight before the first use
// ------------------------------------------------------------------------------------
// Overflow table (cls 32/64): whole mixed keys, 2 x Slot128 per sector, linear chaining.
__device__ __forceinline__ int ovf_lookup(const TableView& t, unsigned long long m, uint32_t& tok) {
    const uint32_t mask = (1u << t.ovf_bbits) - 1;
    uint32_t s = t.ovf_bbits ? (uint32_t)(mix64(m) >> (64 - t.ovf_bbits)) : 0u;
    // the overflow entries of a sector live with the shard that owns the sector
    const uint32_t shard = t.n_shards <= 1 ? 0u : shard_of(t, m);
    const uint4* ovf = t.n_shards <= 1 ? t.ovf : reinterpret_cast<const uint4*>(
        __ldg(reinterpret_cast<const unsigned long long*>(t.shard_ovf) + shard));
    const uint32_t tok0 = t.n_primary_slots + shard * (2u << t.ovf_bbits);
    for (;;) {
        uint4 a, b;
        load_sector(ovf + 2 * (size_t)s, a, b);
        const unsigned long long k0 = u64_of(a.x, a.y), k1 = u64_of(b.x, b.y);
        if (k0 == m) { tok = tok0 + 2 * s + 1; return (int)a.z; }
        if (k1 == m) { tok = tok0 + 2 * s + 2; return (int)b.z; }
        if (k1 == 0) return -1;
        s = (s + 1) & mask;
    }
}

// On return role[i] >= 0 marks a hit and sec[i] holds its de-dup token (global slot index + 1).
template <int CLS, int C>
__device__ __forceinline__ void probe_batch(const TableView& tab,
                                            const typename rem_type<CLS>::type (&rem)[C],
                                            uint32_t (&sec)[C], unsigned okmask, int (&role)[C]) {
    constexpr int S = slots_per_sector<CLS>();
    uint4 a[C], b[C];
#pragma unroll
    for (int i = 0; i < C; i++)
        if (okmask & (1u << i)) load_sector(sector_ptr(tab, sec[i]), a[i], b[i]);
    unsigned pend = 0;
#pragma unroll
    for (int i = 0; i < C; i++) {
        role[i] = -1;
        if (okmask & (1u << i)) {
            uint32_t j = 0;
            bool full;
            const int r = match_sector<CLS>(tab, a[i], b[i], rem[i], j, full);
            if (r >= 0) { role[i] = r; sec[i] = sec[i] * S + j + 1; }
            else if (full) pend |= 1u << i;
        }
    }
    // Rare (the load factor keeps full sectors near 1 %): the home sector has no free slot
    // and no match.
    if (CLS == 128) {
        // whole keys are stored: follow the chain, all pending positions advance together
        const uint32_t sec_mask = (1u << tab.bbits) - 1;
        while (pend) {
#pragma unroll
            for (int i = 0; i < C; i++)
                if (pend & (1u << i)) {
                    sec[i] = (sec[i] + 1) & sec_mask;
                    load_sector(sector_ptr(tab, sec[i]), a[i], b[i]);
                }
#pragma unroll
            for (int i = 0; i < C; i++)
                if (pend & (1u << i)) {
                    uint32_t j = 0;
                    bool full;
                    const int r = match_sector<CLS>(tab, a[i], b[i], rem[i], j, full);
                    if (r >= 0) { role[i] = r; sec[i] = sec[i] * S + j + 1; pend &= ~(1u << i); }
                    else if (!full) pend &= ~(1u << i);
                }
        }
    } else if (pend) {
        // the key, if present, is in the overflow table under its whole mixed value
#pragma unroll
        for (int i = 0; i < C; i++)
            if (pend & (1u << i))
                role[i] = ovf_lookup(tab, ((unsigned long long)sec[i] << tab.rem_bits) | rem[i], sec[i]);
    }
}

// Insert a hit token into the sequence's open-addressed de-dup region; true = first time.
__device__ __forceinline__ bool token_insert(uint32_t* region, uint32_t n, uint32_t token) {
    uint32_t j = (uint32_t)(((unsigned long long)(token * 0x9E3779B1u) * n) >> 32);
    for (;;) {
        uint32_t old = atomicCAS(region + j, TOKEN_EMPTY, token);
        if (old == TOKEN_EMPTY) return true;
        if (old == token) return false;
        j = (j + 1 == n) ? 0 : j + 1;
    }
}

// Wide tables: the de-dup token is the mixed key itself (64 bits), same open addressing.
__device__ __forceinline__ bool token_insert(unsigned long long* region, uint32_t n, unsigned long long token) {
    uint32_t j = (uint32_t)(((unsigned long long)(((uint32_t)token ^ (uint32_t)(token >> 32)) * 0x9E3779B1u) * n) >> 32);
    for (;;) {
        unsigned long long old = atomicCAS(region + j, 0ull, token);
        if (old == 0ull) return true;
        if (old == token) return false;
        j = (j + 1 == n) ? 0 : j + 1;
    }
}

// probe_batch for wide tables (cls 32/64): m[i] is the whole mixed key (sector = m >> rem_bits,
// which may need more than 32 bits); the key is its own token, so only the roles come back.
template <int CLS, int C>
__device__ __forceinline__ void probe_batch_wide(const TableView& tab, const unsigned long long (&m)[C],
                                                 unsigned okmask, int (&role)[C]) {
    uint4 a[C], b[C];
#pragma unroll
    for (int i = 0; i < C; i++)
        if (okmask & (1u << i)) load_sector(sector_ptr(tab, m[i] >> tab.rem_bits), a[i], b[i]);
    unsigned pend = 0;
#pragma unroll
    for (int i = 0; i < C; i++) {
        role[i] = -1;
        if (okmask & (1u << i)) {
            uint32_t j = 0;
            bool full;
            const int r = match_sector<CLS>(tab, a[i], b[i], m[i] & tab.rem_mask, j, full);
            if (r >= 0) role[i] = r;
            else if (full) pend |= 1u << i;
        }
    }
    if (pend) {
#pragma unroll
        for (int i = 0; i < C; i++)
            if (pend & (1u << i)) { uint32_t unused; role[i] = ovf_lookup(tab, m[i], unused); }
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
__device__ __forceinline__ uint32_t first_seq_at(const AnnotParams& p, unsigned long long target) {
    // first sequence i in [0, n_seq) with off[i] >= target
    uint32_t lo = 0, hi = p.n_seq;
    while (lo < hi) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (p.off[mid] < target) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// One descriptor per tile so that the tile kernel starts with a single load:
// {first sequence, number of sequences (a trailing long one removed), residue range [g0, g1)
// relative to the chunk}.  Also the list of long sequences.
__global__ void plan_kernel(AnnotParams p) {
    unsigned long long gid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < p.n_tiles) {
        const uint32_t s0 = first_seq_at(p, p.base + gid * (unsigned long long)p.tile_span);
        uint32_t s1 = first_seq_at(p, p.base + (gid + 1) * (unsigned long long)p.tile_span);
        // Only the last sequence starting in a tile can be a long one (long_seq >= tile_span
        // pushes the next start past the tile): it is left to big_kernel.
        if (s1 > s0 && p.off[s1] - p.off[s1 - 1] > p.long_seq) s1--;
        uint4 d;
        d.x = s0; d.y = s1 - s0;
        d.z = (uint32_t)(p.off[s0] - p.base);
        d.w = (uint32_t)(p.off[s1] - p.base);
        p.first[gid] = d;
    }
    if (gid < p.n_seq) {
        unsigned long long L = p.off[gid + 1] - p.off[gid];
        if (L > p.mid_seq) {
            uint32_t idx = atomicAdd(p.big_count, 1u);
            unsigned long long tb = atomicAdd(p.tok_cursor, 2ull * L);
            BigItem it; it.seq = (uint32_t)gid; it.pad = 0; it.tok_base = tb;
            p.big_list[idx] = it;
        } else if (L > p.long_seq) {
            // a tile of its own, run by the second tile launch (larger shared-memory configuration)
            uint32_t idx = atomicAdd(p.mid_count, 1u);
            uint4 d;
            d.x = (uint32_t)gid; d.y = 1;
            d.z = (uint32_t)(p.off[gid] - p.base);
            d.w = (uint32_t)(p.off[gid + 1] - p.base);
            p.mid_desc[idx] = d;
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
//   [+4*(tok_cap(ext_max)+4*MAX_TILE_SEQ+8))  token set: sequence q starting at residue a owns
//                                        [tok_cap(a)+4q, +tok_cap(L)+4), never full (hits <= L-K+1)
// (tokens are 8 bytes each for wide tables)
size_t tile_smem_bytes(uint32_t ext_max, uint32_t* res_bytes_out, bool wide) {
    uint32_t res_bytes = (ext_max + 16 + 32 + 32 + 15) & ~15u;  // lead slack + K-1 over-read + in-place unpack slack
    if (res_bytes_out) *res_bytes_out = res_bytes;
    return (size_t)res_bytes + 256 + 4 * (MAX_TILE_SEQ + 4) + 3 * 4 * MAX_TILE_SEQ +
           (wide ? 8 : 4) * ((size_t)tok_cap(ext_max) + 4 * MAX_TILE_SEQ + 8);
}

// PACKED: the chunk's residues are the 5-bit code stream of ka_annotate_packed (p.pk); the tile's slice of it is
// bulk-copied to the END of the residue stage and expanded in place to one code byte per residue, so the
// rest of the kernel is unchanged (the LUT then maps digit d to field code d + 1, and 31 to 0).
template <int CLS, int C, int THREADS, int MINB, int MODE = 0, bool WIDE = false, bool PACKED = false>
__global__ void __launch_bounds__(THREADS, MINB) tile_kernel(AnnotParams p) {
    static_assert(!WIDE || CLS != 128, "wide tables are quotiented");
    static_assert(!PACKED || !(MODE == 0 && WIDE), "packed staging: narrow probe kernels and the routed extract / tally kernels");
    typedef typename std::conditional<WIDE, unsigned long long, uint32_t>::type tok_t;
    typedef typename std::conditional<WIDE, unsigned long long, typename rem_type<CLS>::type>::type rem_t;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_bar;

    uint8_t* s_res = smem_raw;
    uint8_t* s_lut = s_res + p.res_bytes;
    uint32_t* s_off = (uint32_t*)(s_lut + 256);
    int* s_cnt = (int*)(s_off + MAX_TILE_SEQ + 4);
    int* s_min = s_cnt + MAX_TILE_SEQ;
    int* s_max = s_min + MAX_TILE_SEQ;
    tok_t* s_tok = (tok_t*)(s_max + MAX_TILE_SEQ);   // 16-byte aligned: res_bytes is, and so is the rest

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31;
    // descriptor and LUT byte are independent loads: both in flight before the first use
    const uint4 desc = p.first[blockIdx.x];
    uint8_t lut_byte[(256 + THREADS - 1) / THREADS];
#pragma unroll
    for (int i = 0; i < (256 + THREADS - 1) / THREADS; i++) {
        const uint32_t b = (tid + i * THREADS) & 255;
        lut_byte[i] = PACKED ? (uint8_t)(b < 31 ? b + 1 : 0) : p.lut[b];
    }
    const uint32_t s0 = desc.x, s1 = desc.x + desc.y;
    if (desc.y == 0) return;

#pragma unroll
    for (int i = 0; i < (256 + THREADS - 1) / THREADS; i++)
        if (tid + i * THREADS < 256) s_lut[tid + i * THREADS] = lut_byte[i];
    if (tid == 0) mbar_init(&s_bar, 1);
    __syncthreads();

    const TableView tab = p.tab;
    const int K = tab.K;
    uint32_t parity = 0;

    for (uint32_t sb = s0; sb < s1; sb += MAX_TILE_SEQ) {
        const uint32_t ns = min((uint32_t)MAX_TILE_SEQ, s1 - sb);
        // the common case (one sub-batch) needs no further global load to know its residue range
        const unsigned long long g0 = (sb == s0) ? desc.z : p.off[sb] - p.base;
        const unsigned long long g1 = (sb + ns == s1) ? desc.w : p.off[sb + ns] - p.base;
        const unsigned long long g0a = PACKED ? g0 : (g0 & ~15ull);
        const uint32_t lead = (uint32_t)(g0 - g0a);
        const uint32_t ext = (uint32_t)(g1 - g0a);           // stage-relative end of the residues
        // packed form: bytes [bt, bt + nbytes) of the chunk's code stream hold the tile; they land at pk_dst
        const unsigned long long pbit0 = 5ull * (g0 + p.pk_lead);
        const unsigned long long bt = (pbit0 >> 3) & ~15ull;
        const uint32_t leadbits = (uint32_t)(pbit0 - 8ull * bt);
        const uint32_t nbytes = PACKED ? (uint32_t)((((5ull * (g1 + p.pk_lead) + 7ull) >> 3) - bt + 15ull) & ~15ull) : ((ext + 15u) & ~15u);
        const uint32_t pk_dst = PACKED ? ((p.res_bytes - nbytes - 16u) & ~15u) : 0u;
        KA_CHECK((PACKED ? ext + 64u : nbytes + 32u) <= p.res_bytes, 1u);        // residue stage holds the tile
        KA_CHECK(!PACKED || 3u * ((ext + 7u) >> 3) <= pk_dst + 24u, 64u);        // in-place expansion never overtakes the unread codes
        KA_CHECK(tok_cap(ext - lead) + 4u * ns + 8u <= tok_cap(p.ext_max) + 4u * MAX_TILE_SEQ + 8u, 2u);  // token set too

        // stage the residues of sequences [sb, sb+ns) with one bulk copy
        if (tid == 0 && nbytes) {
            mbar_expect_tx(&s_bar, nbytes);
            if (PACKED) bulk_g2s(s_res + pk_dst, reinterpret_cast<const unsigned char*>(p.pk) + bt, nbytes, &s_bar);
            else bulk_g2s(s_res, p.res + g0a, nbytes, &s_bar);
        }
        for (uint32_t i = tid; i <= ns; i += THREADS)
            s_off[i] = (uint32_t)(p.off[sb + i] - p.base - g0a);
        for (uint32_t i = tid; i < ns; i += THREADS) {
            s_cnt[i] = 0; s_min[i] = 0x7fffffff; s_max[i] = -1;
        }
        {
            const uint32_t ntok = (tok_cap(ext - lead) + 4u * ns + 4u) * (uint32_t)(sizeof(tok_t) / 4);   // in 32-bit words
            const uint4 z = make_uint4(0, 0, 0, 0);
            for (uint32_t i = tid * 4; i < ntok; i += THREADS * 4)
                *reinterpret_cast<uint4*>(reinterpret_cast<uint32_t*>(s_tok) + i) = z;
        }
        __syncthreads();
        if (nbytes) { mbar_wait(&s_bar, parity); parity ^= 1; }
        if (PACKED) {
            // 8 residues = 40 bits per thread and round: read, barrier, write one byte per code.  Round r writes
            // below 8 (r+1) THREADS, the codes still unread start at pk_dst + 5 (r+1) THREADS: disjoint (check 64).
            const uint32_t n_groups = (ext + 7u) >> 3;
            const uint32_t* pw = reinterpret_cast<const uint32_t*>(s_res + pk_dst);
            for (uint32_t g8 = 0; g8 < n_groups; g8 += THREADS) {
                const uint32_t g = g8 + tid;
                uint32_t lo = 0, hi = 0;
                if (g < n_groups) {
                    const uint32_t bit = leadbits + 40u * g;
                    const uint32_t* w = pw + (bit >> 5);
                    const uint32_t sh = bit & 31u, w0 = w[0], w1 = w[1], w2 = w[2];
                    const uint32_t u0 = __funnelshift_r(w0, w1, sh), u1 = __funnelshift_r(w1, w2, sh);
                    lo = (u0 & 31u) | ((u0 >> 5) & 31u) << 8 | ((u0 >> 10) & 31u) << 16 | ((u0 >> 15) & 31u) << 24;
                    hi = ((u0 >> 20) & 31u) | ((u0 >> 25) & 31u) << 8 | (__funnelshift_r(u0, u1, 30) & 31u) << 16 | ((u1 >> 3) & 31u) << 24;
                }
                __syncthreads();
                if (g < n_groups) *reinterpret_cast<uint2*>(s_res + 8u * g) = make_uint2(lo, hi);
            }
            __syncthreads();
        }

        // passes of up to THREADS*C positions, spread evenly over the threads
        for (uint32_t pb = lead; pb < ext; pb += THREADS * C) {
            const uint32_t pend = min(ext, pb + THREADS * C);
            const uint32_t run = (pend - pb + THREADS - 1) / THREADS;   // positions per thread, <= C
            const uint32_t P0 = pb + tid * run;
            int cur = -1, cnt = 0, mn = 0x7fffffff, mx = -1;
            if (P0 < pend) {
                // sequence containing P0: last i with s_off[i] <= P0
                int lo = 0, hi = (int)ns + 1;
                while (lo < hi) {
                    int mid = (lo + hi) >> 1;
                    if (s_off[mid] <= P0) lo = mid + 1; else hi = mid;
                }
                int si = lo - 1;                       // >= 0: P0 >= lead = s_off[0]
                KA_CHECK(si >= 0 && si < (int)ns, 32u);
                uint32_t nb = s_off[si + 1];           // end of the sequence holding the current position

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

                rem_t rem[C];           // stored remainder; wide tables: the whole mixed key
                uint32_t sec[C];        // home sector, then the de-dup token (narrow tables only)
                uint32_t seqpack = 0;   // sequence index (< MAX_TILE_SEQ = 256) of each position, one byte each
                static_assert(C <= 4 || MAX_TILE_SEQ <= 256, "");
                uint32_t seqpack2 = 0;
                unsigned okmask = 0;
#pragma unroll
                for (int i = 0; i < C; i++) {
                    const uint32_t pos = P0 + i;
                    if (i < (int)run && pos < pend) {
                        uint32_t c = s_lut[r[i]];
                        key = (key << 5) | c;
                        vr = c ? vr + 1 : 0;
                        while (pos >= nb) { si++; KA_CHECK(si < (int)ns, 4u); nb = s_off[si + 1]; }   // pos < ext = s_off[ns]: si stays < ns
                        if (pos + K <= nb && vr >= K) {
                            okmask |= 1u << i;
                            if (i < 4) seqpack |= (uint32_t)si << (8 * (i & 3));
                            else seqpack2 |= (uint32_t)si << (8 * (i & 3));
                            if (WIDE) {
                                rem[i] = mixw(key & tab.key_mask, tab.wbits, tab.key_mask);
                            } else if (MODE != 2) {
                                unsigned long long rm;
                                locate(tab, key & tab.key_mask, sec[i], rm);
                                rem[i] = (rem_t)rm;
                            }
                        }
                    }
                }
                int role[C];
                if (MODE == 1) {
                    // routed mode, step 1: only publish the mixed key of every position of this run
                    // (chunk-relative residue index = stage origin + position)
#pragma unroll
                    for (int i = 0; i < C; i++)
                        if (i < (int)run && P0 + i < pend)
                            p.route_keys[g0a + P0 + i] = !(okmask & (1u << i)) ? ROUTE_INVALID
                                : (WIDE ? (unsigned long long)rem[i] : (((unsigned long long)sec[i] << tab.rem_bits) | (unsigned long long)rem[i]));
                    okmask = 0;
                } else if (MODE == 2) {
                    // routed mode, last step: the owning GPUs have answered (role << 32 | token); the answer of
                    // a position sits at the send slot its key was bucketed into (route_slot)
                    unsigned long long ans[C];
#pragma unroll
                    for (int i = 0; i < C; i++)
                        if (okmask & (1u << i)) ans[i] = __ldg(p.route_ans + __ldg(p.route_slot + g0a + P0 + i));
#pragma unroll
                    for (int i = 0; i < C; i++) {
                        role[i] = -1;
                        if (okmask & (1u << i)) { role[i] = (int)(ans[i] >> 32); sec[i] = (uint32_t)ans[i]; }
                    }
                } else if constexpr (WIDE) {
                    probe_batch_wide<CLS, C>(tab, rem, okmask, role);
                } else {
                    probe_batch<CLS, C>(tab, rem, sec, okmask, role);
                }
#pragma unroll
                for (int i = 0; i < C; i++) {
                    if ((okmask & (1u << i)) && role[i] >= 0) {
                        tok_t token;
                        if constexpr (WIDE) token = rem[i]; else token = sec[i];
                        const int q = (int)(((i < 4 ? seqpack : seqpack2) >> (8 * (i & 3))) & 0xffu);
                        const uint32_t a = s_off[q], b = s_off[q + 1];
                        KA_CHECK(q >= 0 && q < (int)ns && b >= a && a >= lead, 8u);
                        KA_CHECK(tok_cap(a - lead) + 4u * (uint32_t)q + tok_cap(b - a) + 4u <= tok_cap(p.ext_max) + 4u * MAX_TILE_SEQ + 8u, 16u);
                        if (token_insert(s_tok + tok_cap(a - lead) + 4u * (uint32_t)q, tok_cap(b - a) + 4u, token)) {
                            if (q != cur) {
                                if (cur >= 0 && cnt > 0) {
                                    atomicAdd(&s_cnt[cur], cnt);
                                    atomicMin(&s_min[cur], mn);
                                    atomicMax(&s_max[cur], mx);
                                }
                                cur = q; cnt = 0; mn = 0x7fffffff; mx = -1;
                            }
                            cnt++;
                            mn = min(mn, role[i]);
                            mx = max(mx, role[i]);
                        }
                    }
                }
            }
            // segmented reduction over the lanes that ended on the same sequence (skipped when
            // no lane of the warp has a new hit to report)
            if (__any_sync(0xffffffffu, cnt > 0)) {
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
        }
        __syncthreads();
        if (MODE != 1)
            for (uint32_t i = tid; i < ns; i += THREADS)
                emit_call(p, sb + i, s_cnt[i], s_min[i], s_max[i]);
        __syncthreads();
    }
}

template <typename F>
static auto with_tile_kernel(int cls, int variant, bool packed, F f) {
#define KA_VARIANTS(CLS)                                                                  \
    if (packed) {                                                                         \
        switch (variant) {                                                                \
            case 1: return f(tile_kernel<CLS, 4, 256, 3, 0, false, true>, 256);           \
            default: return f(tile_kernel<CLS, 4, 128, 7, 0, false, true>, 128);          \
        }                                                                                 \
    }                                                                                     \
    switch (variant) {                                                                    \
        case 1: return f(tile_kernel<CLS, 4, 256, 3>, 256);                               \
        default: return f(tile_kernel<CLS, 4, 128, 6>, 128);                              \
    }
    if (cls == 32) { KA_VARIANTS(32) }
    if (cls == 64) { KA_VARIANTS(64) }
    KA_VARIANTS(128)
#undef KA_VARIANTS
}

cudaError_t tile_kernel_set_smem(int cls, int variant, size_t bytes) {
    cudaError_t ce = cudaSuccess;
    for (bool packed : {false, true}) {
        if (ce != cudaSuccess) break;
        ce = with_tile_kernel(cls, variant, packed, [&](auto kern, int) {
            cudaError_t c2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
            if (c2 != cudaSuccess) return c2;
            // ask for the largest shared-memory carve-out so that as many CTAs as the registers allow fit
            return cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        });
    }
    return ce;
}

cudaError_t launch_tiles(const AnnotParams& p, int variant, size_t smem, cudaStream_t st) {
    if (p.n_tiles == 0) return cudaSuccess;
    if (p.tab.wide) return launch_tiles_mode(p, variant, 0, smem, st);
    return with_tile_kernel(p.tab.cls, variant, p.pk != nullptr, [&](auto kern, int threads) {
        kern<<<p.n_tiles, threads, smem, st>>>(p);
        return cudaGetLastError();
    });
}

cudaError_t launch_tiles_mode(const AnnotParams& p, int variant, int mode, size_t smem, cudaStream_t st) {
    if (mode == 0 && !p.tab.wide) return launch_tiles(p, variant, smem, st);
    if (p.n_tiles == 0) return cudaSuccess;
    if (p.tab.cls == 128) return cudaErrorInvalidValue;   // routed and wide tables are quotiented
    if (variant > 1) variant = 1;                          // shapes 0 (4 x 128) and 1 (4 x 256)
    const bool packed = p.pk != nullptr && mode != 0;      // routed extract / tally kernels stage the 5-bit stream
#define KA_MODE_LAUNCH(CLS, W)                                                                          \
    if (variant == 0) {                                                                                 \
        if (mode == 0)      tile_kernel<CLS, 4, 128, 6, 0, W><<<p.n_tiles, 128, smem, st>>>(p);          \
        else if (mode == 1 && packed) tile_kernel<CLS, 4, 128, 6, 1, W, true><<<p.n_tiles, 128, smem, st>>>(p);  \
        else if (mode == 1) tile_kernel<CLS, 4, 128, 6, 1, W><<<p.n_tiles, 128, smem, st>>>(p);          \
        else if (packed)    tile_kernel<CLS, 4, 128, 6, 2, W, true><<<p.n_tiles, 128, smem, st>>>(p);    \
        else                tile_kernel<CLS, 4, 128, 6, 2, W><<<p.n_tiles, 128, smem, st>>>(p);          \
    } else {                                                                                            \
        if (mode == 0)      tile_kernel<CLS, 4, 256, 3, 0, W><<<p.n_tiles, 256, smem, st>>>(p);          \
        else if (mode == 1 && packed) tile_kernel<CLS, 4, 256, 3, 1, W, true><<<p.n_tiles, 256, smem, st>>>(p);  \
        else if (mode == 1) tile_kernel<CLS, 4, 256, 3, 1, W><<<p.n_tiles, 256, smem, st>>>(p);          \
        else if (packed)    tile_kernel<CLS, 4, 256, 3, 2, W, true><<<p.n_tiles, 256, smem, st>>>(p);    \
        else                tile_kernel<CLS, 4, 256, 3, 2, W><<<p.n_tiles, 256, smem, st>>>(p);          \
    }
    if (p.tab.wide) { if (p.tab.cls == 32) { KA_MODE_LAUNCH(32, true) } else { KA_MODE_LAUNCH(64, true) } }
    else            { if (p.tab.cls == 32) { KA_MODE_LAUNCH(32, false) } else { KA_MODE_LAUNCH(64, false) } }
#undef KA_MODE_LAUNCH
    return cudaGetLastError();
}

cudaError_t tile_kernel_mode_set_smem(size_t bytes) {
    cudaError_t ce = cudaSuccess;
#define KA_MODE_ATTR(K)                                                                                   \
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); \
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(K, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
#define KA_MODE_ATTR6(CLS, W)                                                                   \
    KA_MODE_ATTR((tile_kernel<CLS, 4, 128, 6, 0, W>)) KA_MODE_ATTR((tile_kernel<CLS, 4, 256, 3, 0, W>))   \
    KA_MODE_ATTR((tile_kernel<CLS, 4, 128, 6, 1, W>)) KA_MODE_ATTR((tile_kernel<CLS, 4, 256, 3, 1, W>))   \
    KA_MODE_ATTR((tile_kernel<CLS, 4, 128, 6, 2, W>)) KA_MODE_ATTR((tile_kernel<CLS, 4, 256, 3, 2, W>))   \
    KA_MODE_ATTR((tile_kernel<CLS, 4, 128, 6, 1, W, true>)) KA_MODE_ATTR((tile_kernel<CLS, 4, 256, 3, 1, W, true>))   \
    KA_MODE_ATTR((tile_kernel<CLS, 4, 128, 6, 2, W, true>)) KA_MODE_ATTR((tile_kernel<CLS, 4, 256, 3, 2, W, true>))
    KA_MODE_ATTR6(32, false) KA_MODE_ATTR6(64, false) KA_MODE_ATTR6(32, true) KA_MODE_ATTR6(64, true)
#undef KA_MODE_ATTR6
#undef KA_MODE_ATTR
    return ce;
}

// ------------------------------------------------------------------------------------
// routed sharded table (table_mode 2): all-to-all of keys, owners probe, answers come back
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t route_owner(const TableView& t, unsigned long long m) {
    return shard_of(t, m);
}

__global__ void route_count_kernel(const unsigned long long* __restrict__ keys, unsigned long long n,
                                   TableView tab, unsigned long long* counts) {
    __shared__ unsigned long long sh[8];
    if (threadIdx.x < 8) sh[threadIdx.x] = 0;
    __syncthreads();
    uint32_t mine[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long m = keys[i];
        const uint32_t o = m == ROUTE_INVALID ? 0xffffffffu : route_owner(tab, m);
#pragma unroll
        for (int k = 0; k < 8; k++) mine[k] += (o == (uint32_t)k);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t w = __reduce_add_sync(0xffffffffu, mine[k]);
        if ((threadIdx.x & 31) == 0 && w) atomicAdd(&sh[k], (unsigned long long)w);
    }
    __syncthreads();
    if (threadIdx.x < 8 && sh[threadIdx.x]) atomicAdd(&counts[threadIdx.x], sh[threadIdx.x]);
}

cudaError_t launch_route_count(const unsigned long long* keys, unsigned long long n, TableView tab,
                               unsigned long long* counts, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    route_count_kernel<<<148 * 8, 256, 0, st>>>(keys, n, tab, counts);
    return cudaGetLastError();
}

// One CTA buckets a tile of 256 x 16 keys.  __match_any_sync groups the lanes of a warp by owner in one
// instruction whatever the number of shards: the lowest lane of a group adds the group's size to the warp's
// shared-memory count of that owner; ONE global atomicAdd per owner and CTA reserves the output ranges.  The keys
// are first ordered by owner in shared memory and then written out linearly, so that consecutive threads store
// consecutive addresses: whole 128-byte segments per owner instead of 32-byte pieces — what matters when the
// destination is another GPU's memory behind NVLink (table_mode 3).
__global__ void __launch_bounds__(256) route_scatter_kernel(const unsigned long long* __restrict__ keys, unsigned long long n,
                                     TableView tab, const unsigned long long* __restrict__ offsets,
                                     unsigned long long* cursor, RouteDst dst, uint32_t* slot_of_pos) {
    constexpr int PER = 16;                       // keys per thread, strided by 32 inside the warp's part of the tile
    __shared__ unsigned long long s_keys[256 * PER];   // the tile's valid keys, ordered by owner
    __shared__ uint32_t s_cnt[8][8];              // [warp][owner]
    __shared__ uint32_t s_next[8][8];             // [warp][owner] next index into s_keys
    __shared__ uint32_t s_start[9];               // first index of every owner in s_keys
    __shared__ unsigned long long s_gbase[8];     // first send slot of every owner for this CTA
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const unsigned long long tile = 256ull * PER;
    for (unsigned long long t0 = (unsigned long long)blockIdx.x * tile; t0 < n; t0 += (unsigned long long)gridDim.x * tile) {
        unsigned long long m[PER];
        uint32_t own[PER];
        if (lane < 8) s_cnt[warp][lane] = 0;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < PER; k++) {
            // a warp reads 32 consecutive keys at a time: ranks follow the position order inside a warp
            const unsigned long long i = t0 + (unsigned long long)warp * (32 * PER) + k * 32 + lane;
            m[k] = i < n ? keys[i] : ROUTE_INVALID;
            own[k] = m[k] == ROUTE_INVALID ? 0xffffffffu : route_owner(tab, m[k]);
            const unsigned grp = __match_any_sync(0xffffffffu, own[k]);
            if (own[k] != 0xffffffffu && (grp & lt) == 0u) atomicAdd(&s_cnt[warp][own[k]], (uint32_t)__popc(grp));
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t run = 0;
            for (uint32_t o = 0; o < 8; o++) {
                s_start[o] = run;
                for (int w = 0; w < 8; w++) { s_next[w][o] = run; run += s_cnt[w][o]; }
            }
            s_start[8] = run;
        }
        __syncthreads();
        if (threadIdx.x < 8) {
            const uint32_t o = threadIdx.x, tot = s_start[o + 1] - s_start[o];
            s_gbase[o] = (tot ? atomicAdd(&cursor[o], (unsigned long long)tot) : 0ull) + offsets[o];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const unsigned long long i = t0 + (unsigned long long)warp * (32 * PER) + k * 32 + lane;
            const unsigned grp = __match_any_sync(0xffffffffu, own[k]);
            const int leader = __ffs(grp) - 1;
            uint32_t first = 0;
            if (own[k] != 0xffffffffu && (int)lane == leader) {
                first = s_next[warp][own[k]];
                s_next[warp][own[k]] = first + (uint32_t)__popc(grp);
            }
            first = __shfl_sync(grp, first, leader);
            if (own[k] != 0xffffffffu) {
                const uint32_t si = first + (uint32_t)__popc(grp & lt);
                s_keys[si] = m[k];
                // answers come back in send order: position i reads send slot (first slot of its owner + rank in the CTA)
                slot_of_pos[i] = (uint32_t)(s_gbase[own[k]] + (si - s_start[own[k]]));
            }
            __syncwarp();
        }
        __syncthreads();
        const uint32_t total = s_start[8];
        for (uint32_t idx = threadIdx.x; idx < total; idx += 256) {
            uint32_t o = 0;
#pragma unroll
            for (uint32_t q = 1; q < 8; q++) o += (idx >= s_start[q]) ? 1u : 0u;
            // the local send buffer, or straight into the owner's receive buffer (NVLink store)
            dst.p[o][s_gbase[o] + (idx - s_start[o])] = s_keys[idx];
        }
        __syncthreads();
    }
}

cudaError_t launch_route_scatter(const unsigned long long* keys, unsigned long long n, TableView tab,
                                 const unsigned long long* offsets, unsigned long long* cursor,
                                 const RouteDst& dst, uint32_t* slot_of_pos, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    unsigned long long want = (n + 256ull * 16 - 1) / (256ull * 16);
    route_scatter_kernel<<<(unsigned)(want < 148ull * 8 ? want : 148ull * 8), 256, 0, st>>>(keys, n, tab, offsets, cursor, dst, slot_of_pos);
    return cudaGetLastError();
}

// every received key belongs to this device's shard: one sector load from local HBM each
template <int CLS>
__global__ void __launch_bounds__(256) route_lookup_kernel(const unsigned long long* __restrict__ keys,
                                                           unsigned long long n, TableView tab,
                                                           unsigned long long* ans, RouteAns ra) {
    constexpr int S = slots_per_sector<CLS>();
    constexpr int U = 4;
    typedef typename rem_type<CLS>::type rem_t;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long local_mask = (1ull << tab.shard_shift) - 1;
    for (unsigned long long i0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += stride * U) {
        uint4 a[U], b[U];
        unsigned long long sec[U];
        rem_t rem[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const unsigned long long i = i0 + u * stride;
            if (i < n) {
                const unsigned long long m = keys[i];
                sec[u] = m >> tab.rem_bits;
                rem[u] = (rem_t)(m & tab.rem_mask);
                load_sector(tab.sectors + 2 * (size_t)(sec[u] & local_mask), a[u], b[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const unsigned long long i = i0 + u * stride;
            if (i < n) {
                uint32_t j = 0, tok = 0;
                bool full;
                int role = match_sector<CLS>(tab, a[u], b[u], rem[u], j, full);
                // (narrow tables: the token is the global slot index + 1; wide tables ignore it, the
                // requester uses the key itself)
                if (role >= 0) tok = (uint32_t)sec[u] * S + j + 1;
                else if (full) {
                    // the overflow entries of a sector live on its owner: local as well
                    const unsigned long long m = (sec[u] << tab.rem_bits) | (unsigned long long)rem[u];
                    const uint32_t mask = (1u << tab.ovf_bbits) - 1;
                    uint32_t s = tab.ovf_bbits ? (uint32_t)(mix64(m) >> (64 - tab.ovf_bbits)) : 0u;
                    const uint32_t tok0 = tab.n_primary_slots + tab.my_shard * (2u << tab.ovf_bbits);
                    for (;;) {
                        uint4 xa, xb;
                        load_sector(tab.ovf + 2 * (size_t)s, xa, xb);
                        const unsigned long long k0 = u64_of(xa.x, xa.y), k1 = u64_of(xb.x, xb.y);
                        if (k0 == m) { tok = tok0 + 2 * s + 1; role = (int)xa.z; break; }
                        if (k1 == m) { tok = tok0 + 2 * s + 2; role = (int)xb.z; break; }
                        if (k1 == 0) break;
                        s = (s + 1) & mask;
                    }
                }
                const unsigned long long v = role >= 0 ? (((unsigned long long)(uint32_t)role << 32) | tok) : ROUTE_MISS;
                if (ra.n_regions == 0) ans[i] = v;
                else {
                    // key i came from the sender whose region holds i: its answer goes straight into that GPU's
                    // answer buffer, in its send order (NVLink store)
                    uint32_t sdr = 0;
#pragma unroll
                    for (uint32_t q = 1; q < 8; q++) sdr += (q < ra.n_regions && i >= ra.first[q]) ? 1u : 0u;
                    ra.p[sdr][i] = v;
                }
            }
        }
    }
}

cudaError_t launch_route_lookup(const unsigned long long* keys, unsigned long long n, TableView tab,
                                unsigned long long* ans, const RouteAns& ra, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    unsigned long long want = (n + 256ull * 4 - 1) / (256ull * 4);
    unsigned blocks = (unsigned)(want < 148ull * 16 ? want : 148ull * 16);
    if (tab.cls == 32) route_lookup_kernel<32><<<blocks, 256, 0, st>>>(keys, n, tab, ans, ra);
    else if (tab.cls == 64) route_lookup_kernel<64><<<blocks, 256, 0, st>>>(keys, n, tab, ans, ra);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// long sequences: one CTA per sequence, token set in global scratch (L2 resident)
// ------------------------------------------------------------------------------------
template <int CLS, int C, bool WIDE>
__global__ void __launch_bounds__(256) big_kernel(AnnotParams p) {
    constexpr int THREADS = 256;
    typedef typename std::conditional<WIDE, unsigned long long, uint32_t>::type tok_t;
    typedef typename std::conditional<WIDE, unsigned long long, typename rem_type<CLS>::type>::type rem_t;
    __shared__ uint8_t s_lut[256];
    __shared__ int sh_cnt, sh_min, sh_max;
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    s_lut[tid] = p.lut[tid];
    const uint32_t nbig = *p.big_count;
    const TableView tab = p.tab;
    const int K = tab.K;

    for (uint32_t bi = blockIdx.x; bi < nbig; bi += gridDim.x) {
        const BigItem it = p.big_list[bi];
        const unsigned long long a = p.off[it.seq] - p.base;
        const unsigned long long L = p.off[it.seq + 1] - p.off[it.seq];
        const unsigned long long W = L - (unsigned long long)K + 1;  // L > long_seq >= K
        tok_t* region = reinterpret_cast<tok_t*>(p.scratch) + it.tok_base;
        const uint32_t nreg = (uint32_t)(2 * L);
        for (uint32_t i = tid; i < nreg; i += THREADS) region[i] = 0;
        if (tid == 0) { sh_cnt = 0; sh_min = 0x7fffffff; sh_max = -1; }
        __syncthreads();

        int cnt = 0, mn = 0x7fffffff, mx = -1;
        for (unsigned long long pb = 0; pb < W; pb += THREADS * C) {
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
            rem_t rem[C];
            uint32_t sec[C];
            unsigned okmask = 0;
#pragma unroll
            for (int i = 0; i < C; i++) {
                if (P0 + i < W) {
                    uint32_t c = s_lut[__ldg(r + i)];
                    key = (key << 5) | c;
                    vr = c ? vr + 1 : 0;
                    if (vr >= K) {
                        okmask |= 1u << i;
                        if (WIDE) {
                            rem[i] = mixw(key & tab.key_mask, tab.wbits, tab.key_mask);
                        } else {
                            unsigned long long rm;
                            locate(tab, key & tab.key_mask, sec[i], rm);
                            rem[i] = (rem_t)rm;
                        }
                    }
                }
            }
            int role[C];
            if constexpr (WIDE) probe_batch_wide<CLS, C>(tab, rem, okmask, role);
            else probe_batch<CLS, C>(tab, rem, sec, okmask, role);
#pragma unroll
            for (int i = 0; i < C; i++) {
                tok_t token;
                if constexpr (WIDE) token = rem[i]; else token = sec[i];
                if ((okmask & (1u << i)) && role[i] >= 0 && token_insert(region, nreg, token)) {
                    cnt++;
                    mn = min(mn, role[i]);
                    mx = max(mx, role[i]);
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
    if (p.tab.wide) {
        if (p.tab.cls == 32) big_kernel<32, 8, true><<<grid, 256, 0, st>>>(p);
        else if (p.tab.cls == 64) big_kernel<64, 8, true><<<grid, 256, 0, st>>>(p);
        else return cudaErrorInvalidValue;
    } else if (p.tab.cls == 32) big_kernel<32, 8, false><<<grid, 256, 0, st>>>(p);
    else if (p.tab.cls == 64) big_kernel<64, 8, false><<<grid, 256, 0, st>>>(p);
    else big_kernel<128, 8, false><<<grid, 256, 0, st>>>(p);
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

// Synthetic DB lines for the oversized-table configuration (ka_db_load_synthetic): line i is the
// K-mer whose j-th residue is AA[(x >> 5j) & 31 mod 20], x = mix64(seed + i * golden), with role
// i mod n_roles.  The host can regenerate any line (kmers.anno_b200/synth.py: synthetic_db_lines).
__global__ void db_generate_kernel(unsigned long long first, unsigned long long n, int K,
                                   unsigned long long seed, uint32_t n_roles, uint8_t* kmers, int32_t* roles) {
    const char* AA = "ACDEFGHIKLMNPQRSTVWY";
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
        const unsigned long long i = first + t;
        unsigned long long x = mix64(seed + i * 0x9E3779B97F4A7C15ull);
        for (int j = 0; j < K; j++) { kmers[t * K + j] = (uint8_t)AA[(x & 31u) % 20u]; x >>= 5; }
        roles[t] = (int32_t)(i % n_roles);
    }
}

cudaError_t launch_db_generate(unsigned long long first, unsigned long long n, int K, unsigned long long seed,
                               uint32_t n_roles, uint8_t* kmers, int32_t* roles, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    db_generate_kernel<<<148 * 16, 256, 0, st>>>(first, n, K, seed, n_roles, kmers, roles);
    return cudaGetLastError();
}

// HashMap.put for every DB line (ApplyKmerProcessor.java:106): a duplicate k-mer keeps one
// slot and the LAST line's role ends up in it: atomicMax on best[slot] = (line + 1) << role_bits | role.
template <int CLS>
__global__ void db_insert_kernel(TableView tab, const uint8_t* __restrict__ kmers,
                                 const int32_t* __restrict__ roles, unsigned long long n,
                                 unsigned long long line_base, const uint8_t* __restrict__ lut,
                                 unsigned long long* best, uint32_t role_bits,
                                 unsigned long long* counters, uint32_t* errs) {
    constexpr int S = slots_per_sector<CLS>();
    __shared__ uint8_t s_lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = lut[i];
    __syncthreads();
    const int K = tab.K;
    const unsigned long long sec_mask = (1ull << tab.bbits) - 1;
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
        if (bad) { atomicAdd(&errs[0], 1u); continue; }
        if (roles[i] < 0) { atomicAdd(&errs[1], 1u); continue; }
        unsigned long long sec, rem;     // 64-bit sector index: wide tables have more than 2^32 sectors
        if (CLS == 128) {
            uint32_t s32;
            locate(tab, key, s32, rem);
            sec = s32;
        } else {
            const unsigned long long m = mixw(key, tab.wbits, tab.key_mask);
            sec = m >> tab.rem_bits;
            rem = m & tab.rem_mask;
        }
        const unsigned long long msec = sec;   // global sector (the overflow key keeps it)
        if (tab.n_shards > 1) {
            // sharded table: this device stores only the sectors of its own shard, at local indices.
            // (cls 128 chains to the NEXT sector, which may belong to another shard: the engine
            // only shards the quotiented classes, whose overflow stays with the home shard.)
            if ((sec >> tab.shard_shift) != tab.my_shard) continue;
            sec &= (1ull << tab.shard_shift) - 1;
        }
        const unsigned long long mine = ((line_base + i + 1) << role_bits) | (unsigned long long)(uint32_t)roles[i];
        uint32_t chain = 1;
        bool done = false;
        if (CLS == 128) {
            while (!done) {
                Slot128* s = reinterpret_cast<Slot128*>(const_cast<uint4*>(tab.sectors)) + 2 * (size_t)sec;
                for (int h = 0; h < 2 && !done; h++) {
                    unsigned long long old = atomicCAS(&s[h].key, 0ull, key);
                    if (old == 0ull || old == key) {
                        if (old == 0ull) atomicAdd(&counters[0], 1ull);
                        atomicMax(&s[h].val, mine);
                        done = true;
                    }
                }
                if (!done) { sec = (sec + 1) & sec_mask; chain++; }
            }
        } else {
            if (CLS == 64) {
                unsigned long long* w = reinterpret_cast<unsigned long long*>(const_cast<uint4*>(tab.sectors)) + (size_t)sec * S;
                const unsigned long long claim = rem | (1ull << tab.rem_bits);  // role field = placeholder
                for (int h = 0; h < S && !done; h++) {
                    unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(w + h);
                    if (cur == 0) cur = atomicCAS(w + h, 0ull, claim);
                    if (cur == 0 || ((cur ^ rem) & tab.rem_mask) == 0) {
                        if (cur == 0) atomicAdd(&counters[0], 1ull);
                        atomicMax(&best[(size_t)sec * S + h], mine);
                        done = true;
                    }
                }
            } else {
                uint32_t* w = reinterpret_cast<uint32_t*>(const_cast<uint4*>(tab.sectors)) + (size_t)sec * S;
                const uint32_t r32 = (uint32_t)rem, m32 = (uint32_t)tab.rem_mask;
                const uint32_t claim = r32 | (1u << tab.rem_bits);
                for (int h = 0; h < S && !done; h++) {
                    uint32_t cur = *reinterpret_cast<volatile uint32_t*>(w + h);
                    if (cur == 0) cur = atomicCAS(w + h, 0u, claim);
                    if (cur == 0 || ((cur ^ r32) & m32) == 0) {
                        if (cur == 0) atomicAdd(&counters[0], 1ull);
                        atomicMax(&best[(size_t)sec * S + h], mine);
                        done = true;
                    }
                }
            }
            if (!done) {
                // home sector full: the whole mixed key goes to the overflow table
                const unsigned long long m = (msec << tab.rem_bits) | rem;
                const uint32_t omask = (1u << tab.ovf_bbits) - 1;
                uint32_t os = tab.ovf_bbits ? (uint32_t)(mix64(m) >> (64 - tab.ovf_bbits)) : 0u;
                Slot128* ovf = reinterpret_cast<Slot128*>(const_cast<uint4*>(tab.ovf));
                while (!done && chain <= omask + 2) {
                    Slot128* s = ovf + 2 * (size_t)os;
                    for (int h = 0; h < 2 && !done; h++) {
                        unsigned long long old = atomicCAS(&s[h].key, 0ull, m);
                        if (old == 0ull || old == m) {
                            if (old == 0ull) atomicAdd(&counters[0], 1ull);
                            atomicMax(&s[h].val, mine);
                            done = true;
                        }
                    }
                    if (!done) { os = (os + 1) & omask; chain++; }
                }
                if (!done) atomicAdd(&errs[2], 1u);  // overflow table full: host rebuilds it larger
            }
        }
        longest = max(longest, chain);
    }
    if (longest) atomicMax(&counters[1], (unsigned long long)longest);
}

cudaError_t launch_db_insert(const TableView& tab, const uint8_t* kmers, const int32_t* roles,
                             unsigned long long n, unsigned long long line_base,
                             const uint8_t* lut, unsigned long long* best, uint32_t role_bits,
                             unsigned long long* counters, uint32_t* errs, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    unsigned long long want = (n + 255) / 256;
    unsigned blocks = (unsigned)(want < 148ull * 32 ? want : 148ull * 32);
    if (tab.cls == 32)
        db_insert_kernel<32><<<blocks, 256, 0, st>>>(tab, kmers, roles, n, line_base, lut, best, role_bits, counters, errs);
    else if (tab.cls == 64)
        db_insert_kernel<64><<<blocks, 256, 0, st>>>(tab, kmers, roles, n, line_base, lut, best, role_bits, counters, errs);
    else
        db_insert_kernel<128><<<blocks, 256, 0, st>>>(tab, kmers, roles, n, line_base, lut, best, role_bits, counters, errs);
    return cudaGetLastError();
}

// Write role+1 of the winning db line into the role field of every occupied slot (cls 32/64).
template <int CLS>
__global__ void db_finalize_kernel(TableView tab, const unsigned long long* __restrict__ best, uint32_t role_bits) {
    const unsigned long long role_mask = (1ull << role_bits) - 1;
    constexpr int S = slots_per_sector<CLS>();
    const unsigned long long n_slots = (unsigned long long)S << (tab.n_shards > 1 ? tab.shard_shift : tab.bbits);
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long s = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; s < n_slots;
         s += stride) {
        if (CLS == 64) {
            unsigned long long* w = reinterpret_cast<unsigned long long*>(const_cast<uint4*>(tab.sectors)) + s;
            const unsigned long long v = *w;
            if (v) *w = (v & tab.rem_mask) | (((best[s] & role_mask) + 1) << tab.rem_bits);
        } else {
            uint32_t* w = reinterpret_cast<uint32_t*>(const_cast<uint4*>(tab.sectors)) + s;
            const uint32_t v = *w;
            if (v) *w = (v & (uint32_t)tab.rem_mask) | ((uint32_t)((best[s] & role_mask) + 1) << tab.rem_bits);
        }
    }
}

// whole-key slots (cls 128 table, overflow table): val = (line + 1) << role_bits | role  ->  role
__global__ void db_finalize_slot128_kernel(Slot128* slots, unsigned long long n_slots, uint32_t role_bits) {
    const unsigned long long role_mask = (1ull << role_bits) - 1;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long s = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += stride)
        if (slots[s].key) slots[s].val &= role_mask;
}

cudaError_t launch_db_finalize(const TableView& tab, const unsigned long long* best, uint32_t role_bits, cudaStream_t st) {
    if (tab.cls == 128) {
        db_finalize_slot128_kernel<<<148 * 16, 256, 0, st>>>(reinterpret_cast<Slot128*>(const_cast<uint4*>(tab.sectors)),
                                                             2ull << tab.bbits, role_bits);
        return cudaGetLastError();
    }
    if (tab.cls == 32) db_finalize_kernel<32><<<148 * 16, 256, 0, st>>>(tab, best, role_bits);
    else db_finalize_kernel<64><<<148 * 16, 256, 0, st>>>(tab, best, role_bits);
    if (tab.ovf)
        db_finalize_slot128_kernel<<<148 * 4, 256, 0, st>>>(reinterpret_cast<Slot128*>(const_cast<uint4*>(tab.ovf)),
                                                            2ull << tab.ovf_bbits, role_bits);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// build: BuildKmerProcessor.java:138-223 on the device
// ------------------------------------------------------------------------------------
// One thread per residue position (grid-stride); the peg of a position is found by binary
// search on the offsets.  The per-peg HashSet of ProteinKmers is irrelevant here: counting
// a window twice changes neither "all the same role" nor "occurs in a zero-role peg".
template <int PASS>
__global__ void build_pass_kernel(const uint8_t* __restrict__ res, const unsigned long long* __restrict__ off,
                                  uint32_t n_seq, const int32_t* __restrict__ n_roles,
                                  const int32_t* __restrict__ peg_role, int K, const uint8_t* __restrict__ lut,
                                  Slot128* table, unsigned long long n_slots) {
    __shared__ uint8_t s_lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = lut[i];
    __syncthreads();
    const unsigned long long base = off[0], total = off[n_seq] - base;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long pos = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; pos < total;
         pos += stride) {
        // peg containing pos: last i with off[i] - base <= pos
        uint32_t lo = 0, hi = n_seq;
        while (lo < hi) {
            uint32_t mid = lo + ((hi - lo) >> 1);
            if (off[mid + 1] - base <= pos) lo = mid + 1; else hi = mid;
        }
        const uint32_t s = lo;
        const int nr = n_roles[s];
        if (PASS == 1 ? nr != 1 : nr != 0) continue;                 // :159, :165
        if (pos + K > off[s + 1] - base) continue;                   // window must stay inside the peg
        unsigned long long key = 0;
        for (int j = 0; j < K; j++) key = (key << 5) | s_lut[res[pos + j]];
        unsigned long long slot = mix64(key) & (n_slots - 1);
        if (PASS == 1) {
            const unsigned long long mine = (unsigned long long)(uint32_t)peg_role[s] + 1ull;
            for (;;) {
                unsigned long long old = atomicCAS(&table[slot].key, 0ull, key);     // computeIfAbsent :170
                if (old == 0ull || old == key) {
                    unsigned long long v = atomicCAS(&table[slot].val, 0ull, mine);  // new RoleCounter(pegRole)
                    if (v != 0ull && (v & 0xffffffffull) != mine) atomicOr(&table[slot].val, BUILD_BAD);  // count(): badCount++ :171
                    break;
                }
                slot = (slot + 1) & (n_slots - 1);
            }
        } else {
            for (;;) {                                                               // kmerMap.remove(kmer) :201
                const unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(&table[slot].key);
                if (k == 0ull) break;
                if (k == key) { atomicOr(&table[slot].val, BUILD_DEAD); break; }
                slot = (slot + 1) & (n_slots - 1);
            }
        }
    }
}

cudaError_t launch_build_pass(int pass, const uint8_t* res, const unsigned long long* off, uint32_t n_seq,
                              const int32_t* n_roles, const int32_t* peg_role, int K, const uint8_t* lut,
                              Slot128* table, unsigned long long n_slots, cudaStream_t st) {
    if (n_seq == 0) return cudaSuccess;
    if (pass == 1) build_pass_kernel<1><<<148 * 16, 256, 0, st>>>(res, off, n_seq, n_roles, peg_role, K, lut, table, n_slots);
    else build_pass_kernel<2><<<148 * 16, 256, 0, st>>>(res, off, n_seq, n_roles, peg_role, K, lut, table, n_slots);
    return cudaGetLastError();
}

__global__ void build_emit_kernel(const Slot128* __restrict__ table, unsigned long long n_slots, int K,
                                  const uint8_t* __restrict__ inv_lut, unsigned long long cap,
                                  uint8_t* out_kmers, int32_t* out_roles, unsigned long long* counter) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long s = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += stride) {
        const unsigned long long key = table[s].key, v = table[s].val;
        if (key == 0ull || (v & (BUILD_BAD | BUILD_DEAD))) continue;   // isGood() :186, removed :201
        const unsigned long long idx = atomicAdd(counter, 1ull);
        if (idx >= cap) continue;
        unsigned long long k = key;
        for (int j = K - 1; j >= 0; j--) { out_kmers[idx * K + j] = inv_lut[k & 31]; k >>= 5; }
        out_roles[idx] = (int32_t)((v & 0xffffffffull) - 1);            // kmer TAB roleId :213-215
    }
}

cudaError_t launch_build_emit(const Slot128* table, unsigned long long n_slots, int K, const uint8_t* inv_lut,
                              unsigned long long cap, uint8_t* out_kmers, int32_t* out_roles,
                              unsigned long long* counter, cudaStream_t st) {
    build_emit_kernel<<<148 * 16, 256, 0, st>>>(table, n_slots, K, inv_lut, cap, out_kmers, out_roles, counter);
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
            if (BYTES == 32) load_sector(buf + 2 * idx, a[u], b[u]);
            else {
                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
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
