// ka_line.cu — kernels of the 128-byte-line table (slot class 16) and of the packed residue stream
// (layout and rationale: ka_line.cuh).  Same reference semantics as ka_kernels.cu:
// ApplyKmerProcessor.java:122-148 (distinct K-windows of a peg :123, kmerRoleMap.get :130,
// unanimous-role tally :131-144, thresholded call :146-147) and :99-110 (HashMap.put, last line wins).
//
//   line_plan_kernel    chunk-relative 32-bit offsets, one descriptor per residue tile, segment descriptors of the
//                       sequences that get tiles of their own (mid), list of those for the long-sequence kernel (big)
//   line_filter_kernel  one warp per tile: a lane rolls the radix-n key halves over a run of <= 8 consecutive
//                       positions (codes read straight from the packed stream), issues the filter loads (L2) of
//                       the whole run, and the warp appends the survivors to the tile's list in global memory
//   line_probe_kernel   one warp per tile: survivors read densely, 3 per lane with all their sector loads (the only
//                       HBM access of a probe) in flight; eight tags matched SIMD-in-register; hits compacted
//                       over the front of the list
//   line_tally_kernel   one CTA per tile: hits de-duplicated per sequence (token set in shared memory), tallied
//                       (warp match + redux), calls written
//   line_big_kernel     sequences beyond the shared-memory tile sizes
//   line_insert/finalize, pack/unpack: table build and stream conversion
#include "ka_line.cuh"
#include "ka_common.cuh"
#include "ka_kernels.cuh"
#include <type_traits>

#ifdef KA_DEBUG
#define KA_CHECK(cond, code) do { if (!(cond)) atomicOr(p.dbg, (code)); } while (0)
#else
#define KA_CHECK(cond, code) do { } while (0)
#endif

namespace ka {

namespace {

__device__ __forceinline__ unsigned long long policy_evict_first() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ unsigned long long policy_evict_normal() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ unsigned long long policy_evict_last() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// one table sector: a single 256-bit load, no L1 allocation, evict-first in L2 (a line is used once)
__device__ __forceinline__ void load_line_sector(const uint4* p, unsigned long long pol, uint4& a, uint4& b) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p), "l"(pol));
}
__device__ __forceinline__ uint32_t load_filter_word(const uint32_t* p, unsigned long long pol) {
    uint32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}

// 0x8000 in every 16-bit half of x that is zero (exact: no carries between the halves)
__device__ __forceinline__ uint32_t zero_halves(uint32_t x) {
    const uint32_t t = (x & 0x7FFF7FFFu) + 0x7FFF7FFFu;
    return ~(t | x | 0x7FFF7FFFu);
}

// compare the eight tags of a sector (a) with `tagdup` (the tag in both halves); role or -1
__device__ __forceinline__ int match16(const uint4& a, const uint4& b, uint32_t tagdup, uint32_t& j) {
    constexpr uint32_t M = TAG_CMP | (TAG_CMP << 16);
    const uint32_t z0 = zero_halves((a.x ^ tagdup) & M), z1 = zero_halves((a.y ^ tagdup) & M);
    const uint32_t z2 = zero_halves((a.z ^ tagdup) & M), z3 = zero_halves((a.w ^ tagdup) & M);
    if ((z0 | z1 | z2 | z3) == 0u) return -1;
    uint32_t z, r, base;
    if (z0) { z = z0; r = b.x; base = 0; }
    else if (z1) { z = z1; r = b.y; base = 2; }
    else if (z2) { z = z2; r = b.z; base = 4; }
    else { z = z3; r = b.w; base = 6; }
    const uint32_t hi = (z & 0x8000u) ? 0u : 1u;
    j = base + hi;
    return (int)((r >> (16u * hi)) & 0xFFFFu);
}

// flag bits of a home sector: bit a-1 = keys of this home live in sector home^a, bit 3 = in the overflow table
__device__ __forceinline__ uint32_t sector_flags(const uint4& a) {
    return ((a.x >> 14) & 1u) | ((a.x >> 29) & 2u) | ((a.y >> 12) & 4u) | ((a.y >> 27) & 8u);
}

__device__ __forceinline__ int line_ovf_lookup(const LineTable& t, uint32_t sector, uint32_t tag, uint32_t& tok) {
    const unsigned long long k = line_ovf_key(sector, tag);
    const uint32_t mask = (1u << t.ovf_bbits) - 1u;
    uint32_t s = t.ovf_bbits ? (uint32_t)(mix64(k) >> (64 - t.ovf_bbits)) : 0u;
    for (;;) {
        uint4 a, b;
        load_sector(t.ovf + 2 * (size_t)s, a, b);
        const unsigned long long k0 = u64_of(a.x, a.y), k1 = u64_of(b.x, b.y);
        if (k0 == k) { tok = t.n_lines * 32u + 2u * s + 1u; return (int)a.z; }
        if (k1 == k) { tok = t.n_lines * 32u + 2u * s + 2u; return (int)b.z; }
        if (k1 == 0) return -1;
        s = (s + 1) & mask;
    }
}

// Full lookup of (sector, tag) given its already loaded home sector; role or -1, tok = de-dup token.
__device__ __forceinline__ int line_resolve(const LineTable& t, unsigned long long pol, uint32_t sector, uint32_t tag,
                                            const uint4& a, const uint4& b, uint32_t& tok) {
    const uint32_t tagdup = tag | (tag << 16);
    uint32_t j = 0;
    int role = match16(a, b, tagdup, j);
    if (role >= 0) { tok = sector * 8u + j + 1u; return role; }
    const uint32_t flags = sector_flags(a);
    if (flags == 0u) return -1;
    const uint32_t line0 = sector & ~3u, home = sector & 3u;
#pragma unroll
    for (uint32_t alt = 1; alt < 4; alt++) {
        if (flags & (1u << (alt - 1))) {
            const uint32_t s2 = line0 | (home ^ alt);
            uint4 xa, xb;
            load_line_sector(t.lines + 2 * (size_t)s2, pol, xa, xb);    // L2 hit: the line was just fetched
            role = match16(xa, xb, tagdup, j);
            if (role >= 0) { tok = s2 * 8u + j + 1u; return role; }
        }
    }
    if (flags & 8u) return line_ovf_lookup(t, sector, tag, tok);
    return -1;
}

__device__ __forceinline__ bool line_token_insert(uint32_t* region, uint32_t n, uint32_t token) {
    uint32_t j = (uint32_t)(((unsigned long long)(token * 0x9E3779B1u) * n) >> 32);
    for (;;) {
        const uint32_t old = atomicCAS(region + j, 0u, token);
        if (old == 0u) return true;
        if (old == token) return false;
        j = (j + 1 == n) ? 0 : j + 1;
    }
}

__device__ __forceinline__ void line_emit(const LineParams& p, uint32_t seq, int cnt, int rmin, int rmax) {
    int role = -1, hits = 0;
    uint8_t flag = 0;                                   // KA_FLAG_NONE
    if (cnt > 0) {
        if (rmin != rmax) { flag = 2; }                 // badPeg: two roles hit (ApplyKmerProcessor.java:140-143)
        else if (cnt >= p.min_hits) { role = rmin; hits = cnt; flag = 1; }  // :146
        else { hits = cnt; flag = 3; }
    }
    p.out_role[seq] = role;
    p.out_hits[seq] = hits;
    if (p.out_flag) p.out_flag[seq] = flag;
}

// 5-bit code number I (compile time) of a 64-bit window held in two registers
template <int I>
__device__ __forceinline__ uint32_t window_code(uint32_t u0, uint32_t u1) {
    constexpr int bit = 5 * I, sh = bit & 31;
    static_assert(bit + 5 <= 64, "code outside the window");
    if (bit + 5 <= 32) return (u0 >> sh) & 31u;
    else if (bit >= 32) return (u1 >> sh) & 31u;
    else return __funnelshift_r(u0, u1, sh) & 31u;
}
// the 64 stream bits starting at stage bit `bit`
__device__ __forceinline__ void load_window(const uint32_t* s_pk, uint32_t bit, uint32_t& u0, uint32_t& u1) {
    const uint32_t* w = s_pk + (bit >> 5);
    const uint32_t sh = bit & 31u, w0 = w[0], w1 = w[1], w2 = w[2];
    u0 = __funnelshift_r(w0, w1, sh);
    u1 = __funnelshift_r(w1, w2, sh);
}

// code of residue g of a packed stream in global memory
__device__ __forceinline__ uint32_t global_code(const uint32_t* pk, unsigned long long g) {
    const unsigned long long bit = 5ull * g;
    const uint32_t* w = pk + (bit >> 5);
    const uint32_t sh = (uint32_t)bit & 31u;
    const uint32_t lo = __ldg(w), hi = sh > 27 ? __ldg(w + 1) : 0u;
    return __funnelshift_r(lo, hi, sh) & 31u;
}

template <typename OffT>
__device__ __forceinline__ uint32_t first_seq_at(const OffT* off, uint32_t n_seq, unsigned long long target) {
    uint32_t lo = 0, hi = n_seq;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if ((unsigned long long)off[mid] < target) lo = mid + 1; else hi = mid;
    }
    return lo;
}

}  // namespace

// ------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------
template <typename OffT>
__global__ void line_plan_kernel(LineParams p, const OffT* __restrict__ off, unsigned long long origin) {
    const unsigned long long gid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long base = off[0];
    if (gid < p.n_tiles) {
        const uint32_t s0 = first_seq_at(off, p.n_seq, base + gid * (unsigned long long)p.tile_span);
        uint32_t s1 = first_seq_at(off, p.n_seq, base + (gid + 1) * (unsigned long long)p.tile_span);
        // only the last sequence starting in a tile can be a long one (long_seq >= tile_span)
        if (s1 > s0 && (unsigned long long)off[s1] - (unsigned long long)off[s1 - 1] > p.long_seq) s1--;
        uint4 d;
        d.x = s0; d.y = s1 - s0;
        d.z = (uint32_t)((unsigned long long)off[s0] - origin);
        d.w = (uint32_t)((unsigned long long)off[s1] - origin);
        p.first[gid] = d;
    }
    if (gid <= p.n_seq) p.off[gid] = (uint32_t)((unsigned long long)off[gid] - origin);
    if (gid < p.n_seq) {
        const unsigned long long L = (unsigned long long)off[gid + 1] - (unsigned long long)off[gid];
        if (L > p.mid_seq) {
            const uint32_t idx = atomicAdd(p.big_count, 1u);
            const unsigned long long tb = atomicAdd(p.tok_cursor, 2ull * L);
            BigItem it; it.seq = (uint32_t)gid; it.pad = 0; it.tok_base = tb;
            reinterpret_cast<BigItem*>(p.big_list)[idx] = it;
        } else if (L > p.long_seq) {
            // a sequence with tiles of its own: equal segments of <= LINE_MID_SEG positions, one warp each in the
            // filter and probe passes; segment k of the sequence has count slot idx + k
            const uint32_t nseg = ((uint32_t)L + LINE_MID_SEG - 1) / LINE_MID_SEG, seglen = ((uint32_t)L + nseg - 1) / nseg;
            const uint32_t idx = atomicAdd(p.mid_count, nseg);
            const uint32_t a0 = (uint32_t)((unsigned long long)off[gid] - origin), a1 = a0 + (uint32_t)L;
            for (uint32_t k = 0; k < nseg; k++) {
                uint4 d;
                d.x = (uint32_t)gid; d.y = idx + k;
                d.z = a0 + k * seglen;
                d.w = min(a1, d.z + seglen);
                p.mid_desc[idx + k] = d;
            }
        }
    }
}

cudaError_t launch_line_plan(const LineParams& p, const unsigned long long* off64, const uint32_t* off32,
                             unsigned long long origin, cudaStream_t st) {
    unsigned long long n = (unsigned long long)p.n_seq + 1;
    if (p.n_tiles > n) n = p.n_tiles;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (off64) line_plan_kernel<unsigned long long><<<blocks, 256, 0, st>>>(p, off64, origin);
    else line_plan_kernel<uint32_t><<<blocks, 256, 0, st>>>(p, off32, origin);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// tile kernels: filter pass, probe pass, tally pass
// ------------------------------------------------------------------------------------
// The work of a tile is three kernels with a list in global memory between them, because its parts want opposite
// things from the SM.  The filter pass is arithmetic over a stream (rolling keys, one L2-resident filter word per
// window): every lane busy, no HBM access besides the 0.625 B per residue of the codes.  The probe pass is nothing
// but random 128-byte lines: with the survivors already compacted, every lane of every warp has PB sector loads in
// flight, which a kernel that discovers its survivors as it goes cannot arrange (profiles/r02_summary.md C: 10-15
// active lanes per instruction and a third of the time at the tile barrier when all of it was one kernel).  The
// tally pass is shared-memory atomics on the hits only.  The list costs 8 bytes per SURVIVOR written and read once
// — ~7 B per window against the 128-byte line saved for every window the filter rejects.  Survivors of the tile
// (sub-batch) whose first position is chunk residue g0 live at surv[g0 ...), their number at surv_cnt[first sequence
// of the sub-batch]; the probe pass writes the hits over the front of the same list: no prefix sum, no zeroing.
constexpr int LF_A = 8;        // window positions per lane and pass
#ifndef KA_LP_PB
#define KA_LP_PB 3
#endif
constexpr int LP_PB = KA_LP_PB;   // sector loads in flight per thread of the probe pass
#ifndef KA_LP_MINW
#define KA_LP_MINW 32            // resident warps per SM of the probe pass (batching the flag-path loads of the PB probes
                                 // at 80 registers / 24 warps measured slower: 59.3 against 64.0 G probes/s)
#endif

// dynamic shared memory of the tally pass: s_off, s_cnt/s_min/s_max, then the token set (region of sequence q at tok_cap(start) + 4q)
constexpr uint32_t LY_OFF_CNT = 4 * (LINE_MAX_SEQ + 4);
constexpr uint32_t LY_OFF_TOK = LY_OFF_CNT + 3 * 4 * LINE_MAX_SEQ;
static_assert(LY_OFF_TOK % 16 == 0, "alignment of the shared-memory parts");

size_t line_tally_smem_bytes(uint32_t ext_max) {
    return (size_t)LY_OFF_TOK + 4 * ((size_t)tok_cap(ext_max) + 4 * LINE_MAX_SEQ + 8);
}
size_t line_tile_smem_bytes(uint32_t ext_max) { return line_tally_smem_bytes(ext_max); }

namespace {
// 5-bit code number J (compile time) of a 96-bit window held in three registers
template <int J>
__device__ __forceinline__ uint32_t wcode(uint32_t u0, uint32_t u1, uint32_t u2) {
    constexpr int bit = 5 * J, w = bit >> 5, sh = bit & 31;
    static_assert(bit + 5 <= 96, "code outside the window");
    const uint32_t lo = w == 0 ? u0 : (w == 1 ? u1 : u2);
    if (sh + 5 <= 32) return (lo >> sh) & 31u;
    const uint32_t hi = w == 0 ? u1 : u2;
    return __funnelshift_r(lo, hi, sh) & 31u;
}
}  // namespace

// Filter pass.  ONE WARP per tile, no shared memory, no barrier, no atomic: the kernel is a chain of dependent
// latencies (offsets -> codes -> filter words), so what it needs is many independent warps, not co-operation.  The
// tile's positions are cut into passes of <= 32 * A; in a pass a lane rolls the two key halves over a run of <= A
// consecutive positions (K is a compile-time constant: every code of the lane's 96-bit window sits at a fixed bit),
// issues the filter loads of the whole run, tests them, and the warp appends its survivors (H | seq << 26, Lo) to
// the tile's list — one ballot per step, the running count in a register.
template <int K, bool FILTER>
__global__ void __launch_bounds__(32, 32) line_filter_kernel(LineParams p) {
    constexpr int A = LF_A;
    constexpr int Kh = K / 2;
    static_assert(A + K - 1 <= 19, "the lane's window holds 19 codes");
    const uint32_t lane = threadIdx.x;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t radix = p.tab.radix, npw_h = 0u - p.tab.pw_h, npw_l = 0u - p.tab.pw_l, n_filt = p.tab.n_filt;
    const uint32_t* const filt = p.tab.filt;
    const uint32_t* __restrict__ const off = p.off;
    const unsigned long long pol_last = policy_evict_last();

    for (uint32_t tile = p.tile0 + blockIdx.x; tile < p.tile1; tile += gridDim.x) {
    // segments of single-sequence (mid) tiles first: the longest items of the launch start earliest
    const bool mid = tile < p.n_mid_tiles;
    const uint4 desc = mid ? p.mid_desc[tile] : p.first[tile - p.n_mid_tiles];
    const uint32_t s0 = desc.x, s1 = mid ? desc.x + 1u : desc.x + desc.y;
    for (uint32_t sb = s0; sb < s1; sb += LINE_MAX_SEQ) {
        const uint32_t ns = min((uint32_t)LINE_MAX_SEQ, s1 - sb);
        const uint32_t g0 = (sb == s0) ? desc.z : off[sb];
        const uint32_t g1 = (sb + ns == s1) ? desc.w : off[sb + ns];
        const uint32_t ext = g1 - g0;                                   // window positions of the sub-batch
        const uint32_t n_pass = (ext + 32 * A - 1) / (32 * A);
        const uint32_t plen = n_pass ? (ext + n_pass - 1) / n_pass : 0;  // <= 32 * A
        uint2* const out = p.surv + g0;
        uint32_t cnt = 0;                                               // survivors so far (warp-uniform)
        uint32_t sw = sb;                                               // a sequence at or before the pass's first position

        for (uint32_t pass = 0; pass < n_pass; pass++) {
            const uint32_t pb = pass * plen, pend = min(ext, pb + plen);
            const uint32_t run = (pend - pb + 31) >> 5;                 // <= A
            const uint32_t P0 = pb + lane * run;
            const uint32_t nrun = P0 < pend ? min(run, pend - P0) : 0u;
            // sequence containing the lane's first position: walk on from the warp's (sequences are mostly longer
            // than a run, so this is one or two cached loads; lanes past the end look at the last position)
            const uint32_t Pc = g0 + min(P0, pend - 1u);
            uint32_t si = sw;
            uint32_t nb = off[si + 1];
            while (Pc >= nb) { si++; nb = off[si + 1]; }
            KA_CHECK(si < sb + ns, 32u);

            uint32_t qh[A], ql[A], fw[A];
            unsigned okm = 0;
            if (nrun) {
                // the 96 stream bits from the run's first position: code J of the window at a fixed bit
                const unsigned long long bit0 = 5ull * (g0 + P0);
                const uint32_t* w = p.pk + (bit0 >> 5);
                const uint32_t sh = (uint32_t)bit0 & 31u, w0 = __ldcs(w), w1 = __ldcs(w + 1), w2 = __ldcs(w + 2), w3 = __ldcs(w + 3);
                const uint32_t u0 = __funnelshift_r(w0, w1, sh), u1 = __funnelshift_r(w1, w2, sh), u2 = __funnelshift_r(w2, w3, sh);
                const uint32_t G0 = g0 + P0;                            // chunk residue of the run's first position

                // warm-up over the K-1 codes in front of the first window end: H complete, Lo short of one digit
                uint32_t H = 0, Lo = 0;
                int okc = 0;                                            // consecutive codes inside the alphabet
                auto warm = [&](auto J_) {
                    constexpr int J = decltype(J_)::value;
                    if (J < K - 1) {
                        const uint32_t c = wcode<(J < K - 1 ? J : 0)>(u0, u1, u2);
                        if (J < Kh) H = H * radix + c; else Lo = Lo * radix + c;
                        okc = (c == CODE_INVALID) ? 0 : okc + 1;
                    }
                };
                warm(std::integral_constant<int, 0>()); warm(std::integral_constant<int, 1>());
                warm(std::integral_constant<int, 2>()); warm(std::integral_constant<int, 3>());
                warm(std::integral_constant<int, 4>()); warm(std::integral_constant<int, 5>());
                warm(std::integral_constant<int, 6>()); warm(std::integral_constant<int, 7>());
                warm(std::integral_constant<int, 8>());
                static_assert(K <= 10, "warm-up unrolled for K - 1 <= 9 codes");

                auto step = [&](auto I_) {
                    constexpr int I = decltype(I_)::value;
                    const uint32_t pos = G0 + I;
                    const uint32_t cnew = wcode<I + K - 1>(u0, u1, u2);
                    if (I > 0) {
                        const uint32_t cmid = wcode<I + Kh - 1>(u0, u1, u2);
                        H = (H + wcode<(I > 0 ? I - 1 : 0)>(u0, u1, u2) * npw_h) * radix + cmid;
                        Lo = (Lo + cmid * npw_l) * radix + cnew;
                    } else {
                        Lo = Lo * radix + cnew;
                    }
                    okc = (cnew == CODE_INVALID) ? 0 : okc + 1;
                    if (pos >= nb && I < (int)nrun) {                   // rare: the run crosses into the next sequence(s)
                        do { si++; KA_CHECK(si < sb + ns, 4u); nb = off[si + 1]; } while (pos >= nb);
                    }
                    // (positions past the lane's run belong to the next lane: never valid here)
                    if (pos + (uint32_t)K <= nb && okc >= K && I < (int)nrun) {
                        qh[I] = H | ((si - sb) << 26);                  // H < 31^5 < 2^25, si - sb < 64
                        ql[I] = Lo;
                        okm |= 1u << I;
                        if (FILTER) fw[I] = load_filter_word(filt + __umulhi(line_filter_hash(H, Lo), n_filt), pol_last);
                    }
                };
                step(std::integral_constant<int, 0>()); step(std::integral_constant<int, 1>());
                step(std::integral_constant<int, 2>()); step(std::integral_constant<int, 3>());
                step(std::integral_constant<int, 4>()); step(std::integral_constant<int, 5>());
                step(std::integral_constant<int, 6>()); step(std::integral_constant<int, 7>());
                static_assert(A == 8, "unrolled for 8 positions per lane");
            }
            // the next pass starts at or after the last lane's sequence
            sw = __shfl_sync(0xffffffffu, si, 31);
            // test and append: the survivors of step i of all lanes go to consecutive list slots
#pragma unroll
            for (int i = 0; i < A; i++) {
                bool ok = (okm >> i) & 1u;
                if (FILTER && ok) {
                    const uint32_t need = line_filter_bits(line_filter_hash(qh[i] & 0x3FFFFFFu, ql[i]));
                    ok = (fw[i] & need) == need;
                }
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                if (ok) __stcs(out + cnt + __popc(m & lt), make_uint2(qh[i], ql[i]));   // streaming: keep L2 for the filter words
                cnt += __popc(m);
            }
        }
        if (lane == 0) p.surv_cnt[mid ? p.n_seq + 2u + desc.y : sb] = cnt;
    }
    }
}

// Probe pass.  One warp per tile, again without shared memory or barriers: the survivors are read densely, PB per
// lane with all their sector loads (the only HBM access of a probe) in flight before the first use; eight tags
// matched SIMD-in-register; a miss in a sector whose flags name other places follows them (L2 hits on the line just
// fetched, or the overflow table).  The hits (de-dup token, role | seq << 16) are written back compacted over the
// front of the tile's list — the warp has consumed an entry before a hit can land on it — and counted in hit_cnt.
__global__ void __launch_bounds__(32, KA_LP_MINW) line_probe_kernel(LineParams p) {
    constexpr int PB = LP_PB;
    const uint32_t lane = threadIdx.x;
    const uint32_t lt = (1u << lane) - 1u;
    const LineTable tab = p.tab;
        // (no filter word is read in this pass: table lines are cached like anything else — 65.8 against 65.3 G probes/s
    // with an evict-first hint)
    const unsigned long long pol_first = policy_evict_normal();

    for (uint32_t tile = p.tile0 + blockIdx.x; tile < p.tile1; tile += gridDim.x) {
    // segments of single-sequence (mid) tiles first: the longest items of the launch start earliest
    const bool mid = tile < p.n_mid_tiles;
    const uint4 desc = mid ? p.mid_desc[tile] : p.first[tile - p.n_mid_tiles];
    const uint32_t s0 = desc.x, s1 = mid ? desc.x + 1u : desc.x + desc.y;
    for (uint32_t sb = s0; sb < s1; sb += LINE_MAX_SEQ) {
        const uint32_t g0 = (sb == s0) ? desc.z : p.off[sb];
        const uint32_t slot = mid ? p.n_seq + 2u + desc.y : sb;
        const uint32_t n = p.surv_cnt[slot];
        uint2* const q = p.surv + g0;
        uint32_t hcnt = 0;
        uint2 kq[PB];
#pragma unroll
        for (int k = 0; k < PB; k++) kq[k] = (lane + k * 32 < n) ? q[lane + k * 32] : make_uint2(0, 0);
        for (uint32_t base = 0; base < n; base += 32 * PB) {
            uint32_t sec[PB], tg[PB];
            uint4 sa[PB], sb2[PB];
#pragma unroll
            for (int k = 0; k < PB; k++) {
                sec[k] = 0; tg[k] = 0;
                sa[k] = make_uint4(0, 0, 0, 0); sb2[k] = sa[k];
                if (base + lane + k * 32 < n) {
                    line_locate(tab, kq[k].x & 0x3FFFFFFu, kq[k].y, sec[k], tg[k]);
                    load_line_sector(tab.lines + 2 * (size_t)sec[k], pol_first, sa[k], sb2[k]);
                }
            }
            // the next round's entries travel while this round's sectors do (they lie behind everything this round writes)
            uint2 nx[PB];
#pragma unroll
            for (int k = 0; k < PB; k++) {
                const uint32_t i = base + 32 * PB + lane + k * 32;
                nx[k] = i < n ? q[i] : make_uint2(0, 0);
            }
#pragma unroll
            for (int k = 0; k < PB; k++) {
                if (base + k * 32 >= n) break;                          // (warp-uniform)
                int role = -1;
                uint32_t tok = 0;
                if (base + lane + k * 32 < n) role = line_resolve(tab, pol_first, sec[k], tg[k], sa[k], sb2[k], tok);
                const unsigned m = __ballot_sync(0xffffffffu, role >= 0);
                if (role >= 0) q[hcnt + __popc(m & lt)] = make_uint2(tok, (uint32_t)role | ((kq[k].x >> 26) << 16));
                hcnt += __popc(m);
            }
#pragma unroll
            for (int k = 0; k < PB; k++) kq[k] = nx[k];
        }
        if (lane == 0) p.hit_cnt[slot] = hcnt;
    }
    }
}

// Tally pass.  One CTA per tile: every hit is de-duplicated against its sequence's token set (a protein contributes
// the SET of its k-mers, ApplyKmerProcessor.java:123) and tallied (warp match + redux, one shared atomic per
// sequence and warp), then the calls of the tile's sequences are written (:131-147).
constexpr int LY_THREADS = 128;
template <bool MID>
__global__ void __launch_bounds__(LY_THREADS, 8) line_tally_kernel(LineParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t* const s_off = reinterpret_cast<uint32_t*>(smem_raw);
    int* const s_cnt = reinterpret_cast<int*>(smem_raw + LY_OFF_CNT);
    int* const s_min = s_cnt + LINE_MAX_SEQ;
    int* const s_max = s_min + LINE_MAX_SEQ;
    uint32_t* const s_tok = reinterpret_cast<uint32_t*>(smem_raw + LY_OFF_TOK);

    const uint32_t tid = threadIdx.x, lane = tid & 31;

    for (uint32_t tile = p.tile0 + blockIdx.x; tile < p.tile1; tile += gridDim.x) {
    const uint4 desc = p.first[tile];
    const uint32_t s0 = desc.x;
    // an ordinary tile has one list of hits per sub-batch; a single-sequence tile one per segment of the sequence,
    // all tallied by the CTA of the first segment
    uint32_t s1 = s0 + desc.y, nlist = 1, seglen = 0, mid_slot = 0;
    if (MID) {
        const uint32_t a0 = p.off[s0], L = p.off[s0 + 1] - a0;
        if (desc.z != a0) continue;                                     // (uniform) not the sequence's first segment
        s1 = s0 + 1;
        nlist = (L + LINE_MID_SEG - 1) / LINE_MID_SEG;
        seglen = (L + nlist - 1) / nlist;
        mid_slot = p.n_seq + 2u + desc.y;
    }
    for (uint32_t sb = s0; sb < s1; sb += LINE_MAX_SEQ) {
        const uint32_t ns = min((uint32_t)LINE_MAX_SEQ, s1 - sb);
        const uint32_t g0 = (sb == s0) ? desc.z : p.off[sb];
        const uint32_t slot0 = MID ? mid_slot : sb;
        const uint32_t n0 = p.hit_cnt[slot0];
        uint32_t total = n0;
        if (MID) for (uint32_t k = 1; k < nlist; k++) total += p.hit_cnt[slot0 + k];
        if (total == 0) {                                               // (uniform) no hit: no call for these sequences
            for (uint32_t i = tid; i < ns; i += LY_THREADS) line_emit(p, sb + i, 0, 0, 0);
            continue;
        }
        const uint2 first = tid < n0 ? p.surv[g0 + tid] : make_uint2(0, 0);   // in flight during the set-up
        const uint32_t g1 = MID ? p.off[s0 + 1] : ((sb + ns == s1) ? desc.w : p.off[sb + ns]);
        const uint32_t ext = g1 - g0;
        KA_CHECK(total <= ext, 16u);
        KA_CHECK(tok_cap(ext) + 4u * ns + 8u <= tok_cap(p.ext_max) + 4u * LINE_MAX_SEQ + 8u, 2u);
        for (uint32_t i = tid; i <= ns; i += LY_THREADS) s_off[i] = p.off[sb + i] - g0;
        for (uint32_t i = tid; i < ns; i += LY_THREADS) { s_cnt[i] = 0; s_min[i] = 0x7fffffff; s_max[i] = -1; }
        {
            const uint32_t ntok = tok_cap(ext) + 4u * ns + 4u;
            const uint4 z = make_uint4(0, 0, 0, 0);
            for (uint32_t i = tid * 4; i < ntok; i += LY_THREADS * 4) *reinterpret_cast<uint4*>(s_tok + i) = z;
        }
        __syncthreads();
        for (uint32_t k = 0; k < (MID ? nlist : 1u); k++) {
            const uint32_t n = (MID && k) ? p.hit_cnt[slot0 + k] : n0;
            const uint2* const q = p.surv + g0 + (MID ? k * seglen : 0u);
            for (uint32_t base = 0; base < n; base += LY_THREADS) {
                if (base + (tid & ~31u) >= n) break;                    // (warp-uniform) the whole warp is past the end
                const uint32_t i = base + tid;
                const uint2 h = (base == 0 && k == 0) ? first : (i < n ? q[i] : make_uint2(0, 0));
                int sq = -1;
                const int role = (int)(h.y & 0xFFFFu);
                if (i < n) {
                    sq = (int)(h.y >> 16);
                    const uint32_t a0 = s_off[sq], a1 = s_off[sq + 1];
                    KA_CHECK(sq < (int)ns && a1 >= a0, 8u);
                    if (!line_token_insert(s_tok + tok_cap(a0) + 4u * (uint32_t)sq, tok_cap(a1 - a0) + 4u, h.x)) sq = -1;
                }
                if (__any_sync(0xffffffffu, sq >= 0)) {
                    const unsigned grp = __match_any_sync(0xffffffffu, sq);
                    const int gmin = __reduce_min_sync(grp, sq >= 0 ? role : 0x7fffffff);
                    const int gmax = __reduce_max_sync(grp, sq >= 0 ? role : -1);
                    if (sq >= 0 && lane == (uint32_t)(__ffs(grp) - 1)) {
                        atomicAdd(&s_cnt[sq], __popc(grp));
                        atomicMin(&s_min[sq], gmin);
                        atomicMax(&s_max[sq], gmax);
                    }
                }
            }
        }
        __syncthreads();
        for (uint32_t i = tid; i < ns; i += LY_THREADS) line_emit(p, sb + i, s_cnt[i], s_min[i], s_max[i]);
        __syncthreads();
    }
    }
}

namespace {
template <int K> void filter_launch(const LineParams& p, unsigned grid, cudaStream_t st) {
    if (p.tab.filt) line_filter_kernel<K, true><<<grid, 32, 0, st>>>(p);
    else line_filter_kernel<K, false><<<grid, 32, 0, st>>>(p);
}
}  // namespace

cudaError_t line_tile_set_smem(size_t bytes) {
    cudaError_t ce = cudaFuncSetAttribute(line_tally_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(line_tally_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(line_tally_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(line_tally_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    return ce;
}

// the three passes over tiles [p.tile0, p.tile1), each with `grid` CTAs striding over them
cudaError_t launch_line_filter(const LineParams& p, unsigned grid, cudaStream_t st) {
    if (p.tile1 <= p.tile0) return cudaSuccess;
    grid = grid < p.tile1 - p.tile0 ? grid : p.tile1 - p.tile0;
    switch (p.tab.K) {
        case 2: filter_launch<2>(p, grid, st); break;
        case 3: filter_launch<3>(p, grid, st); break;
        case 4: filter_launch<4>(p, grid, st); break;
        case 5: filter_launch<5>(p, grid, st); break;
        case 6: filter_launch<6>(p, grid, st); break;
        case 7: filter_launch<7>(p, grid, st); break;
        case 8: filter_launch<8>(p, grid, st); break;
        case 9: filter_launch<9>(p, grid, st); break;
        case 10: filter_launch<10>(p, grid, st); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
cudaError_t launch_line_probe(const LineParams& p, unsigned grid, cudaStream_t st) {
    if (p.tile1 <= p.tile0) return cudaSuccess;
    grid = grid < p.tile1 - p.tile0 ? grid : p.tile1 - p.tile0;
    line_probe_kernel<<<grid, 32, 0, st>>>(p);
    return cudaGetLastError();
}
cudaError_t launch_line_tally(const LineParams& p, unsigned grid, cudaStream_t st) {
    if (p.tile1 <= p.tile0) return cudaSuccess;
    grid = grid < p.tile1 - p.tile0 ? grid : p.tile1 - p.tile0;
    if (p.tally_mid) line_tally_kernel<true><<<grid, LY_THREADS, line_tally_smem_bytes(p.ext_max), st>>>(p);
    else line_tally_kernel<false><<<grid, LY_THREADS, line_tally_smem_bytes(p.ext_max), st>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// long sequences: one CTA per sequence, codes read from global memory, token set in global scratch
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) line_big_kernel(LineParams p) {
    constexpr int THREADS = 256;
    __shared__ int sh_cnt, sh_min, sh_max;
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t nbig = *p.big_count;
    const LineTable tab = p.tab;
    const int K = tab.K;
    const unsigned long long pol_first = policy_evict_first(), pol_last = policy_evict_last();
    const BigItem* list = reinterpret_cast<const BigItem*>(p.big_list);

    for (uint32_t bi = blockIdx.x; bi < nbig; bi += gridDim.x) {
        const BigItem it = list[bi];
        const uint32_t a0 = p.off[it.seq], L = p.off[it.seq + 1] - a0;
        const uint32_t W = L - (uint32_t)K + 1;                       // L > mid_seq >= K
        uint32_t* region = p.scratch + it.tok_base;
        const uint32_t nreg = 2u * L;
        for (uint32_t i = tid; i < nreg; i += THREADS) region[i] = 0;
        if (tid == 0) { sh_cnt = 0; sh_min = 0x7fffffff; sh_max = -1; }
        __syncthreads();
        int cnt = 0, mn = 0x7fffffff, mx = -1;
        for (uint32_t pos = tid; pos < W; pos += THREADS) {
            uint32_t H = 0, Lo = 0;
            bool ok = true;
            for (int j = 0; j < K; j++) {
                const uint32_t c = global_code(p.pk, (unsigned long long)a0 + pos + j);
                ok &= (c != CODE_INVALID);
                if ((uint32_t)j < tab.Kh) H = H * tab.radix + c; else Lo = Lo * tab.radix + c;
            }
            if (!ok) continue;
            if (tab.filt) {
                const uint32_t fh = line_filter_hash(H, Lo), need = line_filter_bits(fh);
                if ((load_filter_word(tab.filt + line_filter_word(fh, tab.n_filt), pol_last) & need) != need) continue;
            }
            uint32_t sector, tag;
            line_locate(tab, H, Lo, sector, tag);
            uint4 a, b;
            load_line_sector(tab.lines + 2 * (size_t)sector, pol_first, a, b);
            uint32_t tok = 0;
            const int role = line_resolve(tab, pol_first, sector, tag, a, b, tok);
            if (role >= 0 && line_token_insert(region, nreg, tok)) {
                cnt++;
                mn = min(mn, role);
                mx = max(mx, role);
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
        if (tid == 0) line_emit(p, it.seq, sh_cnt, sh_min, sh_max);
        __syncthreads();
    }
}

cudaError_t launch_line_big(const LineParams& p, int grid, cudaStream_t st) {
    line_big_kernel<<<grid, 256, 0, st>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// stream conversion
// ------------------------------------------------------------------------------------
// one thread per 32 residues = 160 bits = 5 words
__global__ void __launch_bounds__(256) pack_kernel(const uint8_t* __restrict__ bytes, uint32_t lead, unsigned long long n,
                                                   const uint8_t* __restrict__ lut5, uint32_t* __restrict__ out,
                                                   unsigned long long n_groups) {
    __shared__ uint8_t s_lut[256];
    s_lut[threadIdx.x] = lut5[threadIdx.x];
    __syncthreads();
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_groups; t += stride) {
        unsigned long long acc = 0;
        int nb = 0;
        uint32_t* o = out + 5 * t;
#pragma unroll
        for (int k = 0; k < 32; k++) {
            const unsigned long long g = 32ull * t + k;              // chunk-relative residue
            uint32_t c = CODE_INVALID;
            if (g >= lead && g - lead < n) c = s_lut[bytes[g - lead]];
            acc |= (unsigned long long)c << nb;
            nb += 5;
            if (nb >= 32) { *o++ = (uint32_t)acc; acc >>= 32; nb -= 32; }
        }
    }
}

cudaError_t launch_pack(const uint8_t* bytes, uint32_t lead, unsigned long long n, const uint8_t* lut5,
                        uint32_t* out_words, cudaStream_t st) {
    const unsigned long long groups = (lead + n + 31) / 32;
    if (groups == 0) return cudaSuccess;
    const unsigned long long want = (groups + 255) / 256;
    pack_kernel<<<(unsigned)(want < 148ull * 32 ? want : 148ull * 32), 256, 0, st>>>(bytes, lead, n, lut5, out_words, groups);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) unpack_kernel(const uint32_t* __restrict__ words, uint32_t lead, unsigned long long n,
                                                     const uint8_t* __restrict__ inv32, uint8_t* __restrict__ out) {
    __shared__ uint8_t s_inv[32];
    if (threadIdx.x < 32) s_inv[threadIdx.x] = inv32[threadIdx.x];
    __syncthreads();
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
        out[j] = s_inv[global_code(words, (unsigned long long)lead + j)];
}

cudaError_t launch_unpack(const uint32_t* words, uint32_t lead, unsigned long long n, const uint8_t* inv32,
                          uint8_t* out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const unsigned long long want = (n + 255) / 256;
    unpack_kernel<<<(unsigned)(want < 148ull * 32 ? want : 148ull * 32), 256, 0, st>>>(words, lead, n, inv32, out);
    return cudaGetLastError();
}

__global__ void widen_offsets_kernel(const uint32_t* __restrict__ off32, unsigned long long n, unsigned long long* __restrict__ off64) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) off64[i] = off32[i];
}

cudaError_t launch_widen_offsets(const uint32_t* off32, unsigned long long n, unsigned long long* off64, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const unsigned long long want = (n + 255) / 256;
    widen_offsets_kernel<<<(unsigned)(want < 148ull * 8 ? want : 148ull * 8), 256, 0, st>>>(off32, n, off64);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// table build
// ------------------------------------------------------------------------------------
// HashMap.put for every DB line (ApplyKmerProcessor.java:106).  A key scans its home sector, then the
// sectors home^1, home^2, home^3 of the same line, always in this order and slots are never freed, so
// a duplicate of a stored key reaches the stored copy before any empty slot: one slot per distinct key.
__global__ void __launch_bounds__(256) line_insert_kernel(LineTable t, const uint8_t* __restrict__ kmers,
                                                          const int32_t* __restrict__ roles, unsigned long long n,
                                                          unsigned long long line_base, const uint8_t* __restrict__ lut5,
                                                          unsigned long long* best, uint32_t role_bits,
                                                          unsigned long long* counters, uint32_t* errs) {
    __shared__ uint8_t s_lut[256];
    s_lut[threadIdx.x] = lut5[threadIdx.x];
    __syncthreads();
    const int K = t.K;
    uint32_t* words = reinterpret_cast<uint32_t*>(const_cast<uint4*>(t.lines));
    uint32_t* filt = const_cast<uint32_t*>(t.filt);
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint8_t* km = kmers + i * (unsigned long long)K;
        uint32_t H = 0, Lo = 0;
        bool bad = false;
        for (int j = 0; j < K; j++) {
            const uint32_t c = s_lut[km[j]];
            bad |= (c == CODE_INVALID);
            if ((uint32_t)j < t.Kh) H = H * t.radix + c; else Lo = Lo * t.radix + c;
        }
        if (bad) { atomicAdd(&errs[0], 1u); continue; }
        if (roles[i] < 0) { atomicAdd(&errs[1], 1u); continue; }
        uint32_t sector, tag;
        line_locate(t, H, Lo, sector, tag);
        const unsigned long long mine = ((line_base + i + 1) << role_bits) | (unsigned long long)(uint32_t)roles[i];
        if (filt) {
            const uint32_t fh = line_filter_hash(H, Lo), bits = line_filter_bits(fh);
            uint32_t* fword = filt + line_filter_word(fh, t.n_filt);
            if ((*reinterpret_cast<volatile uint32_t*>(fword) & bits) != bits) atomicOr(fword, bits);
        }
        const uint32_t line0 = sector & ~3u, home = sector & 3u;
        bool done = false;
        for (uint32_t alt = 0; alt < 4 && !done; alt++) {
            const uint32_t s = line0 | (home ^ alt);
            for (uint32_t j = 0; j < 8 && !done; j++) {
                uint32_t* w = words + (size_t)s * 8 + (j >> 1);
                const uint32_t sh = 16u * (j & 1u);
                for (;;) {
                    const uint32_t cur = *reinterpret_cast<volatile uint32_t*>(w);
                    const uint32_t half = (cur >> sh) & 0xFFFFu;
                    if (half == 0u) {
                        if (atomicCAS(w, cur, cur | (tag << sh)) != cur) continue;   // the word changed: look again
                        atomicAdd(&counters[0], 1ull);
                        if (alt) atomicAdd(&counters[1], 1ull);
                        done = true;
                    } else if (((half ^ tag) & TAG_CMP) == 0u) {
                        done = true;                                                  // the key is already stored here
                    }
                    break;
                }
                if (done) {
                    atomicMax(&best[(size_t)s * 8 + j], mine);
                    if (alt) {   // tell lookups that sector home^alt holds keys of this home: flag of slot alt-1
                        uint32_t* fwd = words + (size_t)sector * 8 + ((alt - 1) >> 1);
                        const uint32_t fb = TAG_FLAG << (16u * ((alt - 1) & 1u));
                        if ((*reinterpret_cast<volatile uint32_t*>(fwd) & fb) == 0u) atomicOr(fwd, fb);
                    }
                }
            }
        }
        if (!done) {
            // the whole line is full: (sector, rem) + 1 goes to the overflow table, flag of slot 3
            uint32_t* fwd = words + (size_t)sector * 8 + 1;
            if ((*reinterpret_cast<volatile uint32_t*>(fwd) & (TAG_FLAG << 16)) == 0u) atomicOr(fwd, TAG_FLAG << 16);
            const unsigned long long k = line_ovf_key(sector, tag);
            const uint32_t omask = (1u << t.ovf_bbits) - 1u;
            uint32_t os = t.ovf_bbits ? (uint32_t)(mix64(k) >> (64 - t.ovf_bbits)) : 0u;
            Slot128* ovf = reinterpret_cast<Slot128*>(const_cast<uint4*>(t.ovf));
            for (uint32_t chain = 0; !done && chain <= omask; chain++) {
                Slot128* sl = ovf + 2 * (size_t)os;
                for (int h = 0; h < 2 && !done; h++) {
                    const unsigned long long old = atomicCAS(&sl[h].key, 0ull, k);
                    if (old == 0ull || old == k) {
                        if (old == 0ull) { atomicAdd(&counters[0], 1ull); atomicAdd(&counters[2], 1ull); }
                        atomicMax(&sl[h].val, mine);
                        done = true;
                    }
                }
                os = (os + 1) & omask;
            }
            if (!done) atomicAdd(&errs[2], 1u);
        }
    }
}

cudaError_t launch_line_insert(const LineTable& t, const uint8_t* kmers, const int32_t* roles, unsigned long long n,
                               unsigned long long line_base, const uint8_t* lut5, unsigned long long* best,
                               uint32_t role_bits, unsigned long long* counters, uint32_t* errs, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const unsigned long long want = (n + 255) / 256;
    line_insert_kernel<<<(unsigned)(want < 148ull * 32 ? want : 148ull * 32), 256, 0, st>>>(t, kmers, roles, n, line_base, lut5, best,
                                                                                         role_bits, counters, errs);
    return cudaGetLastError();
}

// role of the winning DB line into the role half of every occupied slot; one thread per pair of slots
__global__ void __launch_bounds__(256) line_finalize_kernel(LineTable t, const unsigned long long* __restrict__ best, uint32_t role_bits) {
    const unsigned long long role_mask = (1ull << role_bits) - 1;
    uint32_t* words = reinterpret_cast<uint32_t*>(const_cast<uint4*>(t.lines));
    const unsigned long long n_pairs = (unsigned long long)t.n_lines * 16;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long pr = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; pr < n_pairs; pr += stride) {
        const unsigned long long sector = pr >> 2, wj = pr & 3;
        const uint32_t tags = words[sector * 8 + wj];
        if (tags == 0u) continue;
        uint32_t r = 0;
        if (tags & 0xFFFFu) r |= (uint32_t)(best[sector * 8 + 2 * wj] & role_mask);
        if (tags >> 16) r |= (uint32_t)(best[sector * 8 + 2 * wj + 1] & role_mask) << 16;
        words[sector * 8 + 4 + wj] = r;
    }
}

__global__ void line_finalize_ovf_kernel(Slot128* slots, unsigned long long n_slots, uint32_t role_bits) {
    const unsigned long long role_mask = (1ull << role_bits) - 1;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long s = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += stride)
        if (slots[s].key) slots[s].val &= role_mask;
}

cudaError_t launch_line_finalize(const LineTable& t, const unsigned long long* best, uint32_t role_bits, cudaStream_t st) {
    line_finalize_kernel<<<148 * 16, 256, 0, st>>>(t, best, role_bits);
    if (t.ovf)
        line_finalize_ovf_kernel<<<148 * 4, 256, 0, st>>>(reinterpret_cast<Slot128*>(const_cast<uint4*>(t.ovf)),
                                                          2ull << t.ovf_bbits, role_bits);
    return cudaGetLastError();
}

}  // namespace ka
