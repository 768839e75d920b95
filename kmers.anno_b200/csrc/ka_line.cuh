// ka_line.cuh — the 128-byte-line table ("slot class 16") and the packed residue stream.
//
// Why: on B200 every L2 miss fetches a whole 128-byte line (profiles/r01_summary.md §D), so the
// unit of the table is the line, and most probes of an annotation batch are ABSENT k-mers, which a
// small L2-resident filter can answer without touching HBM (microbench/filter_probe.cu: 54 -> 93 G
// probes/s memory-side at a 35 % pass rate).
//
//   keys    radix-n digits: the n distinct bytes of the DB get digits 0..n-1 in byte order.  A k-mer is
//           the PAIR (H, Lo) of the radix-n values of its first Kh = K/2 and last Kl = K - Kh digits,
//           bh and bl bits wide (18 + 18 = 36 bits for 20^8 instead of the 40 of 5-bit fields); all
//           arithmetic stays in 32 bits.  A 3-round Feistel network on the two halves (one 32-bit
//           multiply per round) is the bijective mixer (L, R) = mix(H, Lo).
//   lines   B = c * 2^s lines of 128 bytes, c in 8..15 (load factor controllable to ~10 % with
//           power-of-two-cheap indexing): the top u bits U of L choose the c-way part
//           hi = (U * c) >> u — and idx = ((U * c) mod 2^u) / c tells which of the <= 2^(u-3)
//           values of U with the same hi it was —, the other la bits of L and the low a bits of R
//           the line inside the part (la + a = s), the next 2 bits of R the HOME SECTOR of the line,
//           and rem = idx : (R >> a) stays (<= 14 bits).  (hi, idx) is a bijection of U, so
//           (line, rem) identifies the key: quotienting.
//   sector  32 bytes = 8 slots: words 0-3 hold eight 16-bit TAGS, words 4-7 eight 16-bit ROLES.
//           tag = 1 valid | 1 flag | 14 rem.  A lookup loads the home sector (one 256-bit load =
//           the DRAM fetch of the whole line), compares all eight tags with three SIMD-in-register
//           operations per word, and reads the role of the matching slot.
//   spill   a key whose home sector is full lives in another sector of the SAME line (its rem
//           keeps the home bits); the flag bits of slots 0,1,2 of the home sector say which of the
//           sectors home^1, home^2, home^3 hold such keys, so the second-stage loads are L2 hits
//           on the line that was just fetched.  A key whose whole line is full goes to the small
//           overflow table (flag bit of slot 3) under (sector index : rem) + 1.
//   filter  a Bloom filter of n_filt 32-bit words (one per table sector: B*16 bytes, 75 MB for 1e8
//           8-mers), two bits per key in ONE word chosen by a cheap hash of the raw halves
//           (two multiplies and a xor — no mixer, no table geometry), so the windows that are
//           absent cost a few instructions and one L2 access; only the survivors pay
//           the mixer, the line arithmetic and the HBM access.  No false negatives.
//   stream  residues travel and are staged as 5-bit codes (digit 0..n-1, 31 = byte not in the DB
//           alphabet): residue r of the batch occupies bits [5r, 5r+5) of a little-endian byte
//           stream — 0.625 bytes per residue over PCIe instead of 1.
//   passes  a chunk is annotated by three kernels with a list in global memory between them
//           (ka_line.cu): FILTER (one warp per tile: rolling keys, filter words, survivors appended to the
//           tile's list), PROBE (one warp per tile: the survivors' home sectors — the only HBM access of a
//           probe —, hits written back over the front of the list), TALLY (one CTA per tile: distinct
//           k-mers per sequence, unanimous-role call).  Each pass is bound by something else (issue
//           slots / HBM lines / shared-memory atomics) and none waits for another inside a kernel.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ka {

constexpr uint32_t CODE_INVALID = 31u;      // 5-bit code of a byte that is not in the DB alphabet
constexpr uint32_t TAG_VALID = 0x8000u;
constexpr uint32_t TAG_FLAG = 0x4000u;      // per-slot flag bit, ignored by the compare
constexpr uint32_t TAG_CMP = 0xBFFFu;
constexpr uint32_t TAG_REM_BITS = 14;
constexpr int LINE_MAX_SEQ = 64;            // sequences per tile sub-batch of the line kernels
constexpr uint32_t LINE_MID_SEG = 1024;   // a single-sequence (mid) tile is cut into segments of at most this many positions
constexpr int LINE_KMAX = 10;               // the three 64-bit code windows of a lane hold K + 7 <= 17 codes... (Kh <= 5)

struct LineTable {
    const uint4* lines;          // n_lines * 8 uint4
    const uint32_t* filt;        // Bloom filter words, NULL = no filter
    uint32_t n_filt;             // number of filter words (n_lines * 4)
    const uint4* ovf;            // overflow table: 2^ovf_bbits sectors of 2 {key + 1, role} entries
    uint32_t ovf_bbits;
    uint32_t n_lines;            // c << s
    uint32_t c, s, u, inv_c;     // inv_c = ceil(65536 / c)
    uint32_t bh, bl;             // bits of the two key halves
    uint32_t la, a, r;           // la = bh - u bits of L and a = s - la bits of R index the line; r = bl - a bits of R stay
    uint32_t radix;              // number of DB symbols
    uint32_t Kh, Kl;             // digits of the two halves
    uint32_t pw_h, pw_l;         // radix^(Kh-1), radix^(Kl-1)
    int K;
};

// 3-round Feistel network on (bh, bl)-bit halves: a bijection whatever the round functions are
__host__ __device__ __forceinline__ void line_mix(uint32_t& L, uint32_t& R, uint32_t bh, uint32_t bl) {
    L ^= (R * 0x9E3779B1u) >> (32u - bh);       // 3 <= bh, bl <= 30
    R ^= (L * 0x85EBCA6Bu) >> (32u - bl);
    L ^= (R * 0xC2B2AE35u) >> (32u - bh);
}

// key halves -> (global sector index of the home sector, 16-bit tag)
__host__ __device__ __forceinline__ void line_locate(const LineTable& t, uint32_t H, uint32_t Lo,
                                                     uint32_t& sector, uint32_t& tag) {
    uint32_t L = H, R = Lo;
    line_mix(L, R, t.bh, t.bl);
    const uint32_t x = (L >> t.la) * t.c;
    const uint32_t hi = x >> t.u;
    const uint32_t idx = ((x & ((1u << t.u) - 1u)) * t.inv_c) >> 16;   // (x mod 2^u) / c, exact for u <= 12, c < 16
    const uint32_t line = (hi << t.s) | ((L & ((1u << t.la) - 1u)) << t.a) | (R & ((1u << t.a) - 1u));
    const uint32_t rest = R >> t.a;                                     // r bits: home in the low two
    sector = line * 4u + (rest & 3u);
    tag = TAG_VALID | (idx << t.r) | rest;
}

// filter hash of the raw key halves: word = top bits scaled to n_filt, the two bits from the low bits
__host__ __device__ __forceinline__ uint32_t line_filter_hash(uint32_t H, uint32_t Lo) {
    return (H * 0x9E3779B1u) ^ (Lo * 0x85EBCA6Bu);
}
__host__ __device__ __forceinline__ uint32_t line_filter_word(uint32_t fh, uint32_t n_filt) {
    return (uint32_t)(((unsigned long long)fh * n_filt) >> 32);
}
__host__ __device__ __forceinline__ uint32_t line_filter_bits(uint32_t fh) {
    return (1u << (fh & 31u)) | (1u << ((fh >> 5) & 31u));
}

// key under which a (sector, tag) pair lives in the overflow table; never 0
__host__ __device__ __forceinline__ unsigned long long line_ovf_key(uint32_t sector, uint32_t tag) {
    return (((unsigned long long)sector << TAG_REM_BITS) | (tag & ((1u << TAG_REM_BITS) - 1u))) + 1ull;
}

struct LineParams {
    const uint32_t* pk;               // packed codes of the chunk: chunk-relative residue g at bits [5g, 5g+5)
    uint32_t* off;                    // chunk-relative offsets (n_seq + 1), written by the plan kernel
    uint32_t n_seq, n_tiles;
    uint32_t tile0, tile1;            // tiles of this launch of the filter / probe / tally passes
    uint32_t n_mid_tiles;             // filter / probe passes: tiles [0, n_mid_tiles) are mid_desc[] (segments of single
                                      // sequences: {seq, count slot, g0, g1}), the others first[tile - n_mid_tiles]
    uint32_t tally_mid;               // tally pass: first[] holds such segments; the CTA of a sequence's first segment tallies all of them
    uint32_t tile_span, long_seq, mid_seq, ext_max;
    uint4* first;                     // n_tiles descriptors {first seq, n seqs, g0, g1}
    uint2* surv;                      // survivors of the filter pass: (H | seq in tile << 26, Lo), the tile's at surv[g0 ...)
    uint32_t* surv_cnt;               // ... and their number, at the index of the tile's (sub-batch's) first sequence
    uint32_t* hit_cnt;                // hits of the probe pass (token, role | seq << 16), compacted over the front of the tile's list
    uint4* mid_desc;
    uint32_t* mid_count;
    uint32_t* big_count;
    unsigned long long* tok_cursor;
    void* big_list;                   // BigItem[]
    uint32_t* scratch;                // de-dup tokens of the long sequences
    LineTable tab;
    int32_t min_hits;
    int32_t* out_role;
    int32_t* out_hits;
    uint8_t* out_flag;
    uint32_t* dbg;
};

size_t line_tile_smem_bytes(uint32_t ext_max);     // dynamic shared memory of the tally pass for tiles of <= ext_max positions
cudaError_t line_tile_set_smem(size_t bytes);

// offsets -> chunk-relative u32 offsets + tile descriptors + lists of mid / long sequences.
// off64 != NULL: absolute 64-bit offsets; else off32: absolute 32-bit offsets.  `origin` = absolute
// residue index of chunk-relative residue 0.
cudaError_t launch_line_plan(const LineParams& p, const unsigned long long* off64, const uint32_t* off32,
                             unsigned long long origin, cudaStream_t st);
// the three passes over tiles [p.tile0, p.tile1): `grid` CTAs stride over them (32, 32 and 128 threads)
cudaError_t launch_line_filter(const LineParams& p, unsigned grid, cudaStream_t st);
cudaError_t launch_line_probe(const LineParams& p, unsigned grid, cudaStream_t st);
cudaError_t launch_line_tally(const LineParams& p, unsigned grid, cudaStream_t st);
cudaError_t launch_line_big(const LineParams& p, int grid, cudaStream_t st);

// residue bytes -> 5-bit codes.  bytes[0] is chunk-relative residue `lead`; residues outside
// [lead, lead + n) get CODE_INVALID.  Writes ceil((lead + n) / 32) * 5 words.
cudaError_t launch_pack(const uint8_t* bytes, uint32_t lead, unsigned long long n, const uint8_t* lut5,
                        uint32_t* out_words, cudaStream_t st);
// 5-bit codes -> residue bytes (inv[code], inv[31] = a byte outside the alphabet): out[j] = residue lead + j
cudaError_t launch_unpack(const uint32_t* words, uint32_t lead, unsigned long long n, const uint8_t* inv32,
                          uint8_t* out, cudaStream_t st);
// absolute u32 offsets -> absolute u64 offsets (old-class kernels behind the packed entry point)
cudaError_t launch_widen_offsets(const uint32_t* off32, unsigned long long n, unsigned long long* off64, cudaStream_t st);

// table build: errs[0] = k-mers with a byte outside the alphabet, errs[1] = negative role ids,
// errs[2] = overflow table full; counters[0] = distinct keys, counters[1] = keys outside their home sector,
// counters[2] = keys in the overflow table.  best[slot] keeps max((line + 1) << role_bits | role).
cudaError_t launch_line_insert(const LineTable& t, const uint8_t* kmers, const int32_t* roles, unsigned long long n,
                               unsigned long long line_base, const uint8_t* lut5, unsigned long long* best,
                               uint32_t role_bits, unsigned long long* counters, uint32_t* errs, cudaStream_t st);
cudaError_t launch_line_finalize(const LineTable& t, const unsigned long long* best, uint32_t role_bits, cudaStream_t st);

}  // namespace ka
