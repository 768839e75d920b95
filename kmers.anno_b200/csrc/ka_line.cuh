// ka_line.cuh — the 128-byte-line table ("slot class 16") and the packed residue stream.
//
// Why: on B200 every L2 miss fetches a whole 128-byte line (profiles/r01_summary.md §D), so the
// unit of the table is the line, and most probes of an annotation batch are ABSENT k-mers, which a
// small L2-resident filter can answer without touching HBM (microbench/filter_probe.cu: 54 -> 93 G
// probes/s memory-side at a 35 % pass rate).
//
//   keys    radix-n integers: the n distinct bytes of the DB get digits 0..n-1 in byte order,
//           key = sum d_j * n^(K-1-j) < n^K, w = bit length of n^K - 1 (35 bits for 20^8 instead of
//           the 40 of 5-bit fields).  A 4-round Feistel network on the two halves of the w bits
//           (32-bit multiplies) is the bijective mixer m = mix(key).
//   lines   B = c * 2^s lines of 128 bytes, c in 8..15 (so the load factor is controllable to ~10 %
//           with power-of-two-cheap indexing): the top u bits U of m choose the c-way part
//           hi = (U * c) >> u — and idx = ((U * c) mod 2^u) / c tells which of the <= 2^(u-3)
//           values of U with the same hi it was —, the next s bits the line inside the part, the
//           next 2 bits the HOME SECTOR of the line, the rest (rem0) stays.  (hi, idx) is a
//           bijection of U, so (line, home, rem = idx : rem0) identifies the key: quotienting.
//   sector  32 bytes = 8 slots: words 0-3 hold eight 16-bit TAGS, words 4-7 eight 16-bit ROLES.
//           tag = 1 valid | 2 home | 1 flag | 1 spare | 11 rem.  A lookup loads the home sector
//           (one 256-bit load = the DRAM fetch of the whole line), compares all eight tags with three
//           SIMD-in-register operations per word, and reads the role of the matching slot.
//   spill   a key whose home sector is full lives in another sector of the SAME line (its tag
//           keeps the home bits); the flag bits of slots 0,1,2 of the home sector say which of the
//           sectors home^1, home^2, home^3 hold such keys, so the second-stage loads are L2 hits
//           on the line that was just fetched.  A key whose whole line is full goes to the small
//           overflow table (flag bit of slot 3) under (sector index : rem) + 1.
//   filter  one 32-bit word per sector, two bits per key (of its HOME sector): no false negatives.
//           B*16 bytes (77 MB for 1e8 8-mers at load factor 0.65) — L2 resident next to a table
//           whose lines are read with an evict-first hint.
//   stream  residues travel and are staged as 5-bit codes (digit 0..n-1, 31 = byte not in the DB
//           alphabet): residue r of the batch occupies bits [5r, 5r+5) of a little-endian byte
//           stream — 0.625 bytes per residue over PCIe instead of 1.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ka {

constexpr uint32_t CODE_INVALID = 31u;      // 5-bit code of a byte that is not in the DB alphabet
constexpr uint32_t TAG_VALID = 0x8000u;
constexpr uint32_t TAG_FLAG = 0x1000u;      // per-slot flag bit, ignored by the compare
constexpr uint32_t TAG_CMP = 0xEFFFu;
constexpr uint32_t TAG_REM_BITS = 11;
constexpr int LINE_MAX_SEQ = 64;            // sequences per tile sub-batch of the line kernels

struct LineTable {
    const uint4* lines;          // n_lines * 8 uint4
    const uint32_t* filt;        // one word per sector (n_lines * 4), NULL = no filter
    const uint4* ovf;            // overflow table: 2^ovf_bbits sectors of 2 {key + 1, role} entries
    uint32_t ovf_bbits;
    uint32_t n_lines;            // c << s
    uint32_t c, s, u, inv_c;     // inv_c = ceil(65536 / c)
    uint32_t wbits, wl, wh;      // key bits and their Feistel halves (wl low, wh high)
    uint32_t low_bits;           // w - u - s: 2 home bits + rem0
    uint32_t rem0_bits;
    uint32_t radix;              // number of DB symbols
    unsigned long long pow_k1;   // radix^(K-1)
    int K;
};

__host__ __device__ __forceinline__ unsigned long long line_mix(unsigned long long key, uint32_t wl, uint32_t wh) {
    uint32_t L = (uint32_t)(key >> wl), R = (uint32_t)key & ((1u << wl) - 1u);   // wl <= 30
    // (x * C) >> (32 - bits) through a 64-bit shift: well defined for bits == 0
    L ^= (uint32_t)((unsigned long long)(R * 0x9E3779B1u) >> (32 - wh));
    R ^= (uint32_t)((unsigned long long)(L * 0x85EBCA6Bu) >> (32 - wl));
    L ^= (uint32_t)((unsigned long long)(R * 0xC2B2AE35u) >> (32 - wh));
    R ^= (uint32_t)((unsigned long long)(L * 0x27D4EB2Fu) >> (32 - wl));
    return ((unsigned long long)L << wl) | R;
}

// key -> (global sector index of the home sector, 16-bit tag)
__host__ __device__ __forceinline__ void line_locate(const LineTable& t, unsigned long long key,
                                                     uint32_t& sector, uint32_t& tag) {
    const unsigned long long m = line_mix(key, t.wl, t.wh);
    const uint32_t U = (uint32_t)(m >> (t.wbits - t.u));
    const uint32_t x = U * t.c;
    const uint32_t hi = x >> t.u, frac = x & ((1u << t.u) - 1u);
    const uint32_t idx = (frac * t.inv_c) >> 16;                 // frac / c, exact for frac < 4096, c < 16
    const uint32_t low = (uint32_t)m & ((1u << t.low_bits) - 1u);
    const uint32_t line = (hi << t.s) | ((uint32_t)(m >> t.low_bits) & ((1u << t.s) - 1u));
    const uint32_t home = low & 3u;
    tag = TAG_VALID | (home << 13) | (idx << t.rem0_bits) | (low >> 2);
    sector = line * 4u + home;
}

// the two filter bits of a key inside the word of its home sector
__host__ __device__ __forceinline__ uint32_t line_filter_bits(uint32_t tag) {
    const uint32_t h = (tag & ((1u << TAG_REM_BITS) - 1u)) * 0x9E3779B1u;
    return (1u << (h >> 27)) | (1u << ((h >> 22) & 31u));
}

// key under which a (sector, tag) pair lives in the overflow table; never 0
__host__ __device__ __forceinline__ unsigned long long line_ovf_key(uint32_t sector, uint32_t tag) {
    return (((unsigned long long)sector << TAG_REM_BITS) | (tag & ((1u << TAG_REM_BITS) - 1u))) + 1ull;
}

struct LineParams {
    const uint32_t* pk;               // packed codes of the chunk: chunk-relative residue g at bits [5g, 5g+5)
    uint32_t* off;                    // chunk-relative offsets (n_seq + 1), written by the plan kernel
    uint32_t n_seq, n_tiles;
    uint32_t tile_span, long_seq, mid_seq, ext_max;
    uint32_t stage_bytes;             // shared-memory bytes reserved for the packed stage
    uint4* first;                     // n_tiles descriptors {first seq, n seqs, g0, g1}
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

size_t line_tile_smem_bytes(uint32_t ext_max, uint32_t* stage_bytes_out);
cudaError_t line_tile_set_smem(size_t bytes);

// offsets -> chunk-relative u32 offsets + tile descriptors + lists of mid / long sequences.
// off64 != NULL: absolute 64-bit offsets; else off32: absolute 32-bit offsets.  `origin` = absolute
// residue index of chunk-relative residue 0.
cudaError_t launch_line_plan(const LineParams& p, const unsigned long long* off64, const uint32_t* off32,
                             unsigned long long origin, cudaStream_t st);
cudaError_t launch_line_tiles(const LineParams& p, size_t smem, cudaStream_t st);
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
