// ka_kernels.cuh — launcher interface between the engine (ka_engine.cu) and the kernels
// (ka_kernels.cu).  Plain pointers only; every launcher enqueues on the given stream and
// returns the cudaError_t of the launch.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "ka_common.cuh"

namespace ka {

struct BigItem {
    uint32_t seq;                 // sequence index inside the chunk
    uint32_t pad;
    unsigned long long tok_base;  // first token of its de-dup region in the scratch array
};

struct AnnotParams {
    const uint32_t* pk;               // non-NULL: the chunk's residues as the 5-bit code stream (ka_line.cuh); the tile
    uint32_t pk_lead;                 //   kernels stage from it: residue `base` + g sits at bits [5 (g + pk_lead), +5)
    const uint8_t* res;               // residues of the chunk; res[0] is absolute offset `base`
    const unsigned long long* off;    // absolute offsets, n_seq + 1
    unsigned long long base;
    uint32_t n_seq;
    uint32_t n_tiles;                 // floor(R / tile_span) + 1
    uint32_t tile_span;               // residues of sequence starts per tile
    uint32_t long_seq;                // L > long_seq leaves the residue tiles: one tile of its own (mid) or big_kernel
    uint32_t mid_seq;                 // long_seq < L <= mid_seq: a single-sequence tile in the second tile launch
    uint4* mid_desc;                  // descriptors of those single-sequence tiles
    uint32_t* mid_count;              // [0] how many
    uint32_t ext_max;                 // tile_span + long_seq: max residue extent of a tile
    uint32_t res_bytes;               // smem bytes reserved for the residue stage
    uint4* first;                     // n_tiles descriptors {first seq, n seqs (long tail removed), g0, g1}
    TableView tab;
    const uint8_t* lut;               // 256-byte residue -> 5-bit code table (0 = not in DB alphabet)
    int32_t min_hits;
    int32_t* out_role;
    int32_t* out_hits;
    uint8_t* out_flag;
    uint32_t* big_count;              // [0] number of long sequences
    unsigned long long* tok_cursor;   // [0] tokens handed out
    BigItem* big_list;
    uint32_t* scratch;                // de-dup tokens of the long sequences (8-byte tokens for wide tables)
    uint32_t* dbg;                    // KA_DEBUG builds: [0] OR of the codes of failed bounds checks
    // routed mode (table_mode 2): the tile kernel either only EXTRACTS the mixed key of every window
    // position into route_keys[chunk-relative residue index] (ROUTE_INVALID = no window), or TALLIES
    // from route_ans[route_slot[same index]] = (role << 32 | de-dup token) answered by the owning GPU
    // (route_slot = the send slot the key was bucketed into; answers come back in send order).
    unsigned long long* route_keys;
    const unsigned long long* route_ans;
    const uint32_t* route_slot;
};

// tile kernel shapes: 0 = 4 window positions per thread x 128 threads (7 CTAs/SM, the residue tiles),
// 1 = 4 x 256 (the single-sequence tiles of the second launch).  Shapes that were measured and
// dropped (8 x 256, 2 x 128, 64-register caps, 512 threads) are listed in profiles/r01_summary.md.
constexpr int N_VARIANTS = 2;
size_t tile_smem_bytes(uint32_t ext_max, uint32_t* res_bytes_out, bool wide);
// de-dup token capacity of x window positions: x + x/4 (worst-case load factor 0.8)
__host__ __device__ inline uint32_t tok_cap(uint32_t x) { return x + (x >> 2); }
cudaError_t tile_kernel_set_smem(int cls, int variant, size_t bytes);

constexpr unsigned long long ROUTE_INVALID = ~0ull;
constexpr unsigned long long ROUTE_MISS = ~0ull;          // answer of a key that is not in the table
// mode 0 = probe the table, 1 = extract keys only, 2 = tally from routed answers (shapes 0 and 1 only;
// also the entry of every launch on a wide table)
cudaError_t launch_tiles_mode(const AnnotParams& p, int variant, int mode, size_t smem, cudaStream_t st);
cudaError_t tile_kernel_mode_set_smem(size_t bytes);
// Where the bucketed keys go: p[o] is indexed with the sender's send slot w (offsets[o] <= w < offsets[o] + count[o]):
// the local send buffer for every owner (NCCL transport), or the owner's receive buffer shifted so that slot w
// lands in this sender's region of it (peer-store transport: the scatter kernel IS the key exchange).
struct RouteDst { unsigned long long* p[8]; };
// Where the owner's answers go: n_regions = 0: ans[i]; else region q = keys first[q] .. first[q+1) came from sender q
// and p[q] is that sender's answer buffer shifted so that received key i lands at its send slot (peer store).
struct RouteAns { unsigned long long* p[8]; unsigned long long first[8]; uint32_t n_regions; };
// per-owner counts of the valid keys of keys[0..n) (owner = sector >> shard_shift); counts[8] accumulates
cudaError_t launch_route_count(const unsigned long long* keys, unsigned long long n, TableView tab,
                               unsigned long long* counts, cudaStream_t st);
// bucket the valid keys by owner: send_keys at offsets[o] + running cursor[o]; slot_of_pos[i] = that index
cudaError_t launch_route_scatter(const unsigned long long* keys, unsigned long long n, TableView tab,
                                 const unsigned long long* offsets, unsigned long long* cursor,
                                 const RouteDst& dst, uint32_t* slot_of_pos, cudaStream_t st);
// owner side: answer every received key from the local shard
cudaError_t launch_route_lookup(const unsigned long long* keys, unsigned long long n, TableView tab,
                                unsigned long long* ans, const RouteAns& ra, cudaStream_t st);

cudaError_t launch_plan(const AnnotParams& p, cudaStream_t st);
cudaError_t launch_tiles(const AnnotParams& p, int variant, size_t smem, cudaStream_t st);
cudaError_t launch_big(const AnnotParams& p, int grid, cudaStream_t st);

cudaError_t launch_alphabet_scan(const uint8_t* bytes, unsigned long long n, uint32_t* bitmap8,
                                 cudaStream_t st);
// errs[0] = k-mers with a byte outside the alphabet, errs[1] = negative role ids,
// errs[2] = keys that found the overflow table full,
// counters[0] = distinct keys stored, counters[1] = longest sector chain.
// Every slot keeps the maximum of (line + 1) << role_bits | role over the lines of its key while
// the DB streams through: cls 32/64 in best[slot] (8 bytes per primary slot, build-time only),
// whole-key slots (cls 128, overflow table) in their own value word; launch_db_finalize then writes
// the plain role of the last line into every slot.
cudaError_t launch_db_insert(const TableView& tab, const uint8_t* kmers, const int32_t* roles,
                             unsigned long long n, unsigned long long line_base,
                             const uint8_t* lut, unsigned long long* best, uint32_t role_bits,
                             unsigned long long* counters, uint32_t* errs, cudaStream_t st);
cudaError_t launch_db_finalize(const TableView& tab, const unsigned long long* best, uint32_t role_bits, cudaStream_t st);
// synthetic DB lines [first, first + n) for ka_db_load_synthetic (see include/kmeranno.h)
cudaError_t launch_db_generate(unsigned long long first, unsigned long long n, int K, unsigned long long seed,
                               uint32_t n_roles, uint8_t* kmers, int32_t* roles, cudaStream_t st);

// ---- build (BuildKmerProcessor.java:138-223) ----
// table: n_slots (power of two) Slot128 {key, val}; val = role + 1, bit 62 = seen under two
// roles (RoleCounter.badCount > 0), bit 63 = occurs in a zero-role peg.
constexpr unsigned long long BUILD_BAD = 1ull << 62, BUILD_DEAD = 1ull << 63;
// pass = 1: insert the windows of single-role pegs; pass = 2: mark the windows of zero-role pegs
cudaError_t launch_build_pass(int pass, const uint8_t* res, const unsigned long long* off, uint32_t n_seq,
                              const int32_t* n_roles, const int32_t* peg_role, int K, const uint8_t* lut,
                              Slot128* table, unsigned long long n_slots, cudaStream_t st);
// emit the surviving k-mers (unordered); counter[0] = number found (may exceed cap)
cudaError_t launch_build_emit(const Slot128* table, unsigned long long n_slots, int K, const uint8_t* inv_lut,
                              unsigned long long cap, uint8_t* out_kmers, int32_t* out_roles,
                              unsigned long long* counter, cudaStream_t st);

// ---- pairwise k-mer distance (GeneCopyProcessor.java:137-142), ka_distance.cu ----
struct DistParams {
    const uint8_t* res;                    // residues of the batch; res[0] is absolute offset `base`
    const unsigned long long* off;         // absolute offsets, n_seq + 1
    unsigned long long base;
    uint32_t n_seq;
    int K;
    unsigned long long key_mask;           // (1 << 5K) - 1
    const uint8_t* lut;                    // residue byte -> 5-bit code (every byte of the batch has one)
    int32_t* set_size;                     // [n_seq] distinct K-windows of every sequence
    const uint32_t* query_seq;             // [Q] sequence index of every query
    const unsigned long long* group_off;   // [Q + 1] candidates of query q: cand_seq[group_off[q] .. group_off[q+1])
    uint32_t q_begin, q_end;               // queries handled by this launch
    const uint32_t* cand_seq;
    int32_t* common;                       // [M] |A ∩ B|
    double* dist;                          // [M]
    uint32_t smem_cap;                     // entries of the shared-memory hash set (power of two)
    unsigned long long* scratch_keys;      // hash sets of the sequences too long for shared memory
    uint8_t* uniq;                         // [residues] 1 = this window is the marked occurrence of its k-mer
    const unsigned long long* seq_scratch;   // [n_seq] first scratch entry of a sequence's set (set_size_kernel)
    const unsigned long long* query_scratch; // [Q] first scratch entry of a query's set (common_kernel)
};
// hash-set capacity for a sequence with `windows` K-windows: power of two >= 2 * windows, at least 64
__host__ __device__ inline uint32_t dist_set_cap(uint32_t windows) {
    uint32_t c = 64;
    while (c < 2u * windows && c < 0x80000000u) c <<= 1;
    return c;
}
size_t dist_smem_bytes(uint32_t smem_cap);
cudaError_t dist_set_smem(size_t bytes);
cudaError_t launch_set_size(const DistParams& p, int sm_count, cudaStream_t st);
cudaError_t launch_common(const DistParams& p, int sm_count, cudaStream_t st);

cudaError_t launch_random_probe(const uint4* buf, unsigned long long n_slots, int slot_bytes,
                                unsigned long long n_probes, unsigned long long seed,
                                unsigned long long* sink, cudaStream_t st);
cudaError_t launch_fill_random(uint4* buf, unsigned long long n_uint4, cudaStream_t st);

}  // namespace ka
