// ka_common.cuh — shared device/host definitions of the k-mer annotation engine (sm_100a).
//
// Data layout in HBM
//   table   : 2^bbits sectors of 32 bytes.  A lookup reads exactly ONE sector (one 256-bit
//             load); B200 fetches the enclosing 128-byte line from DRAM, so one probe = one
//             DRAM line.  The packed k-mer (5 bits per residue, w = 5K bits, every field
//             1..31 so a key is never 0) goes through a bijective mixer on w bits; the top
//             bbits select the sector and only the remaining rem_bits are stored
//             (quotienting), which makes three slot widths possible:
//               cls 32  : 8 slots/sector, slot = rem | (role+1) << rem_bits   (e.g. K=8, 1e8 keys)
//               cls 64  : 4 slots/sector, same packing in 64 bits             (K=10/12)
//               cls 128 : 2 slots/sector, { u64 key ; u64 (db line << 32 | role) } (fallback)
//             0 = empty slot.  Slots fill in order inside a sector and are never freed.
//             cls 32/64: a key whose home sector is full goes to a small OVERFLOW table that
//             stores the whole mixed key (a remainder alone would be ambiguous outside its
//             home sector); a lookup reads it only when the home sector is full and has no
//             match (~1 % of probes at the default load factor; the table is L2 resident).
//             cls 128 stores whole keys, so it simply chains to the next sector.
//             WIDE tables (more than 2^32 - 16 slots, e.g. > 34 GB of 64-bit slots, or forced by
//             the "wide" option) use 64-bit sector indices and the mixed key itself as the de-dup
//             token; the narrow form keeps 32-bit sector indices and slot-index tokens in registers.
//   batch   : CSR — residues u8[R] (+ padding), offsets u64[N+1], results i32/i32/u8 per sequence.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ka {

constexpr int KMAX = 12;            // 5 bits x 12 = 60 bits
constexpr int MAX_TILE_SEQ = 256;   // sequences handled per tile sub-batch
constexpr uint32_t TOKEN_EMPTY = 0u;

struct Slot128 {
    unsigned long long key;
    unsigned long long val;  // hi 32: db line index, lo 32: role id
};

struct TableView {
    const uint4* sectors;         // 2^bbits sectors, 2 uint4 each
    const uint4* ovf;             // overflow table (cls 32/64): 2^ovf_bbits sectors of 2 Slot128
    uint32_t ovf_bbits;
    uint32_t n_primary_slots;     // slots of the primary table (de-dup tokens of overflow entries start here; narrow only)
    uint32_t wide;                // 1 = 64-bit sector indices and tokens (see above)
    // sharded mode (table larger than one GPU's share): sector s lives on shard s >> shard_shift at
    // local index s & shard_mask; shard_sectors[i] / shard_ovf[i] are PEER pointers (NVLink loads).
    // n_shards == 1: `sectors` / `ovf` are used directly.
    const uint4* const* shard_sectors;
    const uint4* const* shard_ovf;
    uint32_t n_shards;
    uint32_t shard_shift;
    uint32_t my_shard;            // build only: the shard this device fills
    unsigned long long key_mask;  // (1 << 5K) - 1
    unsigned long long rem_mask;  // (1 << rem_bits) - 1   (cls 32 / 64)
    uint32_t bbits;               // log2(number of sectors)
    uint32_t rem_bits;            // 5K - bbits            (cls 32 / 64)
    uint32_t wbits;               // 5K
    int K;
    int cls;                      // 32, 64 or 128
};

// Bijection on w-bit integers (xor-shift and odd multiplication are both invertible
// modulo 2^w), mixing every input bit into the top bits that select the sector.
__host__ __device__ __forceinline__ unsigned long long mixw(unsigned long long k, uint32_t w,
                                                            unsigned long long mask) {
    // two xor-shift-multiply rounds: sector occupancy of 1.25e7 synthetic 8- and 12-mers is
    // indistinguishable from Poisson already after one (var/mean 1.000, tail P(>8) 0.00366)
    const uint32_t s = (w + 1) >> 1;
    k ^= k >> s;
    k = (k * 0xCA5A826395121157ull) & mask;
    k ^= k >> s;
    k = (k * 0x9E3779B97F4A7C15ull) & mask;
    return k;
}

__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    return x;
}

// (sector, stored remainder) of a packed key.  For cls 128 the remainder is the key itself.
__host__ __device__ __forceinline__ void locate(const TableView& t, unsigned long long key,
                                                uint32_t& sector, unsigned long long& rem) {
    if (t.cls == 128) {
        sector = t.bbits ? (uint32_t)(mix64(key) >> (64 - t.bbits)) : 0u;
        rem = key;
    } else {
        const unsigned long long m = mixw(key, t.wbits, t.key_mask);
        sector = (uint32_t)(m >> t.rem_bits);
        rem = m & t.rem_mask;
    }
}

// address of a table sector: local HBM, or the owning GPU's HBM through NVLink peer memory
// (SEC = uint32_t for narrow tables, unsigned long long for wide ones)
template <typename SEC>
__device__ __forceinline__ const uint4* sector_ptr(const TableView& t, SEC sec) {
    if (t.n_shards <= 1) return t.sectors + 2 * (size_t)sec;
    const uint4* base = reinterpret_cast<const uint4*>(
        __ldg(reinterpret_cast<const unsigned long long*>(t.shard_sectors) + (sec >> t.shard_shift)));
    return base + 2 * (size_t)(sec & (((SEC)1 << t.shard_shift) - 1));
}

// shard owning a mixed key (sharded tables are quotiented: sector = mixed >> rem_bits)
__host__ __device__ __forceinline__ uint32_t shard_of(const TableView& t, unsigned long long mixed) {
    return (uint32_t)((mixed >> t.rem_bits) >> t.shard_shift);
}

// One 32-byte sector with a single 256-bit load (LDG.E.256 on sm_100a), read-only path, no
// L1 allocation: a sector is touched once per probe and must not evict the residue tile.
__device__ __forceinline__ void load_sector(const uint4* p, uint4& s0, uint4& s1) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(s0.x), "=r"(s0.y), "=r"(s0.z), "=r"(s0.w), "=r"(s1.x), "=r"(s1.y),
                   "=r"(s1.z), "=r"(s1.w)
                 : "l"(p));
}

__device__ __forceinline__ unsigned long long u64_of(uint32_t lo, uint32_t hi) {
    return (unsigned long long)lo | ((unsigned long long)hi << 32);
}

// Match one loaded sector against a stored remainder.
//   returns role >= 0 and the slot index inside the sector on a hit; -1 on a miss;
//   `full` tells whether the sector has no free slot (the chain continues).
template <int CLS>
__device__ __forceinline__ int match_sector(const TableView& t, const uint4& a, const uint4& b,
                                            unsigned long long rem, uint32_t& slot_in_sector,
                                            bool& full) {
    int role = -1;
    if (CLS == 32) {
        const uint32_t r = (uint32_t)rem, m = (uint32_t)t.rem_mask;
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (w[j] != 0 && ((w[j] ^ r) & m) == 0) { role = (int)(w[j] >> t.rem_bits) - 1; slot_in_sector = j; }
        full = w[7] != 0;
    } else if (CLS == 64) {
        const unsigned long long w[4] = {u64_of(a.x, a.y), u64_of(a.z, a.w), u64_of(b.x, b.y), u64_of(b.z, b.w)};
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (w[j] != 0 && ((w[j] ^ rem) & t.rem_mask) == 0) { role = (int)(w[j] >> t.rem_bits) - 1; slot_in_sector = j; }
        full = w[3] != 0;
    } else {
        const unsigned long long k0 = u64_of(a.x, a.y), k1 = u64_of(b.x, b.y);
        if (k0 == rem) { role = (int)a.z; slot_in_sector = 0; }
        else if (k1 == rem) { role = (int)b.z; slot_in_sector = 1; }
        full = k1 != 0;
    }
    return role;
}

template <int CLS>
__host__ __device__ constexpr int slots_per_sector() { return CLS == 32 ? 8 : (CLS == 64 ? 4 : 2); }

// register type of a stored remainder: 32 bits are enough for 32-bit slots
template <int CLS> struct rem_type { typedef unsigned long long type; };
template <> struct rem_type<32> { typedef uint32_t type; };

// ---- mbarrier + 1-D bulk async copy (TMA engine, UBLKCP) -------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
// dst, src 16-byte aligned; bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

}  // namespace ka
