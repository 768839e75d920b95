// ka_common.cuh — shared device/host definitions of the k-mer annotation engine (sm_100a).
//
// Data layout in HBM
//   table   : n_buckets x 32-byte buckets; a bucket is ONE DRAM sector holding two 16-byte
//             slots { u64 key ; u64 val } with val = (db line index << 32) | role id.
//             key == 0 means empty (a packed k-mer is never 0: every 5-bit code is 1..31).
//             Slots fill in probe order (slot 0, slot 1, next bucket ...) and are never
//             freed, so a lookup stops at the first empty slot.
//   batch   : CSR — residues u8[R] (+ padding), offsets u64[N+1], results i32/i32/u8 per sequence.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ka {

constexpr int KMAX = 12;            // 5 bits x 12 = 60 bits
constexpr int TILE_THREADS = 256;   // threads per CTA of the tile kernel
constexpr int POS_PER_THREAD = 8;   // window positions per thread per pass
constexpr int PASS_POS = TILE_THREADS * POS_PER_THREAD;  // 2048 positions per pass
constexpr int MAX_TILE_SEQ = 512;   // sequences handled per tile sub-batch
constexpr uint32_t TOKEN_EMPTY = 0u;

struct Slot {
    unsigned long long key;
    unsigned long long val;  // hi 32: db line index, lo 32: role id
};

struct TableView {
    const uint4* buckets;     // n_buckets * 2 uint4 (slot0, slot1)
    unsigned long long n_buckets;
    unsigned long long key_mask;  // (1 << 5K) - 1
    int K;
};

// 64-bit finaliser (two xor-shift-multiply rounds); the bucket is the high part of
// hash * n_buckets, so n_buckets need not be a power of two.
__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    return x;
}

__device__ __forceinline__ unsigned long long bucket_of(unsigned long long key,
                                                        unsigned long long n_buckets) {
    return __umul64hi(mix64(key), n_buckets);
}

// One 32-byte bucket with a single 256-bit load (LDG.E.256 on sm_100a), read-only path,
// no L1 allocation: a bucket is touched once per probe and must not evict the residue tile.
__device__ __forceinline__ void load_bucket(const uint4* p, uint4& s0, uint4& s1) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(s0.x), "=r"(s0.y), "=r"(s0.z), "=r"(s0.w), "=r"(s1.x), "=r"(s1.y),
                   "=r"(s1.z), "=r"(s1.w)
                 : "l"(p));
}

__device__ __forceinline__ unsigned long long u64_of(uint32_t lo, uint32_t hi) {
    return (unsigned long long)lo | ((unsigned long long)hi << 32);
}

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
