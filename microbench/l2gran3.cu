// l2gran3.cu — which load flavours fetch less than a full 128-byte line per random probe?
//   l2gran3 <GB> <mode>      prints G probes/s; run under ncu for dram bytes / probe.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__host__ __device__ inline unsigned long long mix64(unsigned long long x) {
    x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32; return x;
}
#define LD8(QUAL) asm volatile("ld." QUAL ".v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" \
    : "=r"(r[u][0]),"=r"(r[u][1]),"=r"(r[u][2]),"=r"(r[u][3]),"=r"(r[u][4]),"=r"(r[u][5]),"=r"(r[u][6]),"=r"(r[u][7]) : "l"(p))
#define LD4(QUAL) asm volatile("ld." QUAL ".v4.u32 {%0,%1,%2,%3}, [%4];" \
    : "=r"(r[u][0]),"=r"(r[u][1]),"=r"(r[u][2]),"=r"(r[u][3]) : "l"(p)); r[u][4]=r[u][5]=r[u][6]=r[u][7]=0
const char* names[] = {"global.nc.L1::no_allocate.v8", "global.cg.v8", "global.cv.v8", "volatile.global.v4(16B)",
    "relaxed.gpu.global.v4(16B)", "global.nc.L2::64B.v8", "global.L2::64B.v8", "global.cs.v8", "global.lu.v8",
    "cp.async.cg 2x16B", "atom.add.u64 0", "cp.async.bulk 32B/lane", "global.L1::evict_first.v8", "global.nc.L2::128B.v8"};
template <int MODE>
__global__ void __launch_bounds__(256) probe(const uint4* __restrict__ buf, unsigned long long n_slots,
                                             unsigned long long n_probes, unsigned long long seed, unsigned long long* sink) {
    constexpr bool STAGED = (MODE == 9 || MODE == 11);
    constexpr int U = STAGED ? 4 : 8;
    __shared__ __align__(128) uint4 stage[STAGED ? 256 * 2 * U : 2];
    __shared__ __align__(8) uint64_t bars[8];
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x * U;
    uint32_t acc = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t parity = 0;
    if (MODE == 11) {
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bars[warp])));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    for (unsigned long long i0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * U; i0 < n_probes; i0 += stride) {
        uint32_t r[U][8];
        if (MODE == 11) {
            uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bars[warp]);
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(32 * 32 * U) : "memory");
            __syncwarp();
#pragma unroll
            for (int u = 0; u < U; u++) {
                const unsigned long long idx = __umul64hi(mix64(seed + i0 + u + 1), n_slots);
                uint32_t dst = (uint32_t)__cvta_generic_to_shared(&stage[(threadIdx.x * U + u) * 2]);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];"
                             ::"r"(dst), "l"(buf + 2 * idx), "r"(bar) : "memory");
            }
            uint32_t done = 0;
            while (!done)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }"
                             : "=r"(done) : "r"(bar), "r"(parity) : "memory");
            parity ^= 1;
#pragma unroll
            for (int u = 0; u < U; u++) { uint4 a = stage[(threadIdx.x * U + u) * 2], b = stage[(threadIdx.x * U + u) * 2 + 1]; acc ^= a.x ^ a.w ^ b.y ^ b.z; }
            __syncwarp();
            continue;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const unsigned long long idx = __umul64hi(mix64(seed + i0 + u + 1), n_slots);
            const uint4* p = buf + 2 * idx;
            if (MODE == 0) LD8("global.nc.L1::no_allocate");
            else if (MODE == 1) LD8("global.cg");
            else if (MODE == 2) LD8("global.cv");
            else if (MODE == 3) { LD4("volatile.global"); }
            else if (MODE == 4) { LD4("relaxed.gpu.global"); }
            else if (MODE == 5) LD8("global.nc.L2::64B");
            else if (MODE == 6) LD8("global.L2::64B");
            else if (MODE == 7) LD8("global.cs");
            else if (MODE == 8) LD8("global.lu");
            else if (MODE == 9) {
                uint32_t dst = (uint32_t)__cvta_generic_to_shared(&stage[(threadIdx.x * U + u) * 2]);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(p) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16), "l"(p + 1) : "memory");
            } else if (MODE == 10) {
                unsigned long long old;
                asm volatile("atom.global.add.u64 %0, [%1], 0;" : "=l"(old) : "l"(p) : "memory");
                r[u][0] = (uint32_t)old; r[u][3] = (uint32_t)(old >> 32); r[u][5] = r[u][6] = 0;
            } else if (MODE == 12) LD8("global.L1::evict_first");
            else if (MODE == 13) LD8("global.nc.L2::128B");
        }
        if (MODE == 9) {
            asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
            for (int u = 0; u < U; u++) { uint4 a = stage[(threadIdx.x * U + u) * 2], b = stage[(threadIdx.x * U + u) * 2 + 1]; acc ^= a.x ^ a.w ^ b.y ^ b.z; }
        } else {
#pragma unroll
            for (int u = 0; u < U; u++) acc ^= r[u][0] ^ r[u][3] ^ r[u][5] ^ r[u][6];
        }
    }
    if (acc == 0x9e3779b9u) atomicAdd(sink, 1ull);
}
__global__ void fill(uint4* buf, unsigned long long n) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        unsigned long long h = mix64(i + 0x1234567ull);
        buf[i] = make_uint4((uint32_t)h, (uint32_t)(h >> 32), (uint32_t)i, (uint32_t)(i >> 32));
    }
}
template <int MODE> float once(const uint4* buf, unsigned long long n_slots, unsigned long long n_probes, unsigned long long seed, unsigned long long* sink) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    probe<MODE><<<148 * 8, 256>>>(buf, n_slots, n_probes, seed, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main(int argc, char** argv) {
    double gb = argc > 1 ? atof(argv[1]) : 3.2;
    int mode = argc > 2 ? atoi(argv[2]) : 0;
    unsigned long long bytes = (unsigned long long)(gb * 1e9) & ~31ull, n16 = bytes / 16, n_slots = n16 / 2, n_probes = 1ull << 27;
    uint4* buf; unsigned long long* sink;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&sink, 8); cudaMemset(sink, 0, 8);
    fill<<<148 * 16, 256>>>(buf, n16);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        float ms = 0;
        switch (mode) {
#define CASE(M) case M: ms = once<M>(buf, n_slots, n_probes, 77ull * (r + 1), sink); break;
            CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13)
        }
        if (r > 0 && ms < best) best = ms;
    }
    printf("mode %2d %-34s: %7.2f G probes/s (%s)\n", mode, names[mode], n_probes / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
