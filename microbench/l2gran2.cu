// l2gran2.cu — one configuration per process so that ncu can attribute DRAM bytes:
//   l2gran2 <GB> <granularity|0> <mode> [reps]
// mode 0: one 32-byte ld.global.nc.v8 per probe; 1: one 8-byte load; 2: 32-byte ld with L2::evict_first
// The L2 fetch-granularity limit is set BEFORE the first allocation.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__host__ __device__ inline unsigned long long mix64(unsigned long long x) {
    x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32; return x;
}
template <int MODE>
__global__ void __launch_bounds__(256) probe(const uint4* __restrict__ buf, unsigned long long n_slots,
                                             unsigned long long n_probes, unsigned long long seed, unsigned long long* sink) {
    constexpr int U = 8;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x * U;
    uint32_t acc = 0;
    for (unsigned long long i0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * U; i0 < n_probes; i0 += stride) {
        uint32_t r[U][8];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const unsigned long long idx = __umul64hi(mix64(seed + i0 + u + 1), n_slots);
            const uint4* p = buf + 2 * idx;
            if (MODE == 0)
                asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                    : "=r"(r[u][0]),"=r"(r[u][1]),"=r"(r[u][2]),"=r"(r[u][3]),"=r"(r[u][4]),"=r"(r[u][5]),"=r"(r[u][6]),"=r"(r[u][7]) : "l"(p));
            else if (MODE == 1) {
                asm volatile("ld.global.v2.u32 {%0,%1}, [%2];" : "=r"(r[u][0]),"=r"(r[u][1]) : "l"(p));
                r[u][2] = r[u][3] = r[u][4] = r[u][5] = r[u][6] = r[u][7] = 0;
            } else {
                asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                    : "=r"(r[u][0]),"=r"(r[u][1]),"=r"(r[u][2]),"=r"(r[u][3]),"=r"(r[u][4]),"=r"(r[u][5]),"=r"(r[u][6]),"=r"(r[u][7]) : "l"(p));
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) acc ^= r[u][0] ^ r[u][3] ^ r[u][5] ^ r[u][6];
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
int main(int argc, char** argv) {
    double gb = argc > 1 ? atof(argv[1]) : 3.2;
    int gran = argc > 2 ? atoi(argv[2]) : 0, mode = argc > 3 ? atoi(argv[3]) : 0, reps = argc > 4 ? atoi(argv[4]) : 4;
    cudaFree(0);
    if (gran) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    unsigned long long bytes = (unsigned long long)(gb * 1e9) & ~31ull, n16 = bytes / 16, n_slots = n16 / 2, n_probes = 1ull << 27;
    uint4* buf; unsigned long long* sink;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&sink, 8); cudaMemset(sink, 0, 8);
    fill<<<148 * 16, 256>>>(buf, n16);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(a);
        if (mode == 0) probe<0><<<148 * 32, 256>>>(buf, n_slots, n_probes, 77ull * (r + 1), sink);
        else if (mode == 1) probe<1><<<148 * 32, 256>>>(buf, n_slots, n_probes, 77ull * (r + 1), sink);
        else probe<2><<<148 * 32, 256>>>(buf, n_slots, n_probes, 77ull * (r + 1), sink);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (r > 0 && ms < best) best = ms;
    }
    printf("GB %.2f gran %zu mode %d: %.2f G probes/s (%s)\n", gb, g, mode, n_probes / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
