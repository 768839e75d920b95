// l2gran.cu — microbenchmark: random 32-byte probes over a large buffer under different
// L2 fetch-granularity limits and load instructions.  Prints probes/s per variant.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__host__ __device__ inline unsigned long long mix64(unsigned long long x) {
    x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32; return x;
}

template <int MODE>
__global__ void __launch_bounds__(256) probe(const uint4* __restrict__ buf, unsigned long long n_slots,
                                             unsigned long long n_probes, unsigned long long seed,
                                             unsigned long long* sink) {
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
                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[u][0]),"=r"(r[u][1]),"=r"(r[u][2]),"=r"(r[u][3]) : "l"(p));
                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[u][4]),"=r"(r[u][5]),"=r"(r[u][6]),"=r"(r[u][7]) : "l"(p + 1));
            } else if (MODE == 2) {   // plain 16-byte load only (half a bucket)
                asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[u][0]),"=r"(r[u][1]),"=r"(r[u][2]),"=r"(r[u][3]) : "l"(p));
                r[u][4] = r[u][5] = r[u][6] = r[u][7] = 0;
            } else if (MODE == 3) {   // 8-byte load
                asm volatile("ld.global.v2.u32 {%0,%1}, [%2];" : "=r"(r[u][0]),"=r"(r[u][1]) : "l"(p));
                r[u][2] = r[u][3] = r[u][4] = r[u][5] = r[u][6] = r[u][7] = 0;
            } else if (MODE == 4) {   // evict-first / no L2 prefetch hint variant
                asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                    : "=r"(r[u][0]),"=r"(r[u][1]),"=r"(r[u][2]),"=r"(r[u][3]),"=r"(r[u][4]),"=r"(r[u][5]),"=r"(r[u][6]),"=r"(r[u][7]) : "l"(p));
            } else if (MODE == 5) {   // ld.global.cv? volatile-ish (cache as volatile)
                asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[u][0]),"=r"(r[u][1]),"=r"(r[u][2]),"=r"(r[u][3]) : "l"(p));
                asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[u][4]),"=r"(r[u][5]),"=r"(r[u][6]),"=r"(r[u][7]) : "l"(p + 1));
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

template <int MODE>
double run(const uint4* buf, unsigned long long n_slots, unsigned long long n_probes, unsigned long long* sink, int grid) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(a);
        probe<MODE><<<grid, 256>>>(buf, n_slots, n_probes, 77ull * (r + 1), sink);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (r > 0 && ms < best) best = ms;
    }
    return n_probes / (best * 1e-3);
}

int main(int argc, char** argv) {
    double gb = argc > 1 ? atof(argv[1]) : 3.2;
    unsigned long long bytes = (unsigned long long)(gb * 1e9) & ~31ull, n16 = bytes / 16, n_slots = n16 / 2;
    unsigned long long n_probes = 1ull << 28;
    uint4* buf; unsigned long long* sink;
    cudaMalloc(&buf, bytes); cudaMalloc(&sink, 8); cudaMemset(sink, 0, 8);
    fill<<<148 * 16, 256>>>(buf, n16);
    cudaDeviceSynchronize();
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    printf("buffer %.2f GB, default cudaLimitMaxL2FetchGranularity = %zu\n", gb, g);
    int grids[3] = {148 * 8, 148 * 16, 148 * 64};
    for (int gran : {0, 32, 64, 128}) {
        if (gran) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran); cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
            printf("set granularity %d -> %s, now %zu\n", gran, cudaGetErrorString(e), g); }
        for (int gi = 0; gi < 3; gi++) {
            int grid = grids[gi];
            printf("  grid %5d: v8.nc %.2f G/s | 2xv4.nc %.2f | v4 16B %.2f | v2 8B %.2f | v8 evict_first %.2f | 2xv4.cg %.2f\n", grid,
                   run<0>(buf, n_slots, n_probes, sink, grid) / 1e9, run<1>(buf, n_slots, n_probes, sink, grid) / 1e9,
                   run<2>(buf, n_slots, n_probes, sink, grid) / 1e9, run<3>(buf, n_slots, n_probes, sink, grid) / 1e9,
                   run<4>(buf, n_slots, n_probes, sink, grid) / 1e9, run<5>(buf, n_slots, n_probes, sink, grid) / 1e9);
        }
    }
    cudaError_t e = cudaGetLastError();
    printf("last error: %s\n", cudaGetErrorString(e));
    return 0;
}
