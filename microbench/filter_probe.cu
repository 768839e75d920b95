// filter_probe.cu — memory-side ceiling of "L2-resident presence filter in front of the HBM table".
//   filter_probe <table_MB> <filter_MB> <mode> <pass_pct> [persist]
// Every probe reads ONE 32-bit filter word at a random index (two-bit Bloom test); the probes that
// pass (hits + false positives, pass_pct %) read one random 32-byte table sector.
//   mode 0  table only, 8 sector loads in flight per thread (R_rand at this table size)
//   mode 1  filter only, 8 word loads in flight per thread
//   mode 2  fused per thread: 8 filter loads, then predicated sector loads of the survivors
//   mode 3  warp queue: survivors are compacted into a per-warp shared-memory queue and popped 32 at a
//           time, so every sector load instruction runs with a full warp
//   mode 4  mode 3 with 2 pops in flight per lane
// persist=1 puts an L2 persisting access-policy window on the filter; table loads carry an
// evict_first cache hint in modes 2-4.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__host__ __device__ inline unsigned long long mix64(unsigned long long x) {
    x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32; return x;
}

__device__ __forceinline__ void ld_sector(const uint4* p, unsigned long long pol, uint4& a, uint4& b) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p), "l"(pol));
}
__device__ __forceinline__ uint32_t ld_word(const uint32_t* p, unsigned long long pol) {
    uint32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}

template <int MODE>
__global__ void __launch_bounds__(128) probe(const uint4* __restrict__ table, unsigned long long n_sectors,
                                             const uint32_t* __restrict__ filt, unsigned long long n_words,
                                             unsigned long long n_probes, unsigned long long seed,
                                             unsigned long long* sink) {
    constexpr int U = 8;
    unsigned long long pol_first, pol_last;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
    __shared__ unsigned long long s_q[4][32 * U + 64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x * U;
    uint32_t acc = 0;
    uint32_t qn = 0;   // warp-uniform queue fill
    const unsigned long long iters = (n_probes + stride - 1) / stride;   // same trip count for every lane (ballots below)
    for (unsigned long long it = 0; it < iters; it++) {
        const unsigned long long i0 = it * stride + ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * U;
        unsigned long long h[U];
        const bool live = i0 < n_probes;
#pragma unroll
        for (int u = 0; u < U; u++) h[u] = mix64(seed + i0 + u + 1);
        if (MODE == 0) {
            uint4 a[U], b[U];
#pragma unroll
            for (int u = 0; u < U; u++) if (live) ld_sector(table + 2 * __umul64hi(h[u], n_sectors), pol_first, a[u], b[u]);
#pragma unroll
            for (int u = 0; u < U; u++) if (live) acc ^= a[u].x ^ a[u].w ^ b[u].y ^ b[u].z;
            continue;
        }
        uint32_t fw[U];
#pragma unroll
        for (int u = 0; u < U; u++) if (live) fw[u] = ld_word(filt + __umul64hi(h[u], n_words), pol_last);
        unsigned pass = 0;
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t need = (1u << (h[u] & 31)) | (1u << ((h[u] >> 5) & 31));
            if (live && (fw[u] & need) == need) pass |= 1u << u;
        }
        if (MODE == 1) { acc ^= pass; continue; }
        if (MODE == 2) {
            uint4 a[U], b[U];
#pragma unroll
            for (int u = 0; u < U; u++) if (pass & (1u << u)) ld_sector(table + 2 * __umul64hi(h[u] * 0x9E3779B97F4A7C15ull, n_sectors), pol_first, a[u], b[u]);
#pragma unroll
            for (int u = 0; u < U; u++) if (pass & (1u << u)) acc ^= a[u].x ^ a[u].w ^ b[u].y ^ b[u].z;
            continue;
        }
        // modes 3/4: compact the survivors of the warp into its queue
#pragma unroll
        for (int u = 0; u < U; u++) {
            const bool ok = pass & (1u << u);
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (ok) s_q[warp][qn + __popc(m & ((1u << lane) - 1))] = h[u] * 0x9E3779B97F4A7C15ull;
            qn += __popc(m);
        }
        __syncwarp();
        constexpr int PB = MODE == 4 ? 2 : 1;
        while (qn >= 32 * PB) {
            uint4 a[PB], b[PB];
#pragma unroll
            for (int k = 0; k < PB; k++) {
                const unsigned long long hh = s_q[warp][qn - 32 * (k + 1) + lane];
                ld_sector(table + 2 * __umul64hi(hh, n_sectors), pol_first, a[k], b[k]);
            }
#pragma unroll
            for (int k = 0; k < PB; k++) acc ^= a[k].x ^ a[k].w ^ b[k].y ^ b[k].z;
            qn -= 32 * PB;
        }
        __syncwarp();
    }
    if (MODE >= 3) {
        while (qn > 0) {
            const uint32_t take = qn < 32 ? qn : 32;
            if (lane < (int)take) {
                uint4 a, b;
                ld_sector(table + 2 * __umul64hi(s_q[warp][qn - take + lane], n_sectors), pol_first, a, b);
                acc ^= a.x ^ a.w ^ b.y ^ b.z;
            }
            qn -= take;
        }
    }
    if (acc == 0x9e3779b9u) atomicAdd(sink, 1ull);
}

__global__ void fill_table(uint4* buf, unsigned long long n) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        unsigned long long h = mix64(i + 0x1234567ull);
        buf[i] = make_uint4((uint32_t)h, (uint32_t)(h >> 32), (uint32_t)i, (uint32_t)(i >> 32));
    }
}
// every bit set with probability fill (so that a two-bit test passes with probability fill^2)
__global__ void fill_filter(uint32_t* f, unsigned long long n, double fill) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long thr = (unsigned long long)(fill * 18446744073709551615.0);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t w = 0;
        for (int b = 0; b < 32; b++) if (mix64(i * 32 + b + 0x777ull) < thr) w |= 1u << b;
        f[i] = w;
    }
}

template <int MODE>
static float run(const uint4* t, unsigned long long ns, const uint32_t* f, unsigned long long nw, unsigned long long np,
                 unsigned long long* sink, cudaStream_t st) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < 6; r++) {
        cudaEventRecord(a, st);
        probe<MODE><<<148 * 16, 128, 0, st>>>(t, ns, f, nw, np, 0x5151ull * (r + 1), sink);
        cudaEventRecord(b, st);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (r > 0 && ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv) {
    const double table_mb = argc > 1 ? atof(argv[1]) : 615, filter_mb = argc > 2 ? atof(argv[2]) : 77;
    const int mode = argc > 3 ? atoi(argv[3]) : 3;
    const double pass = argc > 4 ? atof(argv[4]) / 100.0 : 0.35;
    const int persist = argc > 5 ? atoi(argv[5]) : 0;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    if (mode < 0) {
        printf("device %s: %d SMs, L2 %.1f MB, persistingL2CacheMaxSize %.1f MB, accessPolicyMaxWindowSize %.1f MB, smem/SM %zu, smem/block optin %zu, regs/SM %d\n",
               prop.name, prop.multiProcessorCount, prop.l2CacheSize / 1048576.0, prop.persistingL2CacheMaxSize / 1048576.0,
               prop.accessPolicyMaxWindowSize / 1048576.0, prop.sharedMemPerMultiprocessor, prop.sharedMemPerBlockOptin, prop.regsPerMultiprocessor);
        return 0;
    }
    const unsigned long long ns = (unsigned long long)(table_mb * 1048576.0 / 32), nw = (unsigned long long)(filter_mb * 1048576.0 / 4);
    uint4* t; uint32_t* f; unsigned long long* sink;
    cudaMalloc(&t, ns * 32); cudaMalloc(&f, nw * 4); cudaMalloc(&sink, 8);
    cudaMemset(sink, 0, 8);
    cudaStream_t st;
    cudaStreamCreate(&st);
    fill_table<<<148 * 16, 256, 0, st>>>(t, ns * 2);
    fill_filter<<<148 * 16, 256, 0, st>>>(f, nw, sqrt(pass));
    if (persist) {
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, prop.persistingL2CacheMaxSize);
        cudaStreamAttrValue v = {};
        size_t win = nw * 4 < (size_t)prop.accessPolicyMaxWindowSize ? nw * 4 : (size_t)prop.accessPolicyMaxWindowSize;
        v.accessPolicyWindow.base_ptr = f;
        v.accessPolicyWindow.num_bytes = win;
        double hr = (double)prop.persistingL2CacheMaxSize / (double)win;
        v.accessPolicyWindow.hitRatio = (float)(hr < 1.0 ? hr : 1.0);
        v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v);
    }
    const unsigned long long np = 1ull << 29;
    float ms = 0;
    switch (mode) {
        case 0: ms = run<0>(t, ns, f, nw, np, sink, st); break;
        case 1: ms = run<1>(t, ns, f, nw, np, sink, st); break;
        case 2: ms = run<2>(t, ns, f, nw, np, sink, st); break;
        case 3: ms = run<3>(t, ns, f, nw, np, sink, st); break;
        default: ms = run<4>(t, ns, f, nw, np, sink, st); break;
    }
    cudaError_t ce = cudaGetLastError();
    printf("table %.0f MB filter %.0f MB mode %d pass %.0f%% persist %d: %.3f ms, %.1f G probes/s (%s)\n", table_mb, filter_mb, mode,
           pass * 100, persist, ms, np / (ms * 1e-3) / 1e9, cudaGetErrorString(ce));
    return 0;
}
