// ka_build_api.cu — ka_build: BuildKmerProcessor.java:138-223 on the device (kernels in ka_kernels.cu).
#include "ka_engine_internal.cuh"

using namespace ka;
using namespace kai;

extern "C" {

int ka_build(ka_engine* e, const uint8_t* residues, const uint64_t* offsets, uint64_t N,
             const int32_t* n_roles, const int32_t* peg_role, int K, uint64_t cap,
             uint8_t* out_kmers, int32_t* out_roles, uint64_t* n_out, int load_as_db) {
    if (!e || !n_out) return KA_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    *n_out = 0;
    if (K < 1 || K > KMAX) return fail(e, KA_ERR_K, "K = %d: this engine packs 5 bits per residue, K must be 1..%d", K, KMAX);
    if (N == 0) return KA_OK;
    if (!offsets || !n_roles || !peg_role || (cap && (!out_kmers || !out_roles))) return fail(e, KA_ERR_INVALID, "ka_build: NULL argument");
    if (N > 0xfffffff0ull) return fail(e, KA_ERR_TOO_BIG, "ka_build: too many pegs");
    uint64_t windows = 0;
    int32_t max_role = 0;
    for (uint64_t i = 0; i < N; i++) {
        if (offsets[i + 1] < offsets[i]) return fail(e, KA_ERR_OFFSETS, "ka_build: offsets are not monotone");
        uint64_t L = offsets[i + 1] - offsets[i];
        if (n_roles[i] == 1) {
            if (peg_role[i] < 0) return fail(e, KA_ERR_ROLE, "ka_build: negative role id at peg %llu", (unsigned long long)i);
            if (L >= (uint64_t)K) windows += L - K + 1;
            max_role = std::max(max_role, peg_role[i]);
        }
    }
    Device& d = e->devs[0];
    cudaSetDevice(d.id);
    cudaStream_t st = d.pipe[0].st;
    const uint64_t base = offsets[0], n_res = offsets[N] - base;
    uint64_t n_slots = 1024;
    while (n_slots < 2 * windows) n_slots <<= 1;
    uint8_t *d_res = nullptr, *d_lut = nullptr, *d_inv = nullptr, *d_ok = nullptr;
    unsigned long long *d_off = nullptr, *d_cnt = nullptr;
    int32_t *d_nr = nullptr, *d_pr = nullptr, *d_or = nullptr;
    uint32_t* d_bm = nullptr;
    Slot128* d_tab = nullptr;
    auto cleanup = [&] {
        cudaFree(d_res); cudaFree(d_lut); cudaFree(d_inv); cudaFree(d_ok); cudaFree(d_off); cudaFree(d_cnt);
        cudaFree(d_nr); cudaFree(d_pr); cudaFree(d_or); cudaFree(d_bm); cudaFree(d_tab);
    };
    cudaError_t ce;
    if ((ce = cudaMalloc((void**)&d_res, n_res + K + 64)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&d_off, (N + 1) * 8)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&d_nr, N * 4)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&d_pr, N * 4)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&d_lut, 256)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&d_inv, 32)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&d_bm, 32)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&d_cnt, 8)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&d_tab, n_slots * sizeof(Slot128))) != cudaSuccess ||
        (ce = cudaMalloc((void**)&d_ok, std::max<uint64_t>(cap, 1) * K)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&d_or, std::max<uint64_t>(cap, 1) * 4)) != cudaSuccess) {
        cleanup();
        return fail(e, KA_ERR_OOM, "ka_build: device allocation failed: %s", cudaGetErrorString(ce));
    }
    auto bail = [&](const char* what, cudaError_t c) {
        cleanup();
        return fail(e, KA_ERR_CUDA, "ka_build: %s: %s", what, cudaGetErrorString(c));
    };
    if (n_res && (ce = cudaMemcpyAsync(d_res, residues + base, n_res, cudaMemcpyHostToDevice, st)) != cudaSuccess) return bail("H2D residues", ce);
    if ((ce = cudaMemcpyAsync(d_off, offsets, (N + 1) * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) return bail("H2D offsets", ce);
    if ((ce = cudaMemcpyAsync(d_nr, n_roles, N * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) return bail("H2D n_roles", ce);
    if ((ce = cudaMemcpyAsync(d_pr, peg_role, N * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) return bail("H2D peg_role", ce);
    if ((ce = cudaMemsetAsync(d_bm, 0, 32, st)) != cudaSuccess) return bail("memset", ce);
    if ((ce = cudaMemsetAsync(d_cnt, 0, 8, st)) != cudaSuccess) return bail("memset", ce);
    if ((ce = cudaMemsetAsync(d_tab, 0, n_slots * sizeof(Slot128), st)) != cudaSuccess) return bail("memset table", ce);
    // alphabet of the training proteins (same rule as ka_db_load: at most 31 distinct bytes)
    if ((ce = launch_alphabet_scan(d_res, n_res, d_bm, st)) != cudaSuccess) return bail("alphabet scan", ce);
    uint32_t bitmap[8];
    if ((ce = cudaMemcpyAsync(bitmap, d_bm, 32, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return bail("D2H alphabet", ce);
    if ((ce = cudaStreamSynchronize(st)) != cudaSuccess) return bail("alphabet sync", ce);
    uint8_t lut[256], inv[32];
    memset(lut, 0, 256); memset(inv, 0, 32);
    int nsym = 0;
    for (int b = 0; b < 256; b++)
        if (bitmap[b >> 5] & (1u << (b & 31))) {
            nsym++;
            if (nsym <= 31) { lut[b] = (uint8_t)nsym; inv[nsym] = (uint8_t)b; }
        }
    if (nsym > 31) {
        cleanup();
        return fail(e, KA_ERR_ALPHABET, "ka_build: the proteins use %d distinct residue bytes; at most 31 fit the 5-bit packing", nsym);
    }
    if ((ce = cudaMemcpyAsync(d_lut, lut, 256, cudaMemcpyHostToDevice, st)) != cudaSuccess) return bail("H2D lut", ce);
    if ((ce = cudaMemcpyAsync(d_inv, inv, 32, cudaMemcpyHostToDevice, st)) != cudaSuccess) return bail("H2D inv", ce);
    cudaEventRecord(d.pipe[0].ev_k0, st);
    if ((ce = launch_build_pass(1, d_res, d_off, (uint32_t)N, d_nr, d_pr, K, d_lut, d_tab, n_slots, st)) != cudaSuccess) return bail("build pass 1", ce);
    if ((ce = launch_build_pass(2, d_res, d_off, (uint32_t)N, d_nr, d_pr, K, d_lut, d_tab, n_slots, st)) != cudaSuccess) return bail("build pass 2", ce);
    if ((ce = launch_build_emit(d_tab, n_slots, K, d_inv, cap, d_ok, d_or, d_cnt, st)) != cudaSuccess) return bail("build emit", ce);
    cudaEventRecord(d.pipe[0].ev_k1, st);
    unsigned long long found = 0;
    if ((ce = cudaMemcpyAsync(&found, d_cnt, 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return bail("D2H count", ce);
    if ((ce = cudaStreamSynchronize(st)) != cudaSuccess) return bail("build kernels", ce);
    *n_out = found;
    {
        // ka_get_stats after ka_build: device time of the two passes + emit, window positions inserted or probed
        float ms = 0;
        cudaEventElapsedTime(&ms, d.pipe[0].ev_k0, d.pipe[0].ev_k1);
        e->stats = ka_stats{};
        e->stats.sequences = N; e->stats.residues = n_res; e->stats.kernel_ms = ms; e->stats.kernel_launches = 3;
        uint64_t w_all = 0;
        for (uint64_t i = 0; i < N; i++) {
            const uint64_t L = offsets[i + 1] - offsets[i];
            if (n_roles[i] <= 1 && L >= (uint64_t)K) w_all += L - K + 1;
        }
        e->stats.probes = w_all;
        e->stats.h2d_bytes = n_res + (N + 1) * 8 + N * 8;
    }
    if (found > cap) {
        cleanup();
        return fail(e, KA_ERR_TOO_BIG, "ka_build: %llu k-mers found, output capacity is %llu", found, (unsigned long long)cap);
    }
    if (found) {
        if ((ce = cudaMemcpy(out_kmers, d_ok, found * K, cudaMemcpyDeviceToHost)) != cudaSuccess) return bail("D2H kmers", ce);
        if ((ce = cudaMemcpy(out_roles, d_or, found * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) return bail("D2H roles", ce);
    }
    // load_as_db: the table is built from the k-mers where they are, in HBM (no kmerdb.tbl, no host round trip)
    int rc = KA_OK;
    if (load_as_db && found) rc = db_load_impl(e, d_ok, d_or, found, K, 0, 0, true, max_role);
    cleanup();
    return rc;
}

}  // extern "C"
