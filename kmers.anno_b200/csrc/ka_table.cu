// ka_table.cu — the device table: geometry (slot class, sector count, wide / sharded forms), build
// from host lines or from the synthetic generator, and the two C-ABI loaders.
// Replaces ApplyKmerProcessor.java:99-110 (HashMap.put per DB line, last line wins).
#include "ka_engine_internal.cuh"

using namespace ka;
using namespace kai;

namespace kai {

// Build the table replica (or shard) of one device from the DB lines.
int build_table(ka_engine* e, Device& d, const TableView& geom, const DbSource& src, uint64_t* n_keys, uint32_t* max_probe) {
    const uint8_t* kmers = src.kmers;
    const int32_t* roles = src.roles;
    const uint64_t n = src.n;
    DCK(d, cudaSetDevice(d.id));
    if (d.table) { cudaFree(d.table); d.table = nullptr; }
    if (d.ovf) { cudaFree(d.ovf); d.ovf = nullptr; }
    if (d.filt) { cudaFree(d.filt); d.filt = nullptr; }
    const int K = geom.K;
    const size_t n_sectors = (size_t)1 << (geom.n_shards > 1 ? geom.shard_shift : geom.bbits);  // of this device
    const size_t bytes = n_sectors * 32;
    const size_t n_slots = n_sectors * (geom.cls == 32 ? 8 : (geom.cls == 64 ? 4 : 2));
    cudaError_t ce = cudaMalloc((void**)&d.table, bytes);
    if (ce != cudaSuccess) { d.table = nullptr; return dev_fail(d, KA_ERR_OOM, "table", ce); }
    const size_t ovf_bytes = geom.cls == 128 ? 0 : ((size_t)64 << geom.ovf_bbits);
    if (ovf_bytes) {
        ce = cudaMalloc((void**)&d.ovf, ovf_bytes);
        if (ce != cudaSuccess) { d.ovf = nullptr; return dev_fail(d, KA_ERR_OOM, "overflow table", ce); }
    }
    cudaStream_t st = d.pipe[0].st;
    TableView tab = geom;
    tab.sectors = d.table;
    tab.ovf = d.ovf;
    const uint64_t CH = (src.synthetic ? 64ull : 16ull) << 20;  // k-mers per upload / per generator launch
    uint8_t* dk = nullptr; int32_t* dr = nullptr; unsigned long long* best = nullptr;
    unsigned long long* dc = nullptr; uint32_t* de = nullptr;
    uint64_t ch = std::min<uint64_t>(CH, n ? n : 1);
    const bool packed = geom.cls != 128;
    // cls 32/64: 8 bytes per primary slot hold the winning (line, role) until db_finalize
    if ((ce = cudaMalloc((void**)&dk, ch * K)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&dr, ch * 4)) != cudaSuccess ||
        (packed && (ce = cudaMalloc((void**)&best, n_slots * 8)) != cudaSuccess) ||
        (ce = cudaMalloc((void**)&dc, 16)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&de, 16)) != cudaSuccess) {
        if (dk) cudaFree(dk);
        if (dr) cudaFree(dr);
        if (best) cudaFree(best);
        if (dc) cudaFree(dc);
        return dev_fail(d, KA_ERR_OOM, "DB staging", ce);
    }
    int rc = KA_OK;
    auto step = [&](cudaError_t c, const char* what) {
        if (c != cudaSuccess && rc == KA_OK) rc = dev_fail(d, KA_ERR_CUDA, what, c);
    };
    step(cudaMemsetAsync(d.table, 0, bytes, st), "memset table");
    if (ovf_bytes) step(cudaMemsetAsync(d.ovf, 0, ovf_bytes, st), "memset overflow table");
    if (packed) step(cudaMemsetAsync(best, 0, n_slots * 8, st), "memset best");
    step(cudaMemcpyAsync(d.lut, e->lut, 256, cudaMemcpyHostToDevice, st), "H2D lut");
    step(cudaMemcpyAsync(d.lut5, e->lut5, 256, cudaMemcpyHostToDevice, st), "H2D lut5");
    step(cudaMemcpyAsync(d.inv32, e->inv32, 32, cudaMemcpyHostToDevice, st), "H2D inv32");
    step(cudaMemsetAsync(dc, 0, 16, st), "memset counters");
    step(cudaMemsetAsync(de, 0, 16, st), "memset errs");
    for (uint64_t i = 0; i < n && rc == KA_OK; i += ch) {
        uint64_t m = std::min(ch, n - i);
        if (!src.synthetic) {
            step(cudaMemcpyAsync(dk, kmers + i * K, m * K, cudaMemcpyDefault, st), "copy kmers");
            step(cudaMemcpyAsync(dr, roles + i, m * 4, cudaMemcpyDefault, st), "copy roles");
        } else {
            step(launch_db_generate(i, m, K, src.seed, src.n_roles, dk, dr, st), "db_generate");
        }
        step(launch_db_insert(tab, dk, dr, m, i, d.lut, best, src.role_bits, dc, de, st), "db_insert");
        step(cudaStreamSynchronize(st), "db_insert sync");
    }
    if (rc == KA_OK) step(launch_db_finalize(tab, best, src.role_bits, st), "db_finalize");
    step(cudaStreamSynchronize(st), "db_finalize sync");
    unsigned long long hc[2] = {0, 0};
    uint32_t he[4] = {0, 0, 0, 0};
    step(cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost), "D2H counters");
    step(cudaMemcpy(he, de, 16, cudaMemcpyDeviceToHost), "D2H errs");
    cudaFree(dk); cudaFree(dr); cudaFree(dc); cudaFree(de);
    if (best) cudaFree(best);
    if (rc) return rc;
    if (he[0]) { d.err = KA_ERR_ALPHABET; d.errmsg = "k-mer byte outside the DB alphabet (internal)"; return d.err; }
    if (he[1]) { d.err = KA_ERR_ROLE; d.errmsg = "negative role id in the DB"; return d.err; }
    if (he[2]) { d.err = KA_ERR_TOO_BIG; d.errmsg = "overflow table full"; return d.err; }  // caller retries larger
    *n_keys = hc[0];
    *max_probe = (uint32_t)hc[1];
    return KA_OK;
}

uint32_t ceil_log2(double x) {
    uint32_t b = 0;
    while ((double)(1ull << b) < x && b < 62) b++;
    return b;
}

// Pick slot class and sector count: the smallest table that holds n keys at the requested
// load factor with remainder + role fitting the slot (see ka_common.cuh).
bool choose_geometry(uint64_t n, int K, int32_t max_role, double lf, int force_cls, uint32_t n_shards, bool force_wide, TableView& g) {
    const uint32_t w = 5u * (uint32_t)K;
    uint32_t role_bits = 1;
    while (((uint64_t)max_role + 1) >> role_bits) role_bits++;
    bool found = false;
    uint64_t best_bytes = 0;
    for (int cls : {32, 64, 128}) {
        if (force_cls && cls != force_cls) continue;
        if (n_shards > 1 && cls == 128) continue;   // chaining across shards is not supported: quotiented classes only
        const int S = 256 / cls;
        uint32_t b = ceil_log2((double)(n ? n : 1) / ((double)S * lf));
        if (b < 6) b = 6;
        uint32_t shard_log = 0;
        while ((1u << shard_log) < n_shards) shard_log++;
        if (b < 6 + shard_log) b = 6 + shard_log;
        uint32_t rem_bits = 0;
        if (cls != 128) {
            if ((int)(w + role_bits) - cls > (int)b) b = w + role_bits - (uint32_t)cls;
            if (b > w) b = w;
            if (b < shard_log) continue;
            rem_bits = w - b;
            if (rem_bits + role_bits > (uint32_t)cls) continue;
        }
        uint32_t slot_log = cls == 32 ? 3 : (cls == 64 ? 2 : 1);
        // narrow tables: slot index + 1 must fit the 32-bit de-dup token; beyond that the quotiented
        // classes switch to the wide kernels (64-bit sector indices, the key is the token)
        bool wide = force_wide && cls != 128;
        if (b + slot_log > 31) {
            if (cls == 128 || b > 40) continue;
            wide = true;
        }
        uint64_t bytes = 32ull << b;
        if (!found || bytes < best_bytes) {
            found = true; best_bytes = bytes;
            g.cls = cls; g.bbits = b; g.rem_bits = rem_bits; g.wbits = w; g.K = K;
            g.key_mask = (1ull << w) - 1;
            g.rem_mask = rem_bits ? ((1ull << rem_bits) - 1) : 0;
            g.sectors = nullptr;
            g.ovf = nullptr;
            g.n_primary_slots = wide ? 0u : (uint32_t)((uint64_t)S << b);
            g.wide = wide ? 1u : 0u;
            // expected keys beyond S per sector under Poisson(n / sectors) arrivals
            double lam = (double)n / (double)(1ull << b), pk = std::exp(-lam), over = 0;
            for (int k = 1; k < S + 400; k++) {
                pk *= lam / k;
                if (k > S) over += (k - S) * pk;
            }
            double want = 4.0 * over * (double)(1ull << b) + 4096;
            g.ovf_bbits = cls == 128 ? 0 : ceil_log2(want / 2.0 / (n_shards ? n_shards : 1));
            g.n_shards = n_shards;
            g.shard_shift = b - shard_log;
            g.my_shard = 0;
            g.shard_sectors = nullptr;
            g.shard_ovf = nullptr;
        }
    }
    return found;
}


// ---- line table (slot class 16, ka_line.cuh) ----------------------------------------------------
static uint32_t bit_length(unsigned long long x) {
    uint32_t b = 0;
    while (x) { b++; x >>= 1; }
    return b;
}

// B = c * 2^s lines (c in 8..15) for n keys at load factor lf (ka_line.cuh).  The 14 remainder bits
// of a tag need s >= bh + bl - 17, the line index needs s >= bh - u.  `forced` (option slot_bits = 16)
// accepts a table that is larger than the keys need.
bool choose_line_geometry(uint64_t n, int K, int nsym, double lf, bool forced, LineTable& g) {
    if (nsym < 2 || nsym > 31 || K < 2 || K > LINE_KMAX) return false;
    const uint32_t Kh = (uint32_t)K / 2, Kl = (uint32_t)K - Kh;
    unsigned long long ph = 1, pl = 1;
    for (uint32_t i = 0; i < Kh; i++) ph *= (unsigned long long)nsym;
    for (uint32_t i = 0; i < Kl; i++) pl *= (unsigned long long)nsym;
    const uint32_t bh = bit_length(ph - 1), bl = bit_length(pl - 1);
    if (bh < 3 || bl < 3 || bh > 25 || bl > 30) return false;                // (the survivor queue packs H in 26 bits)
    const uint32_t w = bh + bl;
    const uint32_t u = bh >= 12 ? 12 : 3;                                  // short high halves: power-of-two tables only
    const uint32_t la = bh - u;
    const uint32_t s_min = std::max<uint32_t>(la, w > 17 ? w - 17 : 0), s_max = w - u - 2;
    if (s_min > s_max || s_min > 23) return false;
    // (a DB holds at most nsym^K distinct keys however many lines it has)
    const double n_keys = std::min((double)(n ? n : 1), (double)ph * (double)pl);
    const double want = std::max(8.0, std::ceil(n_keys / (32.0 * lf)));
    uint64_t best = 0;
    uint32_t bc = 0, bs = 0;
    for (uint32_t s = 0; s <= std::min<uint32_t>(s_max, 23); s++)
        for (uint32_t c = 8; c < 16; c++) {
            if (u == 3 && c != 8) continue;
            const uint64_t B = (uint64_t)c << s;
            if ((double)B >= want && (!best || B < best)) { best = B; bc = c; bs = s; }
        }
    if (!best) return false;                                               // more lines than the layout can index
    if (bs < s_min) {
        // the keys would need a padded table: accepted up to 2x (no larger than the 32-bit sector class at its
        // default load factor), or whenever the layout is forced
        if (!forced && (bs + 1 < s_min)) return false;
        bs = s_min; bc = 8;
    }
    if (((uint64_t)bc << bs) >= (1ull << 27)) return false;                // 32-bit slot tokens
    g = LineTable{};
    g.c = bc; g.s = bs; g.u = u;
    g.n_lines = bc << bs;
    g.n_filt = g.n_lines * 4;
    g.inv_c = (65536u + bc - 1) / bc;
    g.bh = bh; g.bl = bl;
    g.la = la; g.a = bs - la; g.r = bl - g.a;
    if (g.r < 2 || (u - 3) + g.r > TAG_REM_BITS) return false;
    g.radix = (uint32_t)nsym;
    g.Kh = Kh; g.Kl = Kl;
    g.pw_h = (uint32_t)(ph / (unsigned long long)nsym);
    g.pw_l = (uint32_t)(pl / (unsigned long long)nsym);
    g.K = K;
    // expected keys beyond 32 per line under Poisson(n / lines) arrivals -> overflow table size
    const double lam = (double)n / (double)g.n_lines;
    double pk = std::exp(-lam), over = 0;
    for (int k = 1; k < 32 + 400; k++) {
        pk *= lam / k;
        if (k > 32) over += (k - 32) * pk;
    }
    const double want_ovf = 4.0 * over * (double)g.n_lines + 4096;
    g.ovf_bbits = ceil_log2(want_ovf / 2.0);
    return true;
}

// Build the line table of one device (replicated on every device of the engine).
// counts3: [0] distinct keys, [1] keys outside their home sector, [2] keys in the overflow table.
int build_line_table(ka_engine* e, Device& d, const LineTable& geom, const DbSource& src, uint64_t* counts3) {
    DCK(d, cudaSetDevice(d.id));
    if (d.table) { cudaFree(d.table); d.table = nullptr; }
    if (d.ovf) { cudaFree(d.ovf); d.ovf = nullptr; }
    if (d.filt) { cudaFree(d.filt); d.filt = nullptr; }
    const int K = geom.K;
    const size_t bytes = (size_t)geom.n_lines * 128, n_slots = (size_t)geom.n_lines * 32;
    const size_t ovf_bytes = (size_t)64 << geom.ovf_bbits, filt_bytes = (size_t)geom.n_lines * 16;
    cudaError_t ce;
    if ((ce = cudaMalloc((void**)&d.table, bytes)) != cudaSuccess) { d.table = nullptr; return dev_fail(d, KA_ERR_OOM, "table", ce); }
    if ((ce = cudaMalloc((void**)&d.ovf, ovf_bytes)) != cudaSuccess) { d.ovf = nullptr; return dev_fail(d, KA_ERR_OOM, "overflow table", ce); }
    if ((ce = cudaMalloc((void**)&d.filt, filt_bytes)) != cudaSuccess) { d.filt = nullptr; return dev_fail(d, KA_ERR_OOM, "presence filter", ce); }
    cudaStream_t st = d.pipe[0].st;
    LineTable tab = geom;
    tab.lines = d.table; tab.ovf = d.ovf; tab.filt = d.filt;
    const uint64_t n = src.n, CH = (src.synthetic ? 64ull : 16ull) << 20;
    const uint64_t ch = std::min<uint64_t>(CH, n ? n : 1);
    uint8_t* dk = nullptr; int32_t* dr = nullptr; unsigned long long* best = nullptr;
    unsigned long long* dc = nullptr; uint32_t* de = nullptr;
    if ((ce = cudaMalloc((void**)&dk, ch * K)) != cudaSuccess || (ce = cudaMalloc((void**)&dr, ch * 4)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&best, n_slots * 8)) != cudaSuccess || (ce = cudaMalloc((void**)&dc, 32)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&de, 16)) != cudaSuccess) {
        for (void* q : {(void*)dk, (void*)dr, (void*)best, (void*)dc}) if (q) cudaFree(q);
        return dev_fail(d, KA_ERR_OOM, "DB staging", ce);
    }
    int rc = KA_OK;
    auto step = [&](cudaError_t c, const char* what) {
        if (c != cudaSuccess && rc == KA_OK) rc = dev_fail(d, KA_ERR_CUDA, what, c);
    };
    step(cudaMemsetAsync(d.table, 0, bytes, st), "memset table");
    step(cudaMemsetAsync(d.ovf, 0, ovf_bytes, st), "memset overflow table");
    step(cudaMemsetAsync(d.filt, 0, filt_bytes, st), "memset filter");
    step(cudaMemsetAsync(best, 0, n_slots * 8, st), "memset best");
    step(cudaMemcpyAsync(d.lut, e->lut, 256, cudaMemcpyHostToDevice, st), "H2D lut");
    step(cudaMemcpyAsync(d.lut5, e->lut5, 256, cudaMemcpyHostToDevice, st), "H2D lut5");
    step(cudaMemcpyAsync(d.inv32, e->inv32, 32, cudaMemcpyHostToDevice, st), "H2D inv32");
    step(cudaMemsetAsync(dc, 0, 32, st), "memset counters");
    step(cudaMemsetAsync(de, 0, 16, st), "memset errs");
    for (uint64_t i = 0; i < n && rc == KA_OK; i += ch) {
        const uint64_t m = std::min(ch, n - i);
        if (!src.synthetic) {
            step(cudaMemcpyAsync(dk, src.kmers + i * K, m * K, cudaMemcpyDefault, st), "copy kmers");
            step(cudaMemcpyAsync(dr, src.roles + i, m * 4, cudaMemcpyDefault, st), "copy roles");
        } else {
            step(launch_db_generate(i, m, K, src.seed, src.n_roles, dk, dr, st), "db_generate");
        }
        step(launch_line_insert(tab, dk, dr, m, i, d.lut5, best, src.role_bits, dc, de, st), "line_insert");
        step(cudaStreamSynchronize(st), "line_insert sync");
    }
    if (rc == KA_OK) step(launch_line_finalize(tab, best, src.role_bits, st), "line_finalize");
    step(cudaStreamSynchronize(st), "line_finalize sync");
    unsigned long long hc[4] = {0, 0, 0, 0};
    uint32_t he[4] = {0, 0, 0, 0};
    step(cudaMemcpy(hc, dc, 32, cudaMemcpyDeviceToHost), "D2H counters");
    step(cudaMemcpy(he, de, 16, cudaMemcpyDeviceToHost), "D2H errs");
    cudaFree(dk); cudaFree(dr); cudaFree(dc); cudaFree(de); cudaFree(best);
    if (rc) return rc;
    if (he[0]) { d.err = KA_ERR_ALPHABET; d.errmsg = "k-mer byte outside the DB alphabet (internal)"; return d.err; }
    if (he[1]) { d.err = KA_ERR_ROLE; d.errmsg = "negative role id in the DB"; return d.err; }
    if (he[2]) { d.err = KA_ERR_TOO_BIG; d.errmsg = "overflow table full"; return d.err; }  // caller retries larger
    counts3[0] = hc[0]; counts3[1] = hc[1]; counts3[2] = hc[2];
    return KA_OK;
}

}  // namespace kai

extern "C" {

int ka_db_load(ka_engine* e, const uint8_t* kmers, const int32_t* role_ids, uint64_t n, int K) {
    if (!e) return KA_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    if (n && (!kmers || !role_ids)) return fail(e, KA_ERR_INVALID, "ka_db_load: NULL input");
    return db_load_impl(e, kmers, role_ids, n, K);
}

int ka_db_load_synthetic(ka_engine* e, uint64_t n, int K, int32_t n_roles, uint64_t seed) {
    if (!e) return KA_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    if (n == 0 || n_roles < 1) return fail(e, KA_ERR_INVALID, "ka_db_load_synthetic: n and n_roles must be positive");
    return db_load_impl(e, nullptr, nullptr, n, K, seed, n_roles);
}

// syn_roles > 0: the lines come from the synthetic generator (syn_seed, syn_roles), kmers/role_ids unused
}  // extern "C"

namespace kai {

int db_load_impl(ka_engine* e, const uint8_t* kmers, const int32_t* role_ids, uint64_t n, int K,
                 uint64_t syn_seed, int32_t syn_roles, bool on_device, int32_t max_role_hint) {
    if (K < 1 || K > KMAX) return fail(e, KA_ERR_K, "K = %d: this engine packs 5 bits per residue, K must be 1..%d", K, KMAX);
    e->have_db = false;
    for (Device& d : e->devs) { cudaSetDevice(d.id); cudaGetLastError(); }   // a stale error of an earlier call is not this call's

    // 1. alphabet: the distinct bytes of the DB, scanned on device 0
    Device& d0 = e->devs[0];
    uint32_t bitmap[8] = {0};
    const bool synthetic = syn_roles > 0;
    if (synthetic) {
        for (const char* a = "ACDEFGHIKLMNPQRSTVWY"; *a; a++) bitmap[(uint8_t)*a >> 5] |= 1u << ((uint8_t)*a & 31);
    } else {
        cudaSetDevice(d0.id);
        cudaStream_t st = d0.pipe[0].st;
        uint32_t* dbm = nullptr; uint8_t* dk = nullptr;
        const uint64_t CH = 256ull << 20;
        uint64_t total = n * (uint64_t)K, ch = std::min<uint64_t>(CH, total ? total : 1);
        if (cudaMalloc((void**)&dbm, 32) != cudaSuccess || cudaMalloc((void**)&dk, ch) != cudaSuccess) {
            if (dbm) cudaFree(dbm);
            return fail(e, KA_ERR_OOM, "ka_db_load: alphabet staging allocation failed");
        }
        cudaError_t ce = cudaMemsetAsync(dbm, 0, 32, st);
        for (uint64_t i = 0; i < total && ce == cudaSuccess; i += ch) {
            uint64_t m = std::min(ch, total - i);
            ce = cudaMemcpyAsync(dk, kmers + i, m, cudaMemcpyDefault, st);   // host lines, or the device-resident output of ka_build
            if (ce == cudaSuccess) ce = launch_alphabet_scan(dk, m, dbm, st);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        }
        if (ce == cudaSuccess) ce = cudaMemcpy(bitmap, dbm, 32, cudaMemcpyDeviceToHost);
        cudaFree(dbm); cudaFree(dk);
        if (ce != cudaSuccess) return fail(e, KA_ERR_CUDA, "ka_db_load: alphabet scan: %s", cudaGetErrorString(ce));
    }
    memset(e->lut, 0, 256);
    int nsym = 0;
    for (int b = 0; b < 256; b++)
        if (bitmap[b >> 5] & (1u << (b & 31))) {
            nsym++;
            if (nsym <= 31) e->lut[b] = (uint8_t)nsym;  // codes 1..31 in byte order; 0 = absent
        }
    if (nsym > 31)
        return fail(e, KA_ERR_ALPHABET,
                    "ka_db_load: the DB uses %d distinct residue bytes; at most 31 fit the 5-bit packing", nsym);
    // radix digits 0..nsym-1 in byte order (line table, packed streams); 31 = byte not in the alphabet
    memset(e->lut5, (int)CODE_INVALID, 256);
    memset(e->inv32, 0, 32);
    {
        int outside = -1;
        for (int b = 0; b < 256; b++) {
            if (e->lut[b]) { e->lut5[b] = (uint8_t)(e->lut[b] - 1); e->inv32[e->lut[b] - 1] = (uint8_t)b; }
            else if (outside < 0) outside = b;
        }
        for (int c = nsym; c < 32; c++) e->inv32[c] = (uint8_t)outside;     // nsym <= 31 < 256: such a byte exists
    }

    // 2. table geometry (slot class, sector count) from n, K and the largest role id
    int32_t max_role = synthetic ? syn_roles - 1 : (on_device ? max_role_hint : 0);
    for (uint64_t i = 0; !synthetic && !on_device && i < n; i++) {
        if (role_ids[i] < 0) return fail(e, KA_ERR_ROLE, "ka_db_load: negative role id %d at line %llu", role_ids[i], (unsigned long long)i);
        if (role_ids[i] > max_role) max_role = role_ids[i];
    }
    DbSource src;
    src.kmers = synthetic ? nullptr : kmers; src.roles = role_ids; src.n = n; src.seed = syn_seed; src.n_roles = (uint32_t)syn_roles;
    src.synthetic = synthetic;
    while (((uint64_t)max_role + 1) >> src.role_bits) src.role_bits++;
    {
        // a slot keeps (line + 1) << role_bits | role in 64 bits while the DB streams in
        uint32_t line_bits = 1;
        while (line_bits < 64 && ((n + 1) >> line_bits)) line_bits++;
        if (line_bits + src.role_bits > 64)
            return fail(e, KA_ERR_TOO_BIG, "ka_db_load: %llu lines with role ids up to %d exceed the 64-bit (line, role) word",
                        (unsigned long long)n, max_role);
    }
    // line table (slot class 16): the default of a replicated table whenever its layout fits (K <= 10, role ids
    // below 65536), or forced by slot_bits = 16.  57 B of DRAM traffic per probe against 95 B, 64 against 53 G probes/s
    // on the C3 batch and 112 against 54 on genomes unrelated to the DB (profiles/r02_summary.md); K = 11, 12, wide
    // and sharded tables use the sector classes.
    if (e->slot_bits == 16 || (e->slot_bits == 0 && e->table_mode == 0 && !e->wide && max_role <= 0xFFFF)) {
        LineTable lg;
        const bool forced = e->slot_bits == 16;
        if (forced && (e->table_mode != 0 || e->wide || max_role > 0xFFFF))
            return fail(e, KA_ERR_INVALID, "ka_db_load: slot_bits = 16 needs a replicated narrow table and role ids below 65536");
        if (choose_line_geometry(n, K, nsym, e->load_factor > 0 ? std::min(e->load_factor, 0.8) : 0.68, forced, lg)) {
            std::vector<std::array<uint64_t, 3>> cnt(e->devs.size());
            auto tb = std::chrono::steady_clock::now();
            int rc = KA_OK;
            for (int attempt = 0; attempt < 6; attempt++) {
                rc = for_each_device(e, [&](Device& d, int i) { return build_line_table(e, d, lg, src, cnt[i].data()); });
                if (rc != KA_ERR_TOO_BIG) break;
                lg.ovf_bbits += 2;
            }
            if (rc) return rc;
            if (getenv("KA_LOAD_TRACE"))
                fprintf(stderr, "[db load] line table (%llu lines of DB, %u x 2^%u table lines, %llu spilled, %llu overflowed): %.2f s\n",
                        (unsigned long long)n, lg.c, lg.s, (unsigned long long)cnt[0][1], (unsigned long long)cnt[0][2],
                        std::chrono::duration<double>(std::chrono::steady_clock::now() - tb).count());
            e->lgeom = lg;
            e->geom = TableView{};
            e->geom.K = K;
            e->line = true;
            e->db_table_mode = 0;
            e->info = ka_db_info{};
            e->info.K = K;
            e->info.n_symbols = nsym;
            e->info.n_lines = n;
            e->info.n_keys = cnt[0][0];
            e->info.n_buckets = (uint64_t)lg.n_lines * 4;
            e->info.table_bytes = (uint64_t)lg.n_lines * 128 + ((uint64_t)64 << lg.ovf_bbits);
            e->info.max_probe = cnt[0][2] ? 3 : (cnt[0][1] ? 2 : 1);
            e->info.slot_bits = 16;
            e->info.filter_bytes = (uint64_t)lg.n_lines * 16;
            e->info.n_spilled = cnt[0][1];
            e->info.n_overflow = cnt[0][2];
            e->have_db = true;
            e->db_serial++;
            for (Device& d : e->devs) {
                cudaSetDevice(d.id);
                for (int k = 0; k < NPIPE; k++) set_l2_window(e, d, d.pipe[k].st);
            }
            return KA_OK;
        }
        if (forced) return fail(e, KA_ERR_TOO_BIG, "ka_db_load: %llu k-mers (K=%d, %d symbols) do not fit the 16-bit line table", (unsigned long long)n, K, nsym);
    }
    TableView geom;
    const uint32_t n_shards = e->table_mode >= 1 ? (uint32_t)e->devs.size() : 1u;
    if (e->table_mode >= 1) {
        if (n_shards != 2 && n_shards != 4 && n_shards != 8)
            return fail(e, KA_ERR_INVALID, "ka_db_load: a sharded table needs an engine on 2, 4 or 8 devices (has %u)", n_shards);
        if (!e->peers_enabled) {
            for (Device& a : e->devs) {
                cudaSetDevice(a.id);
                for (Device& b : e->devs) {
                    if (a.id == b.id) continue;
                    int can = 0;
                    cudaDeviceCanAccessPeer(&can, a.id, b.id);
                    if (!can) return fail(e, KA_ERR_NO_DEVICE, "ka_db_load: device %d cannot access device %d's memory (no NVLink/P2P)", a.id, b.id);
                    cudaError_t pe = cudaDeviceEnablePeerAccess(b.id, 0);
                    if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled)
                        return fail(e, KA_ERR_CUDA, "ka_db_load: cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(pe));
                    cudaGetLastError();
                }
            }
            e->peers_enabled = true;
        }
    }
    // communicators first: NCCL sets up its buffers before the table takes most of the HBM
    if (e->table_mode == 2) { int nrc = route_init_comms(e); if (nrc) return nrc; }   // (table_mode 3 routes with peer stores: no NCCL)
    if (!choose_geometry(n, K, max_role, e->load_factor > 0 ? e->load_factor : 0.4, e->slot_bits == 16 ? 0 : e->slot_bits, n_shards, e->wide != 0, geom))
        return fail(e, KA_ERR_TOO_BIG, "ka_db_load: %llu k-mers (K=%d, max role %d) do not fit %s",
                    (unsigned long long)n, K, max_role, e->slot_bits ? "the forced slot width" : "any slot class of this build");

    // 3. build one replica per device
    auto tb = std::chrono::steady_clock::now();
    std::vector<uint64_t> nk(e->devs.size(), 0);
    std::vector<uint32_t> mp(e->devs.size(), 0);
    int rc = KA_OK;
    for (int attempt = 0; attempt < 6; attempt++) {
        if (!geom.wide && (uint64_t)geom.n_primary_slots + (uint64_t)n_shards * (2ull << geom.ovf_bbits) >= 0xfffffff0ull) {
            if (geom.cls == 128) return fail(e, KA_ERR_TOO_BIG, "ka_db_load: table exceeds the 32-bit slot index of the 128-bit slot class");
            geom.wide = 1; geom.n_primary_slots = 0;   // overflow entries pushed the token range past 32 bits
        }
        rc = for_each_device(e, [&](Device& d, int i) {
            TableView g = geom;
            g.my_shard = n_shards > 1 ? (uint32_t)i : 0u;
            return build_table(e, d, g, src, &nk[i], &mp[i]);
        });
        if (rc != KA_ERR_TOO_BIG) break;
        geom.ovf_bbits += 2;  // overflow table was too small for this key set: rebuild 4x larger
    }
    if (rc) return rc;
    if (getenv("KA_LOAD_TRACE"))
        fprintf(stderr, "[db load] table build (%llu lines, 2^%u sectors of %d-bit slots%s, %u shard(s)): %.2f s\n",
                (unsigned long long)n, geom.bbits, geom.cls, geom.wide ? ", wide" : "", n_shards,
                std::chrono::duration<double>(std::chrono::steady_clock::now() - tb).count());
    if (n_shards > 1) {
        // every device gets the peer pointers of all shards
        std::vector<const uint4*> ps(8, nullptr), po(8, nullptr);
        for (size_t i = 0; i < e->devs.size(); i++) { ps[i] = e->devs[i].table; po[i] = e->devs[i].ovf; nk[0] += i ? nk[i] : 0; mp[0] = std::max(mp[0], mp[i]); }
        for (Device& d : e->devs) {
            cudaSetDevice(d.id);
            if (!d.shard_sectors && cudaMalloc((void**)&d.shard_sectors, 64) != cudaSuccess) return fail(e, KA_ERR_OOM, "shard pointer table");
            if (!d.shard_ovf && cudaMalloc((void**)&d.shard_ovf, 64) != cudaSuccess) return fail(e, KA_ERR_OOM, "shard pointer table");
            cudaMemcpy((void*)d.shard_sectors, ps.data(), 64, cudaMemcpyHostToDevice);
            cudaMemcpy((void*)d.shard_ovf, po.data(), 64, cudaMemcpyHostToDevice);
        }
    }
    e->geom = geom;
    e->line = false;
    e->db_table_mode = e->table_mode;
    e->db_serial++;
    e->info = ka_db_info{};
    e->info.K = K;
    e->info.n_symbols = nsym;
    e->info.n_lines = n;
    e->info.n_keys = nk[0];
    e->info.n_buckets = 1ull << geom.bbits;
    e->info.table_bytes = (32ull << geom.bbits) + (geom.cls == 128 ? 0 : (uint64_t)n_shards * (64ull << geom.ovf_bbits));  // all shards
    e->info.max_probe = mp[0];
    e->info.slot_bits = (uint32_t)geom.cls;
    e->have_db = true;
    for (Device& d : e->devs) {
        cudaSetDevice(d.id);
        for (int k = 0; k < NPIPE; k++) set_l2_window(e, d, d.pipe[k].st);
    }
    return KA_OK;
}

}  // namespace kai
