// ka_synth.cpp — seeded synthetic workloads for tests and bench.py (libkasynth.so).
//
// Shapes follow SURVEY.md §8(d): 20-letter residues with the empirical frequencies of the
// reference fixture src/test/small.gto; role families = one ancestor per role (length
// log-normal, median 264, mean ~311, clipped [35, 3000]) whose members carry i.i.d. 10 %
// substitutions; a proteome = 70 % family members (roles Zipf s=1) + 30 % random proteins,
// 1 % of proteins get an internal tandem repeat so that within-protein duplicate k-mers
// occur; the signature table = every K-window of the first `members_per_role` members of
// each role, reduced by the `build` rule (k-mers seen in two roles are dropped,
// BuildKmerProcessor.java:183-190) and topped up with uniformly random k-mers.
// Everything is a pure function of (seed, indices): no global state, any thread count.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

inline uint64_t splitmix(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline uint64_t stream(uint64_t seed, uint64_t a, uint64_t b, uint64_t c) {
    uint64_t s = seed ^ (a * 0xD1B54A32D192ED03ull) ^ (b * 0x8CB92BA72F3D8DD7ull) ^ (c * 0xABC98388FB8FAC03ull);
    splitmix(s);
    return s;
}
inline double unit(uint64_t& s) { return (double)(splitmix(s) >> 11) * (1.0 / 9007199254740992.0); }

// per-mille residue frequencies of small.gto
const char kLetters[21] = "LAVGDITKSEQNRPFYMHWC";
const int kPermille[20] = {98, 90, 70, 67, 62, 62, 60, 60, 56, 55, 52, 44, 44, 40, 38, 35, 27, 23, 11, 6};

struct ResidueTable {
    uint8_t t[1000];
    ResidueTable() {
        int k = 0;
        for (int i = 0; i < 20; i++)
            for (int j = 0; j < kPermille[i]; j++) t[k++] = (uint8_t)kLetters[i];
    }
};
const ResidueTable kRes;
inline uint8_t residue(uint64_t& s) { return kRes.t[splitmix(s) % 1000]; }

inline uint32_t lognormal_len(uint64_t& s) {
    double u1 = unit(s), u2 = unit(s);
    if (u1 < 1e-300) u1 = 1e-300;
    double z = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    double L = std::exp(std::log(264.0) + 0.5725 * z);
    if (L < 35) L = 35;
    if (L > 3000) L = 3000;
    return (uint32_t)L;
}

}  // namespace

struct kas_families {
    uint64_t seed;
    uint32_t n_roles;
    std::vector<uint64_t> off;   // ancestor CSR
    std::vector<uint8_t> res;
    std::vector<double> zipf;    // cumulative 1/rank
};

namespace {

// member m of role r, written to out (length = ancestor length)
void write_member(const kas_families* f, uint32_t r, uint64_t m, uint8_t* out) {
    uint64_t a = f->off[r], L = f->off[r + 1] - a;
    uint64_t s = stream(f->seed, 0x6d656d62, r, m);
    for (uint64_t i = 0; i < L; i++) {
        uint64_t x = splitmix(s);
        if (x % 10 == 0) { uint64_t t = x >> 8; out[i] = kRes.t[t % 1000]; }
        else out[i] = f->res[a + i];
    }
}

uint32_t zipf_role(const kas_families* f, uint64_t& s) {
    double u = unit(s) * f->zipf.back();
    return (uint32_t)(std::lower_bound(f->zipf.begin(), f->zipf.end(), u) - f->zipf.begin());
}

// One protein of genome g.  mode 0: SURVEY C1/C2/C3 shape.  mode 1/2: config-4 lengths
// (log-uniform 50..5000 / bimodal 90 % 50..300 + 10 % 3000..5000) built by concatenating
// members (same role: stays unanimous; 20 % switch role once: ambiguous) or random residues.
// Returns the length; out may be NULL to only size.  role_out = the family role or -1.
uint32_t make_protein(const kas_families* f, uint64_t seed, uint64_t g, uint64_t p, int mode,
                      int K, uint8_t* out, int32_t* role_out) {
    uint64_t s = stream(seed, 0x70726f74, g, p);
    double kind = unit(s);
    uint32_t L;
    int32_t role = -1;
    if (mode == 0) {
        if (kind < 0.7) {
            role = (int32_t)zipf_role(f, s);
            uint64_t m = 8 + (splitmix(s) >> 16);
            L = (uint32_t)(f->off[role + 1] - f->off[role]);
            if (out) write_member(f, (uint32_t)role, m, out);
        } else {
            L = lognormal_len(s);
            if (out) for (uint32_t i = 0; i < L; i++) out[i] = residue(s);
        }
    } else {
        double u = unit(s);
        if (mode == 1) L = (uint32_t)(50.0 * std::pow(100.0, u));
        else L = unit(s) < 0.9 ? (uint32_t)(50 + u * 250) : (uint32_t)(3000 + u * 2000);
        if (kind < 0.7) {
            role = (int32_t)zipf_role(f, s);
            bool switch_role = kind < 0.2;
            uint32_t done = 0;
            std::vector<uint8_t> tmp;
            int32_t r = role;
            while (done < L) {
                uint64_t m = 8 + (splitmix(s) >> 16);
                uint32_t ml = (uint32_t)(f->off[r + 1] - f->off[r]);
                uint32_t take = std::min(ml, L - done);
                if (out) {
                    tmp.resize(ml);
                    write_member(f, (uint32_t)r, m, tmp.data());
                    memcpy(out + done, tmp.data(), take);
                }
                done += take;
                if (switch_role) { r = (int32_t)zipf_role(f, s); switch_role = false; role = -1; }
            }
        } else if (out) {
            for (uint32_t i = 0; i < L; i++) out[i] = residue(s);
        }
    }
    // 1 % of proteins: internal tandem repeat (duplicate k-mers inside one protein)
    uint64_t s2 = stream(seed, 0x72657065, g, p);
    if (splitmix(s2) % 100 == 0) {
        uint32_t u = (uint32_t)(K + 2 + splitmix(s2) % 30);
        if (L >= 3 * u && out) {
            uint32_t start = (uint32_t)(splitmix(s2) % (L - 2 * u));
            memcpy(out + start + u, out + start, u);
        }
    }
    if (role_out) *role_out = role;
    return L;
}

}  // namespace

extern "C" {

kas_families* kas_families_new(uint64_t seed, uint32_t n_roles) {
    kas_families* f = new kas_families();
    f->seed = seed; f->n_roles = n_roles;
    f->off.resize((size_t)n_roles + 1);
    f->off[0] = 0;
    for (uint32_t r = 0; r < n_roles; r++) {
        uint64_t s = stream(seed, 0x616e6373, r, 0);
        f->off[r + 1] = f->off[r] + lognormal_len(s);
    }
    f->res.resize(f->off[n_roles]);
    for (uint32_t r = 0; r < n_roles; r++) {
        uint64_t s = stream(seed, 0x616e6373, r, 1);
        for (uint64_t i = f->off[r]; i < f->off[r + 1]; i++) f->res[i] = residue(s);
    }
    f->zipf.resize(n_roles);
    double c = 0;
    for (uint32_t r = 0; r < n_roles; r++) { c += 1.0 / (double)(r + 1); f->zipf[r] = c; }
    return f;
}

void kas_families_free(kas_families* f) { delete f; }

uint64_t kas_family_len(const kas_families* f, uint32_t role) { return f->off[role + 1] - f->off[role]; }

// Sizes of genomes [g0, g0+n_genomes) x n_prot proteins: returns total residues.
uint64_t kas_batch_size(const kas_families* f, uint64_t seed, uint64_t g0, uint64_t n_genomes,
                        uint32_t n_prot, int mode, int K) {
    uint64_t total = 0;
    for (uint64_t g = g0; g < g0 + n_genomes; g++)
        for (uint32_t p = 0; p < n_prot; p++) total += make_protein(f, seed, g, p, mode, K, nullptr, nullptr);
    return total;
}

// Fill a CSR batch (offsets[0] = 0).  residues must hold kas_batch_size() bytes, offsets
// n_genomes*n_prot + 1 entries, true_role (may be NULL) one per protein.
void kas_batch_fill(const kas_families* f, uint64_t seed, uint64_t g0, uint64_t n_genomes,
                    uint32_t n_prot, int mode, int K, int n_threads, uint8_t* residues,
                    uint64_t* offsets, int32_t* true_role) {
    uint64_t n = n_genomes * n_prot;
    // pass 1: lengths (parallel over genomes), then exclusive scan
    if (n_threads < 1) n_threads = 1;
    auto run = [&](auto fn) {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++) th.emplace_back(fn, t);
        for (auto& x : th) x.join();
    };
    run([&](int t) {
        for (uint64_t g = t; g < n_genomes; g += n_threads)
            for (uint32_t p = 0; p < n_prot; p++)
                offsets[g * n_prot + p + 1] = make_protein(f, seed, g0 + g, p, mode, K, nullptr, nullptr);
    });
    offsets[0] = 0;
    for (uint64_t i = 0; i < n; i++) offsets[i + 1] += offsets[i];
    run([&](int t) {
        for (uint64_t g = t; g < n_genomes; g += n_threads)
            for (uint32_t p = 0; p < n_prot; p++) {
                uint64_t i = g * n_prot + p;
                int32_t role;
                make_protein(f, seed, g0 + g, p, mode, K, residues + offsets[i], &role);
                if (true_role) true_role[i] = role;
            }
    });
}

// Signature table.  Writes up to `target` lines (kmers_out: target*K bytes, roles_out:
// target ints) and returns the number written.  Lines come out in hash-slot order, i.e.
// unordered, like the reference's kmerdb.tbl (BuildKmerProcessor.java:212-216).
// The key space is split into 64 hash partitions with one private open-addressed set each,
// so any number of threads produces the same table.
uint64_t kas_table(const kas_families* f, uint64_t seed, int K, uint32_t members_per_role,
                   uint64_t target, int n_threads, uint8_t* kmers_out, int32_t* roles_out) {
    if (K < 1 || K > 12 || target == 0) return 0;
    if (n_threads < 1) n_threads = 1;
    constexpr int P = 64;
    uint64_t cap = 64;
    while (cap < target * 3 / P + 1024) cap <<= 1;
    std::vector<std::vector<uint64_t>> keys(P);
    std::vector<std::vector<int32_t>> vals(P);  // role, or -2 = seen in two roles (dropped)
    auto pack = [K](const uint8_t* s) {
        uint64_t k = 0;
        for (int j = 0; j < K; j++) k = (k << 5) | (uint64_t)(s[j] - 'A' + 1);
        return k;
    };
    auto hash = [](uint64_t k) {
        uint64_t h = k * 0x9E3779B97F4A7C15ull;
        h ^= h >> 32;
        h *= 0xD6E8FEB86659FD93ull;
        h ^= h >> 29;
        return h;
    };
    auto run = [&](auto fn) {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++) th.emplace_back(fn, t);
        for (auto& x : th) x.join();
    };
    std::vector<uint64_t> live(P, 0), used(P, 0);
    // phase 1: every K-window of the first members of every role; a k-mer seen in two roles
    // is not discriminating and is dropped (BuildKmerProcessor.java:183-190)
    run([&](int t) {
        for (int p = t; p < P; p += n_threads) { keys[p].assign(cap, 0); vals[p].assign(cap, 0); }
        std::vector<uint8_t> buf;
        for (uint32_t r = 0; r < f->n_roles; r++) {
            uint64_t L = f->off[r + 1] - f->off[r];
            buf.resize(L);
            for (uint32_t m = 0; m < members_per_role; m++) {
                write_member(f, r, m, buf.data());
                for (uint64_t i = 0; i + K <= L; i++) {
                    uint64_t k = pack(buf.data() + i), h = hash(k);
                    int p = (int)(h >> 58);
                    if (p % n_threads != t) continue;
                    if (used[p] * 4 >= cap * 3) continue;  // partition full: later windows are skipped
                    uint64_t sl = h & (cap - 1);
                    while (keys[p][sl] != 0 && keys[p][sl] != k) sl = (sl + 1) & (cap - 1);
                    if (keys[p][sl] == 0) { keys[p][sl] = k; vals[p][sl] = (int32_t)r; live[p]++; used[p]++; }
                    else if (vals[p][sl] >= 0 && vals[p][sl] != (int32_t)r) { vals[p][sl] = -2; live[p]--; }
                }
            }
        }
    });
    // phase 2: top up with uniformly random k-mers (uniform letters) and random roles
    for (int round = 0; round < 8; round++) {
        uint64_t have = 0;
        for (int p = 0; p < P; p++) have += live[p];
        if (have >= target) break;
        uint64_t need = target - have;
        uint64_t n_cand = need + need / 64 + 256;
        std::vector<uint64_t> cand(n_cand);
        std::vector<int32_t> crole(n_cand);
        const uint64_t BLK = 1 << 18;
        uint64_t n_blk = (n_cand + BLK - 1) / BLK;
        run([&](int t) {
            for (uint64_t b = t; b < n_blk; b += n_threads) {
                uint64_t s = stream(seed, 0x746f7075, (uint64_t)round, b);
                uint64_t e = std::min(n_cand, (b + 1) * BLK);
                for (uint64_t i = b * BLK; i < e; i++) {
                    uint64_t k = 0, x = splitmix(s);
                    for (int j = 0; j < K; j++) { k = (k << 5) | (uint64_t)(kLetters[x % 20] - 'A' + 1); x /= 20; }
                    cand[i] = k;
                    crole[i] = (int32_t)(splitmix(s) % f->n_roles);
                }
            }
        });
        run([&](int t) {
            for (uint64_t i = 0; i < n_cand; i++) {
                uint64_t k = cand[i], h = hash(k);
                int p = (int)(h >> 58);
                if (p % n_threads != t) continue;
                if (used[p] * 8 >= cap * 7) continue;
                uint64_t sl = h & (cap - 1);
                while (keys[p][sl] != 0 && keys[p][sl] != k) sl = (sl + 1) & (cap - 1);
                if (keys[p][sl] == 0) { keys[p][sl] = k; vals[p][sl] = crole[i]; live[p]++; used[p]++; }
            }
        });
        if (K < 6) break;  // a small K can exhaust its key space before `target`
    }
    // emit partition-major in slot order, truncated at target
    std::vector<uint64_t> start(P + 1, 0);
    for (int p = 0; p < P; p++) start[p + 1] = start[p] + live[p];
    run([&](int t) {
        for (int p = t; p < P; p += n_threads) {
            uint64_t n = start[p];
            for (uint64_t i = 0; i < cap && n < target; i++) {
                if (keys[p][i] == 0 || vals[p][i] < 0) continue;
                uint64_t k = keys[p][i];
                for (int j = K - 1; j >= 0; j--) { kmers_out[n * K + j] = (uint8_t)('A' - 1 + (k & 31)); k >>= 5; }
                roles_out[n] = vals[p][i];
                n++;
            }
        }
    });
    return std::min(start[P], target);
}

}  // extern "C"
