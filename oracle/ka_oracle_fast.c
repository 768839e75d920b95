/*
 * ka_oracle_fast.c — second, independent CPU restatement of the `apply` decision
 * (ApplyKmerProcessor.java:99-148) with packed integer keys: the "best reasonable CPU"
 * line of BASELINE.md §4 and a cross-check of ka_oracle.c (different data structures, same
 * answers).  TEST INFRASTRUCTURE ONLY — see the header of ka_oracle.c; PARITY UNPINNED for
 * the same reasons.
 *
 * Keys: each residue byte is mapped to a dense code (1..n_sym, by first appearance in the
 * DB) and a k-mer is the base-(n_sym+1) number of its K codes, exact while it fits 64 bits
 * (else orf_db_load fails).  Table: open addressing, linear probing, last line wins.
 * Per protein: hitting (key, role) pairs are collected, sorted, de-duplicated
 * (HashSet semantics of ProteinKmers), then the unanimity / min-hits rule is applied.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct orf_db {
    uint64_t* keys; /* 0 = empty; stored key = packed + 1 */
    int32_t* roles;
    uint64_t cap;   /* power of two */
    uint64_t size;
    uint32_t base;  /* n_sym + 1 */
    int K;
    uint8_t code[256];
} orf_db;

static inline uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

void orf_db_free(orf_db* d) {
    if (!d) return;
    free(d->keys); free(d->roles); free(d);
}

orf_db* orf_db_load(const uint8_t* kmers, const int32_t* roles, uint64_t n, int K) {
    orf_db* d = (orf_db*)calloc(1, sizeof(orf_db));
    if (!d) return NULL;
    d->K = K;
    uint32_t nsym = 0;
    for (uint64_t i = 0; i < n * (uint64_t)K; i++)
        if (!d->code[kmers[i]]) d->code[kmers[i]] = (uint8_t)(++nsym);
    d->base = nsym + 1;
    /* (base^K) must fit 64 bits */
    long double span = 1;
    for (int j = 0; j < K; j++) span *= d->base;
    if (span >= 18446744073709551615.0L) { free(d); return NULL; }
    d->cap = 16;
    while (d->cap < 2 * n + 16) d->cap <<= 1;
    d->keys = (uint64_t*)calloc(d->cap, 8);
    d->roles = (int32_t*)malloc(d->cap * 4);
    if (!d->keys || !d->roles) { orf_db_free(d); return NULL; }
    for (uint64_t i = 0; i < n; i++) {
        uint64_t k = 0;
        for (int j = 0; j < K; j++) k = k * d->base + d->code[kmers[i * (uint64_t)K + j]];
        k += 1;
        uint64_t s = mix(k) & (d->cap - 1);
        while (d->keys[s] && d->keys[s] != k) s = (s + 1) & (d->cap - 1);
        if (!d->keys[s]) { d->keys[s] = k; d->size++; }
        d->roles[s] = roles[i]; /* put(): last line wins (ApplyKmerProcessor.java:106) */
    }
    return d;
}

uint64_t orf_db_size(const orf_db* d) { return d->size; }

typedef struct { uint64_t key; int32_t role; } hit_t;

static int hit_cmp(const void* a, const void* b) {
    uint64_t x = ((const hit_t*)a)->key, y = ((const hit_t*)b)->key;
    return x < y ? -1 : x > y;
}

typedef struct {
    const orf_db* db;
    const uint8_t* residues;
    const uint64_t* offsets;
    uint64_t s0, s1;
    int min_hits, distinct;
    int32_t* role; int32_t* hits; uint8_t* flag;
} fast_job;

static void* fast_worker(void* arg) {
    fast_job* j = (fast_job*)arg;
    const orf_db* d = j->db;
    const int K = d->K;
    hit_t* buf = NULL;
    uint64_t buf_cap = 0;
    uint64_t top = 1; /* base^(K-1) */
    for (int i = 1; i < K; i++) top *= d->base;
    for (uint64_t s = j->s0; s < j->s1; s++) {
        const uint8_t* p = j->residues + j->offsets[s];
        uint64_t L = j->offsets[s + 1] - j->offsets[s];
        uint64_t nh = 0;
        if (L >= (uint64_t)K) {
            if (L > buf_cap) { free(buf); buf_cap = L * 2; buf = (hit_t*)malloc(buf_cap * sizeof(hit_t)); }
            uint64_t k = 0;
            int run = 0; /* consecutive residues that occur in the DB alphabet */
            for (uint64_t i = 0; i < L; i++) {
                uint32_t c = d->code[p[i]];
                if (!c) { k = 0; run = 0; continue; } /* byte not in the DB: no window over it can hit */
                if (run == K) k -= (uint64_t)d->code[p[i - K]] * top; /* drop the oldest code */
                else run++;
                k = k * d->base + c;
                if (run == K) {
                    uint64_t kk = k + 1;
                    uint64_t sl = mix(kk) & (d->cap - 1);
                    while (d->keys[sl] && d->keys[sl] != kk) sl = (sl + 1) & (d->cap - 1);
                    if (d->keys[sl]) { buf[nh].key = kk; buf[nh].role = d->roles[sl]; nh++; }
                }
            }
        }
        int32_t role = -1, cnt = 0;
        int ambiguous = 0;
        if (nh) {
            if (j->distinct) qsort(buf, nh, sizeof(hit_t), hit_cmp);
            for (uint64_t i = 0; i < nh; i++) {
                if (j->distinct && i && buf[i].key == buf[i - 1].key) continue;
                if (cnt == 0) role = buf[i].role;
                else if (buf[i].role != role) ambiguous = 1;
                cnt++;
            }
        }
        uint8_t f;
        if (!cnt) { role = -1; f = 0; }
        else if (ambiguous) { role = -1; cnt = 0; f = 2; }
        else if (cnt >= j->min_hits) f = 1;
        else { role = -1; f = 3; }
        j->role[s] = role; j->hits[s] = cnt;
        if (j->flag) j->flag[s] = f;
    }
    free(buf);
    return NULL;
}

int orf_apply(const orf_db* db, const uint8_t* residues, const uint64_t* offsets, uint64_t N,
              int min_hits, int distinct, int n_threads, int32_t* role, int32_t* hits,
              uint8_t* flag) {
    if (min_hits < 1) return -1;
    if (n_threads < 1) n_threads = 1;
    if ((uint64_t)n_threads > N) n_threads = N ? (int)N : 1;
    fast_job* jobs = (fast_job*)calloc((size_t)n_threads, sizeof(fast_job));
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    uint64_t base = N ? offsets[0] : 0, total = N ? offsets[N] - base : 0, s = 0;
    for (int t = 0; t < n_threads; t++) {
        uint64_t target = base + (total * (uint64_t)(t + 1)) / (uint64_t)n_threads, e = s;
        if (t == n_threads - 1) e = N;
        else while (e < N && offsets[e + 1] <= target) e++;
        jobs[t] = (fast_job){db, residues, offsets, s, e, min_hits, distinct, role, hits, flag};
        s = e;
    }
    if (n_threads == 1) fast_worker(&jobs[0]);
    else {
        for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, fast_worker, &jobs[t]);
        for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    }
    free(jobs); free(th);
    return 0;
}
