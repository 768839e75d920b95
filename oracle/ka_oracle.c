/*
 * ka_oracle.c — CPU ORACLE for the k-mer annotation hot path of SEEDtk/kmers.anno.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (libkmeranno.so) never links, loads or calls anything in oracle/.
 *
 * PARITY UNPINNED: the reference's own tests hold no golden vector for `apply`, `build`
 * or ProteinKmers, no JVM exists in this image, and the k-mer
 * extraction class org.theseed.sequence.ProteinKmers lives in an un-vendored module
 * (org.theseed:sequence:1.0.0 / org.theseed:shared:1.0.0, pom.xml:48-72).  This file
 * restates (a) the in-repo Java line by line and (b) the recalled published behaviour of
 * ProteinKmers, with a switch for each recalled point.  The only reference-held pins it
 * can be checked against are RoleTests.java:15-36 (RoleCounter) and the substring
 * property of AppTest.java:145-161 (countPegKmers); tests/test_oracle.py checks both.
 *
 * It is deliberately Java-shaped (String keys, String.hashCode, HashMap bins, a HashSet
 * per protein) so that it is a fair stand-in for the JVM path when timed.
 *
 * File:line citations are relative to /root/reference/src/main/java/org/theseed/ .
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------
 * java.util.HashMap<String, int> restated: power-of-two table, hash spreading, bins are
 * singly linked lists appended at the tail, resize splits bins preserving order.  (JDK
 * treeification of bins with >= 8 nodes is not restated: it changes neither get/put/remove
 * results nor, for the first 8 nodes, iteration order.)  Iteration order = table index,
 * then bin order, which is what `for (Map.Entry e : map.entrySet())` yields
 * (BuildKmerProcessor.java:212).
 * ---------------------------------------------------------------------------------- */
typedef struct {
    uint32_t hash;
    int32_t next; /* node index or -1 */
    int32_t value;
    int32_t aux;   /* second payload (RoleCounter.badCount in build) */
    uint32_t klen;
    uint64_t koff; /* offset of the key bytes in pool */
} jnode;

typedef struct orc_map {
    int32_t* table; /* head node index per bin, -1 = empty; NULL until first put */
    uint32_t cap;
    int64_t threshold;
    jnode* nodes;
    uint64_t n_nodes, nodes_cap;
    char* pool;
    uint64_t pool_len, pool_cap;
    uint64_t size;
} orc_map;

#define JMAX_CAP (1u << 30)

/* String.hashCode over Latin-1 bytes, then HashMap.hash(): h ^ (h >>> 16). */
static inline uint32_t jstring_hash(const uint8_t* s, uint32_t n) {
    uint32_t h = 0;
    for (uint32_t i = 0; i < n; i++) h = 31u * h + s[i];
    return h ^ (h >> 16);
}

/* HashMap.tableSizeFor */
static uint32_t table_size_for(int64_t c) {
    if (c <= 1) return 1;
    uint64_t n = (uint64_t)c - 1;
    n |= n >> 1; n |= n >> 2; n |= n >> 4; n |= n >> 8; n |= n >> 16;
    n += 1;
    return n >= JMAX_CAP ? JMAX_CAP : (uint32_t)n;
}

/* new HashMap<>(initialCapacity); initial_capacity < 0 selects the default constructor
 * (table of 16 on first put). */
orc_map* orc_map_new(int64_t initial_capacity) {
    orc_map* m = (orc_map*)calloc(1, sizeof(orc_map));
    if (!m) return NULL;
    m->cap = 0;
    m->threshold = initial_capacity < 0 ? 0 : (int64_t)table_size_for(initial_capacity);
    return m;
}

void orc_map_free(orc_map* m) {
    if (!m) return;
    free(m->table); free(m->nodes); free(m->pool); free(m);
}

uint64_t orc_map_size(const orc_map* m) { return m->size; }
uint32_t orc_map_capacity(const orc_map* m) { return m->cap; }

/* reserve node/pool storage up front (pure allocation hint, no semantic effect) */
int orc_map_reserve(orc_map* m, uint64_t n_nodes, uint64_t pool_bytes) {
    if (n_nodes > m->nodes_cap) {
        jnode* p = (jnode*)realloc(m->nodes, n_nodes * sizeof(jnode));
        if (!p) return -1;
        m->nodes = p; m->nodes_cap = n_nodes;
    }
    if (pool_bytes > m->pool_cap) {
        char* p = (char*)realloc(m->pool, pool_bytes);
        if (!p) return -1;
        m->pool = p; m->pool_cap = pool_bytes;
    }
    return 0;
}

static int jresize(orc_map* m) {
    uint32_t old_cap = m->cap, new_cap;
    int64_t new_thr;
    if (old_cap > 0) {
        if (old_cap >= JMAX_CAP) { m->threshold = INT32_MAX; return 0; }
        new_cap = old_cap << 1;
    } else if (m->threshold > 0) {
        new_cap = (uint32_t)m->threshold; /* initial capacity was placed in threshold */
    } else {
        new_cap = 16;
    }
    new_thr = (int64_t)((float)new_cap * 0.75f);
    int32_t* nt = (int32_t*)malloc((size_t)new_cap * sizeof(int32_t));
    if (!nt) return -1;
    memset(nt, 0xff, (size_t)new_cap * sizeof(int32_t));
    if (m->table) {
        /* split every bin into lo/hi lists, preserving relative order (HashMap.resize) */
        for (uint32_t j = 0; j < old_cap; j++) {
            int32_t e = m->table[j];
            int32_t lo_h = -1, lo_t = -1, hi_h = -1, hi_t = -1;
            while (e >= 0) {
                int32_t nx = m->nodes[e].next;
                if ((m->nodes[e].hash & old_cap) == 0) {
                    if (lo_t < 0) lo_h = e; else m->nodes[lo_t].next = e;
                    lo_t = e;
                } else {
                    if (hi_t < 0) hi_h = e; else m->nodes[hi_t].next = e;
                    hi_t = e;
                }
                e = nx;
            }
            if (lo_t >= 0) { m->nodes[lo_t].next = -1; nt[j] = lo_h; }
            if (hi_t >= 0) { m->nodes[hi_t].next = -1; nt[j + old_cap] = hi_h; }
        }
        free(m->table);
    }
    m->table = nt; m->cap = new_cap; m->threshold = new_thr;
    return 0;
}

static inline int key_eq(const orc_map* m, const jnode* nd, const uint8_t* k, uint32_t klen) {
    return nd->klen == klen && memcmp(m->pool + nd->koff, k, klen) == 0;
}

/* returns node index or -1 */
static inline int32_t jfind(const orc_map* m, const uint8_t* k, uint32_t klen, uint32_t h) {
    if (!m->table) return -1;
    int32_t e = m->table[h & (m->cap - 1)];
    while (e >= 0) {
        const jnode* nd = &m->nodes[e];
        if (nd->hash == h && key_eq(m, nd, k, klen)) return e;
        e = nd->next;
    }
    return -1;
}

/* HashMap.putVal: returns node index (existing or new), *created tells which; -2 on OOM.
 * An existing node keeps its key and bin position; only the value changes (caller). */
static int32_t jput_node(orc_map* m, const uint8_t* k, uint32_t klen, int* created) {
    uint32_t h = jstring_hash(k, klen);
    if (!m->table && jresize(m)) return -2;
    uint32_t idx = h & (m->cap - 1);
    int32_t e = m->table[idx], tail = -1;
    while (e >= 0) {
        jnode* nd = &m->nodes[e];
        if (nd->hash == h && key_eq(m, nd, k, klen)) { *created = 0; return e; }
        tail = e; e = nd->next;
    }
    if (m->n_nodes >= (uint64_t)INT32_MAX) return -2;
    if (m->n_nodes == m->nodes_cap) {
        uint64_t nc = m->nodes_cap ? m->nodes_cap * 2 : 1024;
        jnode* p = (jnode*)realloc(m->nodes, nc * sizeof(jnode));
        if (!p) return -2;
        m->nodes = p; m->nodes_cap = nc;
    }
    if (m->pool_len + klen > m->pool_cap) {
        uint64_t nc = m->pool_cap ? m->pool_cap * 2 : 16384;
        while (nc < m->pool_len + klen) nc *= 2;
        char* p = (char*)realloc(m->pool, nc);
        if (!p) return -2;
        m->pool = p; m->pool_cap = nc;
    }
    int32_t ni = (int32_t)m->n_nodes++;
    jnode* nn = &m->nodes[ni];
    nn->hash = h; nn->next = -1; nn->value = 0; nn->aux = 0; nn->klen = klen; nn->koff = m->pool_len;
    memcpy(m->pool + m->pool_len, k, klen);
    m->pool_len += klen;
    if (tail < 0) m->table[idx] = ni; else m->nodes[tail].next = ni;
    *created = 1;
    if ((int64_t)++m->size > m->threshold) { if (jresize(m)) return -2; }
    return ni;
}

/* map.put(key, value): last put wins (ApplyKmerProcessor.java:106) */
int orc_map_put(orc_map* m, const uint8_t* k, uint32_t klen, int32_t value) {
    int created;
    int32_t ni = jput_node(m, k, klen, &created);
    if (ni < 0) return -1;
    m->nodes[ni].value = value;
    return created;
}

/* map.get(key): 1 = found */
int orc_map_get(const orc_map* m, const uint8_t* k, uint32_t klen, int32_t* value) {
    int32_t e = jfind(m, k, klen, jstring_hash(k, klen));
    if (e < 0) return 0;
    if (value) *value = m->nodes[e].value;
    return 1;
}

/* map.remove(key): 1 = removed */
int orc_map_remove(orc_map* m, const uint8_t* k, uint32_t klen) {
    if (!m->table) return 0;
    uint32_t h = jstring_hash(k, klen);
    uint32_t idx = h & (m->cap - 1);
    int32_t e = m->table[idx], prev = -1;
    while (e >= 0) {
        jnode* nd = &m->nodes[e];
        if (nd->hash == h && key_eq(m, nd, k, klen)) {
            if (prev < 0) m->table[idx] = nd->next; else m->nodes[prev].next = nd->next;
            m->size--;
            return 1;
        }
        prev = e; e = nd->next;
    }
    return 0;
}

/* Dump entries in HashMap iteration order.  keys_out receives the key bytes back to back,
 * klens/values one per entry.  Returns the number of entries. */
uint64_t orc_map_dump(const orc_map* m, uint8_t* keys_out, uint32_t* klens, int32_t* values) {
    uint64_t n = 0, off = 0;
    if (!m->table) return 0;
    for (uint32_t j = 0; j < m->cap; j++) {
        for (int32_t e = m->table[j]; e >= 0; e = m->nodes[e].next) {
            const jnode* nd = &m->nodes[e];
            if (keys_out) memcpy(keys_out + off, m->pool + nd->koff, nd->klen);
            off += nd->klen;
            if (klens) klens[n] = nd->klen;
            if (values) values[n] = nd->value;
            n++;
        }
    }
    return n;
}

/* ------------------------------------------------------------------------------------
 * DB load: ApplyKmerProcessor.java:99-110.  kmerRoleMap = new HashMap<>((int)(fileLen/30))
 * (:101); every line put(kmer, role) (:106, last wins); K = length of the LAST k-mer
 * (:108).  Role strings are interned to int32 ids by the caller; equality of ids is
 * equality of strings (contentEquals, :137).
 * ---------------------------------------------------------------------------------- */
orc_map* orc_db_load(const uint8_t* kmers, const int32_t* roles, uint64_t n, int K,
                     int64_t file_len_bytes) {
    int64_t cap = file_len_bytes >= 0 ? (int64_t)(int32_t)(file_len_bytes / 30) : -1;
    orc_map* m = orc_map_new(cap < 0 ? 0 : cap);
    if (!m) return NULL;
    if (orc_map_reserve(m, n ? n : 1, n * (uint64_t)K + 1)) { orc_map_free(m); return NULL; }
    for (uint64_t i = 0; i < n; i++) {
        if (orc_map_put(m, kmers + i * (uint64_t)K, (uint32_t)K, roles[i]) < 0) {
            orc_map_free(m);
            return NULL;
        }
    }
    return m;
}

/* Multi-threaded bulk load that produces EXACTLY the structure the sequential put() loop
 * above produces (same table capacity, same bin contents in the same order), so that a
 * 10^8-line DB loads in seconds.  Why it is the same structure: HashMap.resize() splits
 * bins preserving relative order, so after any number of doublings a bin holds its nodes in
 * insertion order — which is what inserting the lines, in line order, into a table of the
 * final capacity gives.  Each thread owns a contiguous range of bins and walks all lines in
 * order.  The final capacity is the first one of Java's doubling sequence whose threshold
 * holds the number of DISTINCT keys. */
typedef struct {
    orc_map* m;
    const uint8_t* kmers;
    const int32_t* roles;
    const uint32_t* hashes;
    uint64_t n;
    int K, t, T;
    uint64_t distinct;
} bulk_job;

static void* bulk_hash_worker(void* p) {
    bulk_job* j = (bulk_job*)p;
    uint64_t a = j->n * (uint64_t)j->t / (uint64_t)j->T, b = j->n * (uint64_t)(j->t + 1) / (uint64_t)j->T;
    uint32_t* hs = (uint32_t*)j->hashes;
    for (uint64_t i = a; i < b; i++) {
        hs[i] = jstring_hash(j->kmers + i * (uint64_t)j->K, (uint32_t)j->K);
        jnode* nd = &j->m->nodes[i];
        nd->hash = hs[i]; nd->next = -1; nd->value = j->roles[i]; nd->aux = 0;
        nd->klen = (uint32_t)j->K; nd->koff = i * (uint64_t)j->K;
    }
    return NULL;
}

static void* bulk_link_worker(void* p) {
    bulk_job* j = (bulk_job*)p;
    orc_map* m = j->m;
    uint32_t cap = m->cap;
    uint32_t lo = (uint32_t)((uint64_t)cap * (uint64_t)j->t / (uint64_t)j->T);
    uint32_t hi = (uint32_t)((uint64_t)cap * (uint64_t)(j->t + 1) / (uint64_t)j->T);
    uint64_t distinct = 0;
    for (uint64_t i = 0; i < j->n; i++) {
        uint32_t idx = j->hashes[i] & (cap - 1);
        if (idx < lo || idx >= hi) continue;
        const uint8_t* k = j->kmers + i * (uint64_t)j->K;
        int32_t e = m->table[idx], tail = -1;
        int found = 0;
        while (e >= 0) {
            jnode* nd = &m->nodes[e];
            if (nd->hash == j->hashes[i] && memcmp(m->pool + nd->koff, k, (size_t)j->K) == 0) {
                nd->value = j->roles[i]; /* put() on an existing key: value replaced, last wins */
                found = 1;
                break;
            }
            tail = e; e = nd->next;
        }
        if (found) continue;
        if (tail < 0) m->table[idx] = (int32_t)i; else m->nodes[tail].next = (int32_t)i;
        distinct++;
    }
    j->distinct = distinct;
    return NULL;
}

orc_map* orc_db_load_mt(const uint8_t* kmers, const int32_t* roles, uint64_t n, int K,
                        int64_t file_len_bytes, int n_threads) {
    if (n_threads < 2 || n < 100000 || n >= (uint64_t)INT32_MAX)
        return orc_db_load(kmers, roles, n, K, file_len_bytes);
    int64_t cap0 = file_len_bytes >= 0 ? (int64_t)(int32_t)(file_len_bytes / 30) : 0;
    orc_map* m = orc_map_new(cap0 < 0 ? 0 : cap0);
    if (!m) return NULL;
    if (orc_map_reserve(m, n, n * (uint64_t)K + 1)) { orc_map_free(m); return NULL; }
    memcpy(m->pool, kmers, n * (uint64_t)K);
    m->pool_len = n * (uint64_t)K;
    m->n_nodes = n;
    uint32_t* hashes = (uint32_t*)malloc(n * sizeof(uint32_t));
    bulk_job* jobs = (bulk_job*)calloc((size_t)n_threads, sizeof(bulk_job));
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    if (!hashes || !jobs || !th) { free(hashes); free(jobs); free(th); orc_map_free(m); return NULL; }
    for (int t = 0; t < n_threads; t++)
        jobs[t] = (bulk_job){m, kmers, roles, hashes, n, K, t, n_threads, 0};
    for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, bulk_hash_worker, &jobs[t]);
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    /* Java's capacity sequence: initial, then doubling while size > threshold */
    uint32_t cap = (uint32_t)m->threshold;
    if (cap < 1) cap = 1;
    uint64_t want = n; /* first guess: all lines distinct */
    for (int pass = 0; pass < 2; pass++) {
        uint32_t c = cap;
        while ((uint64_t)((float)c * 0.75f) < want && c < JMAX_CAP) c <<= 1;
        if (pass == 1 && c == m->cap) break; /* the guess was right */
        free(m->table);
        m->table = (int32_t*)malloc((size_t)c * sizeof(int32_t));
        if (!m->table) { free(hashes); free(jobs); free(th); orc_map_free(m); return NULL; }
        memset(m->table, 0xff, (size_t)c * sizeof(int32_t));
        m->cap = c;
        m->threshold = (int64_t)((float)c * 0.75f);
        if (pass == 1)
            for (uint64_t i = 0; i < n; i++) { m->nodes[i].next = -1; m->nodes[i].value = roles[i]; }
        for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, bulk_link_worker, &jobs[t]);
        uint64_t d = 0;
        for (int t = 0; t < n_threads; t++) { pthread_join(th[t], NULL); d += jobs[t].distinct; }
        m->size = d;
        want = d;
    }
    free(hashes); free(jobs); free(th);
    return m;
}

/* ------------------------------------------------------------------------------------
 * org.theseed.sequence.ProteinKmers (EXTERNAL, recalled — see header): the set of
 * distinct K-substrings of a protein, built with substring() + HashSet.add for
 * i = 0 .. L-K inclusive; L < K gives the empty set; no residue filtering, no case
 * folding.  Switches:
 *   distinct      1 = HashSet semantics (default); 0 = every window position counts
 *   include_last  1 = windows 0..L-K inclusive (default); 0 = drop the last window as
 *                 KmerReference.countPegKmers does (proteins/kmers/KmerReference.java:134-137)
 * Call sites: ApplyKmerProcessor.java:123,128; BuildKmerProcessor.java:167-168,199-200.
 *
 * The per-protein set is a scratch orc_map reset between proteins.
 * ---------------------------------------------------------------------------------- */
typedef struct {
    orc_map* set;
} kmer_set;

static void set_reset(orc_map* s, int64_t windows) {
    /* new HashSet<String>(capacity for `windows` entries at load 0.75) */
    free(s->table); s->table = NULL; s->cap = 0;
    s->threshold = (int64_t)table_size_for(windows * 4 / 3 + 1);
    s->n_nodes = 0; s->pool_len = 0; s->size = 0;
}

/* Build the ProteinKmers set of one protein into s. */
static void protein_kmers(orc_map* s, const uint8_t* prot, uint64_t L, int K, int distinct,
                          int include_last) {
    int64_t n = (int64_t)L - K + (include_last ? 1 : 0); /* number of windows */
    set_reset(s, n > 0 ? n : 0);
    for (int64_t i = 0; i < n; i++) {
        int created;
        if (distinct) {
            jput_node(s, prot + i, (uint32_t)K, &created);
        } else {
            /* positional variant: make every window unique by appending its position */
            uint8_t tmp[256 + 8];
            memcpy(tmp, prot + i, (size_t)K);
            memcpy(tmp + K, &i, 8);
            jput_node(s, tmp, (uint32_t)K + 8, &created);
        }
    }
}

/* ------------------------------------------------------------------------------------
 * apply: ApplyKmerProcessor.java:122-148, one protein.
 * Flags as include/kmeranno.h: 0 none, 1 called, 2 ambiguous, 3 below min.
 * out_hits: the Java `count` when the peg is unanimous (independent of iteration order);
 * for an ambiguous peg Java's count depends on HashSet order and is never reported, so
 * the contract defines it as 0.
 * ---------------------------------------------------------------------------------- */
static void apply_one(const orc_map* db, orc_map* scratch, const uint8_t* prot, uint64_t L,
                      int K, int min_hits, int distinct, int include_last, int32_t* role,
                      int32_t* hits, uint8_t* flag) {
    protein_kmers(scratch, prot, L, K, distinct, include_last); /* :123 */
    int have_role = 0;                                           /* String roleId = null :125 */
    int32_t role_id = -1;
    int count = 0;                                               /* :126 */
    int bad_peg = 0;                                             /* :127 */
    /* Iterator<String> iter = kmers.iterator(); while (iter.hasNext() && !badPeg) :128-129 */
    if (scratch->table) {
        for (uint32_t j = 0; j < scratch->cap && !bad_peg; j++) {
            for (int32_t e = scratch->table[j]; e >= 0 && !bad_peg; e = scratch->nodes[e].next) {
                const jnode* nd = &scratch->nodes[e];
                int32_t possible;
                /* kmerRoleMap.get(iter.next()) :130 — the key is the K-mer text only */
                if (orc_map_get(db, (const uint8_t*)scratch->pool + nd->koff, (uint32_t)K, &possible)) {
                    if (!have_role) { have_role = 1; role_id = possible; count = 1; } /* :133-136 */
                    else if (possible == role_id) count++;                            /* :137-139 */
                    else bad_peg = 1;                                                 /* :140-143 */
                }
            }
        }
    }
    if (have_role && !bad_peg && count >= min_hits) { /* :146 */
        *role = role_id; *hits = count; *flag = 1;    /* reporter.recordFeature(feat, roleId, count) :147 */
    } else if (bad_peg) {
        *role = -1; *hits = 0; *flag = 2;
    } else if (have_role) {
        *role = -1; *hits = count; *flag = 3;
    } else {
        *role = -1; *hits = 0; *flag = 0;
    }
}

typedef struct {
    const orc_map* db;
    const uint8_t* residues;
    const uint64_t* offsets;
    uint64_t s0, s1;
    int K, min_hits, distinct, include_last;
    int32_t* role; int32_t* hits; uint8_t* flag;
} apply_job;

static void* apply_worker(void* p) {
    apply_job* j = (apply_job*)p;
    orc_map* scratch = orc_map_new(0);
    for (uint64_t s = j->s0; s < j->s1; s++) {
        uint64_t a = j->offsets[s], b = j->offsets[s + 1];
        uint8_t f;
        apply_one(j->db, scratch, j->residues + a, b - a, j->K, j->min_hits, j->distinct,
                  j->include_last, &j->role[s], &j->hits[s], &f);
        if (j->flag) j->flag[s] = f;
    }
    orc_map_free(scratch);
    return NULL;
}

/* Annotate a CSR batch.  n_threads = 1 is the reference's shape (the peg loop is a plain
 * `for`, ApplyKmerProcessor.java:118-122); n_threads > 1 partitions the sequences into
 * residue-balanced contiguous ranges, one per thread, as parallelStream over genomes does
 * in HashAnnotationProcessor.java:208. */
int orc_apply(const orc_map* db, const uint8_t* residues, const uint64_t* offsets, uint64_t N,
              int K, int min_hits, int distinct, int include_last, int n_threads,
              int32_t* role, int32_t* hits, uint8_t* flag) {
    if (K < 1 || K > 256 || min_hits < 1) return -1;
    if (n_threads < 1) n_threads = 1;
    if ((uint64_t)n_threads > N) n_threads = N ? (int)N : 1;
    apply_job* jobs = (apply_job*)calloc((size_t)n_threads, sizeof(apply_job));
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    uint64_t base = N ? offsets[0] : 0, total = N ? offsets[N] - base : 0;
    uint64_t s = 0;
    for (int t = 0; t < n_threads; t++) {
        uint64_t target = base + (total * (uint64_t)(t + 1)) / (uint64_t)n_threads;
        uint64_t e = s;
        if (t == n_threads - 1) e = N;
        else while (e < N && offsets[e + 1] <= target) e++;
        jobs[t] = (apply_job){db, residues, offsets, s, e, K, min_hits, distinct, include_last,
                              role, hits, flag};
        s = e;
    }
    if (n_threads == 1) apply_worker(&jobs[0]);
    else {
        for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, apply_worker, &jobs[t]);
        for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    }
    free(jobs); free(th);
    return 0;
}

/* ------------------------------------------------------------------------------------
 * RoleCounter: proteins/kmers/RoleCounter.java:29-56.  Pinned by RoleTests.java:15-36.
 * ---------------------------------------------------------------------------------- */
typedef struct { int32_t role_id; int32_t good; int32_t bad; } orc_role_counter;

void orc_role_counter_init(orc_role_counter* c, int32_t role_id) { /* :29-33 */
    c->role_id = role_id; c->good = 0; c->bad = 0;
}
int orc_role_counter_count(orc_role_counter* c, int32_t role_hit) { /* :42-49 */
    int ret = c->role_id == role_hit;
    if (ret) c->good++; else c->bad++;
    return ret;
}
int orc_role_counter_is_good(const orc_role_counter* c) { return c->bad == 0; } /* :54-56 */

/* ------------------------------------------------------------------------------------
 * build: BuildKmerProcessor.java:138-223.  Input is the peg stream already classified by
 * the caller (Feature.getUsefulRoles ∩ goodRoles, :158): n_roles[i] = number of good roles
 * of peg i (0, 1, >= 2) and peg_role[i] = the role id when n_roles[i] == 1.
 *   kmerMap = new HashMap<>(goodRoles.size() * 700000)            :140  (int arithmetic)
 *   pass 1  one good role -> computeIfAbsent + count              :165-173
 *           zero good roles -> buffered to the temp FASTA         :159-164
 *           two or more -> ignored                                :165
 *   prune   remove entries with badCount > 0                      :183-190
 *   pass 2  every k-mer of a buffered protein is removed          :196-208
 *   emit    kmer TAB roleId in HashMap iteration order            :212-216
 * node.value = RoleCounter.roleId, node.aux = badCount (goodCount is only logged).
 * Returns the map (caller dumps it with orc_map_dump) or NULL on failure / when the Java
 * capacity expression is negative (HashMap throws IllegalArgumentException).
 * ---------------------------------------------------------------------------------- */
orc_map* orc_build(const uint8_t* residues, const uint64_t* offsets, uint64_t N,
                   const int32_t* n_roles, const int32_t* peg_role, int K, int32_t n_good_roles,
                   int distinct, int include_last, uint64_t* stats /* [4] or NULL */) {
    int32_t cap32 = (int32_t)((uint32_t)n_good_roles * 700000u); /* Java int multiply wraps */
    if (cap32 < 0) return NULL;
    orc_map* m = orc_map_new(cap32);
    orc_map* scratch = orc_map_new(0);
    if (!m || !scratch) return NULL;
    uint64_t buffered = 0, nonunique = 0, deleted2 = 0;
    /* pass 1 */
    for (uint64_t i = 0; i < N; i++) {
        if (n_roles[i] != 1) { if (n_roles[i] == 0) buffered++; continue; }
        uint64_t a = offsets[i], b = offsets[i + 1];
        protein_kmers(scratch, residues + a, b - a, K, distinct, include_last); /* :167 */
        if (!scratch->table) continue;
        for (uint32_t j = 0; j < scratch->cap; j++)
            for (int32_t e = scratch->table[j]; e >= 0; e = scratch->nodes[e].next) {
                int created;
                int32_t ni = jput_node(m, (const uint8_t*)scratch->pool + scratch->nodes[e].koff,
                                       (uint32_t)K, &created); /* computeIfAbsent :170 */
                if (ni < 0) { orc_map_free(m); orc_map_free(scratch); return NULL; }
                if (created) { m->nodes[ni].value = peg_role[i]; m->nodes[ni].aux = 0; }
                if (m->nodes[ni].value != peg_role[i]) m->nodes[ni].aux++; /* counter.count :171 */
            }
    }
    /* prune: iterator.remove() of entries with badCount != 0 (:183-190) */
    if (m->table) {
        for (uint32_t j = 0; j < m->cap; j++) {
            int32_t e = m->table[j], prev = -1;
            while (e >= 0) {
                int32_t nx = m->nodes[e].next;
                if (m->nodes[e].aux != 0) {
                    if (prev < 0) m->table[j] = nx; else m->nodes[prev].next = nx;
                    m->size--; nonunique++;
                } else prev = e;
                e = nx;
            }
        }
    }
    /* pass 2: the buffered proteins, in the order they were written (:196-208) */
    for (uint64_t i = 0; i < N; i++) {
        if (n_roles[i] != 0) continue;
        uint64_t a = offsets[i], b = offsets[i + 1];
        protein_kmers(scratch, residues + a, b - a, K, distinct, include_last); /* :199 */
        if (!scratch->table) continue;
        for (uint32_t j = 0; j < scratch->cap; j++)
            for (int32_t e = scratch->table[j]; e >= 0; e = scratch->nodes[e].next)
                if (orc_map_remove(m, (const uint8_t*)scratch->pool + scratch->nodes[e].koff,
                                   (uint32_t)K)) deleted2++; /* :201-202 */
    }
    orc_map_free(scratch);
    if (stats) { stats[0] = buffered; stats[1] = nonunique; stats[2] = deleted2; stats[3] = m->size; }
    return m;
}

/* ------------------------------------------------------------------------------------
 * KmerReference.countPegKmers window loop (proteins/kmers/KmerReference.java:124-147),
 * restated only to pin conventions against AppTest.java:145-161: windows i = 0 .. L-K-1
 * (`i < end`, end = L - K: the LAST window is dropped), k-mers containing 'X' skipped,
 * 1-based left position i+1.  Writes the 1-based left positions of the emitted windows;
 * returns their number.
 * ---------------------------------------------------------------------------------- */
uint64_t orc_count_peg_kmers_positions(const uint8_t* prot, uint64_t L, int K, uint32_t* left_out) {
    uint64_t n = 0;
    int64_t end = (int64_t)L - K; /* :134 */
    for (int64_t i = 0; i < end; i++) { /* :136 */
        int has_x = 0;
        for (int j = 0; j < K; j++) if (prot[i + j] == 'X') has_x = 1; /* :139 */
        if (!has_x) { if (left_out) left_out[n] = (uint32_t)(i + 1); n++; } /* :140 */
    }
    return n;
}

/* ------------------------------------------------------------------------------------
 * Pairwise k-mer distance: genome/compare/GeneCopyProcessor.java:137-142.
 *   ProteinKmers kmers = new ProteinKmers(target protein)          (:137)
 *   ProteinKmers f2Kmers = new ProteinKmers(candidate protein)     (:141)
 *   double f2Dist = kmers.distance(f2Kmers)                        (:142)
 * ProteinKmers is not in the repository; recalled (docs/SEMANTICS.md): similarity = number of
 * k-mers of the other set contained in this one; distance = 1.0 when similarity is 0, else
 * 1.0 - similarity / ((size() + other.size()) - similarity), all in double.
 * One pair per call; sizes and similarity are returned for the integer parity checks.
 * ---------------------------------------------------------------------------------- */
double orc_kmer_distance(const uint8_t* a, uint64_t la, const uint8_t* b, uint64_t lb, int K,
                         int32_t* size_a, int32_t* size_b, int32_t* common) {
    orc_map* sa = orc_map_new(16);
    orc_map* sb = orc_map_new(16);
    protein_kmers(sa, a, la, K, 1, 1);
    protein_kmers(sb, b, lb, K, 1, 1);
    /* similarity: iterate the other set, count members of this one */
    int32_t sim = 0;
    for (uint32_t i = 0; i < sb->n_nodes; i++) {
        const jnode* nd = &sb->nodes[i];
        if (jfind(sa, (const uint8_t*)sb->pool + nd->koff, nd->klen, nd->hash) >= 0) sim++;
    }
    if (size_a) *size_a = (int32_t)sa->size;
    if (size_b) *size_b = (int32_t)sb->size;
    if (common) *common = sim;
    double ret = 1.0;
    double similarity = (double)sim;
    if (similarity > 0) {
        double uni = (double)((int32_t)sa->size + (int32_t)sb->size) - similarity;
        ret = 1.0 - similarity / uni;
    }
    orc_map_free(sa);
    orc_map_free(sb);
    return ret;
}

/* batch form over a CSR batch: pair m = (qa[m], qb[m]) */
void orc_kmer_distance_pairs(const uint8_t* residues, const uint64_t* offsets, const uint32_t* qa,
                             const uint32_t* qb, uint64_t n_pairs, int K, int32_t* size_a,
                             int32_t* size_b, int32_t* common, double* dist) {
    for (uint64_t m = 0; m < n_pairs; m++) {
        const uint64_t a0 = offsets[qa[m]], a1 = offsets[qa[m] + 1], b0 = offsets[qb[m]], b1 = offsets[qb[m] + 1];
        dist[m] = orc_kmer_distance(residues + a0, a1 - a0, residues + b0, b1 - b0, K,
                                    size_a ? size_a + m : NULL, size_b ? size_b + m : NULL, common ? common + m : NULL);
    }
}

/* window positions P = sum max(0, L_i - K + 1): the metric's unit of work */
uint64_t orc_count_probes(const uint64_t* offsets, uint64_t N, int K) {
    uint64_t p = 0;
    for (uint64_t i = 0; i < N; i++) {
        uint64_t L = offsets[i + 1] - offsets[i];
        if (L >= (uint64_t)K) p += L - (uint64_t)K + 1;
    }
    return p;
}
