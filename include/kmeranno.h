/*
 * kmeranno.h — C ABI of libkmeranno.so, the B200 (sm_100a) engine for the k-mer
 * annotation hot path of SEEDtk/kmers.anno (`apply` command).
 *
 * Every entry point replaces a region of the reference's Java code.  Paths are
 * relative to /root/reference/src/main/java/org/theseed/ :
 *
 *   ka_db_load       <- proteins/kmers/anno/ApplyKmerProcessor.java:99-110
 *                       (kmerRoleMap = HashMap<String,String>; put() = last line wins;
 *                        K taken from the k-mer text, :108)
 *   ka_build         <- proteins/kmers/anno/BuildKmerProcessor.java:138-223 + kmers/RoleCounter.java
 *   ka_annotate      <- proteins/kmers/anno/ApplyKmerProcessor.java:122-148
 *                       (new ProteinKmers(prot) :123, kmerRoleMap.get :130, tally
 *                        :131-144, thresholded call :146-147)
 *   out_role/out_hits feed reports/ApplyKmerReporter.java:75 recordFeature(feat, role, count)
 *
 * The library has no CPU fallback: without a usable CUDA device ka_create fails with
 * KA_ERR_NO_DEVICE.  Nothing here prints to stdout (stdout belongs to the report,
 * ApplyKmerProcessor.java:94); diagnostics go to ka_last_error().
 *
 * Ownership: the caller owns every host buffer and may free it when the call returns.
 * The engine owns all device memory until ka_destroy().  No callbacks.
 * Threading: an engine serialises its entry points internally (one mutex); use one
 * engine per thread for concurrent callers (HashAnnotationProcessor.java:208 style).
 * ka_pack_residues only reads the alphabet and runs concurrently.
 */
#ifndef KMERANNO_H
#define KMERANNO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KA_ABI_VERSION 2

typedef struct ka_engine ka_engine;
typedef struct ka_batch ka_batch;

/* error codes (0 = OK, negative = failure; text via ka_last_error) */
enum {
    KA_OK = 0,
    KA_ERR_INVALID = -1,   /* bad argument (NULL, K out of range, min_hits < 1 ...)       */
    KA_ERR_NO_DEVICE = -2, /* no usable CUDA device / bad device id                       */
    KA_ERR_CUDA = -3,      /* CUDA runtime failure                                        */
    KA_ERR_ALPHABET = -4,  /* DB uses more than 31 distinct residue bytes: not packable   */
    KA_ERR_K = -5,         /* K outside 1..12 (5 bits x K must fit 64 bits)               */
    KA_ERR_NO_DB = -6,     /* annotate before a successful ka_db_load                     */
    KA_ERR_OOM = -7,       /* host or device allocation failed                            */
    KA_ERR_ROLE = -8,      /* negative role id in the DB (-1 is the "no call" value)      */
    KA_ERR_OFFSETS = -9,   /* offsets not monotone / not starting at the batch base       */
    KA_ERR_TOO_BIG = -10   /* table or batch beyond this build's limits                   */
};

/* per-sequence outcome, out_flag[] */
enum {
    KA_FLAG_NONE = 0,      /* no k-mer of the sequence is in the DB                       */
    KA_FLAG_CALLED = 1,    /* unanimous role and hits >= min_hits (recordFeature fires)   */
    KA_FLAG_AMBIGUOUS = 2, /* k-mers hit two or more roles (badPeg, :140-143)             */
    KA_FLAG_BELOW_MIN = 3  /* unanimous role but hits < min_hits (:146)                   */
};

/* ---- engine lifetime ------------------------------------------------------------- */

/* Create an engine on the listed CUDA devices (device_ids == NULL: device 0 only).
 * On failure *out is NULL and ka_last_error(NULL) describes why. */
int ka_create(const int* device_ids, int n_devices, ka_engine** out);
void ka_destroy(ka_engine* e);

/* Last error text of the engine (e == NULL: of the calling thread's last ka_create). */
const char* ka_last_error(const ka_engine* e);

/* Tunables.  Unknown name or a value out of range -> KA_ERR_INVALID and NOTHING changes.
 * Table options take effect at the next ka_db_load (the loaded table keeps the geometry it was
 * built with); tiling options at the next annotate call (a resident batch uploaded under other
 * tiling options is rejected by ka_annotate_resident: upload it again).
 *   "load_factor"   table load factor in (0,0.9]; default: 0.68 for the line table, 0.4 for the sector classes
 *   "slot_bits"     force the table layout: 32 / 64 / 128 = sector classes with slots of that width; 16 = the
 *                   128-byte-line table (16-bit tags + 16-bit roles, spill inside the line, presence filter in
 *                   L2; needs a replicated table, role ids < 65536 and 2 <= K <= 10 with key halves of at
 *                   most 25 bits); 0 (default) = the line table when those conditions hold and its layout fits
 *                   the key space without padding, else the narrowest sector class that holds the DB
 *   "filter"        line table only: 1 (default) = probe the L2-resident presence filter first, 0 = always read
 *                   the table (measurement knob)
 *   "ingest_via"    single-device engines: CUDA device id whose PCIe path carries the H2D copies of ka_annotate /
 *                   ka_annotate_packed (the chunk lands in a staging buffer there and crosses NVLink to the engine's
 *                   GPU); -1 (default) = the engine's own path.  For boxes where some GPUs share a slower host path
 *                   (microbench/pcie_concurrent.py; bench.py probes and pairs the ranks at N = 8)
 *   "resident_packed" 1 (default) = ka_batch_upload keeps the batch as the 5-bit stream where the tile kernels can stage
 *                   it (what ka_annotate_packed ships), 0 = as residue bytes (measurement knob)
 *   "table_mode"    0 = table replicated on every device (default); 1 = table sharded by sector range
 *                   over the engine's 2/4/8 devices, probes load remote sectors through NVLink
 *                   peer memory inside the probe kernel (for tables beyond one GPU); 2 = same sharding,
 *                   but the k-mer keys are ROUTED: NCCL send/recv all-to-all of 8-byte keys to the
 *                   owning GPU, local probe there, 8-byte answers back in request order, rounds pipelined;
 *                   3 = the same routing with the exchanges FUSED into the kernels: the bucket-scatter kernel
 *                   stores the keys straight into the owner's receive buffer and the owner's lookup kernel
 *                   stores the answers straight into the requester's buffer (NVLink peer stores, no NCCL)
 *   "wide"          1 = use the wide-table kernels (64-bit sector indices, the mixed key as de-dup
 *                   token) on any sector-class table; 0 (default) = only beyond 2^32 - 16 slots
 *   "tile_span"     residues of sequence starts per CTA tile, default 1536
 *   "long_seq"      sequences longer than this get a tile of their own (second tile launch), default 2048
 *   "mid_seq"       sequences longer than this use the global-scratch long-sequence kernel, default 8192
 *   "chunk_residues" residues per pipelined H2D chunk; 0 (default) = 64 Mi, 32 Mi on a routed table (table_mode 2/3)
 *   "l2_persist"    sector classes: 1 = L2 persisting access-policy window on the table (default 1)
 */
int ka_set_option(ka_engine* e, const char* name, double value);

/* ---- k-mer database (ApplyKmerProcessor.java:99-110) ------------------------------ */

/* Load n k-mers of K residue bytes each (kmers = n*K bytes, no separators) with their
 * role ids (>= 0; the Java side interns the role strings).  Line order matters only for
 * duplicates: the LAST occurrence of a k-mer wins, as HashMap.put does (:106).
 * Builds the open-addressed table on every device of the engine and replaces any
 * previous DB.  Residue bytes are compared exactly (case-sensitive, no filtering): the
 * distinct bytes of the DB (at most 31) become the 5-bit alphabet. */
int ka_db_load(ka_engine* e, const uint8_t* kmers, const int32_t* role_ids, uint64_t n, int K);

typedef struct ka_db_info {
    int32_t K;               /* k-mer length in residues                                 */
    int32_t n_symbols;       /* distinct residue bytes in the DB                         */
    uint64_t n_lines;        /* k-mer lines given to ka_db_load                          */
    uint64_t n_keys;         /* distinct k-mers stored                                   */
    uint64_t n_buckets;      /* 32-byte sectors (8, 4 or 2 slots each)                   */
    uint64_t table_bytes;    /* device bytes of one table replica                        */
    uint32_t max_probe;      /* longest sector chain seen while building                 */
    uint32_t slot_bits;      /* layout chosen for this DB: 16 (line table), 32, 64 or 128 */
    uint64_t filter_bytes;   /* line table: device bytes of the presence filter (else 0) */
    uint64_t n_spilled;      /* line table: keys stored outside their home sector        */
    uint64_t n_overflow;     /* line table: keys stored in the overflow table            */
} ka_db_info;
int ka_db_get_info(ka_engine* e, ka_db_info* out);

/* ---- annotate (ApplyKmerProcessor.java:122-148) ----------------------------------- */

/* Annotate N sequences given as a CSR batch in HOST memory: residues[offsets[i] ..
 * offsets[i+1]) is sequence i (offsets[0] may be non-zero; residues is indexed from 0).
 * For each sequence: the set of DISTINCT K-windows is probed; if every hit names the
 * same role, hits = number of distinct hitting k-mers.
 *   out_role[i] = role id when called, else -1
 *   out_hits[i] = distinct hitting k-mers when unanimous (called or below min), else 0
 *   out_flag[i] = KA_FLAG_*            (may be NULL)
 * Sequences are sharded over the engine's devices; copies are inside the call. */
int ka_annotate(ka_engine* e, const uint8_t* residues, const uint64_t* offsets, uint64_t N,
                int32_t min_hits, int32_t* out_role, int32_t* out_hits, uint8_t* out_flag);

/* The same call on the PACKED form of the batch, 0.625 bytes per residue over PCIe instead of 1 and
 * 32-bit offsets: residue r of the batch occupies bits [5r, 5r+5) of the little-endian byte stream
 * `codes`; its value is the digit of the residue byte in the DB alphabet (ka_db_get_alphabet: the
 * distinct bytes of the DB numbered 0.. in byte order) or 31 for a byte that is not in the DB — such a
 * byte can never be part of a match (String equality, ApplyKmerProcessor.java:130).  offsets[i] is the
 * residue index of the start of sequence i (N + 1 entries, 32 bits: at most 2^32 - 1 residues per
 * call).  The stream is what the host writes while it touches the residues anyway (FASTA / GTO
 * parser, ka_pack_residues); results are identical to ka_annotate on the unpacked batch. */
int ka_annotate_packed(ka_engine* e, const uint8_t* codes, const uint32_t* offsets, uint64_t N,
                       int32_t min_hits, int32_t* out_role, int32_t* out_hits, uint8_t* out_flag);

/* code_of_byte[256]: the 5-bit code of every byte value for the loaded DB (digit 0..n_symbols-1, or 31). */
int ka_db_get_alphabet(ka_engine* e, uint8_t* code_of_byte);

/* Host helper: write the codes of residues[0..n) as residues first_index .. first_index+n-1 of the
 * stream `codes` (first_index must be a multiple of 8 = a byte boundary of the stream, so that
 * threads can pack disjoint ranges; only whole bytes of the range are written).  Thread-safe. */
int ka_pack_residues(ka_engine* e, const uint8_t* residues, uint64_t n, uint64_t first_index, uint8_t* codes);

/* Device-resident variant, used to time the kernels with inputs already in HBM.
 * ka_batch_upload copies a CSR batch to device `dev_index` (index into the engine's device
 * list); ka_annotate_resident runs the kernels only (results stay on the device);
 * ka_batch_download copies the results of the last run back. */
int ka_batch_upload(ka_engine* e, int dev_index, const uint8_t* residues, const uint64_t* offsets,
                    uint64_t N, ka_batch** out);
int ka_annotate_resident(ka_engine* e, ka_batch* b, int32_t min_hits);
int ka_batch_download(ka_engine* e, ka_batch* b, int32_t* out_role, int32_t* out_hits,
                      uint8_t* out_flag);
void ka_batch_free(ka_engine* e, ka_batch* b);

/* ---- build the discriminating k-mer DB (BuildKmerProcessor.java:138-223) --------------- */

/* GPU restatement of the `build` command's core.  Input: every peg of the training genomes as
 * a CSR batch, already classified by the caller (Feature.getUsefulRoles ∩ goodRoles, :158):
 *   n_roles[i]  number of good roles of peg i (0, 1, or >= 2)
 *   peg_role[i] the role id (>= 0) when n_roles[i] == 1, ignored otherwise
 * Output = the k-mers that occur in at least one single-role peg (:165-173), whose
 * single-role pegs all carry the same role (RoleCounter.isGood, :183-190) and that occur
 * in no zero-role peg (:196-208); pegs with two or more good roles are ignored (:165).
 * At most `cap` k-mers are written (out_kmers: cap*K bytes, out_roles: cap ints) in
 * unspecified order — the reference's order is HashMap iteration order (:212-216).
 * *n_out receives the number of k-mers found; if it exceeds cap the call returns
 * KA_ERR_TOO_BIG and nothing else is guaranteed.  load_as_db != 0 also installs the result
 * as the engine's k-mer database (as ka_db_load would), skipping the kmerdb.tbl round trip.
 * Runs on the engine's first device. */
int ka_build(ka_engine* e, const uint8_t* residues, const uint64_t* offsets, uint64_t N,
             const int32_t* n_roles, const int32_t* peg_role, int K, uint64_t cap,
             uint8_t* out_kmers, int32_t* out_roles, uint64_t* n_out, int load_as_db);

/* ---- pairwise k-mer distance (genome/compare/GeneCopyProcessor.java:137-142) --------- */

/* For every query q (sequence query_seq[q] of the CSR batch) and each of its candidates
 * cand_seq[group_offsets[q] .. group_offsets[q+1]): the ProteinKmers comparison of :137-142,
 *   A = distinct K-windows of the query (new ProteinKmers(...), :137), B = those of the candidate (:141),
 *   out_common[m]   = |A ∩ B|                                 (ProteinKmers.similarity)
 *   out_distance[m] = 1.0 if nothing is shared, else 1.0 - |A ∩ B| / (|A| + |B| - |A ∩ B|)   (:142)
 * and out_set_size[i] = number of distinct K-windows of EVERY sequence i of the batch (may be NULL).
 * K is the -K option of `genes` (:70-71, ProteinKmers.setKmerSize :96), 1..12 here; the alphabet is
 * the distinct bytes of the batch (at most 31).  The caller keeps the reference's selection loop
 * (closest candidate with distance <= maxDist, later candidates win ties, :139-146).
 * Queries are split over the engine's devices; no k-mer database needs to be loaded. */
int ka_kmer_distance(ka_engine* e, const uint8_t* residues, const uint64_t* offsets, uint64_t N, int K,
                     const uint32_t* query_seq, const uint64_t* group_offsets, uint64_t Q,
                     const uint32_t* cand_seq, int32_t* out_set_size, int32_t* out_common,
                     double* out_distance);

/* ---- pinned host memory for zero-staging transfers -------------------------------- */
void* ka_host_alloc(size_t bytes);
void ka_host_free(void* p);

/* ---- measurement ------------------------------------------------------------------ */
typedef struct ka_stats {
    uint64_t sequences;       /* sequences annotated by the last annotate call             */
    uint64_t residues;        /* residues in them                                          */
    uint64_t probes;          /* window positions: sum max(0, L_i - K + 1)                 */
    uint64_t kernel_launches; /* kernels launched by the last annotate call                */
    uint64_t h2d_bytes;       /* host->device bytes of the last annotate call              */
    uint64_t d2h_bytes;       /* device->host bytes of the last annotate call              */
    double kernel_ms;         /* device time of the annotate kernels (CUDA events; max over devices) */
    double tile_kernel_ms;    /* device time of the tile kernel alone (sum over chunks, max over devices) */
    double wall_ms;           /* host wall time of the last annotate call                  */
} ka_stats;
int ka_get_stats(ka_engine* e, ka_stats* out);

/* Random-probe roofline microbenchmark: independent uniformly random `slot_bytes`-wide
 * (16 or 32) loads over a scratch buffer of `table_bytes`, `n_probes` loads per launch,
 * best of `reps` launches.  Returns probes per second on device dev_index. */
int ka_probe_roofline(ka_engine* e, int dev_index, uint64_t table_bytes, uint64_t n_probes,
                      int slot_bytes, int reps, double* probes_per_s);

/* Synthetic k-mer database for the oversized-table configuration (table larger than one GPU,
 * BASELINE.json configs[4]): n lines generated ON THE DEVICES, so the host never holds them.
 * Line i (0-based) is the K-mer whose j-th residue is "ACDEFGHIKLMNPQRSTVWY"[((x >> 5j) & 31) % 20]
 * with x = mix64(seed + i * 0x9E3779B97F4A7C15) (mix64: x ^= x>>32; x *= 0xD6E8FEB86659FD93; twice;
 * x ^= x>>32), and its role id is i % n_roles.  Duplicate k-mers keep the last line, as in
 * ka_db_load.  Measurement / test entry point: it replaces no reference code. */
int ka_db_load_synthetic(ka_engine* e, uint64_t n, int K, int32_t n_roles, uint64_t seed);

int ka_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* KMERANNO_H */
