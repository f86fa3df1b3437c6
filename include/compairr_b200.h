/*
 * compairr_b200.h — C ABI of the B200-native repertoire-overlap engine.
 *
 * This is the drop-in boundary for the hot path of `compairr -m / -x` (CompAIRR 1.13.0):
 * everything the reference does between "both sequence sets are in memory" and "dump the
 * similarity matrix" inside overlap() (reference src/overlap.cc:838-942).  The reference has
 * no FFI; the seam is the set of calls overlap() makes into zobrist.cc / hashtable.cc /
 * bloompat.cc / variants.cc and the sim_thread() fan-out.  Each entry point below names the
 * reference interface it replaces.
 *
 * Conventions
 *   - plain C, no CUDA/torch types; all pointers are HOST pointers unless the name says "device"
 *   - every function returns 0 on success, a negative cb_status on failure; the message is
 *     available from cb_last_error().  The library never calls exit() (the reference's
 *     fatal(), src/util.cc:84-88, is the CLI's job).
 *   - one context drives ONE GPU and must be used from one host thread at a time.  Multi-GPU is
 *     one context (one process, or one host thread) per GPU joined in an NCCL communicator
 *     (cb_comm_*): every context gets the whole of set B — each rank uploads 1/world of it and the
 *     ranks all-gather over NVLink (cb_set_b_sharded) — and a shard of set A; the partial
 *     matrices are summed by cb_allreduce_matrix.
 *   - there is no CPU fallback: if no CUDA device is usable, cb_create() fails.
 */
#ifndef COMPAIRR_B200_H
#define COMPAIRR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB_ABI_VERSION 3

typedef enum cb_status {
  CB_OK = 0,
  CB_ERR_INVALID = -1,   /* bad argument / option combination                          */
  CB_ERR_CUDA = -2,      /* CUDA runtime error (message has the cudaError string)       */
  CB_ERR_NOMEM = -3,     /* host or device allocation failed                            */
  CB_ERR_STATE = -4,     /* call sequence violated (e.g. run before set B was built)    */
  CB_ERR_LIMIT = -5      /* input exceeds an engine limit (sequence too long, ...)      */
} cb_status;

/* Score summands; numbering identical to the reference enum (src/compairr.h:124-133). */
typedef enum cb_score {
  CB_SCORE_PRODUCT = 0,
  CB_SCORE_RATIO = 1,
  CB_SCORE_MIN = 2,
  CB_SCORE_MAX = 3,
  CB_SCORE_MEAN = 4,
  CB_SCORE_MH = 5,       /* summand = product; index computed by the host writer (overlap.cc:548-560) */
  CB_SCORE_JACCARD = 6   /* summand = min;     index computed by the host writer (overlap.cc:562-570) */
} cb_score;

typedef enum cb_mode {
  CB_MODE_MATRIX = 0,    /* -m: rows are set-A repertoires   (overlap.cc:222) */
  CB_MODE_EXISTENCE = 1  /* -x: rows are set-A sequences     (overlap.cc:226) */
} cb_mode;

/*
 * Run options.  These are the reference's global opt_* flags that the hot path reads
 * (src/compairr.h:139-162; read in zobrist.cc:83, overlap.cc:97,146,195,218,232,368).
 */
typedef struct cb_config {
  int32_t abi_version;     /* must be CB_ABI_VERSION                                              */
  int32_t device;          /* CUDA device ordinal                                                 */
  int32_t alphabet_size;   /* 20 (amino acids) or 4 (-n nucleotides), compairr.cc:691-694         */
  int32_t differences;     /* -d, >= 0; d <= 2 uses the hash path, d >= 3 the brute-force path    */
  int32_t indels;          /* -i, only legal with differences == 1 (compairr.cc:639-640)          */
  int32_t ignore_genes;    /* -g                                                                  */
  int32_t ignore_counts;   /* -f: every summand is 1 (overlap.cc:146-147)                         */
  int32_t score;           /* cb_score                                                            */
  int32_t mode;            /* cb_mode                                                             */
  int32_t no_matrix;       /* --no-matrix: do not keep a matrix (overlap.cc:218,876)              */
  int32_t want_pairs;      /* -p: collect (seed, hit) pairs (overlap.cc:232-245)                  */
  uint32_t n_reps_a;       /* repertoires in set A (rows in matrix mode); 1 in existence mode     */
  uint64_t seed;           /* PRNG seed for the Zobrist table; results do not depend on it        */
  /* tuning; 0 selects the default.  Results do not depend on these either. */
  uint32_t bloom_bits_per_key_x16;  /* bits per set-B sequence in EACH of the four class filters, */
                                    /* fixed point 1/16 bit (default 16 bits)                     */
  uint32_t table_load_pct;          /* max hash-table load in percent (default 50)               */
  uint64_t pairs_capacity;          /* device pair-buffer capacity per launch, in pairs          */
  uint32_t flags;                   /* CB_FLAG_*                                                  */
  uint32_t queue_capacity;          /* d = 1, 2: entries of the candidate queue between the      */
                                    /* enumeration and the table kernel (16 B each); 0 = sized    */
                                    /* from the run, at most 2^26.  A queue that overflows only   */
                                    /* costs time: the chunk of seeds is redone in smaller pieces */
} cb_config;

#define CB_FLAG_NO_SMEM_TILE 1u   /* small matrices too: accumulate straight into the global     */
                                  /* matrix, no shared-memory tile (A/B testing)                 */
#define CB_FLAG_NO_BLOOM 2u       /* probe the hash table for every variant (A/B testing)        */
#define CB_FLAG_NO_TENSOR 4u      /* d >= 3: CUDA-core kernel only, no tcgen05 GEMM (A/B testing) */
#define CB_FLAG_NO_PARTITION 8u   /* table build in input order, no radix sort by home slot (A/B testing) */
#define CB_FLAG_FILTERS_IN_BUILD 32u /* large builds: filter bits set by the table-build kernel instead of */
                                    /* the L2-blocked filter passes (A/B testing)                          */
#define CB_FLAG_NO_TILED_BUILD 64u /* large builds: the sorted sweep over a cleared table instead of the */
                                   /* tiled shared-memory build (A/B testing)                            */
#define CB_FLAG_GENERIC_KERNEL 16u /* d = 1, 2: every seed through the any-length enumeration kernel  */
                                   /* (otherwise only seeds longer than 94 residues; A/B testing)     */

/*
 * One sequence set in structure-of-arrays form — what db_read() (src/db.cc:708-901) leaves
 * in memory, minus the strings.  Replaces the getters db_getsequence/db_getsequencelen/
 * db_get_v_gene/db_get_j_gene/db_get_repertoire_id_no/db_get_count (src/db.cc:964-997).
 *
 *   residues   one byte per residue, codes 0..alphabet_size-1 (db.cc:33-71), one arena
 *   offsets    n+1 entries; sequence i is residues[offsets[i] .. offsets[i+1])
 *   v_gene/j_gene  gene numbers, comparable between set A and set B (db.cc:119-125); may be
 *              NULL when ignore_genes is set
 *   rep        repertoire number within this set, 0..n_reps-1, any order
 *   count      duplicate_count (>= 1); may be NULL when ignore_counts is set and the score is
 *              not MH/Jaccard-relevant for the caller
 *   index_base added to sequence indices reported in pairs and to existence-mode rows, so a
 *              shard of a larger set reports global indices
 */
typedef struct cb_set {
  uint64_t n;
  const uint8_t *residues;
  const uint64_t *offsets;
  const uint32_t *v_gene;
  const uint32_t *j_gene;
  const uint32_t *rep;
  const uint64_t *count;
  uint32_t n_reps;
  uint32_t longest;        /* longest sequence in the set, 0 = let the engine scan offsets */
  uint64_t index_base;
} cb_set;

/*
 * The same set with caller-chosen column widths, for hosts that keep narrower types than the
 * reference's seqinfo_s (fewer bytes over PCIe): every per-sequence column is a pointer plus an
 * element width in bytes (1, 2, 4 or 8; data == NULL means "absent").  Sequence boundaries are
 * given EITHER as n+1 offsets (width 8) OR as n lengths (width 1, 2 or 4); with lengths the
 * sequences are contiguous in `residues` starting at byte 0.
 */
typedef struct cb_col {
  const void *data;
  uint32_t width;
  uint32_t reserved;
} cb_col;

typedef struct cb_set_cols {
  uint64_t n;
  const uint8_t *residues;
  cb_col offsets;   /* n+1 entries, width 8, or absent */
  cb_col lengths;   /* n entries, width 1/2/4, or absent */
  cb_col v_gene;
  cb_col j_gene;
  cb_col rep;
  cb_col count;
  uint32_t n_reps;
  uint32_t longest;
  uint64_t index_base;
} cb_set_cols;

/* A matching pair: reference struct pair_s (src/overlap.cc:55-58). */
typedef struct cb_pair {
  uint64_t a;   /* sequence index in set A */
  uint64_t b;   /* sequence index in set B */
} cb_pair;

/* Work and timing counters of the most recent cb_build_b / cb_run_* call. */
typedef struct cb_stats {
  uint64_t seeds;           /* set-A sequences processed                                          */
  uint64_t probes;          /* variant hashes enumerated (= variants the reference generates)     */
  uint64_t bloom_pass;      /* probes that passed the filter stage (reach the table stage)       */
  uint64_t matches;         /* verified (seed, hit) matches (reference all_matches, overlap.cc:230) */
  uint64_t pairs;           /* pairs stored for cb_drain_pairs                                    */
  uint64_t table_slots;     /* hash-table slots                                                   */
  uint64_t bloom_bytes;     /* bytes of ONE of the four class filters                             */
  uint64_t bloom2_bytes;    /* bytes of the other three together                                  */
  float ms_hash_b;          /* device time, CUDA events on the engine's stream                    */
  float ms_build_b;
  float ms_dups_b;
  float ms_hash_a;
  float ms_probe;           /* enumerate + Bloom + probe + verify + accumulate kernel(s)          */
  float ms_total_run;       /* whole cb_run_* device span                                         */
  uint32_t kernel_launches; /* kernels launched by the call                                       */
  float ms_gather_b;        /* cb_set_b_sharded: the NVLink all-gather of records, hashes, residues */
} cb_stats;

typedef struct cb_ctx cb_ctx;
typedef struct cb_dset cb_dset;   /* a sequence set resident in device memory */

/* ---- lifecycle --------------------------------------------------------------------------- */

/* Library-wide message for failures that happen before a context exists. */
const char *cb_global_error(void);
int cb_abi_version(void);
/* Number of usable CUDA devices (0 if none / driver missing). */
int cb_device_count(void);

/* Replaces the option globals + zobrist_init() (src/zobrist.cc:28-67, called overlap.cc:840). */
int cb_create(const cb_config *cfg, cb_ctx **out);
void cb_destroy(cb_ctx *ctx);
const char *cb_last_error(const cb_ctx *ctx);

/* Use a caller-owned CUDA stream (cudaStream_t passed as void*) for all work of this context;
   NULL restores the context's own stream. */
int cb_set_stream(cb_ctx *ctx, void *cuda_stream);

/* ---- device-resident sets ------------------------------------------------------------------ */

/* Copy a set to the GPU and hash it: replaces db_hash() (src/db.cc:903-916) → zobrist_hash()
   (src/zobrist.cc:74-88).  The host arrays may be freed when the call returns. */
int cb_upload(cb_ctx *ctx, const cb_set *set, cb_dset **out);
int cb_upload_cols(cb_ctx *ctx, const cb_set_cols *set, cb_dset **out);
void cb_free_set(cb_ctx *ctx, cb_dset *set);
/* Recompute the hashes of a resident set (db_hash(), src/db.cc:903-916, on data already in
   device memory). */
int cb_rehash(cb_ctx *ctx, cb_dset *set);
/* Copy the per-sequence hashes back (n entries) — for tests. */
int cb_get_hashes(cb_ctx *ctx, const cb_dset *set, uint64_t *out);

/* ---- set B: table + Bloom ------------------------------------------------------------------ */

/* Replaces hash_init() + bloom_init() + the hash_insert() loop (overlap.cc:861-873,
   hashtable.cc:31-54, bloompat.cc:61-78): builds the open-addressing table and the blocked
   Bloom filter over every sequence of the set and counts exact duplicates (dup2). The context
   keeps a reference to the set; free it only after the context is done with it. */
int cb_build_b(cb_ctx *ctx, cb_dset *b);
/* The context's current set B as a resident set (NULL if none), e.g. to run it against itself
   after cb_set_b / cb_set_b_sharded (self-comparison, overlap.cc:799-825).  Owned by the context
   when it came from host arrays: do not free it. */
cb_dset *cb_resident_b(cb_ctx *ctx);
/* Exact duplicates found while building (reference dup2, overlap.cc:861-873). */
uint64_t cb_dups_b(const cb_ctx *ctx);
/* Replaces check_duplicates() (overlap.cc:579-605) for an arbitrary resident set (dup1). */
int cb_count_dups(cb_ctx *ctx, cb_dset *set, uint64_t *out);
/* Replaces the process() loop of dedup() (src/dedup.cc:62-137,176-183) for a resident set:
   sequences with equal (repertoire, V, J unless ignore_genes, residues) form a group.
   leader_out[i] (n entries) = index of the first member of i's group in set order (== i for
   the member the reference reports, dedup.cc:27-59); count_out[i] (n entries) = summed
   duplicate_count of the group at its leader (1 per member with ignore_counts), 0 elsewhere;
   *merged_out (may be NULL) = members merged away = the reference's "Duplicates merged". */
int cb_dedup(cb_ctx *ctx, cb_dset *set, uint32_t *leader_out, uint64_t *count_out, uint64_t *merged_out);
/* Replaces the network and clustering phases of cluster() (src/cluster.cc:57-69,71-274 and
   :277-300,356-411) for a resident set (index_base 0): single-linkage clusters of the graph
   that links two sequences when they match under the context's -d/-i/-g options (the same
   kernels as cb_run, as a self-comparison without self hits).  Builds the set's table first if
   it is not the context's current set B.  Outputs, n entries each, one per OUTPUT ROW in the
   reference's row order (clusters by decreasing size, equal sizes in order of their first
   sequence; inside a cluster the reference's breadth-first order): order_out = sequence index,
   cluster_no_out = 1-based cluster number, cluster_size_out = size of that cluster.
   *n_clusters_out / *n_edges_out (either may be NULL) = clusters / directed edges found. */
int cb_cluster(cb_ctx *ctx, cb_dset *set, uint32_t *order_out, uint32_t *cluster_no_out,
               uint32_t *cluster_size_out, uint64_t *n_clusters_out, uint64_t *n_edges_out);

/* ---- set A: enumerate, probe, verify, accumulate ------------------------------------------- */

/* Replaces ThreadRunner(sim_thread).run() (overlap.cc:926-936 → :376-538): processes
   sequences [first, first+count) of resident set `a` against the built set B and adds the
   scores into the context's matrix.  `a` may be the same handle as set B (self-comparison,
   overlap.cc:799-825).  May be called repeatedly (chunks / shards). */
int cb_run(cb_ctx *ctx, const cb_dset *a, uint64_t first, uint64_t count);

/* Host arrays → built set B in one call.  The copy is chunked and pipelined: while chunk k+1
   crosses PCIe, chunk k is packed, hashed and inserted into the table and the Bloom filter(s). */
int cb_set_b(cb_ctx *ctx, const cb_set *b);
int cb_set_b_cols(cb_ctx *ctx, const cb_set_cols *b);
/* Host arrays → matrix contribution: upload (pipelined with hashing) + cb_run(all) + free. */
int cb_run_a(cb_ctx *ctx, const cb_set *a);
int cb_run_a_cols(cb_ctx *ctx, const cb_set_cols *a);

/* ---- multi-GPU: one context per GPU, NCCL over NVLink ---------------------------------------- */

#define CB_UNIQUE_ID_BYTES 128
/* A fresh communicator id (ncclGetUniqueId) into out[CB_UNIQUE_ID_BYTES]: made by one rank, carried
   to the others by the caller (any host channel), passed by all to cb_comm_init_rank. */
int cb_comm_unique_id(void *out);
/* One process (or host thread) per GPU: join a communicator of `world` contexts as `rank`.
   Collective: returns when all ranks have called it.  All contexts must have been created with
   the same options and seed. */
int cb_comm_init_rank(cb_ctx *ctx, const void *unique_id, int rank, int world);
/* All contexts in ONE process (the CLI's --gpus N): ctxs[i] becomes rank i of n. */
int cb_comm_init_all(cb_ctx **ctxs, int n);
int cb_comm_rank(const cb_ctx *ctx, int *rank, int *world);
/* The shard of a set of n_total sequences that `rank` of `world` holds for cb_set_b_sharded:
   [first, first + count) with equal shards of ceil(n_total / world) (the last may be short or
   empty).  Host helper, no GPU work. */
void cb_shard_range(uint64_t n_total, int rank, int world, uint64_t *first, uint64_t *count);
/* cb_set_b for a communicator: `shard` = this rank's cb_shard_range of set B (columns as in
   cb_set_b_cols; n_reps and index_base are those of the WHOLE set, the same on every rank).  Each rank copies only its shard across
   PCIe, packs and hashes it; records, hashes and residues are all-gathered over NVLink; every
   rank then builds the table and the filters of the whole set (the reference's one shared table,
   src/overlap.cc:861-873, replicated per GPU).  Collective.  Sequence indices (pairs, existence
   rows) are those of the whole set. */
int cb_set_b_sharded(cb_ctx *ctx, const cb_set_cols *shard, uint64_t n_total);
/* Sum the partial matrices of all ranks in place (ncclAllReduce, f64): replaces the merge of the
   per-thread matrices, src/overlap.cc:510-527.  Matrix mode.  Collective; a no-op for world 1. */
int cb_allreduce_matrix(cb_ctx *ctx);

/* ---- results ------------------------------------------------------------------------------- */

/* Matrix mode: n_reps_a x n_reps_b doubles, row-major, indexed by the repertoire numbers the
   caller used in cb_set.rep (reference repertoire_matrix, overlap.cc:44,222).
   Existence mode: rows of the most recent cb_run call, count x n_reps_b (overlap.cc:226). */
int cb_matrix_dims(const cb_ctx *ctx, uint64_t *rows, uint64_t *cols);
int cb_get_matrix(cb_ctx *ctx, double *out, size_t n_values);
int cb_clear_matrix(cb_ctx *ctx);
/* Device pointer of the matrix; valid until the next call that reallocates it (cb_run in
   existence mode) or cb_destroy. */
void *cb_matrix_device(cb_ctx *ctx);
/* Accumulate into a caller-owned DEVICE buffer of rows x cols doubles instead of the context's
   own (matrix mode only; rows must equal n_reps_a).  The buffer is not cleared and not freed by
   the engine; NULL returns to the engine-owned matrix. */
int cb_bind_matrix(cb_ctx *ctx, void *device_ptr, uint64_t rows, uint64_t cols);
/* Host → device: overwrite the matrix (used after an external reduction). */
int cb_set_matrix(cb_ctx *ctx, const double *in, size_t n_values);

/* Pairs collected by cb_run calls since the last drain (reference pairs_list, overlap.cc:385-388,
   455-507).  Order is unspecified (README.md:163).  Copies up to `cap` pairs and removes them. */
int cb_pairs_pending(const cb_ctx *ctx, uint64_t *n);
int cb_drain_pairs(cb_ctx *ctx, cb_pair *buf, size_t cap, size_t *n_out);

int cb_get_stats(const cb_ctx *ctx, cb_stats *out);

/* Closed-form number of variants the reference enumerates for one sequence
   (generate_variants, src/variants.cc:402-428): host helper, no GPU work. */
uint64_t cb_probe_count(const uint8_t *residues, uint32_t len, int alphabet_size,
                        int differences, int indels);

#ifdef __cplusplus
}
#endif
#endif /* COMPAIRR_B200_H */
