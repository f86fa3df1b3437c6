/*
 * overlap_oracle.h — CPU oracle for the repertoire-overlap hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  The product (compairr_b200/) never links, imports or executes anything here.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement against
 *   - the reference's own fixture test/expected.tsv and the README worked examples,
 *   - golden outputs of the unmodified reference binary (oracle/_ref/compairr, built from
 *     /root/reference/src by oracle/Makefile) committed under tests/golden/,
 *   - and, where oracle/_ref/compairr is present, live differential runs on seeded inputs.
 */
#ifndef OVERLAP_ORACLE_H
#define OVERLAP_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_set {
  uint64_t n;
  const uint8_t *residues;  /* codes 0..sigma-1 */
  const uint64_t *offsets;  /* n + 1 */
  const uint32_t *v_gene;   /* may be NULL with ignore_genes */
  const uint32_t *j_gene;
  const uint32_t *rep;
  const uint64_t *count;    /* may be NULL: all 1 */
  uint32_t n_reps;
} orc_set;

typedef struct orc_opts {
  int32_t alphabet_size;  /* 4 or 20 */
  int32_t differences;
  int32_t indels;
  int32_t ignore_genes;
  int32_t ignore_counts;
  int32_t score;          /* reference enum: product ratio min max mean mh jaccard */
  int32_t existence;      /* rows = set-A sequences instead of repertoires */
  int32_t threads;        /* >= 1 */
  int32_t method;         /* 0 = as the reference (hash path for d<=2, pairwise for d>=3),
                             1 = force the pairwise definition (O(N1*N2)) for any d */
  int32_t want_pairs;
} orc_opts;

typedef struct orc_result {
  uint64_t probes;    /* variants enumerated (hash path) */
  uint64_t bloom_pass;
  uint64_t matches;
  uint64_t n_pairs;
  uint64_t *pairs;    /* 2*n_pairs values (a, b), malloc'd; free with orc_free */
  double seconds_build;
  double seconds_probe;
} orc_result;

/* matrix: rows x b->n_reps doubles, zeroed by the caller or not (it is overwritten); rows =
   n_reps_a (matrix mode) or a->n (existence).  May be NULL (no matrix). Returns 0 on success. */
int orc_overlap(const orc_set *a, const orc_set *b, const orc_opts *o, uint32_t n_reps_a,
                double *matrix, orc_result *res);

/* exact duplicates as the reference counts them (overlap.cc:63-128, 579-605) */
uint64_t orc_count_dups(const orc_set *s, int alphabet_size, int ignore_genes);

/* Enumerate the variants the reference generates for one sequence (variants.cc:402-428).
   Writes up to cap records of 5 uint32 {kind,pos1,res1,pos2,res2} into recs (may be NULL) and up
   to cap materialised variant sequences of stride (len+1) bytes + 1 length byte... see .c.
   Returns the number of variants. */
uint64_t orc_enumerate(const uint8_t *seq, uint32_t len, int alphabet_size, int differences,
                       int indels, uint32_t *recs, uint8_t *seqs, uint32_t seq_stride,
                       uint64_t cap);

void orc_free(void *p);

/* Write a set as an AIRR TSV file (inputs of the reference binary / the CLI in bench.py and the
   tests): the same text as SeqSet.write_tsv.  Returns 0 on success. */
int orc_write_tsv(const orc_set *s, const char *path, const char *id_prefix, int nucleotides,
                  uint64_t index_base);

#ifdef __cplusplus
}
#endif
#endif
