/*
 * overlap_oracle.c — plain-C restatement of CompAIRR's overlap hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see overlap_oracle.h): the checker for the CUDA engine and the
 * "port" CPU baseline of bench.py.  Never linked into the product.
 * Parity status: PINNED against the reference's fixtures and binary (tests/test_oracle_golden.py).
 *
 * It follows the reference's ALGORITHM (Zobrist hash -> variant list -> Bloom -> linear-probing
 * table -> exact check -> score -> matrix), each function citing the reference file:line it
 * restates; it is written fresh over SoA inputs.  Hash/Bloom values differ from the reference's
 * (they come from glibc random() there, src/arch.cc:86-104) — results do not depend on them.
 */
#define _POSIX_C_SOURCE 200809L
#include "overlap_oracle.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

enum { K_IDENT = 0, K_SUB = 1, K_DEL = 2, K_INS = 3, K_SUBSUB = 4 }; /* variants.h:24-31 */

typedef struct {
  uint64_t hash;
  uint32_t kind, pos1, pos2;
  uint8_t r1, r2;
} var_t; /* variants.h:65-73 */

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ---- PRNG for table values (ours; the reference uses random(), zobrist.cc:52-63) ---- */
static uint64_t rng_state;
static uint64_t rng_next(void) {
  uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

/* ---- Zobrist (zobrist.cc:28-88, zobrist.h:24-27) ---- */
typedef struct {
  int sigma;
  uint32_t rows;
  uint64_t *tab; /* sigma * rows */
  uint64_t *vtab, *jtab;
  int ignore_genes;
} zob_t;

static uint64_t zv(const zob_t *z, uint32_t pos, uint32_t x) { return z->tab[(size_t)z->sigma * pos + x]; }

static uint64_t zob_hash(const zob_t *z, const uint8_t *s, uint32_t len, uint32_t v, uint32_t j) {
  uint64_t h = 0; /* zobrist.cc:74-88 */
  if (!z->ignore_genes) h ^= z->vtab[v] ^ z->jtab[j];
  for (uint32_t p = 0; p < len; p++) h ^= zv(z, p, s[p]);
  return h;
}

/* ---- linear-probing table (hashtable.h:22-77, hashtable.cc:31-54) ---- */
typedef struct {
  uint64_t size, mask;
  uint64_t *val, *dat;
  uint8_t *occ;
} tab_t;

static int tab_init(tab_t *t, uint64_t n) {
  t->size = 1;
  while (70 * t->size < 100 * n) t->size <<= 1; /* hashtable.cc:24,36-38 */
  t->mask = t->size - 1;
  t->val = malloc(t->size * 8);
  t->dat = malloc(t->size * 8);
  t->occ = calloc((t->size + 63) / 8, 1);
  return (t->val && t->dat && t->occ) ? 0 : -1;
}
static void tab_free(tab_t *t) { free(t->val); free(t->dat); free(t->occ); }
static uint64_t tab_home(const tab_t *t, uint64_t h) { return (h >> 32) & t->mask; } /* hashtable.h:36-41 */
static int tab_occ(const tab_t *t, uint64_t j) { return t->occ[j >> 3] & (1 << (j & 7)); }

/* ---- blocked Bloom filter with pattern table (bloompat.h:26-58, bloompat.cc:36-78) ---- */
typedef struct {
  uint64_t mask;
  uint64_t *bits; /* normal polarity here: 1 = set (the reference stores the complement) */
  uint64_t pat[1024];
} bloom_t;

static int bloom_init(bloom_t *b, uint64_t bytes) {
  if (bytes < 8) bytes = 8;
  b->mask = (bytes >> 3) - 1;
  b->bits = calloc(bytes >> 3, 8);
  for (int i = 0; i < 1024; i++) { /* 8 distinct bits per pattern, bloompat.cc:38-51 */
    uint64_t p = 0;
    for (int k = 0; k < 8; k++) {
      uint64_t one;
      do one = 1ull << (rng_next() & 63); while (p & one);
      p |= one;
    }
    b->pat[i] = p;
  }
  return b->bits ? 0 : -1;
}
static void bloom_set(bloom_t *b, uint64_t h) { b->bits[(h >> 10) & b->mask] |= b->pat[h & 1023]; }
static int bloom_get(const bloom_t *b, uint64_t h) {
  uint64_t p = b->pat[h & 1023];
  return (b->bits[(h >> 10) & b->mask] & p) == p;
}

/* ---- set accessors ---- */
static uint32_t s_len(const orc_set *s, uint64_t i) { return (uint32_t)(s->offsets[i + 1] - s->offsets[i]); }
static const uint8_t *s_seq(const orc_set *s, uint64_t i) { return s->residues + s->offsets[i]; }
static uint32_t s_v(const orc_set *s, uint64_t i) { return s->v_gene ? s->v_gene[i] : 0; }
static uint32_t s_j(const orc_set *s, uint64_t i) { return s->j_gene ? s->j_gene[i] : 0; }
static uint64_t s_cnt(const orc_set *s, uint64_t i) { return s->count ? s->count[i] : 1; }

/* ---- score summand (overlap.cc:144-166) ---- */
static double score_of(const orc_opts *o, uint64_t a, uint64_t b) {
  if (o->ignore_counts) return 1;
  switch (o->score) {
    case 5: case 0: return (double)a * (double)b;
    case 1: return (double)a / (double)b;
    case 6: case 2: return a < b ? a : b;
    case 3: return a > b ? a : b;
    case 4: return ((double)a + (double)b) / 2;
  }
  return 0;
}

/* ---- variant generation (variants.cc:260-428) ---- */
static uint64_t gen_variants(const zob_t *z, uint64_t hash, const uint8_t *s, uint32_t len,
                             uint32_t v, uint32_t j, int d, int indels, var_t *out) {
  uint64_t n = 0;
  const uint32_t S = (uint32_t)z->sigma;
#define EMIT(H, K, P1, R1, P2, R2) do { if (out) { var_t *vr_ = out + n; vr_->hash = (H); vr_->kind = (K); \
    vr_->pos1 = (P1); vr_->r1 = (uint8_t)(R1); vr_->pos2 = (P2); vr_->r2 = (uint8_t)(R2); } n++; } while (0)
  EMIT(hash, K_IDENT, 0, 0, 0, 0); /* :260-268 */
  if (d >= 1) {
    for (uint32_t i = 0; i < len; i++) { /* substitutions :280-293 */
      uint64_t h1 = hash ^ zv(z, i, s[i]);
      for (uint32_t r = 0; r < S; r++)
        if (r != s[i]) EMIT(h1 ^ zv(z, i, r), K_SUB, i, r, 0, 0);
    }
    if (indels) {
      uint64_t g = z->ignore_genes ? 0 : (z->vtab[v] ^ z->jtab[j]);
      if (len > 1) { /* deletions, one per run of equal residues :301-325 */
        uint64_t h = g; /* zobrist_hash_delete_first, zobrist.cc:90-104 */
        for (uint32_t p = 1; p < len; p++) h ^= zv(z, p - 1, s[p]);
        EMIT(h, K_DEL, 0, 0, 0, 0);
        uint8_t deleted = s[0];
        for (uint32_t i = 1; i < len; i++)
          if (s[i] != deleted) {
            h ^= zv(z, i - 1, deleted) ^ zv(z, i - 1, s[i]);
            EMIT(h, K_DEL, i, 0, 0, 0);
            deleted = s[i];
          }
      }
      uint64_t h = g; /* insertions :329-353; zobrist_hash_insert_first, zobrist.cc:122-136 */
      for (uint32_t p = 0; p < len; p++) h ^= zv(z, p + 1, s[p]);
      for (uint32_t r = 0; r < S; r++) EMIT(h ^ zv(z, 0, r), K_INS, 0, r, 0, 0);
      for (uint32_t i = 0; i < len; i++) {
        h ^= zv(z, i, s[i]) ^ zv(z, i + 1, s[i]);
        for (uint32_t r = 0; r < S; r++)
          if (r != s[i]) EMIT(h ^ zv(z, i + 1, r), K_INS, i + 1, r, 0, 0);
      }
    }
  }
  if (d >= 2) /* double substitutions :357-400 */
    for (uint32_t i = 0; i < len; i++) {
      uint64_t h1 = hash ^ zv(z, i, s[i]);
      for (uint32_t r = 0; r < S; r++) {
        if (r == s[i]) continue;
        uint64_t h2 = h1 ^ zv(z, i, r);
        for (uint32_t k = i + 1; k < len; k++) {
          uint64_t h3 = h2 ^ zv(z, k, s[k]);
          for (uint32_t w = 0; w < S; w++)
            if (w != s[k]) EMIT(h3 ^ zv(z, k, w), K_SUBSUB, i, r, k, w);
        }
      }
    }
#undef EMIT
  return n;
}

static uint64_t max_variants(uint64_t L, uint64_t S, int d, int indels) { /* variants.cc:53-107 */
  uint64_t m = 1;
  if (d >= 1) { m += L * (S - 1); if (indels) m += L + (L + 1) * (S - 1) + 1; }
  if (d >= 2) m += L * (L - 1) / 2 * (S - 1) * (S - 1);
  return m;
}

/* exact check "hit == seed with this edit" (variants.cc:166-240) */
static int check_variant(const uint8_t *s, uint32_t sl, const var_t *v, const uint8_t *a, uint32_t al) {
  switch (v->kind) {
    case K_IDENT: return sl == al && !memcmp(s, a, sl);
    case K_SUB:
      return sl == al && a[v->pos1] == v->r1 && !memcmp(s, a, v->pos1) &&
             !memcmp(s + v->pos1 + 1, a + v->pos1 + 1, sl - v->pos1 - 1);
    case K_DEL:
      return sl - 1 == al && !memcmp(s, a, v->pos1) && !memcmp(s + v->pos1 + 1, a + v->pos1, sl - v->pos1 - 1);
    case K_INS:
      return sl + 1 == al && a[v->pos1] == v->r1 && !memcmp(s, a, v->pos1) &&
             !memcmp(s + v->pos1, a + v->pos1 + 1, sl - v->pos1);
    case K_SUBSUB:
      return sl == al && a[v->pos1] == v->r1 && a[v->pos2] == v->r2 && !memcmp(s, a, v->pos1) &&
             !memcmp(s + v->pos1 + 1, a + v->pos1 + 1, v->pos2 - v->pos1 - 1) &&
             !memcmp(s + v->pos2 + 1, a + v->pos2 + 1, sl - v->pos2 - 1);
  }
  return 0;
}

/* insert + duplicate detection (hash_insert, overlap.cc:63-128) */
static int tab_insert(const orc_set *s, const uint64_t *hashes, tab_t *t, bloom_t *b, uint64_t i,
                      int ignore_genes) {
  int dup = 0;
  uint64_t h = hashes[i], j = tab_home(t, h);
  while (tab_occ(t, j)) {
    if ((!b || bloom_get(b, h)) && t->val[j] == h) {
      uint64_t hit = t->dat[j];
      if (s->rep[i] == s->rep[hit] &&
          (ignore_genes || (s_v(s, i) == s_v(s, hit) && s_j(s, i) == s_j(s, hit))) &&
          s_len(s, i) == s_len(s, hit) && !memcmp(s_seq(s, i), s_seq(s, hit), s_len(s, i)))
        dup = 1;
    }
    j = (j + 1) & t->mask;
  }
  t->occ[j >> 3] |= (uint8_t)(1 << (j & 7));
  t->val[j] = h;
  t->dat[j] = i;
  if (b) bloom_set(b, h);
  return dup;
}

/* ---- shared state of one overlap run ---- */
typedef struct {
  const orc_set *a, *b;
  const orc_opts *o;
  zob_t z;
  tab_t t;
  bloom_t bl;
  uint64_t *hash_a, *hash_b;
  uint32_t longest_a;
  double *matrix;
  uint64_t cols;
  pthread_mutex_t mu;
  uint64_t next; /* chunk dispenser (overlap.cc:421-433) */
  uint64_t probes, pass, matches;
  uint64_t *pairs, n_pairs;
  uint64_t rows; /* matrix rows: set-A repertoires, or set-A sequences in existence mode */
} run_t;

typedef struct {
  double *m;
  uint64_t *pairs, n_pairs, cap_pairs;
  uint64_t probes, pass, matches;
} local_t;

static void add_pair(local_t *l, uint64_t a, uint64_t b) {
  if (l->n_pairs == l->cap_pairs) {
    l->cap_pairs = l->cap_pairs ? 2 * l->cap_pairs : 4096;
    l->pairs = realloc(l->pairs, l->cap_pairs * 16);
  }
  l->pairs[2 * l->n_pairs] = a;
  l->pairs[2 * l->n_pairs + 1] = b;
  l->n_pairs++;
}

static void record(run_t *r, local_t *l, uint64_t seed, uint64_t hit) { /* overlap.cc:211-245 */
  if (l->m) {
    uint64_t row = r->o->existence ? seed : r->a->rep[seed];
    l->m[r->cols * row + r->b->rep[hit]] += score_of(r->o, s_cnt(r->a, seed), s_cnt(r->b, hit));
  }
  l->matches++;
  if (r->o->want_pairs) add_pair(l, seed, hit);
}

/* process_variants + find_variant_matches (overlap.cc:168-284) */
static void seed_hash_path(run_t *r, local_t *l, uint64_t seed, var_t *vars) {
  const orc_set *a = r->a, *b = r->b;
  const uint8_t *s = s_seq(a, seed);
  uint32_t sl = s_len(a, seed);
  uint64_t nv = gen_variants(&r->z, r->hash_a[seed], s, sl, s_v(a, seed), s_j(a, seed),
                             r->o->differences, r->o->indels, vars);
  l->probes += nv;
  for (uint64_t k = 0; k < nv; k++) {
    const var_t *v = vars + k;
    if (!bloom_get(&r->bl, v->hash)) continue;
    l->pass++;
    for (uint64_t j = tab_home(&r->t, v->hash); tab_occ(&r->t, j); j = (j + 1) & r->t.mask) {
      if (r->t.val[j] != v->hash) continue;
      uint64_t hit = r->t.dat[j];
      if (!r->o->ignore_genes && (s_v(a, seed) != s_v(b, hit) || s_j(a, seed) != s_j(b, hit))) continue;
      if (check_variant(s, sl, v, s_seq(b, hit), s_len(b, hit))) record(r, l, seed, hit);
    }
  }
}

/* pairwise definition; with method == 0 this is process_trad + seq_diff (overlap.cc:286-359,
   util.cc:172-184), with method == 1 it also covers d <= 2 and the d = 1 indel case */
static int within(const uint8_t *x, uint32_t xl, const uint8_t *y, uint32_t yl, int d, int indels) {
  if (xl == yl) {
    int diff = 0;
    for (uint32_t p = 0; p < xl; p++)
      if (x[p] != y[p] && ++diff > d) return 0;
    return 1;
  }
  if (!indels) return 0;
  if (xl + 1 != yl && yl + 1 != xl) return 0;
  const uint8_t *lo = xl < yl ? x : y, *hi = xl < yl ? y : x; /* hi = lo + one insertion */
  uint32_t n = xl < yl ? xl : yl, p = 0;
  while (p < n && lo[p] == hi[p]) p++;
  return !memcmp(lo + p, hi + p + 1, n - p);
}

static void seed_pairwise(run_t *r, local_t *l, uint64_t seed) {
  const orc_set *a = r->a, *b = r->b;
  for (uint64_t hit = 0; hit < b->n; hit++) {
    if (!r->o->ignore_genes && (s_v(a, seed) != s_v(b, hit) || s_j(a, seed) != s_j(b, hit))) continue;
    if (within(s_seq(a, seed), s_len(a, seed), s_seq(b, hit), s_len(b, hit), r->o->differences, r->o->indels))
      record(r, l, seed, hit);
  }
}

/* worker (sim_thread, overlap.cc:376-538): private matrix, 1000-seed chunks, merge at the end */
static void *worker(void *arg) {
  run_t *r = arg;
  const int hashp = r->o->method == 0 && r->o->differences <= 2;
  local_t l;
  memset(&l, 0, sizeof l);
  var_t *vars = NULL;
  if (hashp) vars = malloc(sizeof(var_t) * max_variants(r->longest_a, (uint64_t)r->o->alphabet_size,
                                                      r->o->differences, r->o->indels));
  const uint64_t mrows = r->rows;
  if (r->matrix) l.m = calloc(mrows * r->cols + 1, sizeof(double));
  for (;;) {
    pthread_mutex_lock(&r->mu);
    uint64_t first = r->next;
    r->next = first + 1000 < r->a->n ? first + 1000 : r->a->n;
    uint64_t last = r->next;
    pthread_mutex_unlock(&r->mu);
    if (first >= r->a->n) break;
    for (uint64_t seed = first; seed < last; seed++)
      if (hashp) seed_hash_path(r, &l, seed, vars); else seed_pairwise(r, &l, seed);
  }
  pthread_mutex_lock(&r->mu);
  if (l.m) for (uint64_t k = 0; k < mrows * r->cols; k++) r->matrix[k] += l.m[k];
  r->probes += l.probes; r->pass += l.pass; r->matches += l.matches;
  if (l.n_pairs) {
    r->pairs = realloc(r->pairs, (r->n_pairs + l.n_pairs) * 16);
    memcpy(r->pairs + 2 * r->n_pairs, l.pairs, l.n_pairs * 16);
    r->n_pairs += l.n_pairs;
  }
  pthread_mutex_unlock(&r->mu);
  free(l.m); free(l.pairs); free(vars);
  return NULL;
}

static int zob_setup(zob_t *z, const orc_set *a, const orc_set *b, const orc_opts *o, uint32_t *longest_a) {
  uint32_t longest = 0, la = 0, vmax = 0, jmax = 0;
  const orc_set *sets[2] = {a, b};
  for (int k = 0; k < 2; k++) {
    const orc_set *s = sets[k];
    if (!s) continue;
    for (uint64_t i = 0; i < s->n; i++) {
      uint32_t L = s_len(s, i);
      if (L > longest) longest = L;
      if (k == 0 && L > la) la = L;
      if (s_v(s, i) > vmax) vmax = s_v(s, i);
      if (s_j(s, i) > jmax) jmax = s_j(s, i);
    }
  }
  if (longest_a) *longest_a = la;
  z->sigma = o->alphabet_size;
  z->rows = longest + 3; /* MAX_INSERTS, overlap.cc:840, compairr.h:111 */
  z->ignore_genes = o->ignore_genes;
  z->tab = malloc(sizeof(uint64_t) * (size_t)z->sigma * z->rows);
  z->vtab = malloc(sizeof(uint64_t) * ((size_t)vmax + 1));
  z->jtab = malloc(sizeof(uint64_t) * ((size_t)jmax + 1));
  if (!z->tab || !z->vtab || !z->jtab) return -1;
  rng_state = 1;
  for (size_t i = 0; i < (size_t)z->sigma * z->rows; i++) z->tab[i] = rng_next();
  for (size_t i = 0; i <= vmax; i++) z->vtab[i] = rng_next();
  for (size_t i = 0; i <= jmax; i++) z->jtab[i] = rng_next();
  return 0;
}
static void zob_free(zob_t *z) { free(z->tab); free(z->vtab); free(z->jtab); }

int orc_overlap(const orc_set *a, const orc_set *b, const orc_opts *o, uint32_t n_reps_a,
                double *matrix, orc_result *res) {
  run_t r;
  memset(&r, 0, sizeof r);
  r.a = a; r.b = b; r.o = o; r.matrix = matrix; r.cols = b->n_reps;
  const int hashp = o->method == 0 && o->differences <= 2;
  r.rows = o->existence ? a->n : n_reps_a;
  if (matrix) memset(matrix, 0, sizeof(double) * r.rows * r.cols);
  double t0 = now_s();
  if (zob_setup(&r.z, a, b, o, &r.longest_a)) return -1;
  if (hashp) { /* overlap.cc:838-873 */
    r.hash_a = malloc(8 * (a->n ? a->n : 1));
    r.hash_b = (a == b) ? r.hash_a : malloc(8 * (b->n ? b->n : 1));
    for (uint64_t i = 0; i < a->n; i++) r.hash_a[i] = zob_hash(&r.z, s_seq(a, i), s_len(a, i), s_v(a, i), s_j(a, i));
    if (a != b)
      for (uint64_t i = 0; i < b->n; i++) r.hash_b[i] = zob_hash(&r.z, s_seq(b, i), s_len(b, i), s_v(b, i), s_j(b, i));
    if (tab_init(&r.t, b->n) || bloom_init(&r.bl, r.t.size)) return -1; /* Bloom bytes = table slots, overlap.cc:863 */
    for (uint64_t i = 0; i < b->n; i++) tab_insert(b, r.hash_b, &r.t, &r.bl, i, o->ignore_genes);
  }
  double t1 = now_s();
  pthread_mutex_init(&r.mu, NULL);
  int nt = o->threads < 1 ? 1 : o->threads;
  if (nt == 1) {
    worker(&r);
  } else {
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)nt);
    for (int k = 0; k < nt; k++) pthread_create(th + k, NULL, worker, &r);
    for (int k = 0; k < nt; k++) pthread_join(th[k], NULL);
    free(th);
  }
  pthread_mutex_destroy(&r.mu);
  double t2 = now_s();
  if (res) {
    res->probes = r.probes; res->bloom_pass = r.pass; res->matches = r.matches;
    res->n_pairs = r.n_pairs; res->pairs = r.pairs;
    res->seconds_build = t1 - t0; res->seconds_probe = t2 - t1;
  } else {
    free(r.pairs);
  }
  if (hashp) {
    tab_free(&r.t); free(r.bl.bits);
    if (r.hash_b != r.hash_a) free(r.hash_b);
    free(r.hash_a);
  }
  zob_free(&r.z);
  return 0;
}

uint64_t orc_count_dups(const orc_set *s, int alphabet_size, int ignore_genes) { /* overlap.cc:579-605 */
  orc_opts o;
  memset(&o, 0, sizeof o);
  o.alphabet_size = alphabet_size; o.ignore_genes = ignore_genes;
  zob_t z;
  if (zob_setup(&z, s, NULL, &o, NULL)) return (uint64_t)-1;
  uint64_t *h = malloc(8 * (s->n ? s->n : 1));
  for (uint64_t i = 0; i < s->n; i++) h[i] = zob_hash(&z, s_seq(s, i), s_len(s, i), s_v(s, i), s_j(s, i));
  tab_t t;
  tab_init(&t, s->n);
  uint64_t dups = 0;
  for (uint64_t i = 0; i < s->n; i++) dups += (uint64_t)tab_insert(s, h, &t, NULL, i, ignore_genes);
  tab_free(&t); free(h); zob_free(&z);
  return dups;
}

static void materialise(const uint8_t *s, uint32_t len, const var_t *v, uint8_t *out, uint32_t *olen) {
  /* the variant's sequence (generate_variant_sequence, variants.cc:109-163) */
  switch (v->kind) {
    case K_IDENT: memcpy(out, s, len); *olen = len; break;
    case K_SUB: memcpy(out, s, len); out[v->pos1] = v->r1; *olen = len; break;
    case K_DEL: memcpy(out, s, v->pos1); memcpy(out + v->pos1, s + v->pos1 + 1, len - v->pos1 - 1); *olen = len - 1; break;
    case K_INS: memcpy(out, s, v->pos1); out[v->pos1] = v->r1; memcpy(out + v->pos1 + 1, s + v->pos1, len - v->pos1); *olen = len + 1; break;
    case K_SUBSUB: memcpy(out, s, len); out[v->pos1] = v->r1; out[v->pos2] = v->r2; *olen = len; break;
  }
}

uint64_t orc_enumerate(const uint8_t *seq, uint32_t len, int alphabet_size, int differences,
                       int indels, uint32_t *recs, uint8_t *seqs, uint32_t seq_stride, uint64_t cap) {
  /* record k: recs[5k..5k+4] = kind,pos1,res1,pos2,res2; seqs[k*stride] = length, then residues */
  orc_opts o;
  memset(&o, 0, sizeof o);
  o.alphabet_size = alphabet_size; o.ignore_genes = 1;
  orc_set s;
  memset(&s, 0, sizeof s);
  uint64_t offs[2] = {0, len};
  s.n = 1; s.residues = seq; s.offsets = offs;
  zob_t z;
  if (zob_setup(&z, &s, NULL, &o, NULL)) return 0;
  uint64_t m = max_variants(len, (uint64_t)alphabet_size, differences, indels);
  var_t *vars = malloc(sizeof(var_t) * m);
  uint64_t n = gen_variants(&z, zob_hash(&z, seq, len, 0, 0), seq, len, 0, 0, differences, indels, vars);
  for (uint64_t k = 0; k < n && k < cap; k++) {
    if (recs) { recs[5*k] = vars[k].kind; recs[5*k+1] = vars[k].pos1; recs[5*k+2] = vars[k].r1; recs[5*k+3] = vars[k].pos2; recs[5*k+4] = vars[k].r2; }
    if (seqs && seq_stride >= len + 2) {
      uint32_t ol = 0;
      materialise(seq, len, vars + k, seqs + k * seq_stride + 1, &ol);
      seqs[k * seq_stride] = (uint8_t)ol;
    }
  }
  free(vars); zob_free(&z);
  return n;
}

void orc_free(void *p) { free(p); }

/* ---- AIRR TSV writer for bench inputs (the reference binary and the CLI read files) ----------- */
/* Same text as SeqSet.write_tsv (compairr_b200/seqset.py): columns repertoire_id, sequence_id,
   duplicate_count, v_call, j_call, junction_aa|junction; names R%04u, <prefix><index>, TRBV%02u,
   TRBJ%02u.  Plain buffered formatting: ~10^7 lines per second. */
static char *put_u64(char *p, uint64_t v) {
  char t[24];
  int n = 0;
  do { t[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  while (n) *p++ = t[--n];
  return p;
}
static char *put_pad(char *p, uint64_t v, int width) {
  char t[24];
  int n = 0;
  do { t[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  for (int k = n; k < width; k++) *p++ = '0';
  while (n) *p++ = t[--n];
  return p;
}
int orc_write_tsv(const orc_set *s, const char *path, const char *id_prefix, int nucleotides,
                  uint64_t index_base) {
  FILE *f = fopen(path, "w");
  if (!f) return -1;
  const char *alpha = nucleotides ? "ACGT" : "ACDEFGHIKLMNPQRSTVWY";
  const size_t cap = 1u << 22, plen = strlen(id_prefix);
  char *buf = malloc(cap + 4096);
  if (!buf) { fclose(f); return -1; }
  char *p = buf;
  p += sprintf(p, "repertoire_id\tsequence_id\tduplicate_count\tv_call\tj_call\t%s\n", nucleotides ? "junction" : "junction_aa");
  int bad = 0;
  for (uint64_t i = 0; i < s->n && !bad; i++) {
    const uint64_t len = s->offsets[i + 1] - s->offsets[i];
    if ((size_t)(p - buf) + len + 256 + plen > cap) {
      bad |= fwrite(buf, 1, (size_t)(p - buf), f) != (size_t)(p - buf);
      p = buf;
      if (len + 256 + plen > cap) { bad = 1; break; }
    }
    *p++ = 'R'; p = put_pad(p, s->rep[i], 4); *p++ = '\t';
    memcpy(p, id_prefix, plen); p += plen; p = put_u64(p, i + index_base); *p++ = '\t';
    p = put_u64(p, s->count ? s->count[i] : 1); *p++ = '\t';
    memcpy(p, "TRBV", 4); p += 4; p = put_pad(p, (uint64_t)(s->v_gene ? s->v_gene[i] : 0) + 1, 2); *p++ = '\t';
    memcpy(p, "TRBJ", 4); p += 4; p = put_pad(p, (uint64_t)(s->j_gene ? s->j_gene[i] : 0) + 1, 2); *p++ = '\t';
    const uint8_t *r = s->residues + s->offsets[i];
    for (uint64_t k = 0; k < len; k++) *p++ = alpha[r[k]];
    *p++ = '\n';
  }
  if (!bad) bad |= fwrite(buf, 1, (size_t)(p - buf), f) != (size_t)(p - buf);
  free(buf);
  if (fclose(f)) bad = 1;
  return bad ? -1 : 0;
}
