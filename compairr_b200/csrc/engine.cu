// engine.cu — the C ABI (include/compairr_b200.h) on top of the kernels: context, device
// memory, launch orchestration, result read-back.  Host-side only; the arithmetic is in
// kernels.cu / brute.cu.  There is no CPU fallback anywhere in this file: every path either
// launches the CUDA kernels or returns an error.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "engine_internal.h"

using namespace cb;

static thread_local std::string g_error;

int cb_fail(cb_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c)
    c->err = buf;
  else
    g_error = buf;
  return code;
}
#define fail cb_fail

static int bind(cb_ctx* c) {
  CU(c, cudaSetDevice(c->device));
  return CB_OK;
}

static int read_counters(cb_ctx* c) {
  CU(c, cudaMemcpyAsync(c->h_counters, c->d_counters, CTR_COUNT * sizeof(unsigned long long),
                        cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return CB_OK;
}

static int zero_counter(cb_ctx* c, int which) {
  CU(c, cudaMemsetAsync(c->d_counters + which, 0, sizeof(unsigned long long), c->stream));
  return CB_OK;
}

// Zobrist table for positions 0..rows-1.  Values depend only on (seed, position, residue), so
// growing the table never changes a hash that was already computed.
static int ensure_ztab(cb_ctx* c, uint32_t rows) {
  if (rows <= c->zrows) return CB_OK;
  rows = (rows + 15) & ~15u;
  const uint32_t sigma = (uint32_t)c->cfg.alphabet_size;
  std::vector<uint64_t> h((size_t)rows * sigma);
  for (uint32_t p = 0; p < rows; p++)
    for (uint32_t r = 0; r < sigma; r++) h[(size_t)p * sigma + r] = zobrist_gen(c->cfg.seed, p, r);
  uint64_t* d = nullptr;
  CU(c, cudaMalloc(&d, h.size() * sizeof(uint64_t)));
  cudaError_t e = cudaMemcpyAsync(d, h.data(), h.size() * sizeof(uint64_t),
                                  cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) {
    cudaFree(d);
    return fail(c, CB_ERR_CUDA, "Zobrist table upload: %s", cudaGetErrorString(e));
  }
  if (c->d_ztab) cudaFree(c->d_ztab);
  c->d_ztab = d;
  c->zrows = rows;
  return CB_OK;
}

DeviceSetView cb_view_of(const cb_dset* s) {
  DeviceSetView v;
  v.meta = s->d_meta;
  v.res = s->d_res;
  v.hash = s->d_hash;
  v.n = s->n;
  v.index_base = s->index_base;
  return v;
}

// ---------------------------------------------------------------------------------------------

extern "C" const char* cb_global_error(void) { return g_error.c_str(); }
extern "C" int cb_abi_version(void) { return CB_ABI_VERSION; }

extern "C" int cb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

static int validate(const cb_config* cfg) {
  if (!cfg) return fail(nullptr, CB_ERR_INVALID, "cb_create: config is NULL");
  if (cfg->abi_version != CB_ABI_VERSION)
    return fail(nullptr, CB_ERR_INVALID, "cb_create: ABI version %d, library is %d",
                cfg->abi_version, CB_ABI_VERSION);
  if (cfg->alphabet_size != 4 && cfg->alphabet_size != 20)
    return fail(nullptr, CB_ERR_INVALID, "cb_create: alphabet_size must be 4 or 20");
  // Same option rules as the reference CLI (compairr.cc:636-689), restated for the library.
  if (cfg->differences < 0)
    return fail(nullptr, CB_ERR_INVALID,
                "Differences specified with -d or -differences cannot be negative.");
  if (cfg->indels && cfg->differences != 1)
    return fail(nullptr, CB_ERR_INVALID, "Indels are only allowed when d=1");
  if (cfg->score < 0 || cfg->score > CB_SCORE_JACCARD)
    return fail(nullptr, CB_ERR_INVALID, "cb_create: unknown score %d", cfg->score);
  if (cfg->mode != CB_MODE_MATRIX && cfg->mode != CB_MODE_EXISTENCE)
    return fail(nullptr, CB_ERR_INVALID, "cb_create: unknown mode %d", cfg->mode);
  if (cfg->mode != CB_MODE_MATRIX && cfg->score == CB_SCORE_MH)
    return fail(nullptr, CB_ERR_INVALID,
                "The Morisita-Horn index is only allowed when computing repertoire overlap");
  if (cfg->mode != CB_MODE_MATRIX && cfg->score == CB_SCORE_JACCARD)
    return fail(nullptr, CB_ERR_INVALID,
                "The Jaccard index is only allowed when computing repertoire overlap");
  if (cfg->differences > 0 && cfg->score == CB_SCORE_MH)
    return fail(nullptr, CB_ERR_INVALID, "The Morisita-Horn index is not defined when d>0");
  if (cfg->differences > 0 && cfg->score == CB_SCORE_JACCARD)
    return fail(nullptr, CB_ERR_INVALID, "The Jaccard index is not defined when d>0");
  if (cfg->mode == CB_MODE_MATRIX && !cfg->no_matrix && cfg->n_reps_a == 0)
    return fail(nullptr, CB_ERR_INVALID, "cb_create: n_reps_a must be > 0 in matrix mode");
  return CB_OK;
}

extern "C" int cb_create(const cb_config* cfg, cb_ctx** out) {
  if (!out) return fail(nullptr, CB_ERR_INVALID, "cb_create: out is NULL");
  *out = nullptr;
  int rc = validate(cfg);
  if (rc) return rc;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, CB_ERR_CUDA,
                "no usable CUDA device (%s); this engine has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  if (cfg->device < 0 || cfg->device >= ndev)
    return fail(nullptr, CB_ERR_INVALID, "cb_create: device %d out of range (%d devices)",
                cfg->device, ndev);
  cb_ctx* c = new (std::nothrow) cb_ctx;
  if (!c) return fail(nullptr, CB_ERR_NOMEM, "cb_create: out of host memory");
  c->cfg = *cfg;
  if (c->cfg.bloom_bits_per_key_x16 == 0) c->cfg.bloom_bits_per_key_x16 = 16 * 16;
  if (c->cfg.table_load_pct == 0 || c->cfg.table_load_pct > 90) c->cfg.table_load_pct = 50;
  if (c->cfg.pairs_capacity == 0) c->cfg.pairs_capacity = 1ull << 24;
  if (c->cfg.bloom_l2_cap_kib == 0) c->cfg.bloom_l2_cap_kib = 40 * 1024;
  if (c->cfg.seed == 0) c->cfg.seed = 1;
  c->device = cfg->device;
#define CU_CREATE(expr)                                                                  \
  do {                                                                                   \
    cudaError_t e2 = (expr);                                                             \
    if (e2 != cudaSuccess) {                                                             \
      fail(nullptr, CB_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e2));               \
      cb_destroy(c);                                                                     \
      return CB_ERR_CUDA;                                                                \
    }                                                                                    \
  } while (0)
  CU_CREATE(cudaSetDevice(c->device));
  cudaDeviceProp prop;
  CU_CREATE(cudaGetDeviceProperties(&prop, c->device));
  c->sm_count = prop.multiProcessorCount;
  CU_CREATE(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  for (auto& ev : c->ev) CU_CREATE(cudaEventCreate(&ev));
  CU_CREATE(cudaMalloc(&c->d_counters, CTR_COUNT * sizeof(unsigned long long)));
  CU_CREATE(cudaMemset(c->d_counters, 0, CTR_COUNT * sizeof(unsigned long long)));
  CU_CREATE(cudaMallocHost(&c->h_counters, CTR_COUNT * sizeof(unsigned long long)));
#undef CU_CREATE
  *out = c;
  return CB_OK;
}

static void free_dset(cb_dset* s) {
  if (!s) return;
  cudaFree(s->d_meta);
  cudaFree(s->d_res);
  cudaFree(s->d_hash);
  cudaFree(s->d_order);
  cudaFree(s->d_packed);
  delete s;
}

extern "C" void cb_destroy(cb_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->own_stream) cudaStreamSynchronize(c->own_stream);
  if (c->b_owned) free_dset(c->b);
  cudaFree(c->d_ztab);
  cudaFree(c->d_counters);
  if (c->h_counters) cudaFreeHost(c->h_counters);
  cudaFree(c->d_table);
  cudaFree(c->d_bloom);
  cudaFree(c->d_bloom2);
  if (!c->matrix_external) cudaFree(c->d_matrix);
  cudaFree(c->d_pairs);
  for (auto& ev : c->ev)
    if (ev) cudaEventDestroy(ev);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
}

extern "C" const char* cb_last_error(const cb_ctx* c) { return c ? c->err.c_str() : g_error.c_str(); }

extern "C" int cb_set_stream(cb_ctx* c, void* s) {
  if (!c) return CB_ERR_INVALID;
  c->stream = s ? (cudaStream_t)s : c->own_stream;
  return CB_OK;
}

// ---- upload + hash -----------------------------------------------------------------------------

extern "C" int cb_upload(cb_ctx* c, const cb_set* set, cb_dset** out) {
  if (!c || !set || !out) return fail(c, CB_ERR_INVALID, "cb_upload: NULL argument");
  *out = nullptr;
  if (set->n && (!set->residues || !set->offsets))
    return fail(c, CB_ERR_INVALID, "cb_upload: residues/offsets are NULL");
  if (set->n && !c->cfg.ignore_genes && (!set->v_gene || !set->j_gene))
    return fail(c, CB_ERR_INVALID, "cb_upload: v_gene/j_gene are NULL but ignore_genes is off");
  if (set->n && !c->cfg.ignore_counts && !set->count)
    return fail(c, CB_ERR_INVALID, "cb_upload: count is NULL but ignore_counts is off");
  if (set->n >= 0xffffffffull && c->cfg.differences > MAXDIFF_HASH)
    return fail(c, CB_ERR_LIMIT, "cb_upload: more than 2^32-1 sequences in one set");
  int rc = bind(c);
  if (rc) return rc;
  cb_dset* s = new (std::nothrow) cb_dset;
  if (!s) return fail(c, CB_ERR_NOMEM, "cb_upload: out of host memory");
  s->n = set->n;
  s->index_base = set->index_base;
  s->n_reps = set->n_reps;
  const uint64_t n = set->n;
  if (n == 0) {
    *out = s;
    return CB_OK;
  }
  const uint64_t off_base = set->offsets[0];
  s->res_bytes = set->offsets[n] - off_base;

  uint64_t* t_off = nullptr;
  uint32_t *t_v = nullptr, *t_j = nullptr, *t_rep = nullptr;
  uint64_t* t_cnt = nullptr;
  auto cleanup = [&]() {
    cudaFree(t_off);
    cudaFree(t_v);
    cudaFree(t_j);
    cudaFree(t_rep);
    cudaFree(t_cnt);
  };
#define CU_UP(expr)                                                                          \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      cleanup();                                                                             \
      free_dset(s);                                                                          \
      return fail(c, e__ == cudaErrorMemoryAllocation ? CB_ERR_NOMEM : CB_ERR_CUDA, "%s: %s", \
                  #expr, cudaGetErrorString(e__));                                           \
    }                                                                                        \
  } while (0)
  const bool genes = !c->cfg.ignore_genes && set->v_gene && set->j_gene;
  CU_UP(cudaMalloc(&s->d_res, s->res_bytes + 16));
  CU_UP(cudaMalloc(&s->d_meta, n * sizeof(SeqMeta)));
  CU_UP(cudaMalloc(&s->d_hash, n * sizeof(uint64_t)));
  CU_UP(cudaMalloc(&t_off, (n + 1) * sizeof(uint64_t)));
  if (genes) {
    CU_UP(cudaMalloc(&t_v, n * sizeof(uint32_t)));
    CU_UP(cudaMalloc(&t_j, n * sizeof(uint32_t)));
  }
  if (set->rep) CU_UP(cudaMalloc(&t_rep, n * sizeof(uint32_t)));
  if (set->count) CU_UP(cudaMalloc(&t_cnt, n * sizeof(uint64_t)));
  cudaStream_t st = c->stream;
  CU_UP(cudaMemcpyAsync(s->d_res, set->residues + off_base, s->res_bytes, cudaMemcpyHostToDevice, st));
  CU_UP(cudaMemcpyAsync(t_off, set->offsets, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  if (genes) {
    CU_UP(cudaMemcpyAsync(t_v, set->v_gene, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    CU_UP(cudaMemcpyAsync(t_j, set->j_gene, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  }
  if (set->rep)
    CU_UP(cudaMemcpyAsync(t_rep, set->rep, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  if (set->count)
    CU_UP(cudaMemcpyAsync(t_cnt, set->count, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  CU_UP(cudaMemsetAsync(c->d_counters + CTR_MAXLEN, 0, sizeof(unsigned long long), st));
  launch_pack_meta(t_off, t_v, t_j, t_rep, t_cnt, n, off_base, s->d_meta, c->d_counters, st);
  CU_UP(cudaGetLastError());
  CU_UP(cudaMemcpyAsync(c->h_counters, c->d_counters, CTR_COUNT * sizeof(unsigned long long),
                        cudaMemcpyDeviceToHost, st));
  CU_UP(cudaStreamSynchronize(st));
  cleanup();
  t_off = nullptr;
  t_v = t_j = t_rep = nullptr;
  t_cnt = nullptr;
  s->longest = (uint32_t)c->h_counters[CTR_MAXLEN];
  if (s->longest >= (1u << 20)) {
    free_dset(s);
    return fail(c, CB_ERR_LIMIT, "cb_upload: sequence longer than 2^20 residues");
  }
  rc = ensure_ztab(c, s->longest + 2);
  if (rc) {
    free_dset(s);
    return rc;
  }
  CU_UP(cudaEventRecord(c->ev[0], st));
  launch_hash(s->d_meta, s->d_res, n, c->d_ztab, s->longest + 1,
              (uint32_t)c->cfg.alphabet_size, c->cfg.seed, c->cfg.ignore_genes != 0, s->d_hash, st);
  CU_UP(cudaGetLastError());
  CU_UP(cudaEventRecord(c->ev[1], st));
  CU_UP(cudaStreamSynchronize(st));
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->stats.ms_hash_a = ms;  // the caller decides whether this was set A or set B
#undef CU_UP
  *out = s;
  return CB_OK;
}

extern "C" void cb_free_set(cb_ctx* c, cb_dset* s) {
  if (!s) return;
  if (c) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->b == s) {
      c->b = nullptr;
      c->b_owned = false;
    }
  }
  free_dset(s);
}

extern "C" int cb_rehash(cb_ctx* c, cb_dset* s) {
  if (!c || !s) return fail(c, CB_ERR_INVALID, "cb_rehash: NULL argument");
  int rc = bind(c);
  if (rc) return rc;
  if (s->n == 0) return CB_OK;
  rc = ensure_ztab(c, s->longest + 2);
  if (rc) return rc;
  CU(c, cudaEventRecord(c->ev[0], c->stream));
  launch_hash(s->d_meta, s->d_res, s->n, c->d_ztab, s->longest + 1, (uint32_t)c->cfg.alphabet_size,
              c->cfg.seed, c->cfg.ignore_genes != 0, s->d_hash, c->stream);
  CU(c, cudaGetLastError());
  CU(c, cudaEventRecord(c->ev[1], c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  cudaEventElapsedTime(&c->stats.ms_hash_a, c->ev[0], c->ev[1]);
  return CB_OK;
}

extern "C" int cb_get_hashes(cb_ctx* c, const cb_dset* s, uint64_t* out) {
  if (!c || !s || !out) return fail(c, CB_ERR_INVALID, "cb_get_hashes: NULL argument");
  int rc = bind(c);
  if (rc) return rc;
  if (s->n == 0) return CB_OK;
  CU(c, cudaMemcpyAsync(out, s->d_hash, s->n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return CB_OK;
}

// ---- set B -------------------------------------------------------------------------------------

struct BuiltTable {
  Slot* table = nullptr;
  uint64_t slots = 0;
  unsigned long long* bloom = nullptr;
  uint32_t blocks = 0;
  bool k2 = false;
  unsigned long long* bloom2 = nullptr;
  uint32_t blocks2 = 0;
  void release() {
    cudaFree(table);
    cudaFree(bloom);
    cudaFree(bloom2);
    *this = BuiltTable();
  }
};

static uint32_t blocks_for_bits(unsigned __int128 bits) {
  unsigned __int128 nb = (bits + 63) / 64;
  if (nb < 16) nb = 16;
  if (nb > 0xffffffffull) nb = 0xffffffffull;
  return (uint32_t)nb;
}

// Table + Bloom filter(s) over a resident set.  Filter policy (measured on B200): a
// Bloom filter probed at random is L2-resident up to ~48 MiB; beyond that every probe is a 64-byte
// DRAM access.  So the filter the enumeration loop tests is capped (default 40 MiB); when the cap
// leaves fewer than 8 bits per key it switches to a 1+1-bit geometry and a second, full-size
// filter in HBM is tested only by the first level's survivors.
static int build_table_for(cb_ctx* c, const cb_dset* s, bool with_bloom, BuiltTable* out) {
  BuiltTable t;
  t.slots = 8;
  while (t.slots * c->cfg.table_load_pct < s->n * 100) t.slots <<= 1;
  CU(c, cudaMalloc(&t.table, t.slots * sizeof(Slot)));
  const unsigned __int128 want_bits = (unsigned __int128)s->n * c->cfg.bloom_bits_per_key_x16 / 16;
  const uint64_t cap_bytes = (uint64_t)c->cfg.bloom_l2_cap_kib << 10;
  cudaError_t e = cudaSuccess;
  if (!with_bloom) {
    t.blocks = 16;  // the build kernel always sets a filter; give it a scratch one
  } else if ((uint64_t)((want_bits + 7) / 8) <= cap_bytes) {
    t.blocks = blocks_for_bits(want_bits);
  } else {
    t.blocks = (uint32_t)(cap_bytes / 8);
    t.k2 = (double)cap_bytes * 8.0 / (double)s->n < 8.0;
    const unsigned __int128 bits2 = t.k2 ? (unsigned __int128)s->n * 12 : want_bits;
    t.blocks2 = blocks_for_bits(bits2);
    e = cudaMalloc(&t.bloom2, (size_t)t.blocks2 * 8);
    if (e == cudaSuccess) e = cudaMemsetAsync(t.bloom2, 0, (size_t)t.blocks2 * 8, c->stream);
  }
  if (e == cudaSuccess) e = cudaMalloc(&t.bloom, (size_t)t.blocks * 8);
  if (e == cudaSuccess) e = cudaMemsetAsync(t.bloom, 0, (size_t)t.blocks * 8, c->stream);
  if (e == cudaSuccess) {
    launch_table_clear(t.table, t.slots, c->stream);
    launch_build(s->d_hash, s->n, t.table, t.slots - 1, t.bloom, t.blocks, t.k2, t.bloom2, t.blocks2, c->stream);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    t.release();
    return fail(c, e == cudaErrorMemoryAllocation ? CB_ERR_NOMEM : CB_ERR_CUDA, "table/Bloom build: %s",
                cudaGetErrorString(e));
  }
  *out = t;
  return CB_OK;
}

extern "C" int cb_build_b(cb_ctx* c, cb_dset* b) {
  if (!c || !b) return fail(c, CB_ERR_INVALID, "cb_build_b: NULL argument");
  int rc = bind(c);
  if (rc) return rc;
  if (c->b_owned && c->b != b) free_dset(c->b);
  c->b = b;
  c->b_owned = false;
  cudaFree(c->d_table);
  cudaFree(c->d_bloom);
  cudaFree(c->d_bloom2);
  c->d_table = nullptr;
  c->d_bloom = c->d_bloom2 = nullptr;
  c->slots = 0;
  c->bloom_blocks = c->bloom2_blocks = 0;
  c->dups_b = 0;
  c->stats.ms_hash_b = c->stats.ms_hash_a;
  c->stats.ms_hash_a = 0;
  c->stats.ms_build_b = c->stats.ms_dups_b = 0;
  c->stats.kernel_launches = 0;
  if (c->cfg.differences > MAXDIFF_HASH) return CB_OK;  // brute-force path: no table, no dup check
  CU(c, cudaEventRecord(c->ev[0], c->stream));
  BuiltTable bt;
  rc = build_table_for(c, b, true, &bt);
  if (rc) return rc;
  c->d_table = bt.table;
  c->slots = bt.slots;
  c->d_bloom = bt.bloom;
  c->bloom_blocks = bt.blocks;
  c->bloom_k2 = bt.k2;
  c->d_bloom2 = bt.bloom2;
  c->bloom2_blocks = bt.blocks2;
  CU(c, cudaEventRecord(c->ev[1], c->stream));
  rc = zero_counter(c, CTR_DUPS);
  if (rc) return rc;
  launch_count_dups(cb_view_of(b), c->d_table, c->slots - 1, c->cfg.ignore_genes != 0, c->d_counters,
                    c->stream);
  CU(c, cudaGetLastError());
  CU(c, cudaEventRecord(c->ev[2], c->stream));
  rc = read_counters(c);
  if (rc) return rc;
  c->dups_b = c->h_counters[CTR_DUPS];
  cudaEventElapsedTime(&c->stats.ms_build_b, c->ev[0], c->ev[1]);
  cudaEventElapsedTime(&c->stats.ms_dups_b, c->ev[1], c->ev[2]);
  c->stats.table_slots = c->slots;
  c->stats.bloom_bytes = (uint64_t)c->bloom_blocks * 8;
  c->stats.bloom2_bytes = (uint64_t)c->bloom2_blocks * 8;
  c->stats.kernel_launches = b->n ? 3 : 1;
  return CB_OK;
}

extern "C" uint64_t cb_dups_b(const cb_ctx* c) { return c ? c->dups_b : 0; }

extern "C" int cb_count_dups(cb_ctx* c, const cb_dset* s, uint64_t* out) {
  if (!c || !s || !out) return fail(c, CB_ERR_INVALID, "cb_count_dups: NULL argument");
  int rc = bind(c);
  if (rc) return rc;
  *out = 0;
  if (s->n == 0) return CB_OK;
  BuiltTable bt;
  rc = build_table_for(c, s, false, &bt);
  if (rc) return rc;
  rc = zero_counter(c, CTR_DUPS);
  if (!rc) {
    launch_count_dups(cb_view_of(s), bt.table, bt.slots - 1, c->cfg.ignore_genes != 0, c->d_counters,
                      c->stream);
    rc = read_counters(c);
  }
  bt.release();
  if (rc) return rc;
  *out = c->h_counters[CTR_DUPS];
  return CB_OK;
}

// ---- set A -------------------------------------------------------------------------------------

static int ensure_matrix(cb_ctx* c, uint64_t rows, uint64_t cols, bool reset) {
  if (c->cfg.no_matrix) return CB_OK;
  if (c->matrix_external) {
    if (c->rows != rows || c->cols != cols)
      return fail(c, CB_ERR_INVALID, "bound matrix is %llu x %llu, this run needs %llu x %llu",
                  (unsigned long long)c->rows, (unsigned long long)c->cols, (unsigned long long)rows,
                  (unsigned long long)cols);
    return CB_OK;
  }
  if (c->d_matrix && c->rows == rows && c->cols == cols && !reset) return CB_OK;
  if (!c->d_matrix || c->rows * c->cols < rows * cols) {
    cudaFree(c->d_matrix);
    c->d_matrix = nullptr;
    CU(c, cudaMalloc(&c->d_matrix, std::max<uint64_t>(rows * cols, 1) * sizeof(double)));
  }
  c->rows = rows;
  c->cols = cols;
  CU(c, cudaMemsetAsync(c->d_matrix, 0, std::max<uint64_t>(rows * cols, 1) * sizeof(double), c->stream));
  return CB_OK;
}

static int ensure_pairs(cb_ctx* c, uint64_t cap) {
  if (c->pairs_cap >= cap) return CB_OK;
  cudaFree(c->d_pairs);
  c->d_pairs = nullptr;
  c->pairs_cap = 0;
  CU(c, cudaMalloc(&c->d_pairs, cap * sizeof(PairOut)));
  c->pairs_cap = cap;
  return CB_OK;
}

extern "C" int cb_run(cb_ctx* c, const cb_dset* a, uint64_t first, uint64_t count) {
  if (!c || !a) return fail(c, CB_ERR_INVALID, "cb_run: NULL argument");
  if (!c->b) return fail(c, CB_ERR_STATE, "cb_run: set B has not been built (cb_build_b / cb_set_b)");
  if (count >= 0xffffffffull)
    return fail(c, CB_ERR_LIMIT, "cb_run: at most 2^32-1 sequences per call; run the set in chunks");
  if (first > a->n || count > a->n - first)
    return fail(c, CB_ERR_INVALID, "cb_run: range [%llu, +%llu) outside the set (%llu sequences)",
                (unsigned long long)first, (unsigned long long)count, (unsigned long long)a->n);
  int rc = bind(c);
  if (rc) return rc;
  const bool existence = c->cfg.mode == CB_MODE_EXISTENCE;
  const uint64_t cols = c->b->n_reps;
  if (existence)
    rc = ensure_matrix(c, count, cols, true);
  else
    rc = ensure_matrix(c, c->cfg.n_reps_a, cols, false);
  if (rc) return rc;
  if (!existence && a->n_reps > c->cfg.n_reps_a)
    return fail(c, CB_ERR_INVALID, "cb_run: set A has %u repertoires, config says %u", a->n_reps,
                c->cfg.n_reps_a);
  if (c->cfg.want_pairs) {
    rc = ensure_pairs(c, c->cfg.pairs_capacity);
    if (rc) return rc;
  }
  cb_stats& S = c->stats;
  S.seeds = count;
  S.probes = S.bloom_pass = S.matches = S.pairs = 0;
  S.ms_probe = S.ms_total_run = 0;
  S.kernel_launches = 0;
  if (count == 0 || c->b->n == 0) return CB_OK;

  const bool hash_path = c->cfg.differences <= MAXDIFF_HASH;
  CU(c, cudaMemsetAsync(c->d_counters, 0, CTR_COUNT * sizeof(unsigned long long), c->stream));
  if (hash_path)  // bookkeeping for the probes/s metric, outside the timed span
    launch_count_probes(cb_view_of(a), first, count, (uint32_t)c->cfg.alphabet_size,
                        c->cfg.differences, c->cfg.indels != 0, c->d_counters, c->stream);
  CU(c, cudaEventRecord(c->ev[3], c->stream));
  int launches = 0;
  ProbeParams p{};
  if (hash_path) {
    rc = ensure_ztab(c, a->longest + 2);
    if (rc) return rc;
    p.a = cb_view_of(a);
    p.b = cb_view_of(c->b);
    p.a_first = first;
    p.a_count = count;
    p.table = c->d_table;
    p.table_mask = c->slots - 1;
    p.bloom = c->d_bloom;
    p.bloom_blocks = c->bloom_blocks;
    p.bloom_k2 = c->bloom_k2;
    p.bloom2 = c->d_bloom2;
    p.bloom2_blocks = c->bloom2_blocks;
    p.ztab = c->d_ztab;
    p.zrows = a->longest + 2;
    p.sigma = (uint32_t)c->cfg.alphabet_size;
    p.seed = c->cfg.seed;
    p.matrix = c->d_matrix;
    p.n_cols = cols;
    p.pairs = c->d_pairs;
    p.pairs_cap = c->pairs_cap;
    p.counters = c->d_counters;
    p.lmax = a->longest;
    // one matrix row in shared memory per CTA (d=1 kernel) when it is small enough
    p.tile_cols = (!existence && !c->cfg.no_matrix && !(c->cfg.flags & CB_FLAG_NO_SMEM_TILE) && cols <= 4096)
                      ? (uint32_t)cols : 0;
    p.score = c->cfg.score;
    p.ignore_counts = c->cfg.ignore_counts != 0;
    p.ignore_genes = c->cfg.ignore_genes != 0;
    p.existence = existence;
    p.no_matrix = c->cfg.no_matrix != 0;
    p.want_pairs = c->cfg.want_pairs != 0;
    p.use_bloom = !(c->cfg.flags & CB_FLAG_NO_BLOOM);
    p.count_bloom = 1;
    p.differences = c->cfg.differences;
    p.indels = c->cfg.indels != 0;
    // d = 2: split each seed's outer (position, residue) space over several warps when there
    // are too few seeds to fill the machine
    p.split = 1;
    if (c->cfg.differences == 2) {
      const uint64_t want_items = (uint64_t)c->sm_count * 64 * 4;
      while (p.split < 64 && count * p.split < want_items) p.split <<= 1;
    }
    const char* kerr = nullptr;
    launches = launch_probe(p, c->sm_count, c->stream, &kerr);
    if (launches < 0) return fail(c, CB_ERR_LIMIT, "cb_run: %s", kerr ? kerr : "launch failed");
    CU(c, cudaGetLastError());
  } else {
    rc = cb_run_brute(c, a, first, count, false, &launches);
    if (rc) return rc;
  }
  CU(c, cudaEventRecord(c->ev[4], c->stream));
  rc = read_counters(c);
  if (rc) return rc;
  cudaEventElapsedTime(&S.ms_probe, c->ev[3], c->ev[4]);
  S.matches = c->h_counters[CTR_MATCHES];
  S.probes = c->h_counters[CTR_PROBES];
  S.bloom_pass = c->h_counters[CTR_BLOOM_PASS];
  S.kernel_launches = (uint32_t)launches;
  uint64_t np = c->h_counters[CTR_PAIRS];

  if (c->cfg.want_pairs && np > c->pairs_cap) {
    // Pair buffer overflowed.  The matrix and the counters are complete; redo this range for the
    // pairs alone, now that the exact count is known.
    rc = ensure_pairs(c, np);
    if (rc) return rc;
    CU(c, cudaMemsetAsync(c->d_counters, 0, 4 * sizeof(unsigned long long), c->stream));
    if (hash_path) {
      p.pairs = c->d_pairs;
      p.pairs_cap = c->pairs_cap;
      p.no_matrix = 1;
      p.count_bloom = 0;
      const char* kerr = nullptr;
      int l2 = launch_probe(p, c->sm_count, c->stream, &kerr);
      if (l2 < 0) return fail(c, CB_ERR_LIMIT, "cb_run: %s", kerr ? kerr : "launch failed");
      launches += l2;
    } else {
      int l2 = 0;
      rc = cb_run_brute(c, a, first, count, true, &l2);
      if (rc) return rc;
      launches += l2;
    }
    CU(c, cudaGetLastError());
    rc = read_counters(c);
    if (rc) return rc;
    if (c->h_counters[CTR_PAIRS] != np)
      return fail(c, CB_ERR_CUDA, "cb_run: pair count changed between passes (%llu vs %llu)",
                  (unsigned long long)np, (unsigned long long)c->h_counters[CTR_PAIRS]);
    S.kernel_launches = (uint32_t)launches;
  }
  if (c->cfg.want_pairs && np) {
    const size_t old = c->pending.size();
    try {
      c->pending.resize(old + np);
    } catch (...) {
      return fail(c, CB_ERR_NOMEM, "cb_run: out of host memory for %llu pairs", (unsigned long long)np);
    }
    static_assert(sizeof(cb_pair) == sizeof(PairOut), "pair layout");
    CU(c, cudaMemcpyAsync(c->pending.data() + old, c->d_pairs, np * sizeof(PairOut),
                          cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
  }
  S.pairs = c->cfg.want_pairs ? np : 0;
  CU(c, cudaEventRecord(c->ev[5], c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  cudaEventElapsedTime(&S.ms_total_run, c->ev[3], c->ev[5]);
  return CB_OK;
}

extern "C" int cb_set_b(cb_ctx* c, const cb_set* b) {
  cb_dset* d = nullptr;
  int rc = cb_upload(c, b, &d);
  if (rc) return rc;
  rc = cb_build_b(c, d);
  if (rc) {
    free_dset(d);
    if (c->b == d) c->b = nullptr;
    return rc;
  }
  c->b_owned = true;
  return CB_OK;
}

extern "C" int cb_run_a(cb_ctx* c, const cb_set* a) {
  cb_dset* d = nullptr;
  int rc = cb_upload(c, a, &d);
  if (rc) return rc;
  const float ms_hash = c->stats.ms_hash_a;
  rc = cb_run(c, d, 0, d->n);
  c->stats.ms_hash_a = ms_hash;
  cudaStreamSynchronize(c->stream);
  free_dset(d);
  return rc;
}

// ---- results -----------------------------------------------------------------------------------

extern "C" int cb_matrix_dims(const cb_ctx* c, uint64_t* rows, uint64_t* cols) {
  if (!c) return CB_ERR_INVALID;
  if (rows) *rows = c->rows;
  if (cols) *cols = c->cols;
  return CB_OK;
}

extern "C" int cb_get_matrix(cb_ctx* c, double* out, size_t n_values) {
  if (!c || !out) return fail(c, CB_ERR_INVALID, "cb_get_matrix: NULL argument");
  if (c->cfg.no_matrix) return fail(c, CB_ERR_STATE, "cb_get_matrix: context was created with no_matrix");
  if (!c->d_matrix) return fail(c, CB_ERR_STATE, "cb_get_matrix: no matrix yet (call cb_run first)");
  if (n_values != c->rows * c->cols)
    return fail(c, CB_ERR_INVALID, "cb_get_matrix: expected %llu values, caller gave %llu",
                (unsigned long long)(c->rows * c->cols), (unsigned long long)n_values);
  int rc = bind(c);
  if (rc) return rc;
  if (n_values == 0) return CB_OK;
  CU(c, cudaMemcpyAsync(out, c->d_matrix, n_values * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return CB_OK;
}

extern "C" int cb_set_matrix(cb_ctx* c, const double* in, size_t n_values) {
  if (!c || !in) return fail(c, CB_ERR_INVALID, "cb_set_matrix: NULL argument");
  if (!c->d_matrix || n_values != c->rows * c->cols)
    return fail(c, CB_ERR_INVALID, "cb_set_matrix: size mismatch");
  int rc = bind(c);
  if (rc) return rc;
  CU(c, cudaMemcpyAsync(c->d_matrix, in, n_values * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return CB_OK;
}

extern "C" int cb_bind_matrix(cb_ctx* c, void* device_ptr, uint64_t rows, uint64_t cols) {
  if (!c) return CB_ERR_INVALID;
  if (c->cfg.mode != CB_MODE_MATRIX || c->cfg.no_matrix)
    return fail(c, CB_ERR_STATE, "cb_bind_matrix: only in matrix mode with a matrix");
  int rc = bind(c);
  if (rc) return rc;
  CU(c, cudaStreamSynchronize(c->stream));
  if (!c->matrix_external) cudaFree(c->d_matrix);
  c->d_matrix = nullptr;
  c->rows = c->cols = 0;
  c->matrix_external = false;
  if (!device_ptr) return CB_OK;
  if (rows != c->cfg.n_reps_a) return fail(c, CB_ERR_INVALID, "cb_bind_matrix: rows must equal n_reps_a");
  c->d_matrix = (double*)device_ptr;
  c->rows = rows;
  c->cols = cols;
  c->matrix_external = true;
  return CB_OK;
}

extern "C" int cb_clear_matrix(cb_ctx* c) {
  if (!c) return CB_ERR_INVALID;
  int rc = bind(c);
  if (rc) return rc;
  if (c->d_matrix && c->rows * c->cols)
    CU(c, cudaMemsetAsync(c->d_matrix, 0, c->rows * c->cols * sizeof(double), c->stream));
  return CB_OK;
}

extern "C" void* cb_matrix_device(cb_ctx* c) {
  if (!c) return nullptr;
  if (!c->d_matrix && c->cfg.mode == CB_MODE_MATRIX && c->b && !c->cfg.no_matrix) {
    if (bind(c) || ensure_matrix(c, c->cfg.n_reps_a, c->b->n_reps, false)) return nullptr;
    cudaStreamSynchronize(c->stream);
  }
  return c->d_matrix;
}

extern "C" int cb_pairs_pending(const cb_ctx* c, uint64_t* n) {
  if (!c || !n) return CB_ERR_INVALID;
  *n = c->pending.size();
  return CB_OK;
}

extern "C" int cb_drain_pairs(cb_ctx* c, cb_pair* buf, size_t cap, size_t* n_out) {
  if (!c || !n_out || (cap && !buf)) return fail(c, CB_ERR_INVALID, "cb_drain_pairs: NULL argument");
  const size_t have = c->pending.size();
  const size_t take = std::min(cap, have);
  // hand out from the tail so the vector shrinks without moving the remainder
  if (take) memcpy(buf, c->pending.data() + (have - take), take * sizeof(cb_pair));
  c->pending.resize(have - take);
  if (c->pending.empty()) std::vector<cb_pair>().swap(c->pending);
  *n_out = take;
  return CB_OK;
}

extern "C" int cb_get_stats(const cb_ctx* c, cb_stats* out) {
  if (!c || !out) return CB_ERR_INVALID;
  *out = c->stats;
  return CB_OK;
}

extern "C" uint64_t cb_probe_count(const uint8_t* residues, uint32_t len, int alphabet_size,
                                   int differences, int indels) {
  if (differences > MAXDIFF_HASH) return 0;
  return probe_count(residues, len, (uint32_t)alphabet_size, differences, indels != 0);
}
