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

#include <cub/device/device_radix_sort.cuh>

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
#define ensure_ztab cb_ensure_ztab

static thread_local cudaStream_t g_alloc_stream = nullptr;
static thread_local cudaMemPool_t g_alloc_pool = nullptr;

cudaError_t cb_dmalloc_raw(void** p, size_t bytes) {
  *p = nullptr;
  if (!g_alloc_pool) return cudaErrorInvalidValue;  // no context bound to this thread
  return cudaMallocFromPoolAsync(p, bytes ? bytes : 1, g_alloc_pool, g_alloc_stream);
}

cudaError_t cb_dfree(void* p) { return p ? cudaFreeAsync(p, g_alloc_stream) : cudaSuccess; }

int cb_bind_device(cb_ctx* c) {
  CU(c, cudaSetDevice(c->device));
  g_alloc_stream = c->stream;
  g_alloc_pool = c->pool;
  return CB_OK;
}
#define bind cb_bind_device

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
int cb_ensure_ztab(cb_ctx* c, uint32_t rows) {
  if (rows <= c->zrows) return CB_OK;
  rows = (rows + 15) & ~15u;
  const uint32_t sigma = (uint32_t)c->cfg.alphabet_size;
  std::vector<uint64_t> h((size_t)rows * sigma);
  for (uint32_t p = 0; p < rows; p++)
    for (uint32_t r = 0; r < sigma; r++) h[(size_t)p * sigma + r] = zobrist_gen(c->cfg.seed, p, r);
  uint64_t* d = nullptr;
  CU(c, cb_dmalloc(&d, h.size() * sizeof(uint64_t)));
  cudaError_t e = cudaMemcpyAsync(d, h.data(), h.size() * sizeof(uint64_t),
                                  cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) {
    cb_dfree(d);
    return fail(c, CB_ERR_CUDA, "Zobrist table upload: %s", cudaGetErrorString(e));
  }
  if (c->d_ztab) cb_dfree(c->d_ztab);
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
  // bits per key in EACH of the four class filters.  16: 0.6 % of the candidates reach the table
  // stage (0.33 % at 24, for 1.5x the memory and the same four random updates per key)
  if (c->cfg.bloom_bits_per_key_x16 == 0) c->cfg.bloom_bits_per_key_x16 = 16 * 16;
  if (c->cfg.table_load_pct == 0 || c->cfg.table_load_pct > 90) c->cfg.table_load_pct = 50;
  if (c->cfg.pairs_capacity == 0) c->cfg.pairs_capacity = 1ull << 24;
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

  {
    // the engine's own stream-ordered pool (the device's default pool and its attributes belong
    // to the host application); cached blocks are never handed back between calls
    cudaMemPoolProps props{};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = c->device;
    CU_CREATE(cudaMemPoolCreate(&c->pool, &props));
    uint64_t keep = ~0ull;
    CU_CREATE(cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  // (cudaLimitMaxL2FetchGranularity 32 / 64 / 128 was tried for the random 8-byte filter reads, which
  // cost ~110 B of DRAM traffic each: no change in any kernel's time on B200.)
  CU_CREATE(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  CU_CREATE(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  CU_CREATE(cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  g_alloc_stream = c->stream;
  g_alloc_pool = c->pool;
  for (auto& ev : c->ev) CU_CREATE(cudaEventCreate(&ev));
  CU_CREATE(cb_dmalloc(&c->d_counters, CTR_COUNT * sizeof(unsigned long long)));
  CU_CREATE(cudaMemset(c->d_counters, 0, CTR_COUNT * sizeof(unsigned long long)));
  CU_CREATE(cudaMallocHost(&c->h_counters, CTR_COUNT * sizeof(unsigned long long)));
#undef CU_CREATE
  *out = c;
  return CB_OK;
}

void cb_free_dset(cb_dset* s) {
  if (!s) return;
  cb_dfree(s->d_meta);
  cb_dfree(s->d_res);
  cb_dfree(s->d_hash);
  cb_dfree(s->d_order);
  cb_dfree(s->d_packed);
  delete s;
}

extern "C" void cb_destroy(cb_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  g_alloc_stream = c->stream ? c->stream : c->own_stream;
  g_alloc_pool = c->pool;
  if (c->own_stream) cudaStreamSynchronize(c->own_stream);
  if (c->stream && c->stream != c->own_stream) cudaStreamSynchronize(c->stream);
  cb_comm_release(c);
  if (c->b_owned) cb_free_dset(c->b);
  cb_dfree(c->d_ztab);
  cb_dfree(c->d_counters);
  if (c->h_counters) cudaFreeHost(c->h_counters);
  for (auto& q : c->pin)
    if (q) cudaFreeHost(q);
  cb_dfree(c->d_table);
  cb_dfree(c->d_bloom);
  if (!c->matrix_external) cb_dfree(c->d_matrix);
  cb_dfree(c->d_pairs);
  cb_dfree(c->d_gq_hv);
  cb_dfree(c->d_gq_vs);
  cb_dfree(c->d_overflow);
  for (auto& ev : c->ev)
    if (ev) cudaEventDestroy(ev);
  if (g_alloc_stream) cudaStreamSynchronize(g_alloc_stream);
  if (c->pool) cudaMemPoolDestroy(c->pool);
  g_alloc_stream = nullptr;
  g_alloc_pool = nullptr;
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->side_stream) cudaStreamDestroy(c->side_stream);
  delete c;
}

extern "C" const char* cb_last_error(const cb_ctx* c) { return c ? c->err.c_str() : g_error.c_str(); }

extern "C" int cb_set_stream(cb_ctx* c, void* s) {
  if (!c) return CB_ERR_INVALID;
  cudaStreamSynchronize(c->stream);
  c->stream = s ? (cudaStream_t)s : c->own_stream;
  g_alloc_stream = c->stream;
  g_alloc_pool = c->pool;
  return CB_OK;
}

// ---- upload + hash: see upload.cu ----------------------------------------------------------------

extern "C" void cb_free_set(cb_ctx* c, cb_dset* s) {
  if (!s) return;
  if (c) {
    cb_bind_device(c);
    cudaStreamSynchronize(c->stream);
    if (c->b == s) {
      c->b = nullptr;
      c->b_owned = false;
    }
  }
  cb_free_dset(s);
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

static uint32_t blocks_for_bits(unsigned __int128 bits) {
  unsigned __int128 nb = (bits + 63) / 64;
  if (nb < 16) nb = 16;
  if (nb > 0xffffffffull) nb = 0xffffffffull;
  return (uint32_t)nb;
}

// Table + the four class filters (common.cuh) sized for n keys: bloom_bits_per_key bits per key in
// EACH filter.  Their size does not have to fit L2: a filter word is fetched once per slot, not once
// per candidate, so the filters are read a few dozen sectors per seed.
int cb_table_alloc(cb_ctx* c, uint64_t n, bool with_bloom, BuiltTable* out, bool clear_table) {
  BuiltTable t;
  t.slots = 8;
  while (t.slots * c->cfg.table_load_pct < n * 100) t.slots <<= 1;
  const unsigned __int128 want_bits = (unsigned __int128)n * c->cfg.bloom_bits_per_key_x16 / 16;
  t.blocks = with_bloom ? blocks_for_bits(want_bits) : 0;  // no filters: a table for duplicate counting / -z only
  cudaError_t e = cudaSuccess;
  if (with_bloom && c->d_table && c->slots == t.slots && c->bloom_blocks == t.blocks) {
    // rebuilding a set-B structure of the same geometry: clear and refill the buffers in place
    // instead of growing the memory pool by another table
    t.table = c->d_table;
    t.bloom = c->d_bloom;
    c->d_table = nullptr;
    c->d_bloom = nullptr;
    c->slots = 0;
    c->bloom_blocks = 0;
  } else {
    e = cb_dmalloc(&t.table, t.slots * sizeof(Slot));
    if (e == cudaSuccess && t.blocks) e = cb_dmalloc(&t.bloom, (size_t)t.blocks * 8 * CB_CLASSES);
  }
  if (e == cudaSuccess && t.blocks) e = cudaMemsetAsync(t.bloom, 0, (size_t)t.blocks * 8 * CB_CLASSES, c->stream);
  if (e == cudaSuccess && clear_table) {
    launch_table_clear(t.table, t.slots, c->stream);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    t.release();
    return fail(c, e == cudaErrorMemoryAllocation ? CB_ERR_NOMEM : CB_ERR_CUDA, "table/Bloom allocation: %s",
                cudaGetErrorString(e));
  }
  *out = t;
  return CB_OK;
}

// Insert sequences [first, first + n) (hashes at d_hash[first..]) into the table and filter(s).
//
// A table much larger than L2 filled in input order is bound by DRAM row activations: every CAS,
// every slot store and every filter update is a random 32-byte sector (measured on B200 at 10^8
// keys / 4.3 GB: 21 G CAS/s, 22 G stores/s, 49 G RED/s, together 11.4 ms; the same operations in
// address order 2.4 ms — tools/bench_atomics.cu).  So a batch that is a sizeable part of the
// table is first sorted by the hash bits that pick the home slot — 8-bit radix passes over
// (h * CB_HOME_MUL, index) pairs, 0.85 ms each at 10^8 — and then either
//   - a whole set into an empty table: built tile by tile in shared memory and streamed out, the
//     table written exactly once and never cleared or read (build_tile_kernel), or
//   - a batch into a table that already holds keys: the build kernel sweeps the table in address
//     order, keys sorted down to segments of 16 slots.
// The class filters are indexed by other functions of the hash: their own L2-blocked passes
// (launch_filters).  Small batches (the chunks of the upload pipeline, which hide behind the PCIe
// copy anyway) keep the direct path.
// Does a set of n keys in a table of `slots` slots take the sorted (partitioned) build?
static bool sorted_build(const cb_ctx* c, uint64_t slots, uint64_t n) {
  return !(c->cfg.flags & CB_FLAG_NO_PARTITION) && slots * sizeof(Slot) >= (256ull << 20) && n >= (1ull << 22) &&
         n * 64 >= slots && n < 0xffffffffull - 4096;  // 32-bit sorted positions, a CTA's stride of headroom
}
bool cb_tiled_build(const cb_ctx* c, uint64_t slots, uint64_t n) {
  return sorted_build(c, slots, n) && !(c->cfg.flags & CB_FLAG_NO_TILED_BUILD);
}

// whole: the table is EMPTY BUT NOT CLEARED (cb_table_alloc(..., clear_table = false)) and this batch
// is everything it will hold — the tiled build (kernels.cu build_tile_kernel) writes every slot once,
// keys sorted by their tile only (two radix passes at 10^8 instead of three); if its scratch buffers
// cannot be had the table is cleared here and filled the direct way.
void cb_table_insert(cb_ctx* c, const BuiltTable& t, cb_dset* s, uint64_t first, uint64_t n, bool whole) {
  s->links_dirty = true;
  uint64_t *key_in = nullptr, *part_hash = nullptr;
  uint32_t *iota = nullptr, *part_idx = nullptr, *tile_first = nullptr;
  void* temp = nullptr;
  c->insert_launches = n ? 1 : 0;
  bool filters_done = false, filters_forked = false;
  int tbits = 0;
  while ((1ull << tbits) < t.slots) tbits++;
  const bool tiled = whole && cb_tiled_build(c, t.slots, n) && !(t.bloom && (c->cfg.flags & CB_FLAG_FILTERS_IN_BUILD));
  if (sorted_build(c, t.slots, n)) {
    // tiled: down to tiles of BUILD_TILE_SLOTS; swept: down to segments of 16 slots, whole radix passes
    const int pbits = tiled ? tbits - BUILD_TILE_BITS : std::max(8, (tbits - 4) / 8 * 8);
    size_t temp_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, key_in, part_hash, iota, part_idx, n,
                                                    CB_PARTITION_TOP_BIT - pbits, CB_PARTITION_TOP_BIT, c->stream);
    if (e == cudaSuccess) e = cb_dmalloc(&key_in, n * sizeof(uint64_t));
    if (e == cudaSuccess) e = cb_dmalloc(&part_hash, n * sizeof(uint64_t));
    if (e == cudaSuccess) e = cb_dmalloc(&iota, n * sizeof(uint32_t));
    if (e == cudaSuccess) e = cb_dmalloc(&part_idx, n * sizeof(uint32_t));
    if (e == cudaSuccess) e = cb_dmalloc(&temp, temp_bytes ? temp_bytes : 1);
    if (e == cudaSuccess && tiled) e = cb_dmalloc(&tile_first, ((t.slots >> BUILD_TILE_BITS) + 1) * sizeof(uint32_t));
    if (e == cudaSuccess && t.bloom && !(c->cfg.flags & CB_FLAG_FILTERS_IN_BUILD)) {
      // the filters in their own L2-blocked passes, on the side stream: compute- and L2-bound, they
      // run beside the sort and the table build, which wait on DRAM
      // (COMPAIRR_B200_FILTER_OVERLAP=0: on the main stream, one after the other; measurements)
      static const bool beside = [] {
        const char* v = getenv("COMPAIRR_B200_FILTER_OVERLAP");
        return !(v && v[0] == '0');
      }();
      cudaStream_t fs = c->stream;
      if (beside && cudaEventRecord(c->ev[8], c->stream) == cudaSuccess &&
          cudaStreamWaitEvent(c->side_stream, c->ev[8], 0) == cudaSuccess)
        fs = c->side_stream;  // (not the copy stream: the upload pipeline's next chunk is on its way there)
      c->insert_launches += launch_filters(s->d_hash + first, n, t.bloom, t.blocks, c->sm_count, fs);
      if (fs != c->stream) {
        cudaEventRecord(c->ev[9], fs);
        filters_forked = true;
      }
      filters_done = true;
    }
    if (e == cudaSuccess) {
      launch_partition_keys(s->d_hash + first, n, key_in, iota, c->stream);  // h * CB_HOME_MUL: top bits = home slot
      e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, key_in, part_hash, iota, part_idx, n,
                                          CB_PARTITION_TOP_BIT - pbits, CB_PARTITION_TOP_BIT, c->stream);
    }
    if (e == cudaSuccess) c->insert_launches += 1 + 2 + (pbits + 7) / 8;  // keys, histogram, scan, one sweep per digit
    if (e != cudaSuccess) {  // no memory for the sort buffers: the direct path still works
      (void)cudaGetLastError();
      cb_dfree(part_hash);
      cb_dfree(part_idx);
      part_hash = nullptr;
      part_idx = nullptr;
    }
  }
  // a sorted (large) batch: the build kernels write the table only
  unsigned long long* bloom_in_build = filters_done ? nullptr : t.bloom;  // nullptr also for a table without filters
  if (tiled && part_hash && !bloom_in_build) {
    // the sort's inputs are dead: its keys hold the keys set aside tile by tile, its values the spill list
    cudaMemsetAsync(c->d_counters + CTR_SPILL, 0, sizeof(unsigned long long), c->stream);
    c->insert_launches += launch_build_tiled(s->d_meta, s->d_res, part_hash, part_idx, first, n, c->cfg.ignore_genes != 0,
                                             t.table, (uint32_t)tbits, tile_first, reinterpret_cast<uint32_t*>(key_in), iota,
                                             c->d_counters, c->sm_count, c->stream) - 1;
  } else {
    if (whole) {
      launch_table_clear(t.table, t.slots, c->stream);
      c->insert_launches++;
    }
    launch_build(s->d_meta, s->d_res, s->d_hash, part_hash, part_idx, first, n, c->cfg.ignore_genes != 0, t.table,
                 t.slots - 1, bloom_in_build, t.blocks, c->stream);
  }
  if (filters_forked) cudaStreamWaitEvent(c->stream, c->ev[9], 0);
  cb_dfree(key_in);
  cb_dfree(part_hash);
  cb_dfree(iota);
  cb_dfree(part_idx);
  cb_dfree(tile_first);
  cb_dfree(temp);
}

static int build_table_for(cb_ctx* c, cb_dset* s, bool with_bloom, BuiltTable* out) {
  int rc = cb_table_alloc(c, s->n, with_bloom, out, false);  // cb_table_insert(whole) clears it if it has to
  if (rc) return rc;
  if (s->links_dirty) launch_reset_next(s->d_meta, s->n, c->stream);  // the set has been inserted before
  cb_table_insert(c, *out, s, 0, s->n, true);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    out->release();
    return fail(c, CB_ERR_CUDA, "table build launch: %s", cudaGetErrorString(e));
  }
  return CB_OK;
}

// Make t the context's set-B structure and count the exact duplicates (dup2).
int cb_adopt_table(cb_ctx* c, cb_dset* b, BuiltTable& t, bool owned) {
  if (c->b_owned && c->b != b) cb_free_dset(c->b);
  c->b = b;
  c->b_owned = owned;
  cb_dfree(c->d_table);
  cb_dfree(c->d_bloom);
  c->d_table = t.table;
  c->slots = t.slots;
  c->d_bloom = t.bloom;
  c->bloom_blocks = t.blocks;
  t = BuiltTable();
  c->dups_b = 0;
  c->stats.ms_dups_b = 0;
  if (!c->d_table) return CB_OK;
  CU(c, cudaEventRecord(c->ev[1], c->stream));
  int rc = zero_counter(c, CTR_DUPS);
  if (rc) return rc;
  launch_count_dups(cb_view_of(b), c->d_counters, c->stream);
  CU(c, cudaGetLastError());
  CU(c, cudaEventRecord(c->ev[2], c->stream));
  rc = read_counters(c);
  if (rc) return rc;
  c->dups_b = c->h_counters[CTR_DUPS];
  cudaEventElapsedTime(&c->stats.ms_dups_b, c->ev[1], c->ev[2]);
  c->stats.table_slots = c->slots;
  c->stats.bloom_bytes = (uint64_t)c->bloom_blocks * 8;                     // one class filter
  c->stats.bloom2_bytes = (uint64_t)c->bloom_blocks * 8 * (CB_CLASSES - 1);  // the other three
  return CB_OK;
}

extern "C" int cb_build_b(cb_ctx* c, cb_dset* b) {
  if (!c || !b) return fail(c, CB_ERR_INVALID, "cb_build_b: NULL argument");
  int rc = bind(c);
  if (rc) return rc;
  c->stats.ms_hash_b = c->stats.ms_hash_a;
  c->stats.ms_hash_a = 0;
  c->stats.ms_build_b = c->stats.ms_dups_b = 0;
  c->stats.kernel_launches = 0;
  BuiltTable bt;
  const bool reset_links = b->links_dirty;
  if (c->cfg.differences <= MAXDIFF_HASH) {  // the brute-force path has no table and no dup check
    CU(c, cudaEventRecord(c->ev[0], c->stream));
    rc = build_table_for(c, b, true, &bt);
    if (rc) return rc;
    CU(c, cudaEventRecord(c->ev[6], c->stream));
  }
  rc = cb_adopt_table(c, b, bt, c->b == b ? c->b_owned : false);
  if (rc) return rc;
  if (c->d_table) {
    cudaEventElapsedTime(&c->stats.ms_build_b, c->ev[0], c->ev[6]);
    c->stats.kernel_launches = b->n ? 1 + (reset_links ? 1 : 0) + c->insert_launches : 1;  // [reset links,] [clear,] [sort, filters,] insert, duplicates
  }
  return CB_OK;
}

extern "C" uint64_t cb_dups_b(const cb_ctx* c) { return c ? c->dups_b : 0; }

extern "C" cb_dset* cb_resident_b(cb_ctx* c) { return c ? c->b : nullptr; }

extern "C" int cb_count_dups(cb_ctx* c, cb_dset* s, uint64_t* out) {
  if (!c || !s || !out) return fail(c, CB_ERR_INVALID, "cb_count_dups: NULL argument");
  int rc = bind(c);
  if (rc) return rc;
  *out = 0;
  if (s->n == 0) return CB_OK;
  if (s == c->b && c->d_table) {  // its occurrence lists are the live set-B structure: already counted
    *out = c->dups_b;
    return CB_OK;
  }
  if (s->n >= 0xffffffffull) return fail(c, CB_ERR_LIMIT, "more than 2^32-1 sequences in one set");
  BuiltTable bt;
  rc = build_table_for(c, s, false, &bt);
  if (rc) return rc;
  rc = zero_counter(c, CTR_DUPS);
  if (!rc) {
    launch_count_dups(cb_view_of(s), c->d_counters, c->stream);
    rc = read_counters(c);
  }
  bt.release();
  if (rc) return rc;
  *out = c->h_counters[CTR_DUPS];
  return CB_OK;
}

extern "C" int cb_dedup(cb_ctx* c, cb_dset* s, uint32_t* leader_out, uint64_t* count_out, uint64_t* merged_out) {
  if (!c || !s || !leader_out || !count_out) return fail(c, CB_ERR_INVALID, "cb_dedup: NULL argument");
  int rc = bind(c);
  if (rc) return rc;
  if (merged_out) *merged_out = 0;
  if (s->n == 0) return CB_OK;
  if (s->n >= 0xffffffffull) return fail(c, CB_ERR_LIMIT, "more than 2^32-1 sequences in one set");
  const bool live = s == c->b && c->d_table;  // its occurrence lists are already built
  BuiltTable bt;
  if (!live) {
    rc = build_table_for(c, s, false, &bt);
    if (rc) return rc;
  }
  uint32_t* d_lead = nullptr;
  unsigned long long* d_sum = nullptr;
  cudaError_t e = cb_dmalloc(&d_lead, s->n * sizeof(uint32_t));
  if (e == cudaSuccess) e = cb_dmalloc(&d_sum, s->n * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemsetAsync(d_sum, 0, s->n * sizeof(unsigned long long), c->stream);
  if (e == cudaSuccess) {
    rc = zero_counter(c, CTR_DUPS);
    if (!rc) {
      launch_dedup(cb_view_of(s), c->cfg.ignore_counts != 0, d_lead, d_sum, c->d_counters, c->stream);
      e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaMemcpyAsync(leader_out, d_lead, s->n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream);
      if (e == cudaSuccess) e = cudaMemcpyAsync(count_out, d_sum, s->n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream);
      if (e == cudaSuccess) rc = read_counters(c);  // synchronises the stream
    }
  }
  cb_dfree(d_lead);
  cb_dfree(d_sum);
  bt.release();
  if (e != cudaSuccess)
    return fail(c, e == cudaErrorMemoryAllocation ? CB_ERR_NOMEM : CB_ERR_CUDA, "cb_dedup: %s", cudaGetErrorString(e));
  if (rc) return rc;
  if (merged_out) *merged_out = c->h_counters[CTR_DUPS];
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
    cb_dfree(c->d_matrix);
    c->d_matrix = nullptr;
    CU(c, cb_dmalloc(&c->d_matrix, std::max<uint64_t>(rows * cols, 1) * sizeof(double)));
  }
  c->rows = rows;
  c->cols = cols;
  CU(c, cudaMemsetAsync(c->d_matrix, 0, std::max<uint64_t>(rows * cols, 1) * sizeof(double), c->stream));
  return CB_OK;
}

static int ensure_pairs(cb_ctx* c, uint64_t cap) {
  if (c->pairs_cap >= cap) return CB_OK;
  cb_dfree(c->d_pairs);
  c->d_pairs = nullptr;
  c->pairs_cap = 0;
  CU(c, cb_dmalloc(&c->d_pairs, cap * sizeof(PairOut)));
  c->pairs_cap = cap;
  return CB_OK;
}

// (Re)allocate the global candidate queue (16 B per entry) and point the launch parameters at it.
static int size_queue(cb_ctx* c, ProbeParams& p, uint64_t cap) {
  if (c->gq_cap != cap) {
    cb_dfree(c->d_gq_hv);
    cb_dfree(c->d_gq_vs);
    c->d_gq_hv = nullptr;
    c->d_gq_vs = nullptr;
    c->gq_cap = 0;
    CU(c, cb_dmalloc(&c->d_gq_hv, cap * sizeof(uint64_t)));
    CU(c, cb_dmalloc(&c->d_gq_vs, cap * sizeof(uint2)));
    c->gq_cap = cap;
  }
  if (!c->d_overflow) CU(c, cb_dmalloc(&c->d_overflow, 64 * sizeof(uint32_t)));
  p.gq_hv = c->d_gq_hv;
  p.gq_vs = c->d_gq_vs;
  p.gq_cap = c->gq_cap;
  p.overflow_chunks = c->d_overflow;
  return CB_OK;
}

// d <= 2.  d = 0 is one kernel.  d = 1, 2 run in chunks of seeds: the enumeration kernel fills the
// global candidate queue, the table kernel drains it.  A chunk whose candidates did not fit is
// skipped by the table kernel (nothing accumulated) and redone here in smaller pieces; if a single
// seed does not fit (a hub sequence with more neighbours than the queue has entries) the queue
// grows instead.
static int run_chunks(cb_ctx* c, ProbeParams& p, uint64_t w_first, uint64_t w_count, uint64_t per_chunk,
                      int depth, int* launches) {
  std::vector<std::pair<uint64_t, uint64_t>> chunks;
  for (uint64_t at = 0; at < w_count; at += per_chunk) chunks.emplace_back(w_first + at, std::min(per_chunk, w_count - at));
  if (chunks.size() > 64) return fail(c, CB_ERR_LIMIT, "cb_run: internal: more than 64 chunks in one pass");
  CU(c, cudaMemsetAsync(c->d_counters + CTR_OVERFLOW, 0, sizeof(unsigned long long), c->stream));
  for (size_t k = 0; k < chunks.size(); k++) {
    p.w_first = chunks[k].first;
    p.w_count = chunks[k].second;
    p.split = 1;
    if (p.differences == 2) {  // too few seeds to fill the machine: split each seed's outer space
      const uint64_t want_items = (uint64_t)c->sm_count * 64 * 4;
      while (p.split < 64 && p.w_count * p.split < want_items) p.split <<= 1;
    }
    CU(c, cudaMemsetAsync(c->d_counters + CTR_GQ, 0, sizeof(unsigned long long), c->stream));
    CU(c, cudaMemsetAsync(c->d_counters + CTR_WORK, 0, sizeof(unsigned long long), c->stream));
    const char* kerr = nullptr;
    const int l = launch_probe(p, c->sm_count, c->stream, &kerr);
    if (l < 0) return fail(c, CB_ERR_LIMIT, "cb_run: %s", kerr ? kerr : "launch failed");
    launch_table_stage(p, c->sm_count, (uint32_t)k, c->stream);
    CU(c, cudaGetLastError());
    *launches += l + 1;
  }
  int rc = read_counters(c);
  if (rc) return rc;
  const uint64_t over = c->h_counters[CTR_OVERFLOW];
  if (over == 0) return CB_OK;
  std::vector<uint32_t> ids(std::min<uint64_t>(over, 64));  // <= 64 chunks per pass: every id was recorded
  CU(c, cudaMemcpyAsync(ids.data(), c->d_overflow, ids.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  for (uint32_t id : ids) {
    if (per_chunk <= 1 || depth >= 8) {  // cannot cut finer: a bigger queue, same pieces
      if (c->gq_cap >= (1ull << 30))
        return fail(c, CB_ERR_LIMIT, "cb_run: one sequence has more than 2^30 candidate matches");
      rc = size_queue(c, p, c->gq_cap * 4);
      c->gq_cap_grown = true;
      if (!rc) rc = run_chunks(c, p, chunks[id].first, chunks[id].second, per_chunk, depth, launches);
    } else {
      rc = run_chunks(c, p, chunks[id].first, chunks[id].second, std::max<uint64_t>(per_chunk / 8, 1), depth + 1, launches);
    }
    if (rc) return rc;
  }
  return CB_OK;
}

static int run_hash_path(cb_ctx* c, ProbeParams& p, uint64_t count, int* launches) {
  CU(c, cudaMemsetAsync(c->d_counters + CTR_WORK, 0, sizeof(unsigned long long), c->stream));
  if (p.differences == 0) {
    p.w_first = 0;
    p.w_count = count;
    p.split = 1;
    const char* kerr = nullptr;
    const int l = launch_probe(p, c->sm_count, c->stream, &kerr);
    if (l < 0) return fail(c, CB_ERR_LIMIT, "cb_run: %s", kerr ? kerr : "launch failed");
    CU(c, cudaGetLastError());
    *launches += l;
    return CB_OK;
  }
  const double L = p.a.n ? (double)c->run_res_bytes / (double)p.a.n : 1.0;
  const double s1 = p.sigma - 1.0;
  double probes = 1.0 + s1 * L + (p.indels ? (L + p.sigma * (L + 1.0)) : 0.0);
  if (p.differences == 2) probes += s1 * s1 * L * (L - 1.0) / 2.0;
  // Global candidate queue (16 B per entry).  cfg.queue_capacity if given; else sized for the whole
  // run at ~3 % of its probes reaching the table stage, between 2^16 and 2^26 entries (1 GiB).  It
  // only grows: a context that has seen a large run keeps its queue.
  uint64_t cap = c->cfg.queue_capacity;
  if (cap == 0) {
    const double want = (double)count * probes * 0.03 * 2.0;
    cap = 1ull << 16;
    while (cap < (1ull << 26) && (double)cap < want) cap <<= 1;
  }
  cap = std::max<uint64_t>(cap, 64);
  if (!c->cfg.queue_capacity) cap = std::max(cap, c->gq_cap);  // it only grows
  else if (c->gq_cap > cap && c->gq_cap_grown) cap = c->gq_cap;  // ... also past a configured size, once a seed needed it
  int rcq = size_queue(c, p, cap);
  if (rcq) return rcq;
  cap = c->gq_cap;
  // Seeds per chunk.  A small first chunk measures how many candidates a seed produces on this
  // data (it depends on d, on the filters' false-positive rate and on how much the sets overlap);
  // the rest runs in chunks sized to fill the queue at most half.  Few, large chunks matter for
  // d = 2, where one seed is ~36 000 probes and every kernel tail costs.
  const uint64_t guess = (uint64_t)std::max(1.0, (double)(cap / 2) / (probes * 0.03));
  const uint64_t first_n = std::min<uint64_t>(count, std::min<uint64_t>(guess, std::max<uint64_t>(count / 32, 4096)));
  int rc = run_chunks(c, p, 0, first_n, first_n, 0, launches);
  if (rc || first_n == count) return rc;
  const double per_seed = std::max(1.0, (double)c->h_counters[CTR_GQ] / (double)first_n);
  uint64_t per_chunk = (uint64_t)std::max(1.0, (double)(cap / 2) / (per_seed * 1.5));
  per_chunk = std::max<uint64_t>(per_chunk, (count - first_n + 63) / 64);
  return run_chunks(c, p, first_n, count - first_n, per_chunk, 0, launches);
}

extern "C" int cb_run(cb_ctx* c, const cb_dset* a, uint64_t first, uint64_t count) {
  if (!c || !a) return fail(c, CB_ERR_INVALID, "cb_run: NULL argument");
  if (!c->b) return fail(c, CB_ERR_STATE, "cb_run: set B has not been built (cb_build_b / cb_set_b)");
  if (c->cfg.differences <= MAXDIFF_HASH && !c->d_table && c->b->n)
    return fail(c, CB_ERR_STATE, "cb_run: set B has no table (the last cb_build_b / cb_set_b failed)");
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
    c->run_res_bytes = a->res_bytes;
    p.table = c->d_table;
    p.table_mask = c->slots - 1;
    p.bloom = c->d_bloom;
    p.bloom_blocks = c->bloom_blocks;
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
    p.force_generic = (c->cfg.flags & CB_FLAG_GENERIC_KERNEL) ? 1u : 0u;
    // a small matrix is accumulated in CTA-private shared-memory tiles with warp-aggregated atomics
    p.tile_cells = (!existence && !c->cfg.no_matrix && !(c->cfg.flags & CB_FLAG_NO_SMEM_TILE) &&
                    c->rows * cols <= MATRIX_TILE_MAX_CELLS)
                       ? (uint32_t)(c->rows * cols) : 0;
    p.score = c->cfg.score;
    p.ignore_counts = c->cfg.ignore_counts != 0;
    p.ignore_genes = c->cfg.ignore_genes != 0;
    p.existence = existence;
    p.no_matrix = c->cfg.no_matrix != 0;
    p.want_pairs = c->cfg.want_pairs != 0;
    p.pair_variant = c->network_mode;
    p.use_bloom = !(c->cfg.flags & CB_FLAG_NO_BLOOM);
    p.count_bloom = 1;
    p.differences = c->cfg.differences;
    p.indels = c->cfg.indels != 0;
    rc = run_hash_path(c, p, count, &launches);
    if (rc) return rc;
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
      int l2 = 0;
      rc = run_hash_path(c, p, count, &l2);
      if (rc) return rc;
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
  if (!c->matrix_external) cb_dfree(c->d_matrix);
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
