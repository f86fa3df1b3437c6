// upload.cu — host arrays -> device-resident set: chunked, double-buffered H2D copy pipelined
// with the pack + hash (+ table/Bloom insert) kernels.
//
// Replaces db_hash() (src/db.cc:903-916) and, for set B, the hash_insert() loop
// (src/overlap.cc:861-873) — with the copy across PCIe hidden behind them (or, more often, them
// hidden behind the copy: at 10^8 sequences the copy is ~4 GB and the kernels ~35 ms).
#include <cuda_runtime.h>

#include <algorithm>
#include <cub/cub.cuh>
#include <new>
#include <thread>
#include <vector>

#include "engine_internal.h"

using namespace cb;

namespace {

struct HostCols {  // normalised view of cb_set / cb_set_cols
  uint64_t n = 0;
  const uint8_t* residues = nullptr;
  cb_col offsets{}, lengths{}, v{}, j{}, rep{}, count{};
  uint32_t n_reps = 0, longest = 0;
  uint64_t index_base = 0;
};

HostCols from_set(const cb_set* s) {
  HostCols h;
  h.n = s->n;
  h.residues = s->residues;
  h.offsets = {s->offsets, 8, 0};
  h.v = {s->v_gene, 4, 0};
  h.j = {s->j_gene, 4, 0};
  h.rep = {s->rep, 4, 0};
  h.count = {s->count, 8, 0};
  h.n_reps = s->n_reps;
  h.longest = s->longest;
  h.index_base = s->index_base;
  return h;
}

HostCols from_cols(const cb_set_cols* s) {
  HostCols h;
  h.n = s->n;
  h.residues = s->residues;
  h.offsets = s->offsets;
  h.lengths = s->lengths;
  h.v = s->v_gene;
  h.j = s->j_gene;
  h.rep = s->rep;
  h.count = s->count;
  h.n_reps = s->n_reps;
  h.longest = s->longest;
  h.index_base = s->index_base;
  return h;
}

bool width_ok(const cb_col& c, bool allow8) {
  return !c.data || c.width == 1 || c.width == 2 || c.width == 4 || (allow8 && c.width == 8);
}

uint64_t sum_lengths(const void* p, uint32_t w, uint64_t first, uint64_t n) {
  uint64_t s = 0;
  if (w == 1) {
    const uint8_t* q = (const uint8_t*)p + first;
    for (uint64_t i = 0; i < n; i++) s += q[i];
  } else if (w == 2) {
    const uint16_t* q = (const uint16_t*)p + first;
    for (uint64_t i = 0; i < n; i++) s += q[i];
  } else {
    const uint32_t* q = (const uint32_t*)p + first;
    for (uint64_t i = 0; i < n; i++) s += q[i];
  }
  return s;
}

// Is p ordinary pageable host memory?  A cudaMemcpyAsync from it is staged by the driver at 5 GB/s
// (measured: 4.3 GB of std::vector columns in 0.88 s); from pinned memory the copy engine reads at
// PCIe speed.  So pageable columns go through the context's own pinned buffers, filled by a few
// host threads one chunk ahead of the copy engine.
bool is_pageable(const void* p) {
  if (!p) return false;
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return at.type == cudaMemoryTypeUnregistered;
}

void parallel_copy(void* dst, const void* src, size_t bytes) {
  const size_t n_thr = bytes >= (8u << 20) ? 8 : 1;
  if (n_thr == 1) {
    memcpy(dst, src, bytes);
    return;
  }
  std::vector<std::thread> pool;
  for (size_t t = 0; t < n_thr; t++) {
    const size_t a = bytes * t / n_thr, b = bytes * (t + 1) / n_thr;
    pool.emplace_back([=] { memcpy((char*)dst + a, (const char*)src + a, b - a); });
  }
  for (auto& th : pool) th.join();
}

struct Staging {  // one of the two device staging buffers of the pipeline
  uint64_t* starts = nullptr;  // offsets (chunk+1) or scan output
  uint64_t* wide = nullptr;    // widened lengths (lengths mode)
  void* lengths = nullptr;
  void *v = nullptr, *j = nullptr, *rep = nullptr, *count = nullptr;
  void* scan_tmp = nullptr;
  size_t scan_bytes = 0;
  cudaEvent_t copied = nullptr, consumed = nullptr;
  void release() {
    cb_dfree(starts); cb_dfree(wide); cb_dfree(lengths); cb_dfree(v); cb_dfree(j); cb_dfree(rep);
    cb_dfree(count); cb_dfree(scan_tmp);
    if (copied) cudaEventDestroy(copied);
    if (consumed) cudaEventDestroy(consumed);
    *this = Staging();
  }
};

// The pipeline.  If `table` is given, every chunk is also inserted into it (set B).  With a
// placement the host columns are one SHARD of a larger set: the device arrays are sized for the
// whole set and the shard is packed and hashed at its place in them (comm.cu fills in the rest).
int upload_pipeline(cb_ctx* c, const HostCols& h, BuiltTable* table, cb_dset** out, const cb_placement* pl = nullptr) {
  *out = nullptr;
  const bool len_mode = h.lengths.data != nullptr;
  if (h.n && !h.residues) return cb_fail(c, CB_ERR_INVALID, "upload: residues is NULL");
  if (h.n && !len_mode && !(h.offsets.data && h.offsets.width == 8))
    return cb_fail(c, CB_ERR_INVALID, "upload: need offsets (width 8) or lengths (width 1, 2 or 4)");
  if (!width_ok(h.lengths, false) || !width_ok(h.v, false) || !width_ok(h.j, false) || !width_ok(h.rep, false) ||
      !width_ok(h.count, true))
    return cb_fail(c, CB_ERR_INVALID, "upload: column width must be 1, 2, 4 (or 8 for counts)");
  if (h.n && !c->cfg.ignore_genes && (!h.v.data || !h.j.data))
    return cb_fail(c, CB_ERR_INVALID, "upload: v_gene/j_gene are NULL but ignore_genes is off");
  if (h.n && !c->cfg.ignore_counts && !h.count.data)
    return cb_fail(c, CB_ERR_INVALID, "upload: count is NULL but ignore_counts is off");
  if (h.n >= 0xffffffffull)
    return cb_fail(c, CB_ERR_LIMIT, "upload: more than 2^32-1 sequences in one set");
  int rc = cb_bind_device(c);
  if (rc) return rc;
  cb_dset* s = new (std::nothrow) cb_dset;
  if (!s) return cb_fail(c, CB_ERR_NOMEM, "upload: out of host memory");
  s->n = pl ? pl->n_total : h.n;
  s->index_base = h.index_base;
  s->n_reps = h.n_reps;
  const uint64_t n = h.n;
  const uint64_t seq0 = pl ? pl->seq_first : 0, res0 = pl ? pl->res_first : 0;
  if (pl ? pl->n_total == 0 : n == 0) {
    *out = s;
    return CB_OK;
  }
  const bool genes = !c->cfg.ignore_genes && h.v.data && h.j.data;
  const uint64_t* off = (const uint64_t*)h.offsets.data;

  // chunking: ~16 chunks, between 256 Ki and 8 Mi sequences (2 Mi when the columns are pageable and
  // pass through the pinned buffers: 2 x ~100 MB of pinned memory)
  const bool stage_host = n >= (1u << 16) && is_pageable(h.residues);
  uint64_t chunk = std::min<uint64_t>(std::max<uint64_t>((n + 15) / 16, 1ull << 18), stage_host ? 1ull << 21 : 1ull << 23);
  const uint64_t n_chunks = n ? (n + chunk - 1) / chunk : 0;
  std::vector<uint64_t> res_begin(n_chunks + 1, 0);  // residue index where each chunk starts
  if (len_mode) {
    // per-chunk residue counts on the host, one thread per chunk (10^8 one-byte lengths are 16 ms
    // of wall clock on one core: more than a third of the whole PCIe copy they precede)
    std::vector<uint64_t> part(n_chunks, 0);
    auto sum_range = [&](uint64_t k0, uint64_t k1) {
      for (uint64_t k = k0; k < k1; k++)
        part[k] = sum_lengths(h.lengths.data, h.lengths.width, k * chunk, std::min(chunk, n - k * chunk));
    };
    const uint64_t n_thr = n >= (1ull << 22) ? std::min<uint64_t>(n_chunks, 16) : 1;
    if (n_thr > 1) {
      std::vector<std::thread> pool;
      for (uint64_t t = 0; t < n_thr; t++)
        pool.emplace_back(sum_range, t * n_chunks / n_thr, (t + 1) * n_chunks / n_thr);
      for (auto& th : pool) th.join();
    } else {
      sum_range(0, n_chunks);
    }
    for (uint64_t k = 0; k < n_chunks; k++) res_begin[k + 1] = res_begin[k] + part[k];
  } else if (n) {
    for (uint64_t k = 0; k <= n_chunks; k++) res_begin[k] = off[std::min(k * chunk, n)] - off[0];
  }
  const uint64_t res_base = (len_mode || !n) ? 0 : off[0];
  s->res_bytes = res_begin[n_chunks];
  if (pl && res0 + s->res_bytes > pl->res_total) {
    delete s;
    return cb_fail(c, CB_ERR_INVALID, "upload: the shard's residues do not fit its place in the arena");
  }

  Staging st[2];
  cudaStream_t cs = c->copy_stream, ks = c->stream;
  uint32_t zrows_used = 0;
  auto bail = [&](int code, const char* what, cudaError_t e) {
    cudaStreamSynchronize(cs);
    cudaStreamSynchronize(ks);
    st[0].release();
    st[1].release();
    cb_free_dset(s);
    return cb_fail(c, code, "%s: %s", what, cudaGetErrorString(e));
  };
#define UP(expr)                                                                                      \
  do {                                                                                                \
    cudaError_t e__ = (expr);                                                                         \
    if (e__ != cudaSuccess)                                                                           \
      return bail(e__ == cudaErrorMemoryAllocation ? CB_ERR_NOMEM : CB_ERR_CUDA, #expr, e__);         \
  } while (0)

  UP(cb_dmalloc(&s->d_res, (pl ? pl->res_total : s->res_bytes) + 16));
  UP(cb_dmalloc(&s->d_meta, (pl ? pl->n_alloc : n) * sizeof(SeqRec)));
  UP(cb_dmalloc(&s->d_hash, (pl ? pl->n_alloc : n) * sizeof(uint64_t)));
  for (auto& b : st) {
    UP(cb_dmalloc(&b.starts, (chunk + 1) * 8));
    if (len_mode) {
      UP(cb_dmalloc(&b.wide, chunk * 8));
      UP(cb_dmalloc(&b.lengths, chunk * h.lengths.width));
      UP(cub::DeviceScan::ExclusiveSum(nullptr, b.scan_bytes, b.wide, b.starts, (int64_t)chunk, ks));
      UP(cb_dmalloc(&b.scan_tmp, b.scan_bytes));
    }
    if (genes) {
      UP(cb_dmalloc(&b.v, chunk * h.v.width));
      UP(cb_dmalloc(&b.j, chunk * h.j.width));
    }
    if (h.rep.data) UP(cb_dmalloc(&b.rep, chunk * h.rep.width));
    if (h.count.data) UP(cb_dmalloc(&b.count, chunk * h.count.width));
    UP(cudaEventCreateWithFlags(&b.copied, cudaEventDisableTiming));
    UP(cudaEventCreateWithFlags(&b.consumed, cudaEventDisableTiming));
  }
  // the hash kernel needs Zobrist rows for the longest sequence, which is only known after the
  // pack kernels ran: start with the hint (or 64 rows) and redo the hashes below if that was short
  rc = cb_ensure_ztab(c, std::max<uint32_t>(h.longest + 2, 64));
  if (rc) {
    st[0].release();
    st[1].release();
    cb_free_dset(s);
    return rc;
  }
  zrows_used = c->zrows;
  UP(cudaMemsetAsync(c->d_counters + CTR_MAXLEN, 0, sizeof(unsigned long long), ks));
  UP(cudaEventRecord(c->ev[0], ks));
  // the buffers above were allocated in stream order on the compute stream: the copy stream may
  // touch them only after that point
  UP(cudaStreamWaitEvent(cs, c->ev[0], 0));
  // everything queued on the compute stream so far (table clear, counters) precedes the copies'
  // consumers by stream order; the copy stream only has to respect buffer reuse
  const uint32_t sigma = (uint32_t)c->cfg.alphabet_size;
  auto col_at = [](const cb_col& col, uint64_t i) { return (const char*)col.data + i * col.width; };
  // pinned pass-through buffers (one per pipeline slot), kept by the context across calls
  size_t pin_need = 0;
  if (stage_host) {
    uint64_t rmax = 0;
    for (uint64_t k = 0; k < n_chunks; k++) rmax = std::max(rmax, res_begin[k + 1] - res_begin[k]);
    pin_need = ((rmax + 63) & ~63ull) + (chunk + 1) * 8 + chunk * (size_t)(h.v.width + h.j.width + h.rep.width + h.count.width + 8) + 512;
    for (int q = 0; q < 2; q++)
      if (c->pin_bytes[q] < pin_need) {
        if (c->pin[q]) cudaFreeHost(c->pin[q]);
        c->pin[q] = nullptr;
        c->pin_bytes[q] = 0;
        UP(cudaHostAlloc(&c->pin[q], pin_need, cudaHostAllocDefault));
        c->pin_bytes[q] = pin_need;
      }
  }
  for (uint64_t k = 0; k < n_chunks; k++) {
    Staging& b = st[k & 1];
    const uint64_t first = k * chunk, cn = std::min(chunk, n - first);
    if (k >= 2) UP(cudaStreamWaitEvent(cs, b.consumed, 0));
    const uint64_t rb = res_begin[k], rn = res_begin[k + 1] - rb;
    char* pin_at = nullptr;
    if (stage_host) {
      if (k >= 2) UP(cudaEventSynchronize(b.copied));  // the copy engine is done with this slot's pinned buffer
      pin_at = (char*)c->pin[k & 1];
    }
    // host source of one column piece: the caller's memory, or its copy in the pinned buffer
    auto src = [&](const void* p, size_t bytes) -> const void* {
      if (!stage_host || bytes == 0) return p;
      char* dst = pin_at;
      parallel_copy(dst, p, bytes);
      pin_at += (bytes + 63) & ~(size_t)63;
      return dst;
    };
    UP(cudaMemcpyAsync(s->d_res + res0 + rb, src(h.residues + res_base + rb, rn), rn, cudaMemcpyHostToDevice, cs));
    if (len_mode)
      UP(cudaMemcpyAsync(b.lengths, src(col_at(h.lengths, first), cn * h.lengths.width), cn * h.lengths.width, cudaMemcpyHostToDevice, cs));
    else
      UP(cudaMemcpyAsync(b.starts, src(off + first, (cn + 1) * 8), (cn + 1) * 8, cudaMemcpyHostToDevice, cs));
    if (genes) {
      UP(cudaMemcpyAsync(b.v, src(col_at(h.v, first), cn * h.v.width), cn * h.v.width, cudaMemcpyHostToDevice, cs));
      UP(cudaMemcpyAsync(b.j, src(col_at(h.j, first), cn * h.j.width), cn * h.j.width, cudaMemcpyHostToDevice, cs));
    }
    if (h.rep.data)
      UP(cudaMemcpyAsync(b.rep, src(col_at(h.rep, first), cn * h.rep.width), cn * h.rep.width, cudaMemcpyHostToDevice, cs));
    if (h.count.data)
      UP(cudaMemcpyAsync(b.count, src(col_at(h.count, first), cn * h.count.width), cn * h.count.width, cudaMemcpyHostToDevice, cs));
    UP(cudaEventRecord(b.copied, cs));
    UP(cudaStreamWaitEvent(ks, b.copied, 0));
    PackCols pc{};
    pc.starts = b.starts;
    if (len_mode) {
      launch_widen(b.lengths, h.lengths.width, cn, b.wide, ks);
      UP(cub::DeviceScan::ExclusiveSum(b.scan_tmp, b.scan_bytes, b.wide, b.starts, (int64_t)cn, ks));
      pc.lengths = b.lengths;
      pc.len_w = h.lengths.width;
      pc.res_add = res0 + rb;
    } else {
      pc.off_sub = res_base - res0;  // wraps when the shard sits further up the arena than in the host's: off = o - off_sub still holds
    }
    pc.v = b.v; pc.v_w = h.v.width;
    pc.j = b.j; pc.j_w = h.j.width;
    pc.rep = b.rep; pc.rep_w = h.rep.width;
    pc.count = b.count; pc.count_w = h.count.width;
    launch_pack_meta(pc, cn, s->d_meta + seq0 + first, c->d_counters, ks);
    launch_hash(s->d_meta + seq0 + first, s->d_res, cn, c->d_ztab, zrows_used, sigma, c->cfg.seed,
                c->cfg.ignore_genes != 0, s->d_hash + seq0 + first, ks);
    if (table) cb_table_insert(c, *table, s, first, cn);
    UP(cudaGetLastError());
    UP(cudaEventRecord(b.consumed, ks));
  }
  UP(cudaEventRecord(c->ev[1], ks));
  UP(cudaMemcpyAsync(c->h_counters, c->d_counters, CTR_COUNT * sizeof(unsigned long long),
                     cudaMemcpyDeviceToHost, ks));
  UP(cudaStreamSynchronize(cs));
  UP(cudaStreamSynchronize(ks));
  st[0].release();
  st[1].release();
  cudaEventElapsedTime(&c->stats.ms_hash_a, c->ev[0], c->ev[1]);  // copy + pack + hash (+ insert) span
  s->longest = (uint32_t)c->h_counters[CTR_MAXLEN];
  if (s->longest >= (1u << 20)) {
    cb_free_dset(s);
    return cb_fail(c, CB_ERR_LIMIT, "upload: sequence longer than 2^20 residues");
  }
  if (s->res_bytes >= (1ull << 40)) {
    cb_free_dset(s);
    return cb_fail(c, CB_ERR_LIMIT, "upload: more than 2^40 residues in one set");
  }
  if (s->longest > zrows_used) {
    // a sequence was longer than the Zobrist rows we had: extend the table (same values for the
    // old rows) and redo hashes — and the inserts, into a cleared table
    rc = cb_ensure_ztab(c, s->longest + 2);
    if (rc) {
      cb_free_dset(s);
      return rc;
    }
    launch_hash(s->d_meta + seq0, s->d_res, n, c->d_ztab, c->zrows, sigma, c->cfg.seed, c->cfg.ignore_genes != 0,
                s->d_hash + seq0, ks);
    if (table) {
      BuiltTable fresh;
      rc = cb_table_alloc(c, n, true, &fresh);
      if (rc) {
        cb_free_dset(s);
        return rc;
      }
      table->release();
      *table = fresh;
      launch_reset_next(s->d_meta, n, ks);
      cb_table_insert(c, *table, s, 0, n);
    }
    UP(cudaGetLastError());
    UP(cudaStreamSynchronize(ks));
  }
#undef UP
  *out = s;
  return CB_OK;
}

int set_b_impl(cb_ctx* c, const HostCols& h) {
  int rc = cb_bind_device(c);
  if (rc) return rc;
  c->stats.ms_hash_b = c->stats.ms_build_b = c->stats.ms_dups_b = 0;
  if (c->b_owned && c->b) {  // replacing our own copy of set B: give its memory back to the pool first
    cudaStreamSynchronize(c->stream);
    cb_free_dset(c->b);
    c->b = nullptr;
    c->b_owned = false;
  }
  BuiltTable bt;
  const bool hash_path = c->cfg.differences <= MAXDIFF_HASH;
  if (hash_path) {
    rc = cb_table_alloc(c, h.n, true, &bt);
    if (rc) return rc;
  }
  cb_dset* d = nullptr;
  rc = upload_pipeline(c, h, hash_path ? &bt : nullptr, &d);
  if (rc) {
    bt.release();
    return rc;
  }
  c->stats.ms_build_b = c->stats.ms_hash_a;  // one fused span: copy + pack + hash + insert
  c->stats.ms_hash_a = 0;
  rc = cb_adopt_table(c, d, bt, true);
  if (rc) return rc;
  c->stats.kernel_launches = 0;
  return CB_OK;
}

}  // namespace

// Sum of n lengths of width w (1, 2 or 4 bytes) on up to 16 host threads.
uint64_t cb_sum_lengths(const void* p, uint32_t w, uint64_t n) {
  const uint64_t n_thr = n >= (1ull << 22) ? 16 : 1;
  if (n_thr == 1) return sum_lengths(p, w, 0, n);
  std::vector<uint64_t> part(n_thr, 0);
  std::vector<std::thread> pool;
  for (uint64_t t = 0; t < n_thr; t++)
    pool.emplace_back([&, t] { part[t] = sum_lengths(p, w, t * n / n_thr, (t + 1) * n / n_thr - t * n / n_thr); });
  for (auto& th : pool) th.join();
  uint64_t s = 0;
  for (uint64_t x : part) s += x;
  return s;
}

// One shard of a larger set into its place (comm.cu).
int cb_upload_shard(cb_ctx* c, const cb_set_cols* shard, const cb_placement* pl, cb_dset** out) {
  return upload_pipeline(c, from_cols(shard), nullptr, out, pl);
}

namespace {

int run_a_impl(cb_ctx* c, const HostCols& h) {
  cb_dset* d = nullptr;
  int rc = upload_pipeline(c, h, nullptr, &d);
  if (rc) return rc;
  const float ms_hash = c->stats.ms_hash_a;
  rc = cb_run(c, d, 0, d->n);
  c->stats.ms_hash_a = ms_hash;
  cudaStreamSynchronize(c->stream);
  cb_free_dset(d);
  return rc;
}

}  // namespace

extern "C" int cb_upload(cb_ctx* c, const cb_set* set, cb_dset** out) {
  if (!c || !set || !out) return cb_fail(c, CB_ERR_INVALID, "cb_upload: NULL argument");
  return upload_pipeline(c, from_set(set), nullptr, out);
}

extern "C" int cb_upload_cols(cb_ctx* c, const cb_set_cols* set, cb_dset** out) {
  if (!c || !set || !out) return cb_fail(c, CB_ERR_INVALID, "cb_upload_cols: NULL argument");
  return upload_pipeline(c, from_cols(set), nullptr, out);
}

extern "C" int cb_set_b(cb_ctx* c, const cb_set* b) {
  if (!c || !b) return cb_fail(c, CB_ERR_INVALID, "cb_set_b: NULL argument");
  return set_b_impl(c, from_set(b));
}

extern "C" int cb_set_b_cols(cb_ctx* c, const cb_set_cols* b) {
  if (!c || !b) return cb_fail(c, CB_ERR_INVALID, "cb_set_b_cols: NULL argument");
  return set_b_impl(c, from_cols(b));
}

extern "C" int cb_run_a(cb_ctx* c, const cb_set* a) {
  if (!c || !a) return cb_fail(c, CB_ERR_INVALID, "cb_run_a: NULL argument");
  return run_a_impl(c, from_set(a));
}

extern "C" int cb_run_a_cols(cb_ctx* c, const cb_set_cols* a) {
  if (!c || !a) return cb_fail(c, CB_ERR_INVALID, "cb_run_a_cols: NULL argument");
  return run_a_impl(c, from_cols(a));
}
