// comm.cu — multi-GPU: one context per GPU, NCCL over NVLink / NVSwitch.
//
// The path shards by set-A sequences: every GPU holds the whole set-B structure
// and works on its own range of set A, exactly as the reference's worker threads share one table
// and take chunks of seeds (src/overlap.cc:421-448).  Two exchanges are needed, both here:
//
//   cb_set_b_sharded     each rank copies only ITS 1/world of set B across its own PCIe link, packs
//                        and hashes it, and the ranks all-gather records, hashes and residues over
//                        NVLink (ncclAllGather); every rank then builds its own table + filters.
//                        Without it every rank pulls the whole of set B through one host's PCIe
//                        complex: at 8 GPUs that was 20.6 GB per step and 0.55 e2e efficiency.
//   cb_allreduce_matrix  sum of the per-rank partial matrices (ncclAllReduce, f64) — the merge of
//                        the per-thread matrices at the end of sim_thread (src/overlap.cc:510-527).
//
// Communicators: cb_comm_init_rank (one process per GPU: the caller carries the 128-byte unique id
// to the other processes over whatever host channel it has) or cb_comm_init_all (all contexts in
// one process, the CLI's --gpus N).
#include <cuda_runtime.h>
#include <nccl.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "engine_internal.h"

using namespace cb;

#define NC(c, expr)                                                                              \
  do {                                                                                           \
    ncclResult_t r__ = (expr);                                                                   \
    if (r__ != ncclSuccess) return cb_fail((c), CB_ERR_CUDA, "%s: %s", #expr, ncclGetErrorString(r__)); \
  } while (0)

static_assert(CB_UNIQUE_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "unique id size");

extern "C" int cb_comm_unique_id(void* out) {
  if (!out) return cb_fail(nullptr, CB_ERR_INVALID, "cb_comm_unique_id: NULL argument");
  ncclUniqueId id;
  NC(nullptr, ncclGetUniqueId(&id));
  memcpy(out, &id, sizeof id);
  return CB_OK;
}

static void drop_comm(cb_ctx* c) {
  if (c->comm) ncclCommDestroy((ncclComm_t)c->comm);
  c->comm = nullptr;
  c->rank = 0;
  c->world = 1;
}

void cb_comm_release(cb_ctx* c) { drop_comm(c); }

extern "C" int cb_comm_init_rank(cb_ctx* c, const void* unique_id, int rank, int world) {
  if (!c || !unique_id) return cb_fail(c, CB_ERR_INVALID, "cb_comm_init_rank: NULL argument");
  if (world < 1 || rank < 0 || rank >= world)
    return cb_fail(c, CB_ERR_INVALID, "cb_comm_init_rank: rank %d of %d", rank, world);
  int rc = cb_bind_device(c);
  if (rc) return rc;
  drop_comm(c);
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof id);
  ncclComm_t comm = nullptr;
  NC(c, ncclCommInitRank(&comm, world, id, rank));
  c->comm = comm;
  c->rank = rank;
  c->world = world;
  return CB_OK;
}

extern "C" int cb_comm_init_all(cb_ctx** ctxs, int n) {
  if (!ctxs || n < 1) return cb_fail(nullptr, CB_ERR_INVALID, "cb_comm_init_all: no contexts");
  std::vector<int> devs(n);
  for (int i = 0; i < n; i++) {
    if (!ctxs[i]) return cb_fail(nullptr, CB_ERR_INVALID, "cb_comm_init_all: context %d is NULL", i);
    devs[i] = ctxs[i]->device;
    for (int k = 0; k < i; k++)
      if (devs[k] == devs[i])
        return cb_fail(ctxs[i], CB_ERR_INVALID, "cb_comm_init_all: device %d appears twice", devs[i]);
    drop_comm(ctxs[i]);
  }
  std::vector<ncclComm_t> comms(n, nullptr);
  NC(ctxs[0], ncclCommInitAll(comms.data(), n, devs.data()));
  for (int i = 0; i < n; i++) {
    ctxs[i]->comm = comms[i];
    ctxs[i]->rank = i;
    ctxs[i]->world = n;
  }
  return CB_OK;
}

extern "C" int cb_comm_rank(const cb_ctx* c, int* rank, int* world) {
  if (!c) return CB_ERR_INVALID;
  if (rank) *rank = c->rank;
  if (world) *world = c->world;
  return CB_OK;
}

extern "C" void cb_shard_range(uint64_t n_total, int rank, int world, uint64_t* first, uint64_t* count) {
  const uint64_t per = world > 0 ? (n_total + (uint64_t)world - 1) / (uint64_t)world : n_total;
  const uint64_t f = std::min<uint64_t>(n_total, per * (uint64_t)std::max(rank, 0));
  if (first) *first = f;
  if (count) *count = std::min<uint64_t>(per, n_total - f);
}

extern "C" int cb_allreduce_matrix(cb_ctx* c) {
  if (!c) return CB_ERR_INVALID;
  if (c->cfg.mode != CB_MODE_MATRIX || c->cfg.no_matrix)
    return cb_fail(c, CB_ERR_STATE, "cb_allreduce_matrix: only in matrix mode with a matrix");
  int rc = cb_bind_device(c);
  if (rc) return rc;
  if (c->world == 1) return CB_OK;
  if (!c->comm) return cb_fail(c, CB_ERR_STATE, "cb_allreduce_matrix: no communicator (cb_comm_init_*)");
  if (!cb_matrix_device(c)) return cb_fail(c, CB_ERR_STATE, "cb_allreduce_matrix: no matrix yet (set B first)");
  const size_t cells = (size_t)(c->rows * c->cols);
  if (cells == 0) return CB_OK;
  NC(c, ncclAllReduce(c->d_matrix, c->d_matrix, cells, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return CB_OK;
}

// Set B from shards.  Rank r holds sequences [r * per, min((r + 1) * per, n_total)), per =
// ceil(n_total / world) (cb_shard_range).  Layout on every GPU after the call: records and hashes
// of the whole set in sequence order (so indices are global), residues in `world` regions of equal
// size (the largest shard's byte count, rounded up) — offsets in the records are absolute, the gaps
// between regions are never addressed.  Equal-sized regions make all three exchanges a plain
// in-place ncclAllGather.
extern "C" int cb_set_b_sharded(cb_ctx* c, const cb_set_cols* shard, uint64_t n_total) {
  if (!c || !shard) return cb_fail(c, CB_ERR_INVALID, "cb_set_b_sharded: NULL argument");
  int rc = cb_bind_device(c);
  if (rc) return rc;
  if (c->world == 1) {
    if (shard->n != n_total) return cb_fail(c, CB_ERR_INVALID, "cb_set_b_sharded: one rank must hold the whole set");
    return cb_set_b_cols(c, shard);
  }
  if (!c->comm) return cb_fail(c, CB_ERR_STATE, "cb_set_b_sharded: no communicator (cb_comm_init_*)");
  if (n_total >= 0xffffffffull) return cb_fail(c, CB_ERR_LIMIT, "more than 2^32-1 sequences in one set");
  ncclComm_t comm = (ncclComm_t)c->comm;
  const int world = c->world, rank = c->rank;
  uint64_t first = 0, count = 0;
  cb_shard_range(n_total, rank, world, &first, &count);
  if (shard->n != count)
    return cb_fail(c, CB_ERR_INVALID, "cb_set_b_sharded: rank %d of %d must pass %llu sequences of %llu, got %llu", rank,
                   world, (unsigned long long)count, (unsigned long long)n_total, (unsigned long long)shard->n);
  const uint64_t per = (n_total + world - 1) / world;
  cudaStream_t st = c->stream;
  c->stats.ms_hash_b = c->stats.ms_build_b = c->stats.ms_dups_b = 0;
  if (c->b_owned && c->b) {  // replacing our own copy of set B: give its memory back to the pool first
    cudaStreamSynchronize(st);
    cb_free_dset(c->b);
    c->b = nullptr;
    c->b_owned = false;
  }

  // residue bytes of this shard -> everybody's (one 8-byte all-gather); the longest sequence rides along
  uint64_t my_res = 0;
  if (shard->n) {
    if (shard->lengths.data) {
      const uint32_t w = shard->lengths.width;
      if (w != 1 && w != 2 && w != 4) return cb_fail(c, CB_ERR_INVALID, "cb_set_b_sharded: lengths must be 1, 2 or 4 bytes wide");
      my_res = cb_sum_lengths(shard->lengths.data, w, shard->n);
    } else if (shard->offsets.data && shard->offsets.width == 8) {
      const uint64_t* o = (const uint64_t*)shard->offsets.data;
      my_res = o[shard->n] - o[0];
    } else {
      return cb_fail(c, CB_ERR_INVALID, "cb_set_b_sharded: need offsets (width 8) or lengths (width 1, 2 or 4)");
    }
  }
  unsigned long long* d_x = nullptr;
  CU(c, cb_dmalloc(&d_x, (size_t)world * 8));
  std::vector<unsigned long long> sizes(world, 0);
  cudaError_t e = cudaMemcpyAsync(d_x + rank, &my_res, 8, cudaMemcpyHostToDevice, st);
  ncclResult_t nr = ncclSuccess;
  if (e == cudaSuccess) nr = ncclAllGather(d_x + rank, d_x, 1, ncclUint64, comm, st);
  if (e == cudaSuccess && nr == ncclSuccess)
    e = cudaMemcpyAsync(sizes.data(), d_x, (size_t)world * 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && nr == ncclSuccess) e = cudaStreamSynchronize(st);
  cb_dfree(d_x);
  if (nr != ncclSuccess) return cb_fail(c, CB_ERR_CUDA, "shard sizes all-gather: %s", ncclGetErrorString(nr));
  if (e != cudaSuccess) return cb_fail(c, CB_ERR_CUDA, "shard sizes all-gather: %s", cudaGetErrorString(e));
  uint64_t res_per = 0, res_sum = 0;
  for (int r = 0; r < world; r++) {
    res_per = std::max<uint64_t>(res_per, sizes[r]);
    res_sum += sizes[r];
  }
  res_per = (res_per + 15) & ~15ull;

  cb_placement pl{};
  pl.n_total = n_total;
  pl.n_alloc = per * world;
  pl.seq_first = first;
  pl.res_total = res_per * world;
  pl.res_first = res_per * rank;
  cb_dset* d = nullptr;
  rc = cb_upload_shard(c, shard, &pl, &d);
  if (rc) return rc;
  const float ms_upload = c->stats.ms_hash_a;
  c->stats.ms_hash_a = 0;

  // the exchange: three in-place all-gathers over NVLink, and the longest sequence of any shard
  unsigned long long* d_len = c->d_counters + CTR_MAXLEN;
  unsigned long long my_longest = d->longest;
  e = cudaMemcpyAsync(d_len, &my_longest, 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaEventRecord(c->ev[0], st);
  if (e == cudaSuccess) {
    nr = ncclGroupStart();
    if (nr == ncclSuccess) nr = ncclAllGather(d->d_meta + first, d->d_meta, per * sizeof(SeqRec), ncclUint8, comm, st);
    if (nr == ncclSuccess) nr = ncclAllGather(d->d_hash + first, d->d_hash, per, ncclUint64, comm, st);
    if (nr == ncclSuccess) nr = ncclAllGather(d->d_res + pl.res_first, d->d_res, res_per, ncclUint8, comm, st);
    if (nr == ncclSuccess) nr = ncclAllReduce(d_len, d_len, 1, ncclUint64, ncclMax, comm, st);
    const ncclResult_t ge = ncclGroupEnd();
    if (nr == ncclSuccess) nr = ge;
  }
  if (e == cudaSuccess && nr == ncclSuccess) e = cudaEventRecord(c->ev[1], st);
  if (e == cudaSuccess && nr == ncclSuccess)
    e = cudaMemcpyAsync(&my_longest, d_len, 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && nr == ncclSuccess) e = cudaStreamSynchronize(st);
  if (nr != ncclSuccess || e != cudaSuccess) {
    cb_free_dset(d);
    return cb_fail(c, CB_ERR_CUDA, "set-B all-gather: %s", nr != ncclSuccess ? ncclGetErrorString(nr) : cudaGetErrorString(e));
  }
  float ms_gather = 0;
  cudaEventElapsedTime(&ms_gather, c->ev[0], c->ev[1]);
  d->longest = (uint32_t)my_longest;
  d->res_bytes = res_sum;
  d->n_reps = shard->n_reps;
  rc = cb_ensure_ztab(c, d->longest + 2);  // rows for the longest sequence of ANY shard (values are a function of the seed)
  if (rc) {
    cb_free_dset(d);
    return rc;
  }
  c->b = d;        // cb_build_b adopts it
  c->b_owned = true;
  rc = cb_build_b(c, d);
  if (rc) return rc;
  c->stats.ms_hash_b = ms_upload;   // H2D of the shard + pack + hash
  c->stats.ms_gather_b = ms_gather;
  return CB_OK;
}
