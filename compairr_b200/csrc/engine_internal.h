// engine_internal.h — context and device-set structs shared by engine.cu and brute.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "../../include/compairr_b200.h"
#include "kernels.cuh"

struct cb_dset {
  uint64_t n = 0;
  uint64_t index_base = 0;
  uint64_t res_bytes = 0;
  uint32_t n_reps = 0;
  uint32_t longest = 0;
  bool links_dirty = false;  // SeqRec.next written by a table insert (a fresh upload leaves them SEQ_NIL)
  cb::SeqRec* d_meta = nullptr;
  uint8_t* d_res = nullptr;
  uint64_t* d_hash = nullptr;
  // d >= 3 only (set B): bucket order by (length[, V, J]), packed words, host bucket directory
  uint32_t* d_order = nullptr;
  uint32_t* d_packed = nullptr;
  std::vector<uint64_t> bucket_key;    // sorted unique keys
  std::vector<uint64_t> bucket_start;  // n_buckets + 1 positions into d_order
  std::vector<uint64_t> pack_off;      // n_buckets + 1 word offsets into d_packed
};

// Device memory comes from a stream-ordered pool owned by the context (cudaMallocFromPoolAsync)
// with an unlimited release threshold: repeated set-B builds and set-A uploads reuse cached blocks instead of paying
// cudaMalloc/cudaFree of multi-GB buffers every call (that was ~100 ms per bench step).  Ordered
// on the stream of the context bound to the calling thread (cb_bind_device).
cudaError_t cb_dmalloc_raw(void** p, size_t bytes);
template <typename T>
static inline cudaError_t cb_dmalloc(T** p, size_t bytes) { return cb_dmalloc_raw(reinterpret_cast<void**>(p), bytes); }
cudaError_t cb_dfree(void* p);

struct BuiltTable {
  cb::Slot* table = nullptr;
  uint64_t slots = 0;
  unsigned long long* bloom = nullptr;  // the four class filters back to back, CB_CLASSES * blocks words
  uint32_t blocks = 0;                  // 64-bit words per filter
  void release() {
    cb_dfree(table);
    cb_dfree(bloom);
    *this = BuiltTable();
  }
};

struct cb_ctx {
  int insert_launches = 0;  // kernels launched by the most recent cb_table_insert (ours + CUB radix sort passes)
  cb_config cfg{};
  int device = 0;
  int sm_count = 148;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // H2D side of the pipelined upload
  cudaStream_t side_stream = nullptr;  // filter passes of a sorted table insert, beside sort and table build
  cudaMemPool_t pool = nullptr;        // the context's own stream-ordered memory pool
  void* pin[2] = {nullptr, nullptr};   // pinned pass-through buffers of the upload pipeline (pageable callers)
  size_t pin_bytes[2] = {0, 0};
  cudaEvent_t ev[10]{};  // [8], [9]: fork / join of the filter passes running beside the table build
  std::string err;

  uint64_t* d_ztab = nullptr;
  uint32_t zrows = 0;

  unsigned long long* d_counters = nullptr;
  unsigned long long* h_counters = nullptr;  // pinned

  cb_dset* b = nullptr;
  bool b_owned = false;
  cb::Slot* d_table = nullptr;
  uint64_t slots = 0;
  unsigned long long* d_bloom = nullptr;   // the four class filters (CB_CLASSES * bloom_blocks words)
  uint32_t bloom_blocks = 0;
  uint64_t dups_b = 0;

  double* d_matrix = nullptr;
  bool matrix_external = false;
  uint64_t rows = 0, cols = 0;

  uint64_t* d_gq_hv = nullptr;      // global candidate queue (enumeration kernel -> table kernel)
  uint2* d_gq_vs = nullptr;
  uint32_t* d_overflow = nullptr;
  uint64_t gq_cap = 0;              // entries
  bool gq_cap_grown = false;        // a single seed overflowed the configured queue: it was enlarged
  uint64_t run_res_bytes = 0;
  cb::PairOut* d_pairs = nullptr;
  uint64_t pairs_cap = 0;
  std::vector<cb_pair> pending;
  bool network_mode = false;  // cb_cluster: pairs carry their variant descriptor (cluster.cu)

  // multi-GPU (comm.cu): NCCL communicator over the contexts that share the work
  void* comm = nullptr;  // ncclComm_t
  int rank = 0, world = 1;

  cb_stats stats{};
};

int cb_fail(cb_ctx* c, int code, const char* fmt, ...);
cb::DeviceSetView cb_view_of(const cb_dset* s);

#define CU(c, expr)                                                                        \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess)                                                                \
      return cb_fail((c), e__ == cudaErrorMemoryAllocation ? CB_ERR_NOMEM : CB_ERR_CUDA,   \
                     "%s: %s", #expr, cudaGetErrorString(e__));                            \
  } while (0)

// engine.cu
int cb_bind_device(cb_ctx* c);
int cb_ensure_ztab(cb_ctx* c, uint32_t rows);
int cb_table_alloc(cb_ctx* c, uint64_t n, bool with_bloom, BuiltTable* out, bool clear_table = true);
void cb_table_insert(cb_ctx* c, const BuiltTable& t, cb_dset* s, uint64_t first, uint64_t n, bool whole = false);
int cb_adopt_table(cb_ctx* c, cb_dset* b, BuiltTable& t, bool owned);  // sets ctx fields, counts dups
void cb_free_dset(cb_dset* s);

// upload.cu: pack + hash one shard of a larger set at its place in device arrays sized for the whole
struct cb_placement {
  uint64_t n_total;    // sequences of the whole set (cb_dset.n)
  uint64_t n_alloc;    // entries the record / hash arrays are allocated for (>= n_total)
  uint64_t seq_first;  // index of the shard's first sequence
  uint64_t res_total;  // bytes of the residue arena
  uint64_t res_first;  // arena offset of the shard's first residue
};
uint64_t cb_sum_lengths(const void* p, uint32_t w, uint64_t n);
int cb_upload_shard(cb_ctx* c, const cb_set_cols* shard, const cb_placement* pl, cb_dset** out);

// comm.cu
void cb_comm_release(cb_ctx* c);

// brute.cu
int cb_run_brute(cb_ctx* c, const cb_dset* a, uint64_t first, uint64_t count, bool pairs_only,
                 int* launches);
