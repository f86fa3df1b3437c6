// hamming_tc.cuh — launch interface of the tcgen05 one-hot int8 GEMM kernel (hamming_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace cb {

constexpr uint32_t TC_ROWS = 128;  // rows of every shared-memory tile (MMA M, and MMA N)
constexpr uint32_t TC_MA = 256;    // set-A sequences per work item: two accumulator row blocks
constexpr uint32_t TC_NB = 128;    // set-B sequences per tile (accumulator columns per row block)
constexpr uint32_t TC_KMAX = 416;  // widest row image (bytes): 4 tiles x 128 rows x 416 B = 208 KB of shared memory
constexpr uint32_t TC_BSTAGES = 4;  // most B tile stages in shared memory (fewer when rows are wide)
constexpr uint32_t TC_QCAP = 64;   // candidate queue entries per epilogue warp
constexpr uint32_t TC_PACK_PAD = 20;  // byte that fills the last packed word of a sequence (no residue has this code)

struct TcItem {       // one set-A tile against a run of set-B tiles of the same bucket
  uint64_t a_start;   // position in the A bucket order
  uint64_t b_start;   // position in the B bucket order
  uint64_t a_pack;    // word offset of the A bucket in the packed array (word-major inside the bucket)
  uint64_t b_pack;
  uint32_t a_pos;     // position of the tile's first sequence inside its bucket
  uint32_t b_pos;
  uint32_t a_bucket;  // sequences in the whole A bucket (stride between packed words)
  uint32_t b_bucket;
  uint32_t a_n;       // <= TC_MA
  uint32_t b_n;       // any (processed TC_NB at a time)
  uint32_t len;       // sequence length of the bucket
  uint32_t kpad;      // columns per position * len, rounded up to 32
};

struct TcLaunch {
  DeviceSetView a, b;
  const uint32_t* a_order;
  const uint32_t* b_order;
  const uint32_t* a_packed;  // residues packed 4 per word, bucket order
  const uint32_t* b_packed;
  const TcItem* items;
  uint32_t n_items;
  uint32_t kmax;      // largest kpad among the items
  uint32_t b_stages;  // set by launch_hamming_tc
  uint32_t aa;        // 1: 8-byte residue code (filter + exact verify, alphabet <= 20), 0: 4-byte one-hot (exact, alphabet <= 4)
  uint64_t a_first;
  double* matrix;
  uint64_t n_cols;
  PairOut* pairs;
  uint64_t pairs_cap;
  unsigned long long* counters;
  int32_t score, differences;
  uint8_t ignore_counts, existence, no_matrix, want_pairs;
};

size_t tc_smem_bytes(uint32_t kmax, uint32_t b_stages);
// bytes one sequence position occupies in a tile row (0: alphabet not supported by this kernel)
uint32_t tc_cols_per_position(uint32_t sigma);
int launch_hamming_tc(TcLaunch p, int sm_count, cudaStream_t st, const char** err);

}  // namespace cb
