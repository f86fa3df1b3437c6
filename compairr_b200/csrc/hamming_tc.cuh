// hamming_tc.cuh — launch interface of the tcgen05 one-hot int8 GEMM kernel (hamming_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace cb {

constexpr int TC_M = 128;  // set-A sequences per tile (accumulator rows = TMEM lanes)
constexpr int TC_N = 256;  // set-B sequences per tile (accumulator columns)
constexpr uint32_t TC_KMAX = 352;  // largest one-hot width (bytes): (128 + 2 * 256) rows x 352 B = 220 KB of shared memory

struct TcItem {       // one set-A tile against a run of set-B tiles of the same bucket
  uint64_t a_start;   // position in the A bucket order
  uint64_t b_start;   // position in the B bucket order
  uint64_t a_pack;    // word offset of the A bucket in the packed array (word-major inside the bucket)
  uint64_t b_pack;
  uint32_t a_pos;     // position of the tile's first sequence inside its bucket
  uint32_t b_pos;
  uint32_t a_bucket;  // sequences in the whole A bucket (stride between packed words)
  uint32_t b_bucket;
  uint32_t a_n;       // <= TC_M
  uint32_t b_n;       // any (processed TC_N at a time)
  uint32_t len;       // sequence length of the bucket
  uint32_t kpad;      // sigma * len rounded up to 32
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
  uint32_t sigma;
  uint64_t a_first;
  double* matrix;
  uint64_t n_cols;
  PairOut* pairs;
  uint64_t pairs_cap;
  unsigned long long* counters;
  int32_t score, differences;
  uint8_t ignore_counts, existence, no_matrix, want_pairs;
};

size_t tc_smem_bytes(uint32_t kmax);
int launch_hamming_tc(const TcLaunch& p, int sm_count, cudaStream_t st, const char** err);

}  // namespace cb
