// kernels.cuh — launch interface between the engine (engine.cu) and the sm_100a kernels
// (kernels.cu).  Internal; the public boundary is include/compairr_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace cb {

struct PairOut {
  uint64_t a, b;
};

// Device counters, one block of 8 u64 per context.
enum Counter : int {
  CTR_MATCHES = 0,
  CTR_BLOOM_PASS = 1,
  CTR_PAIRS = 2,    // pair cursor (may exceed capacity: overflow is detected from it)
  CTR_WORK = 3,     // work-item dispenser of the probe kernels
  CTR_DUPS = 4,
  CTR_MAXLEN = 5,
  CTR_PROBES = 6,
  CTR_GQ = 7,       // cursor of the global candidate queue (may exceed capacity: overflow)
  CTR_OVERFLOW = 8, // chunks whose candidates did not fit the queue (they are redone smaller)
  CTR_SPILL = 9,    // keys the tiled build passed on to the direct insert
  CTR_COUNT = 16
};

struct DeviceSetView {
  const SeqRec* meta;
  const uint8_t* res;
  const uint64_t* hash;
  uint64_t n;
  uint64_t index_base;
};

// Shared-memory matrix tile: up to 1024 cells (8 KB, 32 x 32 repertoires).  Measured on 10^6
// low-complexity sequences, self-comparison, d = 1 -i (1.35e10 matches; tools/tile_ab.py): with the
// tile vs plain global REDs — 2 repertoires (every match on one of 4 cells) 0.74 s vs 12.8 s, 8
// repertoires 0.42 vs 2.04 s, 32 x 32 repertoires 0.68 vs 0.68 s, 64 x 64 0.68 vs 0.44 s: beyond
// ~10^3 cells the global REDs are spread thinly enough and shared-memory f64 atomics (CAS loops)
// lose.
constexpr uint32_t MATRIX_TILE_MAX_CELLS = 1024;

struct ProbeParams {
  DeviceSetView a;
  DeviceSetView b;
  uint64_t a_first;  // first seed of the run: matrix rows / queue entries are relative to it
  uint64_t a_count;  // seeds in the run
  uint64_t w_first;  // this launch works on seeds [a_first + w_first, a_first + w_first + w_count)
  uint64_t w_count;
  // global candidate queue between the enumeration kernel and the table kernel
  uint64_t* gq_hv;
  uint2* gq_vs;      // {variant descriptor, seed number relative to a_first}
  uint64_t gq_cap;
  uint32_t* overflow_chunks;  // ids of chunks that overflowed the queue (first 64)
  const Slot* table;
  uint64_t table_mask;
  const unsigned long long* bloom;   // the four class filters back to back (common.cuh)
  uint32_t bloom_blocks;             // 64-bit words per filter
  const uint64_t* ztab;  // global copy of the Zobrist table, zrows x sigma
  uint32_t zrows;        // rows staged in shared memory by the variant kernel (>= longest A + 1)
  uint32_t sigma;
  uint64_t seed;
  double* matrix;  // rows x n_cols
  uint64_t n_cols;
  PairOut* pairs;
  uint64_t pairs_cap;
  unsigned long long* counters;
  uint32_t lmax;   // longest seed of the set (decides which enumeration kernels are launched)
  uint32_t len_lo, len_hi;  // an enumeration launch handles the seeds with len_lo <= length <= len_hi
  uint32_t force_generic;   // A/B and tests: every seed through the generic (any-length) kernel
  uint32_t tile_cells;  // > 0: the matrix (rows x n_cols = tile_cells doubles) is small: CTAs of the table stage
                        // accumulate into a private copy in shared memory (device_utils.cuh accumulate_warp)
  uint32_t split;  // d=2: work items per seed
  int32_t score;
  uint8_t ignore_counts, ignore_genes, existence, no_matrix;
  uint8_t want_pairs, use_bloom, count_bloom;
  uint8_t pair_variant;  // network mode (-c): pair.b carries the 31-bit variant descriptor in its high half
  int32_t differences;
  uint8_t indels;
};

// pack one uploaded chunk of columns into SeqMeta records, track the longest sequence
struct PackCols {
  const uint64_t* starts;  // offsets (n+1) or exclusive scan of lengths (n)
  const void* lengths;     // non-null selects lengths mode
  const void* v;
  const void* j;
  const void* rep;
  const void* count;
  uint64_t off_sub;        // offsets mode: subtracted from every offset
  uint64_t res_add;        // lengths mode: residue index of the chunk's first sequence
  uint32_t len_w, v_w, j_w, rep_w, count_w;
};
void launch_widen(const void* src, uint32_t w, uint64_t n, uint64_t* dst, cudaStream_t st);
void launch_pack_meta(const PackCols& k, uint64_t n, SeqRec* out, unsigned long long* counters,
                      cudaStream_t st);

// K1: batched Zobrist hashing
void launch_hash(const SeqRec* meta, const uint8_t* res, uint64_t n, const uint64_t* ztab,
                 uint32_t zrows, uint32_t sigma, uint64_t seed, bool ignore_genes, uint64_t* out,
                 cudaStream_t st);

// K2: table + Bloom build, duplicate count
void launch_table_clear(Slot* table, uint64_t slots, cudaStream_t st);
void launch_reset_next(SeqRec* meta, uint64_t n, cudaStream_t st);
// Inserts sequences [first, first + n) of the set; writes their SeqRec.next links (which must be
// SEQ_NIL on entry).  bloom == nullptr: table only (the filters are built by launch_filters).
// part_hash/part_idx (both or neither): the keys h * CB_HOME_MUL sorted by their top bits (= by home
// slot), position t = sequence first + part_idx[t].
void launch_build(SeqRec* meta, const uint8_t* res, const uint64_t* hash, const uint64_t* part_hash,
                  const uint32_t* part_idx, uint64_t first, uint64_t n, bool ignore_genes, Slot* table,
                  uint64_t mask, unsigned long long* bloom, uint32_t bloom_blocks, cudaStream_t st);
// A whole set into an EMPTY, UNCLEARED table of 2^tbits slots (tbits >= BUILD_TILE_BITS): one CTA per tile of
// BUILD_TILE_SLOTS slots builds it in shared memory and streams it out.  part_key sorted on (at least) its top
// tbits - BUILD_TILE_BITS bits; tile_first (2^(tbits - BUILD_TILE_BITS) + 1 words), defer and spill (n words
// each) are scratch; counters[CTR_SPILL] must be zero on entry.  Returns the launches made.
constexpr int BUILD_TILE_BITS = 12;
constexpr uint32_t BUILD_TILE_SLOTS = 1u << BUILD_TILE_BITS;  // 64 KiB of shared memory
// CTA shape of the tile kernel (whole build at 10^8 keys: 256x3 11.9, 512x2 10.43, 384x3 10.28, 512x3 with spills
// 11.3, 1024x1 12.8 ms): three 64-KiB tiles per SM in flight
#ifndef CB_TILE_THREADS
#define CB_TILE_THREADS 384
#endif
#ifndef CB_TILE_CTAS
#define CB_TILE_CTAS 3  // per SM
#endif
constexpr int BUILD_TILE_THREADS = CB_TILE_THREADS;
int launch_build_tiled(SeqRec* meta, const uint8_t* res, const uint64_t* part_key, const uint32_t* part_idx,
                       uint64_t first, uint64_t n, bool ignore_genes, Slot* table, uint32_t tbits, uint32_t* tile_first,
                       uint32_t* defer, uint32_t* spill, unsigned long long* counters, int sm_count, cudaStream_t st);
// the four class filters of hashes [0, n), L2-sized word ranges at a time; returns the launches made
int launch_filters(const uint64_t* hash, uint64_t n, unsigned long long* bloom, uint32_t bloom_blocks, int sm_count,
                   cudaStream_t st);
void launch_iota(uint32_t* p, uint64_t n, cudaStream_t st);
void launch_partition_keys(const uint64_t* hash, uint64_t n, uint64_t* key, uint32_t* idx, cudaStream_t st);
void launch_count_dups(DeviceSetView s, unsigned long long* counters, cudaStream_t st);
// -z: lead[i] = first member (file order) of i's (repertoire, V, J, sequence) group, sums[lead] =
// the group's count (sums must be zero on entry), counters[CTR_DUPS] += members merged away
void launch_dedup(DeviceSetView s, bool ignore_counts, uint32_t* lead, unsigned long long* sums,
                  unsigned long long* counters, cudaStream_t st);

// K3+K4: enumerate variants, Bloom, probe, verify, accumulate.  Returns launches made, <0 on
// a configuration the kernels cannot take (message in *err).
int launch_probe(const ProbeParams& p, int sm_count, cudaStream_t st, const char** err);
// K4 as its own kernel: consumes the global candidate queue (d = 1, 2 paths).
void launch_table_stage(const ProbeParams& p, int sm_count, uint32_t chunk_id, cudaStream_t st);

// Closed-form variant count of seeds [first, first+count) summed into counters[CTR_PROBES]
// (bookkeeping for the probes/s metric; not part of the timed hot path).
void launch_count_probes(DeviceSetView a, uint64_t first, uint64_t count, uint32_t sigma, int d,
                         bool indels, unsigned long long* counters, cudaStream_t st);

}  // namespace cb
