// brute.cu — d >= 3: exhaustive Hamming comparison (replaces process_trad + seq_diff,
// overlap.cc:286-359, util.cc:172-184).
//
// The reference tests every (seed, hit) pair: V/J equal (unless -g), lengths equal, Hamming
// distance <= d.  The first two conditions are equalities, so we bucket both sets by
// (length, V, J) (or by length alone with -g) with a device radix sort and only ever compare
// inside matching buckets — same result set, a fraction of the pair tests.  Inside a bucket pair
// a CTA holds a tile of set-A sequences packed 4 residues per 32-bit word in shared memory and
// streams the bucket's set-B sequences (pre-packed, word-major so the loads coalesce); each
// thread keeps one set-B sequence in registers and counts differing bytes against every
// set-A sequence of the tile with XOR + carry-less byte test + POPC.
//
// This version runs on the CUDA cores.  A one-hot int8 tcgen05 GEMM formulation is analysed in
// DESIGN.md (its epilogue, one TMEM read per pair, bounds it near this kernel's rate for
// CDR3-length sequences); it is not built yet.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <stdio.h>

#include <algorithm>
#include <cub/cub.cuh>
#include <string>
#include <vector>

#include "device_utils.cuh"
#include "engine_internal.h"
#include "hamming_tc.cuh"

using namespace cb;

namespace cb {

static constexpr unsigned FULLM = FULL;
constexpr int BK_THREADS = 256;
constexpr int BK_TA = 128;  // set-A sequences per tile

// bucket key: injective in (len, v, j); len < 2^20, v and j < 2^22 (checked on the host)
__global__ void __launch_bounds__(256)
bucket_key_kernel(const SeqRec* __restrict__ meta, uint64_t first, uint64_t n, bool ignore_genes,
                  uint64_t* __restrict__ keys, uint32_t* __restrict__ idx,
                  unsigned long long* __restrict__ gene_max) {
  uint32_t gm = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const SeqMeta m = ld_meta(meta + first + i);
    uint64_t k = (uint64_t)m.len << 44;
    if (!ignore_genes) {
      k |= ((uint64_t)(m.v & 0x3fffff) << 22) | (m.j & 0x3fffff);
      gm = max(gm, max(m.v, m.j));
    }
    keys[i] = k;
    idx[i] = (uint32_t)(first + i);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gm = max(gm, __shfl_xor_sync(FULLM, gm, o));
  if ((threadIdx.x & 31) == 0 && gm) atomicMax(gene_max, (unsigned long long)gm);
}

// Pack the sequences of a bucket-sorted order into words (4 residues each, the last word filled up
// with TC_PACK_PAD), word-major inside each bucket:
// word k of the s-th sequence of a bucket lives at pack_off[bucket] + k * bucket_n + s.
__global__ void __launch_bounds__(256)
pack_words_kernel(const SeqRec* __restrict__ meta, const uint8_t* __restrict__ res,
                  const uint32_t* __restrict__ order, const uint64_t* __restrict__ bstart,
                  const uint64_t* __restrict__ pack_off, uint32_t n_buckets, uint64_t n,
                  uint32_t* __restrict__ packed) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    // bucket of sorted position i: binary search in bstart
    uint32_t lo = 0, hi = n_buckets;
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (bstart[mid] <= i) lo = mid; else hi = mid;
    }
    const uint64_t bn = bstart[lo + 1] - bstart[lo];
    const uint64_t s = i - bstart[lo];
    const SeqMeta m = ld_meta(meta + order[i]);
    const uint32_t words = (m.len + 3) >> 2;
    const uint8_t* r = res + m.off;
    for (uint32_t k = 0; k < words; k++) {
      uint32_t w = 0;
      for (uint32_t b = 0; b < 4; b++) {
        const uint32_t p = k * 4 + b;
        w |= (p < m.len ? (uint32_t)r[p] : TC_PACK_PAD) << (8 * b);
      }
      packed[pack_off[lo] + (uint64_t)k * bn + s] = w;
    }
  }
}

struct BruteJoin {     // one matching (set-A bucket, set-B bucket) pair
  uint64_t a_start;    // position in the A order
  uint64_t b_start;    // position in the B order
  uint64_t b_pack;     // word offset of the B bucket in the packed array
  uint32_t a_n, b_n;
  uint32_t len, tiles_a;
  uint32_t tiles_b, b_chunk;
  uint64_t tile_base;  // first global tile id of this join
};

struct BruteLaunch {
  DeviceSetView a, b;
  const uint32_t* a_order;
  const uint32_t* b_order;
  const uint32_t* b_packed;
  const BruteJoin* joins;
  uint32_t n_joins;
  uint64_t n_tiles;
  uint64_t a_first;
  double* matrix;
  uint64_t n_cols;
  PairOut* pairs;
  uint64_t pairs_cap;
  unsigned long long* counters;
  int32_t score, differences;
  uint8_t ignore_counts, existence, no_matrix, want_pairs;
};

__device__ __forceinline__ uint32_t diff_bytes(uint32_t a, uint32_t b) {
  // residue codes are < 32, so (x + 0x7f) sets bit 7 of a byte iff the byte is non-zero and never
  // carries into the next byte
  const uint32_t x = a ^ b;
  return __popc((x + 0x7f7f7f7fu) & 0x80808080u);
}

template <int W>
__global__ void __launch_bounds__(BK_THREADS) brute_kernel(const __grid_constant__ BruteLaunch P) {
  extern __shared__ __align__(16) uint32_t sm_words[];  // [BK_TA][Wrt] then a_idx[BK_TA]
  for (uint64_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
    uint32_t lo = 0, hi = P.n_joins;
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (P.joins[mid].tile_base <= tile) lo = mid; else hi = mid;
    }
    const BruteJoin J = P.joins[lo];
    const uint32_t words = (J.len + 3) >> 2;
    const uint32_t Wrt = (W > 0) ? (uint32_t)W : ((words + 3) & ~3u);
    uint32_t* const a_idx = sm_words + (size_t)BK_TA * Wrt;
    const uint64_t t = tile - J.tile_base;
    const uint32_t ta = (uint32_t)(t % J.tiles_a), tb = (uint32_t)(t / J.tiles_a);
    const uint32_t a0 = ta * BK_TA;
    const uint32_t an = min((uint32_t)BK_TA, J.a_n - a0);
    const uint32_t b0 = tb * J.b_chunk;
    const uint32_t bn = min(J.b_chunk, J.b_n - b0);

    __syncthreads();  // previous tile's readers are done
    for (uint32_t i = threadIdx.x; i < an * Wrt; i += BK_THREADS) {
      const uint32_t a = i / Wrt, k = i - a * Wrt;
      const uint32_t seq = P.a_order[J.a_start + a0 + a];
      const SeqMeta m = ld_meta(P.a.meta + seq);
      uint32_t w = 0;  // same padding as pack_words_kernel inside the sequence's words, zero words past them
      for (uint32_t bb = 0; bb < 4; bb++) {
        const uint32_t p = k * 4 + bb;
        if (p < m.len) w |= (uint32_t)__ldg(P.a.res + m.off + p) << (8 * bb);
        else if (k < words) w |= TC_PACK_PAD << (8 * bb);
      }
      sm_words[i] = w;
      if (k == 0) a_idx[a] = seq;
    }
    __syncthreads();

    uint32_t nmatch = 0;
    for (uint32_t b = threadIdx.x; b < bn; b += BK_THREADS) {
      const uint32_t* bp = P.b_packed + J.b_pack + (b0 + b);
      uint32_t bw[W > 0 ? W : 1];
      if (W > 0) {
#pragma unroll
        for (int k = 0; k < W; k++) bw[k] = (k < (int)words) ? __ldg(bp + (uint64_t)k * J.b_n) : 0u;
      }
      for (uint32_t a = 0; a < an; a++) {
        uint32_t diff = 0;
        if (W > 0) {
          const uint4* aw = reinterpret_cast<const uint4*>(sm_words + (size_t)a * Wrt);
#pragma unroll
          for (int q = 0; q < W / 4; q++) {
            const uint4 x = aw[q];  // same address in every lane: broadcast
            diff += diff_bytes(x.x, bw[4 * q]) + diff_bytes(x.y, bw[4 * q + 1]) +
                    diff_bytes(x.z, bw[4 * q + 2]) + diff_bytes(x.w, bw[4 * q + 3]);
          }
        } else {
          for (uint32_t k = 0; k < words; k++)
            diff += diff_bytes(sm_words[(size_t)a * Wrt + k], __ldg(bp + (uint64_t)k * J.b_n));
        }
        if (diff <= (uint32_t)P.differences) {
          const uint32_t seed = a_idx[a];
          const uint32_t hit = P.b_order[J.b_start + b0 + b];
          const SeqMeta am = ld_meta(P.a.meta + seed);
          const SeqMeta bm = ld_meta(P.b.meta + hit);
          nmatch++;
          if (!P.no_matrix) {
            const uint64_t row = P.existence ? (uint64_t)seed - P.a_first : am.rep;
            atomicAdd(P.matrix + row * P.n_cols + bm.rep,
                      score_of(P.score, P.ignore_counts, am.count, bm.count));
          }
          if (P.want_pairs) {
            const unsigned long long at = atomicAdd(P.counters + CTR_PAIRS, 1ull);
            if (at < P.pairs_cap) {
              PairOut po;
              po.a = seed + P.a.index_base;
              po.b = hit + P.b.index_base;
              P.pairs[at] = po;
            }
          }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nmatch += __shfl_xor_sync(FULLM, nmatch, o);
    if ((threadIdx.x & 31) == 0 && nmatch)
      atomicAdd(P.counters + CTR_MATCHES, (unsigned long long)nmatch);
  }
}

}  // namespace cb

// ---- host side ---------------------------------------------------------------------------------

#define BCU(c, expr) CU(c, expr)

// Sort sequences [first, first+n) of a set by bucket key; returns the device order and the host
// bucket directory (unique keys ascending + start positions).
static int bucket_sort(cb_ctx* c, DeviceSetView v, uint64_t first, uint64_t n, uint32_t** d_order,
                       std::vector<uint64_t>& keys, std::vector<uint64_t>& starts) {
  cudaStream_t st = c->stream;
  const bool ig = c->cfg.ignore_genes != 0;
  keys.clear();
  starts.clear();
  *d_order = nullptr;
  if (n == 0) {
    starts.push_back(0);
    return CB_OK;
  }
  uint64_t *k_in = nullptr, *k_out = nullptr, *u_keys = nullptr, *u_cnt = nullptr;
  uint32_t *i_in = nullptr, *i_out = nullptr;
  uint64_t* d_nruns = nullptr;
  void* tmp = nullptr;
  unsigned long long* d_gmax = nullptr;
  auto cleanup = [&]() {
    cb_dfree(k_in); cb_dfree(k_out); cb_dfree(u_keys); cb_dfree(u_cnt);
    cb_dfree(i_in); cb_dfree(d_nruns); cb_dfree(tmp); cb_dfree(d_gmax);
  };
#define BCU2(expr)                                                                           \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      cleanup();                                                                             \
      cb_dfree(i_out);                                                                       \
      std::string m__ = std::string(#expr) + ": " + cudaGetErrorString(e__);                 \
      return cb_fail(c, e__ == cudaErrorMemoryAllocation ? CB_ERR_NOMEM : CB_ERR_CUDA, "%s", \
                     m__.c_str());                                                           \
    }                                                                                        \
  } while (0)
  BCU2(cb_dmalloc(&k_in, n * 8));
  BCU2(cb_dmalloc(&k_out, n * 8));
  BCU2(cb_dmalloc(&i_in, n * 4));
  BCU2(cb_dmalloc(&i_out, n * 4));
  BCU2(cb_dmalloc(&d_gmax, 8));
  BCU2(cudaMemsetAsync(d_gmax, 0, 8, st));
  const uint64_t blocks = (n + 255) / 256;
  bucket_key_kernel<<<(unsigned)std::min<uint64_t>(blocks, 148 * 16), 256, 0, st>>>(
      v.meta, first, n, ig, k_in, i_in, d_gmax);
  BCU2(cudaGetLastError());
  size_t tmp_bytes = 0;
  BCU2(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, i_in, i_out, (int64_t)n, 0, 64, st));
  BCU2(cb_dmalloc(&tmp, tmp_bytes));
  BCU2(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, i_in, i_out, (int64_t)n, 0, 64, st));
  cb_dfree(tmp);
  tmp = nullptr;
  // run-length encode the sorted keys -> bucket directory
  BCU2(cb_dmalloc(&u_keys, n * 8));
  BCU2(cb_dmalloc(&u_cnt, n * 8));
  BCU2(cb_dmalloc(&d_nruns, 8));
  tmp_bytes = 0;
  BCU2(cub::DeviceRunLengthEncode::Encode(nullptr, tmp_bytes, k_out, u_keys, u_cnt, d_nruns, (int64_t)n, st));
  BCU2(cb_dmalloc(&tmp, tmp_bytes));
  BCU2(cub::DeviceRunLengthEncode::Encode(tmp, tmp_bytes, k_out, u_keys, u_cnt, d_nruns, (int64_t)n, st));
  uint64_t nruns = 0;
  unsigned long long gmax = 0;
  BCU2(cudaMemcpyAsync(&nruns, d_nruns, 8, cudaMemcpyDeviceToHost, st));
  BCU2(cudaMemcpyAsync(&gmax, d_gmax, 8, cudaMemcpyDeviceToHost, st));
  BCU2(cudaStreamSynchronize(st));
  if (gmax >= (1ull << 22)) {
    cleanup();
    cb_dfree(i_out);
    return cb_fail(c, CB_ERR_LIMIT, "more than 2^22 distinct V or J genes on the d>=3 path");
  }
  keys.resize(nruns);
  std::vector<uint64_t> cnt(nruns);
  BCU2(cudaMemcpyAsync(keys.data(), u_keys, nruns * 8, cudaMemcpyDeviceToHost, st));
  BCU2(cudaMemcpyAsync(cnt.data(), u_cnt, nruns * 8, cudaMemcpyDeviceToHost, st));
  BCU2(cudaStreamSynchronize(st));
  starts.resize(nruns + 1);
  starts[0] = 0;
  for (uint64_t i = 0; i < nruns; i++) starts[i + 1] = starts[i] + cnt[i];
  cleanup();
#undef BCU2
  *d_order = i_out;
  return CB_OK;
}

// Set B: sort + pack, once.
// Pack the sequences of a bucket-sorted order (4 residues per word, word-major inside a bucket).
static int pack_order(cb_ctx* c, DeviceSetView v, const uint32_t* order, uint64_t n,
                      const std::vector<uint64_t>& keys, const std::vector<uint64_t>& starts,
                      std::vector<uint64_t>& pack_off, uint32_t** packed) {
  cudaStream_t st = c->stream;
  const size_t nb = keys.size();
  pack_off.assign(nb + 1, 0);
  for (size_t i = 0; i < nb; i++) {
    const uint64_t len = keys[i] >> 44;
    pack_off[i + 1] = pack_off[i] + ((len + 3) / 4) * (starts[i + 1] - starts[i]);
  }
  *packed = nullptr;
  if (n == 0) return CB_OK;
  uint64_t *d_starts = nullptr, *d_poff = nullptr;
  BCU(c, cb_dmalloc(packed, std::max<uint64_t>(pack_off[nb], 1) * 4));
  BCU(c, cb_dmalloc(&d_starts, (nb + 1) * 8));
  BCU(c, cb_dmalloc(&d_poff, (nb + 1) * 8));
  BCU(c, cudaMemcpyAsync(d_starts, starts.data(), (nb + 1) * 8, cudaMemcpyHostToDevice, st));
  BCU(c, cudaMemcpyAsync(d_poff, pack_off.data(), (nb + 1) * 8, cudaMemcpyHostToDevice, st));
  const uint64_t blocks = (n + 255) / 256;
  pack_words_kernel<<<(unsigned)std::min<uint64_t>(blocks, 148 * 16), 256, 0, st>>>(
      v.meta, v.res, order, d_starts, d_poff, (uint32_t)nb, n, *packed);
  BCU(c, cudaGetLastError());
  BCU(c, cudaStreamSynchronize(st));
  cb_dfree(d_starts);
  cb_dfree(d_poff);
  return CB_OK;
}

static int cb_build_brute_b(cb_ctx* c, cb_dset* b) {
  uint32_t** order = &b->d_order;
  std::vector<uint64_t>* keys = &b->bucket_key;
  std::vector<uint64_t>* starts = &b->bucket_start;
  uint32_t** packed = &b->d_packed;
  std::vector<uint64_t>* pack_off = &b->pack_off;
  if (*order) return CB_OK;  // already prepared
  DeviceSetView v = cb_view_of(b);
  int rc = bucket_sort(c, v, 0, v.n, order, *keys, *starts);
  if (rc) return rc;
  return pack_order(c, v, *order, v.n, *keys, *starts, *pack_off, packed);
}

int cb_run_brute(cb_ctx* c, const cb_dset* a, uint64_t first, uint64_t count, bool pairs_only,
                 int* launches) {
  *launches = 0;
  cb_dset* b = c->b;
  int rc = cb_build_brute_b(c, b);
  if (rc) return rc;
  uint32_t** b_order = &b->d_order;
  std::vector<uint64_t>* b_keys = &b->bucket_key;
  std::vector<uint64_t>* b_starts = &b->bucket_start;
  uint32_t** b_packed = &b->d_packed;
  std::vector<uint64_t>* b_poff = &b->pack_off;
  cudaStream_t st = c->stream;
  const cb_config& cfg = c->cfg;

  uint32_t* a_order = nullptr;
  std::vector<uint64_t> a_keys, a_starts;
  rc = bucket_sort(c, cb_view_of(a), first, count, &a_order, a_keys, a_starts);
  if (rc) return rc;
  *launches += 3;
  std::vector<uint64_t> a_poff;
  uint32_t* a_packed = nullptr;
  const bool use_tc = !(cfg.flags & CB_FLAG_NO_TENSOR);
  if (use_tc) {  // the tensor-core kernel builds its one-hot tiles from packed words on both sides
    rc = pack_order(c, cb_view_of(a), a_order, count, a_keys, a_starts, a_poff, &a_packed);
    if (rc) {
      cb_dfree(a_order);
      return rc;
    }
    (*launches)++;
  }

  // merge-join the two sorted bucket directories.  Bucket pairs that are a dense contraction worth
  // the tensor cores (one-hot width fits the shared-memory tiles, enough rows to fill a 128 x 256
  // tile several times) go to the tcgen05 kernel; the rest to the CUDA-core kernel, grouped by
  // packed width.
  const uint32_t sigma = (uint32_t)cfg.alphabet_size;
  std::vector<TcItem> tc_items;
  const uint32_t tc_kp = tc_cols_per_position(sigma);
  uint32_t tc_kmax = 0;
  std::vector<BruteJoin> joins[4];  // W = 4, 8, 16, generic
  uint32_t max_words[4] = {4, 8, 16, 0};
  const uint64_t sm_target = (uint64_t)c->sm_count * 8;
  size_t ia = 0, ib = 0;
  while (ia < a_keys.size() && ib < b_keys->size()) {
    if (a_keys[ia] < (*b_keys)[ib]) { ia++; continue; }
    if (a_keys[ia] > (*b_keys)[ib]) { ib++; continue; }
    BruteJoin J{};
    J.a_start = a_starts[ia];
    J.a_n = (uint32_t)(a_starts[ia + 1] - a_starts[ia]);
    J.b_start = (*b_starts)[ib];
    J.b_n = (uint32_t)((*b_starts)[ib + 1] - (*b_starts)[ib]);
    J.b_pack = (*b_poff)[ib];
    J.len = (uint32_t)(a_keys[ia] >> 44);
    J.tiles_a = (J.a_n + BK_TA - 1) / BK_TA;
    J.b_chunk = 1u << 14;
    J.tiles_b = (J.b_n + J.b_chunk - 1) / J.b_chunk;
    const uint32_t words = (J.len + 3) / 4;
    const int cls = words <= 4 ? 0 : words <= 8 ? 1 : words <= 16 ? 2 : 3;
    const uint32_t kpad = (tc_kp * J.len + 31) & ~31u;
    const bool dense = use_tc && tc_kp != 0 && J.len > 0 && kpad <= TC_KMAX && (int)J.len > cfg.differences &&
                       (uint64_t)J.a_n * J.b_n >= (uint64_t)128 * 256 * 4 && J.a_n >= 32 && J.b_n >= 64;
    if (dense) {
      // one item = one 128-row A tile against a chunk of B; chunk sized so the whole join gives at
      // least a few items per SM
      const uint32_t a_tiles = (J.a_n + TC_MA - 1) / TC_MA;
      uint32_t b_chunk = 128 * TC_NB;
      while (b_chunk > TC_NB && (uint64_t)a_tiles * ((J.b_n + b_chunk - 1) / b_chunk) < sm_target / 2) b_chunk >>= 1;
      for (uint32_t a0 = 0; a0 < J.a_n; a0 += TC_MA)
        for (uint32_t b0 = 0; b0 < J.b_n; b0 += b_chunk) {
          TcItem it{};
          it.a_start = J.a_start + a0;
          it.a_n = std::min<uint32_t>(TC_MA, J.a_n - a0);
          it.a_pack = a_poff[ia];
          it.a_pos = a0;
          it.a_bucket = J.a_n;
          it.b_start = J.b_start + b0;
          it.b_n = std::min<uint32_t>(b_chunk, J.b_n - b0);
          it.b_pack = J.b_pack;
          it.b_pos = b0;
          it.b_bucket = J.b_n;
          it.len = J.len;
          it.kpad = kpad;
          tc_items.push_back(it);
        }
      tc_kmax = std::max(tc_kmax, kpad);
    } else {
      if (cls == 3) max_words[3] = std::max(max_words[3], (words + 3) & ~3u);
      if (J.len > 0) joins[cls].push_back(J);
    }
    ia++;
    ib++;
  }
  int ret = CB_OK;
  if (!tc_items.empty()) {
    // largest first: the kernel's CTAs pull items from a counter
    std::stable_sort(tc_items.begin(), tc_items.end(), [](const TcItem& x, const TcItem& y) {
      return (uint64_t)x.a_n * x.b_n * x.kpad > (uint64_t)y.a_n * y.b_n * y.kpad;
    });
    TcItem* d_items = nullptr;
    BCU(c, cb_dmalloc(&d_items, tc_items.size() * sizeof(TcItem)));
    BCU(c, cudaMemcpyAsync(d_items, tc_items.data(), tc_items.size() * sizeof(TcItem), cudaMemcpyHostToDevice, st));
    TcLaunch T{};
    T.a = cb_view_of(a);
    T.b = cb_view_of(b);
    T.a_order = a_order;
    T.b_order = *b_order;
    T.a_packed = a_packed;
    T.b_packed = *b_packed;
    T.items = d_items;
    T.n_items = (uint32_t)tc_items.size();
    T.kmax = tc_kmax;
    T.aa = tc_kp == 8;
    T.a_first = first;
    T.matrix = c->d_matrix;
    T.n_cols = c->cols;
    T.pairs = c->d_pairs;
    T.pairs_cap = c->pairs_cap;
    T.counters = c->d_counters;
    T.score = cfg.score;
    T.differences = cfg.differences;
    T.ignore_counts = cfg.ignore_counts != 0;
    T.existence = cfg.mode == CB_MODE_EXISTENCE;
    T.no_matrix = (cfg.no_matrix != 0) || pairs_only;
    T.want_pairs = cfg.want_pairs != 0;
    const char* kerr = nullptr;
    if (launch_hamming_tc(T, c->sm_count, st, &kerr) < 0) {
      ret = cb_fail(c, CB_ERR_LIMIT, "d>=3 tensor-core kernel: %s", kerr ? kerr : "launch failed");
    } else {
      cudaError_t e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) ret = cb_fail(c, CB_ERR_CUDA, "d>=3 tensor-core kernel: %s", cudaGetErrorString(e));
      (*launches)++;
    }
    cb_dfree(d_items);
  }
  for (int cls = 0; cls < 4 && ret == CB_OK; cls++) {
    std::vector<BruteJoin>& js = joins[cls];
    if (js.empty()) continue;
    // shrink the B chunk while there are too few tiles to fill the GPU
    uint64_t tiles = 0;
    for (auto& J : js) tiles += (uint64_t)J.tiles_a * J.tiles_b;
    uint32_t chunk = 1u << 14;
    while (tiles < sm_target && chunk > BK_THREADS) {
      chunk >>= 1;
      tiles = 0;
      for (auto& J : js) {
        J.b_chunk = chunk;
        J.tiles_b = (J.b_n + chunk - 1) / chunk;
        tiles += (uint64_t)J.tiles_a * J.tiles_b;
      }
    }
    tiles = 0;
    for (auto& J : js) {
      J.tile_base = tiles;
      tiles += (uint64_t)J.tiles_a * J.tiles_b;
    }
    BruteJoin* d_joins = nullptr;
    BCU(c, cb_dmalloc(&d_joins, js.size() * sizeof(BruteJoin)));
    BCU(c, cudaMemcpyAsync(d_joins, js.data(), js.size() * sizeof(BruteJoin), cudaMemcpyHostToDevice, st));
    BruteLaunch L{};
    L.a = cb_view_of(a);
    L.b = cb_view_of(b);
    L.a_order = a_order;
    L.b_order = *b_order;
    L.b_packed = *b_packed;
    L.joins = d_joins;
    L.n_joins = (uint32_t)js.size();
    L.n_tiles = tiles;
    L.a_first = first;
    L.matrix = c->d_matrix;
    L.n_cols = c->cols;
    L.pairs = c->d_pairs;
    L.pairs_cap = c->pairs_cap;
    L.counters = c->d_counters;
    L.score = cfg.score;
    L.differences = cfg.differences;
    L.ignore_counts = cfg.ignore_counts != 0;
    L.existence = cfg.mode == CB_MODE_EXISTENCE;
    L.no_matrix = (cfg.no_matrix != 0) || pairs_only;
    L.want_pairs = cfg.want_pairs != 0;
    const unsigned grid = (unsigned)std::min<uint64_t>(tiles, (uint64_t)c->sm_count * 16);
    const uint32_t wrt = max_words[cls];
    const size_t smem = (size_t)BK_TA * wrt * 4 + BK_TA * 4;
    cudaError_t e = cudaSuccess;
    if (smem > 200 * 1024) {
      ret = cb_fail(c, CB_ERR_LIMIT, "sequence too long for the d>=3 kernel");
    } else {
      switch (cls) {
        case 0: brute_kernel<4><<<grid, BK_THREADS, smem, st>>>(L); break;
        case 1: brute_kernel<8><<<grid, BK_THREADS, smem, st>>>(L); break;
        case 2: brute_kernel<16><<<grid, BK_THREADS, smem, st>>>(L); break;
        default:
          e = cudaFuncSetAttribute(brute_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          if (e == cudaSuccess) brute_kernel<0><<<grid, BK_THREADS, smem, st>>>(L);
          break;
      }
      if (e == cudaSuccess) e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) {
        std::string m = std::string("d>=3 kernel: ") + cudaGetErrorString(e);
        ret = cb_fail(c, CB_ERR_CUDA, "%s", m.c_str());
      }
      (*launches)++;
    }
    cb_dfree(d_joins);
  }
  cb_dfree(a_order);
  cb_dfree(a_packed);
  return ret;
}
