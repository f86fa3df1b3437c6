// device_utils.cuh — device-side helpers shared by kernels.cu and variant.cu.
#pragma once
#include "kernels.cuh"

namespace cb {

static constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ Slot ld_slot(const Slot* p) {
  const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(p));  // one 128-bit load
  Slot s;
  s.hash = v.x;
  s.idx = v.y;
  return s;
}

__device__ __forceinline__ SeqMeta ld_meta(const SeqRec* p) {
  const ulonglong2 lo = __ldg(reinterpret_cast<const ulonglong2*>(p));
  const uint4 hi = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  return unpack_rec(lo.x, lo.y, hi.x, hi.y, hi.z, hi.w);
}

// Are the `len` residues at res + oa and res + ob equal?  Aligned 64-bit loads and funnel shifts
// instead of a byte loop with two dependent loads per residue: all loads of a sequence of up to
// 24 residues are issued at once (one memory round trip).  Reads whole aligned words, i.e. up to 7
// bytes past a sequence: residue buffers are allocated with 16 bytes of slack (upload.cu).
__device__ __forceinline__ bool seq_equal(const uint8_t* __restrict__ res, uint64_t oa, uint64_t ob, uint32_t len) {
  const uint64_t* pa = reinterpret_cast<const uint64_t*>(res + (oa & ~7ull));
  const uint64_t* pb = reinterpret_cast<const uint64_t*>(res + (ob & ~7ull));
  const uint32_t sa = (uint32_t)(oa & 7) * 8, sb = (uint32_t)(ob & 7) * 8;
  const uint32_t na = ((uint32_t)(oa & 7) + len + 7) >> 3, nb = ((uint32_t)(ob & 7) + len + 7) >> 3;
  auto chunk = [](uint64_t lo, uint64_t hi, uint32_t s) { return s ? (lo >> s) | (hi << (64 - s)) : lo; };
  auto mask = [](uint32_t rem) { return rem >= 8 ? ~0ull : (1ull << (rem * 8)) - 1; };
  uint64_t diff = 0;
  if (len <= 24) {
    uint64_t wa[4], wb[4];
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) {
      wa[k] = k < na ? __ldg(pa + k) : 0ull;
      wb[k] = k < nb ? __ldg(pb + k) : 0ull;
    }
#pragma unroll
    for (uint32_t k = 0; k < 3; k++)
      if (k * 8 < len) diff |= (chunk(wa[k], wa[k + 1], sa) ^ chunk(wb[k], wb[k + 1], sb)) & mask(len - k * 8);
    return diff == 0;
  }
  uint64_t a_lo = __ldg(pa), b_lo = __ldg(pb);
  for (uint32_t k = 0; k * 8 < len; k++) {
    const uint64_t a_hi = k + 1 < na ? __ldg(pa + k + 1) : 0ull, b_hi = k + 1 < nb ? __ldg(pb + k + 1) : 0ull;
    diff |= (chunk(a_lo, a_hi, sa) ^ chunk(b_lo, b_hi, sb)) & mask(len - k * 8);
    a_lo = a_hi;
    b_lo = b_hi;
  }
  return diff == 0;
}

// Class-filter lookup of hash h (common.cuh) in the filter that serves free positions of class c.
__device__ __forceinline__ bool pfilter_test(const unsigned long long* __restrict__ bloom,
                                             uint32_t nblocks, uint64_t h, uint32_t c) {
  return pattern_hit(__ldg(bloom + pfilter_word(h, nblocks, c)), pattern_field(h, c));
}

// Variant descriptor in 31 bits: kind(3) | res1(5) | res2(5) | pos1(9) | pos2(9).  Positions up to
// 511, which the shared-memory budget of the variant kernels enforces anyway.
constexpr uint32_t VAR_MAX_POS = 511;
__device__ __forceinline__ uint32_t pack_var(uint32_t kind, uint32_t pos1, uint32_t r1, uint32_t pos2,
                                             uint32_t r2) {
  return kind | (r1 << 3) | (r2 << 8) | (pos1 << 13) | (pos2 << 22);
}

// K4 accumulate (matrix[R2 * i + j] += s, overlap.cc:218-228), called by the whole warp; `ok` lanes
// carry a match.  Two forms:
//   tile == nullptr   one fire-and-forget RED.F64 per match into the global matrix.  At the match
//                     densities of repertoire data the matrix is nowhere near a limit: 25 M matches
//                     per 4 ms step on 10^6 cells (C3) is 0.1 % of the L2 atomic rate.
//   tile != nullptr   the matrix is small enough for a CTA-private copy in shared memory
//                     (cells <= MATRIX_TILE_MAX_CELLS): lanes that hit the same cell are combined
//                     first (__match_any_sync + shuffles: one atomic per distinct cell per warp
//                     step), the tile is flushed with one global RED per non-zero cell when the CTA
//                     ends.  This is what keeps few-repertoire, many-match runs (self-comparisons
//                     of low-complexity sets: 10^9 matches on a handful of cells) off the
//                     same-address atomic rate of L2.
// f64 sums of integer-valued summands are exact and order-free below 2^53; for `ratio` the order
// moves the last bits, far inside the 1e-12 budget.
__device__ __forceinline__ void accumulate_warp(const ProbeParams* __restrict__ P, double* tile, bool ok,
                                                uint32_t row, uint32_t col, double sc) {
  if (tile == nullptr) {
    if (ok) atomicAdd(P->matrix + (uint64_t)row * P->n_cols + col, sc);
    return;
  }
  const unsigned active = __ballot_sync(FULL, ok);
  if (!ok) return;
  const uint32_t cell = row * (uint32_t)P->n_cols + col;
  const unsigned peers = __match_any_sync(active, cell);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t leader = (uint32_t)__ffs((int)peers) - 1u;
  double acc = sc;
  unsigned rem = peers & ~(1u << leader);  // the lanes whose summands the leader still has to collect
  while (__any_sync(active, rem != 0)) {
    const int src = rem ? __ffs((int)rem) - 1 : (int)lane;
    const double v = __shfl_sync(active, sc, src);
    if (rem) {
      acc += v;
      rem &= rem - 1;
    }
  }
  if (lane == leader) atomicAdd(tile + cell, acc);
}

// K4 for the warp's candidate groups (one distinct set-B sequence with its occurrence list per
// lane): V/J compare and exact verify of the edit ONCE against the head (byte-wise; a word-wide
// verify for sequences up to 32 residues removed 40 % of this kernel's instructions and none of its
// time, 14.80 vs 14.83 ms at C3 geometry: the kernel waits on dependent loads, not on issue slots), then score + accumulate +
// pair append for every occurrence (overlap.cc:189-245).  Called by the whole warp; `cand` lanes
// hold a group.  The walk over the occurrences is warp-synchronous so that the accumulation can
// combine lanes.
__device__ __forceinline__ uint32_t verify_and_record(const ProbeParams* __restrict__ P, bool cand,
                                                      uint64_t seed_idx, const SeqMeta& sm,
                                                      uint32_t row, uint32_t var, uint64_t head,
                                                      double* tile) {
  SeqMeta hm = {};
  bool ok = false;
  if (cand) {
    hm = ld_meta(P->b.meta + head);
    ok = P->ignore_genes || (hm.v == sm.v && hm.j == sm.j);
    if (ok) {
      const uint32_t kind = var & 7, r1 = (var >> 3) & 31, r2 = (var >> 8) & 31;
      const uint32_t pos1 = (var >> 13) & 511, pos2 = (var >> 22) & 511;
      ok = verify_variant(P->a.res + sm.off, sm.len, P->b.res + hm.off, hm.len, kind, pos1, r1, pos2, r2);
    }
  }
  uint32_t found = 0;
  uint64_t node = head;
  while (__any_sync(FULL, ok)) {
    if (!P->no_matrix)
      accumulate_warp(P, tile, ok, row, hm.rep, ok ? score_of(P->score, P->ignore_counts, sm.count, hm.count) : 0.0);
    if (ok) {
      found++;
      // network mode (-c): one set against itself, the self hit is not an edge (cluster.cc:105)
      if (P->want_pairs && !(P->pair_variant && node == seed_idx)) {
        const unsigned long long at = atomicAdd(P->counters + CTR_PAIRS, 1ull);
        if (at < P->pairs_cap) {
          PairOut po;
          po.a = seed_idx + P->a.index_base;
          po.b = node + P->b.index_base;
          if (P->pair_variant) po.b |= (uint64_t)var << 32;
          P->pairs[at] = po;
        }
      }
      if (hm.next == SEQ_NIL) {
        ok = false;
      } else {
        node = hm.next;
        hm = ld_meta(P->b.meta + node);
      }
    }
  }
  return found;
}

// Linear probing for up to 32 (hash, variant, seed) candidates, one per lane, executed by the
// whole warp: lane-divergent chain walk (cheap: one 128-bit load per step, to the first empty
// slot, visiting every slot with an equal stored hash, overlap.cc:181-250), then — re-converged by
// warp votes, once per round — metadata, V/J, exact verify, score, atomics for every lane that
// holds a candidate.
__device__ __forceinline__ uint32_t probe_chains(const ProbeParams* __restrict__ P, bool walking,
                                                 uint64_t hv, uint32_t var, uint64_t sidx,
                                                 uint32_t exist_row, double* tile) {
  const uint64_t mask = P->table_mask;
  const Slot* __restrict__ table = P->table;
  uint64_t slot = table_home(hv, mask);
  SeqMeta sm = {};
  bool have_meta = false;
  uint32_t found = 0;
  while (__any_sync(FULL, walking)) {
    bool cand = false;
    uint64_t hit = 0;
    while (walking) {
      const Slot s = ld_slot(table + slot);
      slot = (slot + 1) & mask;
      if (s.idx == SLOT_EMPTY) {
        walking = false;
      } else if (s.hash == hv) {
        cand = true;
        hit = (uint32_t)s.idx;  // low half = head of the occurrence list (high half = hash tag)
        break;
      }
    }
    __syncwarp();
    if (cand && !have_meta) {
      sm = ld_meta(P->a.meta + sidx);
      have_meta = true;
    }
    found += verify_and_record(P, cand, sidx, sm, P->existence ? exist_row : sm.rep, var, hit, tile);
    __syncwarp();
  }
  return found;
}

// CTA-private matrix tile (accumulate_warp): zero it / add it into the global matrix.
__device__ __forceinline__ double* matrix_tile_begin(const ProbeParams& P, unsigned char* smem) {
  if (!P.tile_cells) return nullptr;
  double* tile = reinterpret_cast<double*>(smem);
  for (uint32_t i = threadIdx.x; i < P.tile_cells; i += blockDim.x) tile[i] = 0.0;
  __syncthreads();
  return tile;
}
__device__ __forceinline__ void matrix_tile_flush(const ProbeParams& P, double* tile) {
  if (!tile) return;
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < P.tile_cells; i += blockDim.x) {
    const double v = tile[i];
    if (v != 0.0) atomicAdd(P.matrix + i, v);
  }
}

__device__ __forceinline__ void flush_counters(const ProbeParams& P, uint32_t nmatch, uint32_t npass) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nmatch += __shfl_xor_sync(FULL, nmatch, o);
    npass += __shfl_xor_sync(FULL, npass, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (nmatch) atomicAdd(P.counters + CTR_MATCHES, (unsigned long long)nmatch);
    if (npass) atomicAdd(P.counters + CTR_BLOOM_PASS, (unsigned long long)npass);
  }
}

}  // namespace cb
