// kernels.cu — hand-written sm_100a kernels of the overlap hot path.
//
//   pack_meta_kernel   SoA upload -> 32-byte SeqMeta records (+ longest length)
//   hash_kernel        K1  batched Zobrist hashing                    (replaces zobrist.cc:74-88, db.cc:903-916)
//   build_kernel       K2  open-addressing insert + Bloom set         (replaces overlap.cc:63-128 insert part,
//                                                                      hashtable.h:48-77, bloompat.h:50-53)
//   dups_kernel        K2b exact-duplicate count                      (replaces overlap.cc:63-128 dup part, :579-605)
//   identical_kernel   K3/K4 for d = 0: one thread per seed           (replaces overlap.cc:253-284 with variants.cc:260-268)
//   variant_kernel     K3/K4 for d = 1,2: one warp per seed (part): on-the-fly variant hashes by
//                      incremental XOR, Bloom prefilter, table probe, exact verify, score,
//                      matrix accumulation, pair append               (replaces variants.cc:270-428, overlap.cc:168-284)
//
// All of this is integer/byte work bound by random 8-byte Bloom reads (one 32-byte sector per
// probe); nothing here is a GEMM, so no tensor cores (see DESIGN.md for the roofline).
#include <stdio.h>

#include "kernels.cuh"

namespace cb {

static constexpr unsigned FULL = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ Slot ld_slot(const Slot* p) {
  const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(p));  // one 128-bit load
  Slot s;
  s.hash = v.x;
  s.idx = v.y;
  return s;
}

__device__ __forceinline__ SeqMeta ld_meta(const SeqMeta* p) {
  const ulonglong2 lo = __ldg(reinterpret_cast<const ulonglong2*>(p));
  const uint4 hi = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  SeqMeta m;
  m.off = lo.x;
  m.count = lo.y;
  m.len = hi.x;
  m.v = hi.y;
  m.j = hi.z;
  m.rep = hi.w;
  return m;
}


// First-level filter test.  K2 = true: 1 bit per 32-bit half (the low-bits-per-key geometry used
// when the filter is capped to stay L2-resident), else 3 + 3 bits.
__device__ __forceinline__ bool bloom_test(const unsigned long long* __restrict__ bloom,
                                           uint32_t nblocks, uint64_t h, bool k2) {
  const unsigned long long w = __ldg(bloom + bloom_block(h, nblocks));
  const uint32_t plo = k2 ? bloom1_pat_lo(h) : bloom_pat_lo(h);
  const uint32_t phi = k2 ? bloom1_pat_hi(h) : bloom_pat_hi(h);
  return (((uint32_t)w & plo) == plo) & (((uint32_t)(w >> 32) & phi) == phi);
}

// Variant descriptor packed into one register pair for the (rare) slow path.
__device__ __forceinline__ uint64_t pack_variant(uint32_t kind, uint32_t pos1, uint32_t r1,
                                                 uint32_t pos2, uint32_t r2) {
  return (uint64_t)kind | ((uint64_t)r1 << 8) | ((uint64_t)r2 << 16) | ((uint64_t)pos1 << 24) |
         ((uint64_t)pos2 << 44);
}

constexpr int VK_QCAP = 64;  // per-warp survivor queue entries (ring); drained 32 at a time

// K4 for one lane's candidate hit: V/J compare, exact verify of the edit, score, accumulate,
// pair append (overlap.cc:189-245).
__device__ __forceinline__ uint32_t verify_and_record(const ProbeParams* __restrict__ P,
                                                      uint64_t seed_idx, const SeqMeta& sm,
                                                      uint32_t row, uint64_t var, uint64_t hit) {
  const SeqMeta hm = ld_meta(P->b.meta + hit);
  if (!P->ignore_genes && (hm.v != sm.v || hm.j != sm.j)) return 0;
  const uint32_t kind = (uint32_t)(var & 0xff);
  const uint32_t r1 = (uint32_t)(var >> 8) & 0xff, r2 = (uint32_t)(var >> 16) & 0xff;
  const uint32_t pos1 = (uint32_t)(var >> 24) & 0xfffff, pos2 = (uint32_t)(var >> 44);
  if (!verify_variant(P->a.res + sm.off, sm.len, P->b.res + hm.off, hm.len, kind, pos1, r1, pos2, r2))
    return 0;
  if (!P->no_matrix) {
    const double sc = score_of(P->score, P->ignore_counts, sm.count, hm.count);
    atomicAdd(P->matrix + (uint64_t)row * P->n_cols + hm.rep, sc);
  }
  if (P->want_pairs) {
    const unsigned long long at = atomicAdd(P->counters + CTR_PAIRS, 1ull);
    if (at < P->pairs_cap) {
      PairOut po;
      po.a = seed_idx + P->a.index_base;
      po.b = hit + P->b.index_base;
      P->pairs[at] = po;
    }
  }
  return 1;
}

// K4: the slow path for up to 32 survivors of the first-level Bloom test, one per lane, always
// called by the whole warp (a call from divergent code leaves the warp split for the rest of
// the kernel — measured 4 of 32 lanes active, profiles/r01_*).  Per lane: optional second-level
// Bloom test (the big filter in HBM), then linear probing to the first empty slot visiting every
// slot with an equal stored hash (overlap.cc:181-250).  The chain walk is lane-divergent but
// cheap; the expensive part — metadata + residue compare + atomics — runs re-converged, once per
// round, for all lanes that hold a candidate.
__device__ __noinline__ uint32_t drain32(const ProbeParams* __restrict__ P, const uint64_t* qhv,
                                         const uint64_t* qvar, const uint32_t* qseed, uint32_t head,
                                         uint32_t n) {
  const uint32_t lane = threadIdx.x & 31;
  bool walking = lane < n;
  const uint32_t e = (head + lane) & (VK_QCAP - 1);
  uint64_t hv = 0, var = 0;
  uint32_t slocal = 0;
  if (walking) {
    hv = qhv[e];
    var = qvar[e];
    slocal = qseed[e];
  }
  if (P->bloom2 != nullptr && walking) {
    const unsigned long long w = __ldg(P->bloom2 + bloom_block(hv, P->bloom2_blocks));
    const uint32_t plo = bloom_pat_lo(hv), phi = bloom_pat_hi(hv);
    walking = (((uint32_t)w & plo) == plo) & (((uint32_t)(w >> 32) & phi) == phi);
  }
  const uint64_t mask = P->table_mask;
  const Slot* __restrict__ table = P->table;
  uint64_t slot = table_home(hv, mask);
  const uint64_t sidx = P->a_first + slocal;
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
        hit = s.idx;
        break;
      }
    }
    __syncwarp();
    if (cand) {
      if (!have_meta) {
        sm = ld_meta(P->a.meta + sidx);
        have_meta = true;
      }
      found += verify_and_record(P, sidx, sm, P->existence ? slocal : sm.rep, var, hit);
    }
    __syncwarp();
  }
  return found;
}

__device__ __forceinline__ void flush_counters(const ProbeParams& P, uint32_t nmatch,
                                               uint32_t npass) {
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

// ---------------------------------------------------------------------------------------------
// pack_meta
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
pack_meta_kernel(const uint64_t* __restrict__ offsets, const uint32_t* __restrict__ v,
                 const uint32_t* __restrict__ j, const uint32_t* __restrict__ rep,
                 const uint64_t* __restrict__ count, uint64_t n, uint64_t off_base,
                 SeqMeta* __restrict__ out, unsigned long long* counters) {
  uint32_t mymax = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t o0 = offsets[i], o1 = offsets[i + 1];
    SeqMeta m;
    m.off = o0 - off_base;
    m.len = (uint32_t)(o1 - o0);
    m.count = count ? count[i] : 1ull;
    m.v = v ? v[i] : 0u;
    m.j = j ? j[i] : 0u;
    m.rep = rep ? rep[i] : 0u;
    out[i] = m;
    mymax = max(mymax, m.len);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mymax = max(mymax, __shfl_xor_sync(FULL, mymax, o));
  if ((threadIdx.x & 31) == 0 && mymax) atomicMax(counters + CTR_MAXLEN, (unsigned long long)mymax);
}

void launch_pack_meta(const uint64_t* offsets, const uint32_t* v, const uint32_t* j,
                      const uint32_t* rep, const uint64_t* count, uint64_t n, uint64_t off_base,
                      SeqMeta* out, unsigned long long* counters, cudaStream_t st) {
  if (n == 0) return;
  const uint64_t blocks = (n + 255) / 256;
  pack_meta_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(
      offsets, v, j, rep, count, n, off_base, out, counters);
}

// ---------------------------------------------------------------------------------------------
// K1: Zobrist hash.  One thread per sequence; the table (sigma x rows u64) is staged in shared
// memory when it fits, else read through L1.  Streaming: (L + 32) bytes in, 8 bytes out.
// ---------------------------------------------------------------------------------------------

template <bool ZSMEM>
__global__ void __launch_bounds__(256)
hash_kernel(const SeqMeta* __restrict__ meta, const uint8_t* __restrict__ res, uint64_t n,
            const uint64_t* __restrict__ ztab, uint32_t zrows, uint32_t sigma, uint64_t seed,
            bool ignore_genes, uint64_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* zs = reinterpret_cast<uint64_t*>(smem_raw);
  if (ZSMEM) {
    for (uint32_t i = threadIdx.x; i < zrows * sigma; i += blockDim.x) zs[i] = ztab[i];
    __syncthreads();
  }
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const SeqMeta m = ld_meta(meta + i);
    uint64_t h = ignore_genes ? 0ull : vj_hash(seed, m.v, m.j);
    const uint8_t* s = res + m.off;
    for (uint32_t p = 0; p < m.len; p++) {
      const uint32_t r = __ldg(s + p);
      h ^= ZSMEM ? zs[p * sigma + r] : __ldg(ztab + p * sigma + r);
    }
    out[i] = h;
  }
}

void launch_hash(const SeqMeta* meta, const uint8_t* res, uint64_t n, const uint64_t* ztab,
                 uint32_t zrows, uint32_t sigma, uint64_t seed, bool ignore_genes, uint64_t* out,
                 cudaStream_t st) {
  if (n == 0) return;
  const uint64_t blocks = (n + 255) / 256;
  const unsigned grid = (unsigned)(blocks < 148 * 8 ? blocks : 148 * 8);
  const size_t zbytes = (size_t)zrows * sigma * sizeof(uint64_t);
  if (zbytes <= 40 * 1024) {
    hash_kernel<true><<<grid, 256, zbytes, st>>>(meta, res, n, ztab, zrows, sigma, seed,
                                                 ignore_genes, out);
  } else {
    hash_kernel<false><<<grid, 256, 0, st>>>(meta, res, n, ztab, zrows, sigma, seed,
                                              ignore_genes, out);
  }
}

// ---------------------------------------------------------------------------------------------
// K2: table + Bloom build.  One thread per set-B sequence.  Slot order inside a chain differs
// from the serial reference; harmless, every equal-hash slot of a chain is visited on probe.
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) table_clear_kernel(Slot* table, uint64_t slots) {
  const ulonglong2 e = make_ulonglong2(0ull, SLOT_EMPTY);
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < slots;
       i += (uint64_t)gridDim.x * blockDim.x)
    reinterpret_cast<ulonglong2*>(table)[i] = e;
}

void launch_table_clear(Slot* table, uint64_t slots, cudaStream_t st) {
  const uint64_t blocks = (slots + 255) / 256;
  table_clear_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(table,
                                                                                        slots);
}

__global__ void __launch_bounds__(256)
build_kernel(const uint64_t* __restrict__ hash, uint64_t n, Slot* table, uint64_t mask,
             unsigned long long* bloom, uint32_t bloom_blocks, bool k2, unsigned long long* bloom2,
             uint32_t bloom2_blocks) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t h = hash[i];
    uint64_t slot = table_home(h, mask);
    for (;;) {
      const unsigned long long prev = atomicCAS(
          reinterpret_cast<unsigned long long*>(&table[slot].idx), SLOT_EMPTY, (unsigned long long)i);
      if (prev == SLOT_EMPTY) {
        table[slot].hash = h;
        break;
      }
      slot = (slot + 1) & mask;
    }
    atomicOr(bloom + bloom_block(h, bloom_blocks), k2 ? bloom1_pattern(h) : bloom_pattern(h));
    if (bloom2) atomicOr(bloom2 + bloom_block(h, bloom2_blocks), bloom_pattern(h));
  }
}

void launch_build(const uint64_t* hash, uint64_t n, Slot* table, uint64_t mask,
                  unsigned long long* bloom, uint32_t bloom_blocks, bool k2, unsigned long long* bloom2,
                  uint32_t bloom2_blocks, cudaStream_t st) {
  if (n == 0) return;
  const uint64_t blocks = (n + 255) / 256;
  build_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(
      hash, n, table, mask, bloom, bloom_blocks, k2, bloom2, bloom2_blocks);
}

// Exact duplicates: sequence i is a duplicate iff an identical sequence (same repertoire, same
// V/J unless -g, same residues) with a SMALLER index is in its probe chain.  Sum over groups of
// (size - 1) — the same number the serial reference counts (overlap.cc:63-128, 865-873).
__global__ void __launch_bounds__(256)
dups_kernel(DeviceSetView s, const Slot* __restrict__ table, uint64_t mask, bool ignore_genes,
            unsigned long long* counters) {
  uint32_t dups = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < s.n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t h = s.hash[i];
    uint64_t slot = table_home(h, mask);
    bool have_meta = false, dup = false;
    SeqMeta m;
    for (;;) {
      const Slot sl = ld_slot(table + slot);
      if (sl.idx == SLOT_EMPTY) break;
      if (sl.hash == h && sl.idx < i) {
        if (!have_meta) {
          m = ld_meta(s.meta + i);
          have_meta = true;
        }
        const SeqMeta o = ld_meta(s.meta + sl.idx);
        if (o.rep == m.rep && o.len == m.len && (ignore_genes || (o.v == m.v && o.j == m.j))) {
          bool same = true;
          for (uint32_t p = 0; p < m.len && same; p++)
            same = s.res[m.off + p] == s.res[o.off + p];
          if (same) {
            dup = true;
            break;
          }
        }
      }
      slot = (slot + 1) & mask;
    }
    dups += dup;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dups += __shfl_xor_sync(FULL, dups, o);
  if ((threadIdx.x & 31) == 0 && dups) atomicAdd(counters + CTR_DUPS, (unsigned long long)dups);
}

void launch_count_dups(DeviceSetView s, const Slot* table, uint64_t mask, bool ignore_genes,
                       unsigned long long* counters, cudaStream_t st) {
  if (s.n == 0) return;
  const uint64_t blocks = (s.n + 255) / 256;
  dups_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(
      s, table, mask, ignore_genes, counters);
}

// Bookkeeping: closed-form number of variants for a range of seeds (SURVEY section 8d "unit of work").
__global__ void __launch_bounds__(256)
count_probes_kernel(DeviceSetView a, uint64_t first, uint64_t count, uint32_t sigma, int d,
                    bool indels, unsigned long long* counters) {
  unsigned long long sum = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const SeqMeta m = ld_meta(a.meta + first + i);
    sum += probe_count(a.res + m.off, m.len, sigma, d, indels);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
  if ((threadIdx.x & 31) == 0 && sum) atomicAdd(counters + CTR_PROBES, sum);
}

void launch_count_probes(DeviceSetView a, uint64_t first, uint64_t count, uint32_t sigma, int d,
                         bool indels, unsigned long long* counters, cudaStream_t st) {
  if (count == 0) return;
  const uint64_t blocks = (count + 255) / 256;
  count_probes_kernel<<<(unsigned)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, st>>>(
      a, first, count, sigma, d, indels, counters);
}

// ---------------------------------------------------------------------------------------------
// K3/K4, d = 0: one thread per seed (one probe per seed: a hash join).
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) identical_kernel(const __grid_constant__ ProbeParams P) {
  uint32_t nmatch = 0, npass = 0;
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t mask = P.table_mask;
  // warp-uniform trip count: the chain walk below re-converges with warp-wide votes
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x + (threadIdx.x & ~31u); i0 < P.a_count;
       i0 += stride) {
    const uint64_t i = i0 + lane;
    const bool in = i < P.a_count;
    const uint64_t sidx = P.a_first + (in ? i : 0);
    const uint64_t h = P.a.hash[sidx];
    bool walking = in;
    if (P.use_bloom && in) {
      walking = bloom_test(P.bloom, P.bloom_blocks, h, P.bloom_k2);
      if (walking && P.bloom2 != nullptr) walking = bloom_test(P.bloom2, P.bloom2_blocks, h, false);
    }
    npass += walking;
    uint64_t slot = table_home(h, mask);
    SeqMeta sm = {};
    bool have_meta = false;
    while (__any_sync(FULL, walking)) {
      bool cand = false;
      uint64_t hit = 0;
      while (walking) {
        const Slot s = ld_slot(P.table + slot);
        slot = (slot + 1) & mask;
        if (s.idx == SLOT_EMPTY) {
          walking = false;
        } else if (s.hash == h) {
          cand = true;
          hit = s.idx;
          break;
        }
      }
      __syncwarp();
      if (cand) {
        if (!have_meta) {
          sm = ld_meta(P.a.meta + sidx);
          have_meta = true;
        }
        nmatch += verify_and_record(&P, sidx, sm, P.existence ? (uint32_t)i : sm.rep,
                                    pack_variant(VK_IDENTICAL, 0, 0, 0, 0), hit);
      }
      __syncwarp();
    }
  }
  flush_counters(P, nmatch, P.count_bloom ? npass : 0);
}

// ---------------------------------------------------------------------------------------------
// K3/K4, d = 1 and d = 2: one warp per work item (a seed, or 1/split of a seed's d=2 space).
//
// Shared memory per CTA:   Zobrist rows 0..zrows-1 (sigma u64 each), then per warp
//   zo[p]  = Z(p, s[p])                       the seed's own table values
//   pre[p] = xor_{q<p} Z(q, s[q])             INDELS only: prefix/suffix scans replace the serial
//   sm[p]  = xor_{q>=p} Z(q-1, s[q])          incremental walks of variants.cc:311-324,341-353
//   sp[p]  = xor_{q>=p} Z(q+1, s[q])
//   sres[p] = s[p]
// Every lane decodes one candidate per step, XORs its hash together from these arrays, tests
// the Bloom block, and only survivors (<1 %) leave the loop for table_probe().
// ---------------------------------------------------------------------------------------------

constexpr int VK_THREADS = 256;
constexpr int VK_WARPS = VK_THREADS / 32;
__host__ __device__ inline uint32_t vk_lpad(uint32_t lmax) { return (lmax + 2 + 7) & ~7u; }
__host__ __device__ inline size_t vk_warp_u64(uint32_t lmax, bool indels) {
  return (size_t)vk_lpad(lmax) * (indels ? 4 : 1) + 2 * VK_QCAP;  // scratch + queue hv/var
}
__host__ __device__ inline size_t vk_warp_bytes(uint32_t lmax) {
  return (size_t)vk_lpad(lmax) + VK_QCAP * 4;  // residues + queue seed numbers
}
static size_t vk_smem_bytes(uint32_t zrows, uint32_t sigma, uint32_t lmax, bool indels) {
  return (size_t)zrows * sigma * 8 + VK_WARPS * (vk_warp_u64(lmax, indels) * 8 + vk_warp_bytes(lmax));
}

// Per-warp ring of Bloom survivors waiting for the table probe.
struct WarpQueue {
  uint64_t* hv;
  uint64_t* var;
  uint32_t* seed;  // seed number relative to a_first
  uint32_t head, count;
};

// Probe up to 32 queued survivors, one per lane; all 32 lanes make the call.
__device__ __forceinline__ uint32_t queue_drain(const ProbeParams& P, WarpQueue& q, uint32_t n) {
  const uint32_t found = drain32(&P, q.hv, q.var, q.seed, q.head, n);
  __syncwarp();
  q.head = (q.head + n) & (VK_QCAP - 1);
  q.count -= n;
  return found;
}

// Append this step's survivors (ballot + prefix popcount), drain when 32 are waiting.
__device__ __forceinline__ uint32_t queue_push(const ProbeParams& P, WarpQueue& q, uint32_t lane,
                                               bool pass, uint64_t hv, uint64_t var,
                                               uint32_t slocal) {
  const unsigned m = __ballot_sync(FULL, pass);
  if (m == 0) return 0;
  if (pass) {
    const uint32_t e = (q.head + q.count + __popc(m & ((1u << lane) - 1))) & (VK_QCAP - 1);
    q.hv[e] = hv;
    q.var[e] = var;
    q.seed[e] = slocal;
    // the drain will test the second-level filter (HBM): start that fetch now, towards L2
    if (P.bloom2 != nullptr)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(P.bloom2 + bloom_block(hv, P.bloom2_blocks)));
  }
  q.count += __popc(m);
  __syncwarp();
  return q.count >= 32 ? queue_drain(P, q, 32) : 0;
}

template <int SIGMA, bool INDELS, int D>
__global__ void __launch_bounds__(VK_THREADS, 3)
variant_kernel(const __grid_constant__ ProbeParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* const z = reinterpret_cast<uint64_t*>(smem_raw);
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lpad = vk_lpad(P.lmax);
  const size_t wu64 = vk_warp_u64(P.lmax, INDELS);
  uint64_t* const wbase = z + (size_t)P.zrows * SIGMA + warp * wu64;
  WarpQueue q;
  q.hv = wbase;
  q.var = wbase + VK_QCAP;
  uint64_t* const zo = wbase + 2 * VK_QCAP;
  uint64_t* const pre = zo + lpad;
  uint64_t* const sm = pre + lpad;
  uint64_t* const sp = sm + lpad;
  unsigned char* const bbase = reinterpret_cast<unsigned char*>(z + (size_t)P.zrows * SIGMA + VK_WARPS * wu64) +
                               warp * vk_warp_bytes(P.lmax);
  q.seed = reinterpret_cast<uint32_t*>(bbase);
  uint8_t* const sres = bbase + VK_QCAP * 4;
  q.head = 0;
  q.count = 0;

  for (uint32_t i = threadIdx.x; i < P.zrows * SIGMA; i += VK_THREADS) z[i] = P.ztab[i];
  __syncthreads();

  constexpr uint32_t S1 = SIGMA - 1;
  constexpr uint32_t BATCH = (D >= 2) ? 1 : 4;
  const uint64_t total_items = P.a_count * P.split;
  const uint32_t split_mask = P.split - 1;
  const uint32_t split_shift = 31 - __clz(P.split);
  const bool use_bloom = P.use_bloom;
  const bool k2 = P.bloom_k2;
  uint32_t nmatch = 0, npass = 0;

  for (;;) {
    unsigned long long item0 = 0;
    if (lane == 0) item0 = atomicAdd(P.counters + CTR_WORK, (unsigned long long)BATCH);
    item0 = __shfl_sync(FULL, item0, 0);
    if (item0 >= total_items) break;
    const uint64_t item_end = (item0 + BATCH < total_items) ? item0 + BATCH : total_items;

    for (uint64_t item = item0; item < item_end; ++item) {
      const uint32_t slocal = (uint32_t)(item >> split_shift);
      const uint32_t part = (uint32_t)item & split_mask;
      const uint64_t sidx = P.a_first + slocal;
      const SeqMeta m = ld_meta(P.a.meta + sidx);  // same address in all lanes: one broadcast
      const uint64_t h = __ldg(P.a.hash + sidx);
      const uint32_t L = m.len;

      __syncwarp();  // all lanes are done with the previous item's scratch
      for (uint32_t p = lane; p < L; p += 32) {
        const uint32_t r = __ldg(P.a.res + m.off + p);
        sres[p] = (uint8_t)r;
        zo[p] = z[p * SIGMA + r];
      }
      __syncwarp();

      uint64_t vjh = 0;
      if (INDELS) {
        // exclusive prefix XOR of zo[] in chunks of 32 positions
        uint64_t carry = 0;
        for (uint32_t base = 0; base < L; base += 32) {
          const uint32_t p = base + lane;
          uint64_t x = p < L ? zo[p] : 0ull;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint64_t y = __shfl_up_sync(FULL, x, o);
            if ((int)lane >= o) x ^= y;
          }
          if (p < L) pre[p + 1] = carry ^ x;
          carry ^= __shfl_sync(FULL, x, 31);
        }
        if (lane == 0) pre[0] = 0ull;
        vjh = h ^ carry;  // h = VJ ^ pre[L]
        // suffix XORs of the shifted-left / shifted-right values, walking from the end
        uint64_t cm = 0, cp = 0;
        for (uint32_t base = 0; base < L; base += 32) {
          const uint32_t t = base + lane;
          const bool ok = t < L;
          const uint32_t qq = ok ? L - 1 - t : 0;
          const uint32_t r = sres[qq];
          uint64_t xm = (ok && qq >= 1) ? z[(qq - 1) * SIGMA + r] : 0ull;
          uint64_t xp = ok ? z[(qq + 1) * SIGMA + r] : 0ull;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint64_t ym = __shfl_up_sync(FULL, xm, o);
            const uint64_t yp = __shfl_up_sync(FULL, xp, o);
            if ((int)lane >= o) {
              xm ^= ym;
              xp ^= yp;
            }
          }
          if (ok) {
            sm[qq] = cm ^ xm;
            sp[qq] = cp ^ xp;
          }
          cm ^= __shfl_sync(FULL, xm, 31);
          cp ^= __shfl_sync(FULL, xp, 31);
        }
        if (lane == 0) {
          sm[L] = 0ull;
          sp[L] = 0ull;
        }
        __syncwarp();
      }

      // ---- phase A: identical + single substitutions (+ deletions + insertions) -------------
      if (part == 0) {
        const uint32_t nsub = S1 * L;
        const uint32_t T = 1 + nsub + (INDELS ? L + SIGMA * (L + 1) : 0);
        for (uint32_t base = 0; base < T; base += 64) {
          uint64_t hv[2], var[2];
          bool pass[2];
#pragma unroll
          for (int u = 0; u < 2; u++) {
            const uint32_t idx = base + u * 32 + lane;
            pass[u] = idx < T;
            hv[u] = h;
            var[u] = pack_variant(VK_IDENTICAL, 0, 0, 0, 0);
            if (pass[u] && idx >= 1) {
              uint32_t t = idx - 1;
              if (t < nsub) {
                const uint32_t pos = t / S1, rp = t - pos * S1;
                const uint32_t r = sub_residue(rp, sres[pos]);
                hv[u] = h ^ zo[pos] ^ z[pos * SIGMA + r];
                var[u] = pack_variant(VK_SUBSTITUTION, pos, r, 0, 0);
              } else if (INDELS) {
                t -= nsub;
                if (t < L) {  // deletion of residue t, only at the start of a run, only if L > 1
                  pass[u] = (L > 1) && (t == 0 || sres[t] != sres[t - 1]);
                  hv[u] = vjh ^ pre[t] ^ sm[t + 1];
                  var[u] = pack_variant(VK_DELETION, t, 0, 0, 0);
                } else {  // insertion of residue r before seed position pos
                  t -= L;
                  const uint32_t pos = t / SIGMA, r = t - pos * SIGMA;
                  pass[u] = (pos == 0) || (r != sres[pos - 1]);
                  hv[u] = vjh ^ pre[pos] ^ z[pos * SIGMA + r] ^ sp[pos];
                  var[u] = pack_variant(VK_INSERTION, pos, r, 0, 0);
                }
              }
            }
          }
          if (use_bloom) {
#pragma unroll
            for (int u = 0; u < 2; u++)
              if (pass[u]) pass[u] = bloom_test(P.bloom, P.bloom_blocks, hv[u], k2);
          }
#pragma unroll
          for (int u = 0; u < 2; u++) {
            npass += pass[u];
            nmatch += queue_push(P, q, lane, pass[u], hv[u], var[u], slocal);
          }
        }
      }

      // ---- phase B: double substitutions i < j ------------------------------------------------
      if (D >= 2) {
        const uint32_t nouter = S1 * L;
        for (uint32_t o = part; o < nouter; o += P.split) {
          const uint32_t i = o / S1, vp = o - i * S1;
          const uint32_t v = sub_residue(vp, sres[i]);
          const uint64_t base2 = h ^ zo[i] ^ z[i * SIGMA + v];
          const uint64_t var_iv = pack_variant(VK_SUB_SUB, i, v, 0, 0);
          const uint32_t ninner = S1 * (L - 1 - i);
          for (uint32_t tb = 0; tb < ninner; tb += 64) {
            uint64_t hv[2];
            uint32_t jw[2];
            bool pass[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
              const uint32_t t = tb + u * 32 + lane;
              pass[u] = t < ninner;
              hv[u] = 0;
              jw[u] = 0;
              if (pass[u]) {
                const uint32_t jj = t / S1, wp = t - jj * S1;
                const uint32_t j = i + 1 + jj;
                const uint32_t w = sub_residue(wp, sres[j]);
                hv[u] = base2 ^ zo[j] ^ z[j * SIGMA + w];
                jw[u] = (j << 8) | w;
              }
            }
            if (use_bloom) {
#pragma unroll
              for (int u = 0; u < 2; u++)
                if (pass[u]) pass[u] = bloom_test(P.bloom, P.bloom_blocks, hv[u], k2);
            }
#pragma unroll
            for (int u = 0; u < 2; u++) {
              npass += pass[u];
              const uint64_t var = var_iv | ((uint64_t)(jw[u] & 0xff) << 16) | ((uint64_t)(jw[u] >> 8) << 44);
              nmatch += queue_push(P, q, lane, pass[u], hv[u], var, slocal);
            }
          }
        }
      }
    }
  }
  __syncwarp();
  if (q.count) nmatch += queue_drain(P, q, q.count);  // q.count < 32 here
  flush_counters(P, nmatch, P.count_bloom ? npass : 0);
}

template <int SIGMA, bool INDELS, int D>
static int launch_variant(const ProbeParams& p, int sm_count, cudaStream_t st, const char** err) {
  const size_t smem = vk_smem_bytes(p.zrows, SIGMA, p.lmax, INDELS);
  if (smem > 200 * 1024) {
    *err = "sequence too long for the shared-memory variant kernel";
    return -1;
  }
  auto kern = variant_kernel<SIGMA, INDELS, D>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
      cudaSuccess) {
    *err = "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed";
    return -1;
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, VK_THREADS, smem) !=
          cudaSuccess ||
      per_sm < 1) {
    *err = "variant kernel does not fit on an SM";
    return -1;
  }
  const uint64_t items = p.a_count * p.split;
  const uint64_t want = (items + VK_WARPS - 1) / VK_WARPS;
  uint64_t grid = (uint64_t)sm_count * per_sm;  // persistent: whole waves of resident CTAs
  if (want < grid) grid = want;
  kern<<<(unsigned)grid, VK_THREADS, smem, st>>>(p);
  return 1;
}

int launch_probe(const ProbeParams& p, int sm_count, cudaStream_t st, const char** err) {
  if (p.a_count == 0) return 0;
  if (p.differences == 0) {
    const uint64_t blocks = (p.a_count + 255) / 256;
    const uint64_t cap = (uint64_t)sm_count * 8;
    identical_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(p);
    return 1;
  }
  if (p.sigma == 20) {
    if (p.differences == 1)
      return p.indels ? launch_variant<20, true, 1>(p, sm_count, st, err)
                      : launch_variant<20, false, 1>(p, sm_count, st, err);
    return launch_variant<20, false, 2>(p, sm_count, st, err);
  }
  if (p.sigma == 4) {
    if (p.differences == 1)
      return p.indels ? launch_variant<4, true, 1>(p, sm_count, st, err)
                      : launch_variant<4, false, 1>(p, sm_count, st, err);
    return launch_variant<4, false, 2>(p, sm_count, st, err);
  }
  *err = "alphabet size must be 4 or 20";
  return -1;
}

}  // namespace cb
