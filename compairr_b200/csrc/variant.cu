// variant.cu — K3/K4 for d = 1 and d = 2: on-the-fly variant enumeration by incremental XOR,
// Bloom prefilter(s), table probe, exact verify, score, matrix accumulation, pair append
// (replaces generate_variants_1/_2, variants.cc:270-400; bloom_get, bloompat.h:55-58;
// find_variant_matches, overlap.cc:168-251; check_variant, variants.cc:166-240).
//
// Structure of both kernels (one warp works on one seed at a time):
//
//   enumerate  every lane decodes VK_U candidates per step from the seed's index spaces and XORs
//              their hashes together from shared-memory Zobrist values
//   filter 1   VK_U independent 8-byte loads of the L2-resident Bloom filter per lane
//   Q1         survivors are compacted (ballot + prefix popcount) into a per-warp ring; when 32
//              wait, their second-level filter words (HBM) are REQUESTED and the previous batch's
//              words, requested one stage earlier, are tested — the HBM latency is never waited on
//   Q2         survivors of filter 2 (or of filter 1 when there is only one level) are compacted
//              again; 32 at a time they walk their probe chains, and verification + atomics run
//              re-converged (device_utils.cuh: probe_chains)
//
// variant1_kernel (d = 1): warps stage batches of 8 consecutive seeds (metadata, hashes, residues)
// into shared memory with coalesced loads, so the per-seed dependent global loads are paid once per
// batch.  variant2_kernel (d = 2): seeds are heavy (~36 000 probes), warps take (seed, part) items
// from a global dispenser.
#include <algorithm>

#include "device_utils.cuh"
#include "kernels.cuh"

namespace cb {

constexpr int VK_THREADS = 256;
constexpr int VK_WARPS = VK_THREADS / 32;
constexpr int VK_QCAP = 64;  // ring entries per queue per warp
constexpr int VK_U = 4;      // probes per lane per step: 4 independent filter loads in flight per lane (2: -8 %, 8: register-bound, 2x slower)
constexpr int VK_WB = 8;     // seeds per warp batch (d = 1)

// Per-warp shared-memory block, addressed from ONE base pointer to keep the register footprint of
// the enumeration loop small (separate pointers per array pushed the kernels over their 80-register
// budget and the queue counters into local memory):
//   [0, 1024)     queue 1: hv[64] u64 | var[64] u32 | seed[64] u32
//   [1024, 2048)  queue 2: same
//   [2048, ...)   seed scratch: zo[lpad] (+ pre[lpad], sm[lpad], sp[lpad] with indels), u64 each
constexpr uint32_t VK_Q_BYTES = VK_QCAP * 16;

struct Ring {  // queue state; the arrays live at wb + which * VK_Q_BYTES
  uint32_t head, count;
};

struct Pend {  // one batch of first-level survivors whose second-level words are in flight
  uint64_t hv;
  unsigned long long w;
  uint32_t var, seed;
  bool valid;
};

struct WarpCtx {
  unsigned char* wb;  // per-warp block
  Ring q1, q2;
  Pend pd;
  uint32_t lane;
  uint32_t nmatch, npass;
};

__device__ __forceinline__ uint64_t* q_hv(unsigned char* wb, int which) {
  return reinterpret_cast<uint64_t*>(wb + which * VK_Q_BYTES);
}
__device__ __forceinline__ uint32_t* q_var(unsigned char* wb, int which) {
  return reinterpret_cast<uint32_t*>(wb + which * VK_Q_BYTES + VK_QCAP * 8);
}
__device__ __forceinline__ uint32_t* q_seed(unsigned char* wb, int which) {
  return reinterpret_cast<uint32_t*>(wb + which * VK_Q_BYTES + VK_QCAP * 12);
}

template <int WHICH>
__device__ __forceinline__ void ring_push(WarpCtx& c, bool pass, uint64_t hv, uint32_t var, uint32_t seed) {
  Ring& q = WHICH ? c.q2 : c.q1;
  const unsigned m = __ballot_sync(FULL, pass);
  if (m == 0) return;
  if (pass) {
    const uint32_t e = (q.head + q.count + __popc(m & ((1u << c.lane) - 1))) & (VK_QCAP - 1);
    q_hv(c.wb, WHICH)[e] = hv;
    q_var(c.wb, WHICH)[e] = var;
    q_seed(c.wb, WHICH)[e] = seed;
  }
  q.count += __popc(m);
  __syncwarp();
}

// Hand n <= 32 queued candidates to the table stage: one cursor bump per warp, coalesced stores
// into the global candidate queue.  The table stage is a separate kernel (table_kernel) — thread
// per candidate, whole GPU's worth of parallelism behind its dependent loads — so the enumeration
// kernel contains no call, no verify code and no matrix atomics, and its registers are its own.
// If the queue is full the entries are dropped and the cursor shows it: the host redoes that
// chunk of seeds in smaller pieces (the enumeration kernel has no other side effect).
__device__ __forceinline__ void q2_drain(const ProbeParams& P, WarpCtx& c, uint32_t n) {
  unsigned long long pos = 0;
  if (c.lane == 0) pos = atomicAdd(P.counters + CTR_GQ, (unsigned long long)n);
  pos = __shfl_sync(FULL, pos, 0);
  if (c.lane < n && pos + n <= P.gq_cap) {
    const uint32_t e = (c.q2.head + c.lane) & (VK_QCAP - 1);
    P.gq_hv[pos + c.lane] = q_hv(c.wb, 1)[e];
    P.gq_vs[pos + c.lane] = make_uint2(q_var(c.wb, 1)[e], q_seed(c.wb, 1)[e]);
  }
  __syncwarp();
  c.q2.head = (c.q2.head + n) & (VK_QCAP - 1);
  c.q2.count -= n;
}

// Second-level stage: test the words requested one stage ago, then pop n entries of Q1 and
// request theirs.
__device__ __forceinline__ void f2_stage(const ProbeParams& P, WarpCtx& c, uint32_t n) {
  const bool pass2 = c.pd.valid && bloom_word_test(c.pd.w, c.pd.hv, false);
  ring_push<1>(c, pass2, c.pd.hv, c.pd.var, c.pd.seed);
  c.pd.valid = c.lane < n;
  if (c.pd.valid) {
    const uint32_t e = (c.q1.head + c.lane) & (VK_QCAP - 1);
    c.pd.hv = q_hv(c.wb, 0)[e];
    c.pd.var = q_var(c.wb, 0)[e];
    c.pd.seed = q_seed(c.wb, 0)[e];
    c.pd.w = __ldg(P.bloom2 + bloom_block(c.pd.hv, P.bloom2_blocks));
  }
  __syncwarp();
  c.q1.head = (c.q1.head + n) & (VK_QCAP - 1);
  c.q1.count -= n;
  if (c.q2.count >= 32) q2_drain(P, c, 32);
}

__device__ __forceinline__ bool is_two_level(const ProbeParams& P) { return P.bloom2 != nullptr && P.use_bloom; }

// One lane-step's verdicts into the pipeline.
__device__ __forceinline__ void submit(const ProbeParams& P, WarpCtx& c, bool pass, uint64_t hv,
                                       uint32_t var, uint32_t seed) {
  c.npass += pass;
  if (is_two_level(P)) {
    ring_push<0>(c, pass, hv, var, seed);
    if (c.q1.count >= 32) f2_stage(P, c, 32);
  } else {
    ring_push<1>(c, pass, hv, var, seed);
    if (c.q2.count >= 32) q2_drain(P, c, 32);
  }
}

__device__ __forceinline__ void finish(const ProbeParams& P, WarpCtx& c) {
  __syncwarp();
  if (is_two_level(P)) {
    f2_stage(P, c, c.q1.count);  // tests the batch in flight, requests the tail of Q1
    f2_stage(P, c, 0);           // tests the tail
  }
  while (c.q2.count) q2_drain(P, c, c.q2.count < 32 ? c.q2.count : 32);
}

// Per-warp scratch for one seed, addressed from the per-warp block.
template <int SIGMA, bool INDELS>
struct SeedScratch {
  uint64_t* zo;    // Z(p, s[p]); the scans follow at multiples of lpad:
  uint32_t lpad;   //   pre[p] = xor_{q<p} Z(q, s[q])        INDELS only: prefix/suffix scans replace the
                   //   sm[p]  = xor_{q>=p} Z(q-1, s[q])     serial incremental walks of
                   //   sp[p]  = xor_{q>=p} Z(q+1, s[q])     variants.cc:311-324,341-353
  __device__ __forceinline__ uint64_t* pre() const { return zo + lpad; }
  __device__ __forceinline__ uint64_t* sm() const { return zo + 2 * lpad; }
  __device__ __forceinline__ uint64_t* sp() const { return zo + 3 * lpad; }
};

// Fill zo[] (and the three scans) for the seed whose residues are at sres[0..L).  Returns VJ.
template <int SIGMA, bool INDELS>
__device__ __forceinline__ uint64_t prepare_seed(const uint64_t* __restrict__ z, const uint8_t* sres,
                                                 uint32_t L, uint64_t h, uint32_t lane,
                                                 SeedScratch<SIGMA, INDELS>& s) {
  for (uint32_t p = lane; p < L; p += 32) s.zo[p] = z[p * SIGMA + sres[p]];
  __syncwarp();
  if (!INDELS) return 0;
  uint64_t carry = 0;
  for (uint32_t base = 0; base < L; base += 32) {
    const uint32_t p = base + lane;
    uint64_t x = p < L ? s.zo[p] : 0ull;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint64_t y = __shfl_up_sync(FULL, x, o);
      if ((int)lane >= o) x ^= y;
    }
    if (p < L) s.pre()[p + 1] = carry ^ x;
    carry ^= __shfl_sync(FULL, x, 31);
  }
  if (lane == 0) s.pre()[0] = 0ull;
  uint64_t cm = 0, cp = 0;
  for (uint32_t base = 0; base < L; base += 32) {  // suffix XORs, walking from the end
    const uint32_t t = base + lane;
    const bool ok = t < L;
    const uint32_t q = ok ? L - 1 - t : 0;
    const uint32_t r = sres[q];
    uint64_t xm = (ok && q >= 1) ? z[(q - 1) * SIGMA + r] : 0ull;
    uint64_t xp = ok ? z[(q + 1) * SIGMA + r] : 0ull;
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
      s.sm()[q] = cm ^ xm;
      s.sp()[q] = cp ^ xp;
    }
    cm ^= __shfl_sync(FULL, xm, 31);
    cp ^= __shfl_sync(FULL, xp, 31);
  }
  if (lane == 0) {
    s.sm()[L] = 0ull;
    s.sp()[L] = 0ull;
  }
  __syncwarp();
  return h ^ carry;  // h = VJ ^ pre[L]
}

// Phase A: identical + single substitutions (+ deletions + insertions); flat index space
// [0, T): 0 identical | (S-1)L substitutions | L deletion candidates | S(L+1) insertion candidates.
template <int SIGMA, bool INDELS>
__device__ __forceinline__ void phase_a(const ProbeParams& P, WarpCtx& c, const uint64_t* __restrict__ z,
                                        const uint8_t* sres, const SeedScratch<SIGMA, INDELS>& s,
                                        uint32_t L, uint64_t h, uint64_t vjh, uint32_t slocal) {
  constexpr uint32_t S1 = SIGMA - 1;
  const uint32_t nsub = S1 * L;
  const uint32_t T = 1 + nsub + (INDELS ? L + SIGMA * (L + 1) : 0);
  for (uint32_t base = 0; base < T; base += 32 * VK_U) {
    uint64_t hv[VK_U];
    uint32_t var[VK_U];
    bool pass[VK_U];
#pragma unroll
    for (int u = 0; u < VK_U; u++) {
      const uint32_t idx = base + u * 32 + c.lane;
      pass[u] = idx < T;
      hv[u] = h;
      var[u] = pack_var(VK_IDENTICAL, 0, 0, 0, 0);
      if (pass[u] && idx >= 1) {
        uint32_t t = idx - 1;
        if (t < nsub) {
          const uint32_t pos = t / S1, rp = t - pos * S1;
          const uint32_t r = sub_residue(rp, sres[pos]);
          hv[u] = h ^ s.zo[pos] ^ z[pos * SIGMA + r];
          var[u] = pack_var(VK_SUBSTITUTION, pos, r, 0, 0);
        } else if (INDELS) {
          t -= nsub;
          if (t < L) {  // deletion of residue t: only at the start of a run, only if L > 1
            pass[u] = (L > 1) && (t == 0 || sres[t] != sres[t - 1]);
            hv[u] = vjh ^ s.pre()[t] ^ s.sm()[t + 1];
            var[u] = pack_var(VK_DELETION, t, 0, 0, 0);
          } else {  // insertion of residue r before seed position pos
            t -= L;
            const uint32_t pos = t / SIGMA, r = t - pos * SIGMA;
            pass[u] = (pos == 0) || (r != sres[pos - 1]);
            hv[u] = vjh ^ s.pre()[pos] ^ z[pos * SIGMA + r] ^ s.sp()[pos];
            var[u] = pack_var(VK_INSERTION, pos, r, 0, 0);
          }
        }
      }
    }
    if (P.use_bloom) {
      unsigned long long w[VK_U];
#pragma unroll
      for (int u = 0; u < VK_U; u++)  // all loads first: VK_U sectors in flight per lane
        w[u] = pass[u] ? __ldg(P.bloom + bloom_block(hv[u], P.bloom_blocks)) : 0ull;
#pragma unroll
      for (int u = 0; u < VK_U; u++) pass[u] = pass[u] && bloom_word_test(w[u], hv[u], P.bloom_k2);
    }
#pragma unroll
    for (int u = 0; u < VK_U; u++) submit(P, c, pass[u], hv[u], var[u], slocal);
  }
}

// Phase B: double substitutions i < j.  Outer (i, v) warp-uniform, lanes over (j > i, w).
template <int SIGMA, bool INDELS>
__device__ __forceinline__ void phase_b(const ProbeParams& P, WarpCtx& c, const uint64_t* __restrict__ z,
                                        const uint8_t* sres, const SeedScratch<SIGMA, INDELS>& s,
                                        uint32_t L, uint64_t h, uint32_t slocal, uint32_t part,
                                        uint32_t split) {
  constexpr uint32_t S1 = SIGMA - 1;
  const uint32_t nouter = S1 * L;
  for (uint32_t o = part; o < nouter; o += split) {
    const uint32_t i = o / S1, vp = o - i * S1;
    const uint32_t v = sub_residue(vp, sres[i]);
    const uint64_t base2 = h ^ s.zo[i] ^ z[i * SIGMA + v];
    const uint32_t var_iv = pack_var(VK_SUB_SUB, i, v, 0, 0);
    const uint32_t ninner = S1 * (L - 1 - i);
    for (uint32_t tb = 0; tb < ninner; tb += 32 * VK_U) {
      uint64_t hv[VK_U];
      uint32_t var[VK_U];
      bool pass[VK_U];
#pragma unroll
      for (int u = 0; u < VK_U; u++) {
        const uint32_t t = tb + u * 32 + c.lane;
        pass[u] = t < ninner;
        hv[u] = 0;
        var[u] = var_iv;
        if (pass[u]) {
          const uint32_t jj = t / S1, wp = t - jj * S1;
          const uint32_t j = i + 1 + jj;
          const uint32_t w = sub_residue(wp, sres[j]);
          hv[u] = base2 ^ s.zo[j] ^ z[j * SIGMA + w];
          var[u] = var_iv | (w << 8) | (j << 22);
        }
      }
      if (P.use_bloom) {
        unsigned long long w[VK_U];
#pragma unroll
        for (int u = 0; u < VK_U; u++)
          w[u] = pass[u] ? __ldg(P.bloom + bloom_block(hv[u], P.bloom_blocks)) : 0ull;
#pragma unroll
        for (int u = 0; u < VK_U; u++) pass[u] = pass[u] && bloom_word_test(w[u], hv[u], P.bloom_k2);
      }
#pragma unroll
      for (int u = 0; u < VK_U; u++) submit(P, c, pass[u], hv[u], var[u], slocal);
    }
  }
}

// ---- shared-memory carve-up (host and device must agree) ------------------------------------------

struct VkLayout {
  uint32_t lpad;        // per-seed scratch entries (>= lmax + 2, multiple of 8)
  size_t z_u64;         // Zobrist rows
  size_t warp_bytes;    // per-warp block: two queues + seed scratch
  size_t blk_u64;       // staged seed batches, all warps: metas (4 u64 each) + hashes
  size_t res_per_warp;  // bytes of staged residues per warp
  size_t total;
};

__host__ __device__ inline VkLayout vk_layout(uint32_t zrows, uint32_t sigma, uint32_t lmax, bool indels,
                                              bool staged) {
  VkLayout l;
  l.lpad = (lmax + 2 + 7) & ~7u;
  l.z_u64 = (size_t)zrows * sigma;
  l.warp_bytes = 2 * VK_Q_BYTES + (size_t)l.lpad * (indels ? 4 : 1) * 8;
  l.blk_u64 = staged ? (size_t)VK_WARPS * VK_WB * 5 : 0;
  l.res_per_warp = staged ? (((size_t)VK_WB * lmax + 15) & ~(size_t)15) : l.lpad;
  l.total = l.z_u64 * 8 + VK_WARPS * l.warp_bytes + l.blk_u64 * 8 + VK_WARPS * l.res_per_warp;
  return l;
}

template <int SIGMA, bool INDELS>
__device__ __forceinline__ void carve_warp(const VkLayout& l, unsigned char* smem, uint32_t warp, uint32_t lane,
                                           uint64_t*& z, SeedScratch<SIGMA, INDELS>& s, WarpCtx& c,
                                           uint64_t*& blk, uint8_t*& bytes) {
  z = reinterpret_cast<uint64_t*>(smem);
  c.wb = smem + l.z_u64 * 8 + warp * l.warp_bytes;
  s.zo = reinterpret_cast<uint64_t*>(c.wb + 2 * VK_Q_BYTES);
  s.lpad = l.lpad;
  blk = reinterpret_cast<uint64_t*>(smem + l.z_u64 * 8 + VK_WARPS * l.warp_bytes);
  bytes = reinterpret_cast<uint8_t*>(blk + l.blk_u64);
  c.q1.head = c.q1.count = c.q2.head = c.q2.count = 0;
  c.pd.valid = false;
  c.pd.hv = 0;
  c.pd.w = 0;
  c.pd.var = c.pd.seed = 0;
  c.lane = lane;
  c.nmatch = c.npass = 0;
}

// ---- d = 1 -------------------------------------------------------------------------------------------
//
// Warps take batches of VK_WB consecutive seeds from a global dispenser and stage the batch's
// metadata, hashes and residues (contiguous in the arena) into per-warp shared memory with
// coalesced loads: the dependent global loads (dispenser -> metadata -> residues) are paid once
// per batch, not once per seed, and nothing waits on a CTA-wide barrier (a block-synchronous
// variant lost a third of its time at __syncthreads behind whichever warp was in the slow path,
// profiles/r01_d_*).

template <int SIGMA, bool INDELS>
__global__ void __launch_bounds__(VK_THREADS, 3) variant1_kernel(const __grid_constant__ ProbeParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const VkLayout lay = vk_layout(P.zrows, SIGMA, P.lmax, INDELS, true);
  uint64_t* z;
  SeedScratch<SIGMA, INDELS> sc;
  WarpCtx c;
  uint64_t* blk;
  uint8_t* bytes;
  carve_warp<SIGMA, INDELS>(lay, smem_raw, warp, lane, z, sc, c, blk, bytes);
  uint64_t* const my = blk + warp * (VK_WB * 5);                 // this warp's staging area
  SeqRec* const b_meta = reinterpret_cast<SeqRec*>(my);          // VK_WB records of 32 B
  uint64_t* const b_hash = my + VK_WB * 4;
  uint8_t* const b_res = bytes + warp * lay.res_per_warp;

  for (uint32_t i = threadIdx.x; i < P.zrows * SIGMA; i += VK_THREADS) z[i] = P.ztab[i];
  __syncthreads();
  const uint64_t n_batches = (P.w_count + VK_WB - 1) / VK_WB;

  for (;;) {
    unsigned long long b = 0;
    if (lane == 0) b = atomicAdd(P.counters + CTR_WORK, 1ull);
    b = __shfl_sync(FULL, b, 0);
    if (b >= n_batches) break;
    const uint64_t first = P.w_first + b * VK_WB;  // relative to a_first
    const uint32_t nb = (uint32_t)((P.w_first + P.w_count - first < VK_WB) ? P.w_first + P.w_count - first : VK_WB);
    __syncwarp();  // previous batch fully consumed
    if (lane < nb * 2)
      reinterpret_cast<uint4*>(b_meta)[lane] =
          __ldg(reinterpret_cast<const uint4*>(P.a.meta + P.a_first + first) + lane);
    if (lane >= 16 && lane < 16 + nb) b_hash[lane - 16] = __ldg(P.a.hash + P.a_first + first + (lane - 16));
    __syncwarp();
    const uint64_t res0 = b_meta[0].off_len & ((1ull << 40) - 1);
    const uint64_t last = b_meta[nb - 1].off_len;
    const uint32_t res_n = (uint32_t)((last & ((1ull << 40) - 1)) + (last >> 40) - res0);
    for (uint32_t i = lane; i < res_n; i += 32) b_res[i] = __ldg(P.a.res + res0 + i);
    __syncwarp();

    for (uint32_t k = 0; k < nb; k++) {
      const uint64_t off_len = b_meta[k].off_len;  // the enumeration needs only offset and length
      const uint32_t L = (uint32_t)(off_len >> 40);
      const uint64_t h = b_hash[k];
      const uint8_t* sres = b_res + (uint32_t)((off_len & ((1ull << 40) - 1)) - res0);
      __syncwarp();  // all lanes are done with the previous seed's scratch
      const uint64_t vjh = prepare_seed<SIGMA, INDELS>(z, sres, L, h, lane, sc);
      phase_a<SIGMA, INDELS>(P, c, z, sres, sc, L, h, vjh, (uint32_t)(first + k));
    }
  }
  finish(P, c);
  flush_counters(P, 0, P.count_bloom ? c.npass : 0);
}

// ---- d = 2 -------------------------------------------------------------------------------------------

template <int SIGMA>
__global__ void __launch_bounds__(VK_THREADS, 3) variant2_kernel(const __grid_constant__ ProbeParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const VkLayout lay = vk_layout(P.zrows, SIGMA, P.lmax, false, false);
  uint64_t* z;
  SeedScratch<SIGMA, false> sc;
  WarpCtx c;
  uint64_t* blk;
  uint8_t* bytes;
  carve_warp<SIGMA, false>(lay, smem_raw, warp, lane, z, sc, c, blk, bytes);
  uint8_t* const sres = bytes + warp * lay.res_per_warp;

  for (uint32_t i = threadIdx.x; i < P.zrows * SIGMA; i += VK_THREADS) z[i] = P.ztab[i];
  __syncthreads();

  const uint64_t total_items = P.w_count * P.split;
  const uint32_t split_mask = P.split - 1;
  const uint32_t split_shift = 31 - __clz(P.split);
  for (;;) {
    unsigned long long item = 0;
    if (lane == 0) item = atomicAdd(P.counters + CTR_WORK, 1ull);
    item = __shfl_sync(FULL, item, 0);
    if (item >= total_items) break;
    const uint32_t slocal = (uint32_t)(P.w_first + (item >> split_shift));
    const uint32_t part = (uint32_t)item & split_mask;
    const uint64_t sidx = P.a_first + slocal;
    const uint64_t off_len = __ldg(&P.a.meta[sidx].off_len);  // same address in all lanes: one broadcast
    const uint32_t L = (uint32_t)(off_len >> 40);
    const uint64_t h = __ldg(P.a.hash + sidx);
    __syncwarp();
    for (uint32_t p = lane; p < L; p += 32) sres[p] = __ldg(P.a.res + (off_len & ((1ull << 40) - 1)) + p);
    __syncwarp();
    prepare_seed<SIGMA, false>(z, sres, L, h, lane, sc);
    if (part == 0) phase_a<SIGMA, false>(P, c, z, sres, sc, L, h, 0, slocal);
    phase_b<SIGMA, false>(P, c, z, sres, sc, L, h, slocal, part, P.split);
  }
  finish(P, c);
  flush_counters(P, 0, P.count_bloom ? c.npass : 0);
}

// ---- K4: the table stage ----------------------------------------------------------------------------
//
// One thread per queued candidate (hash, variant, seed): probe chain, exact verify against the head
// of the occurrence list, score + matrix atomics + pair append per occurrence.  A queue that
// overflowed is not touched at all: the chunk is recorded and redone by the host.
__global__ void __launch_bounds__(256) table_kernel(const __grid_constant__ ProbeParams P, uint32_t chunk_id) {
  const unsigned long long filled = P.counters[CTR_GQ];
  if (filled > P.gq_cap) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      const unsigned long long k = atomicAdd(P.counters + CTR_OVERFLOW, 1ull);
      if (k < 64) P.overflow_chunks[k] = chunk_id;
    }
    return;
  }
  uint32_t nmatch = 0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x + (threadIdx.x & ~31u); i0 < filled; i0 += stride) {
    const uint64_t i = i0 + (threadIdx.x & 31);
    const bool act = i < filled;
    uint64_t hv = 0;
    uint2 vs = make_uint2(0, 0);
    if (act) {
      hv = P.gq_hv[i];
      vs = P.gq_vs[i];
    }
    nmatch += probe_chains(&P, act, hv, vs.x, P.a_first + vs.y, vs.y, nullptr, 0);
  }
  flush_counters(P, nmatch, 0);
}

void launch_table_stage(const ProbeParams& p, int sm_count, uint32_t chunk_id, cudaStream_t st) {
  table_kernel<<<sm_count * 8, 256, 0, st>>>(p, chunk_id);
}

// ---- launch ------------------------------------------------------------------------------------------

template <typename K>
static int launch_one(K kern, const ProbeParams& p, size_t smem, uint64_t work_ctas, int sm_count,
                      cudaStream_t st, const char** err) {
  if (smem > 200 * 1024) {
    *err = "sequence too long for the shared-memory variant kernels";
    return -1;
  }
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    *err = "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed";
    return -1;
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, VK_THREADS, smem) != cudaSuccess || per_sm < 1) {
    *err = "variant kernel does not fit on an SM";
    return -1;
  }
  // The rate of random 8-byte loads an SM sustains has a cliff in the shared-memory carve-out
  // (tools/bench_l2_random.cu on B200: 268-290 G loads/s chip-wide up to a 100 KB carve-out,
  // 139-158 G loads/s from 132 KB on, whatever the occupancy) — the L1 side that tracks the
  // misses in flight shrinks with it.  The Bloom stage is nothing but such loads, so: no more
  // resident CTAs than fit 100 KB of shared memory (1 KB per CTA is the system's), and ask for
  // exactly that carve-out instead of the driver's "room for the most CTAs" default.
  constexpr size_t kCarveCliff = 100 * 1024;
  const int fit = (int)(kCarveCliff / (smem + 1024));
  if (fit >= 2 || (fit == 1 && per_sm == 1)) {
    per_sm = std::min(per_sm, fit);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)(kCarveCliff * 100 / (228 * 1024)));
  }
  uint64_t grid = (uint64_t)sm_count * per_sm;  // persistent: whole waves of resident CTAs
  if (work_ctas < grid) grid = work_ctas;
  kern<<<(unsigned)grid, VK_THREADS, smem, st>>>(p);
  return 1;
}

int launch_variant_kernels(const ProbeParams& p_in, int sm_count, cudaStream_t st, const char** err) {
  // only the Zobrist rows the longest seed can touch are staged in shared memory (positions
  // 0..lmax, +1 for the shifted rows of the indel scans): the table itself may be longer
  ProbeParams p = p_in;
  p.zrows = std::min(p_in.zrows, p_in.lmax + 2);
  if (p.lmax + 1 > VAR_MAX_POS) {
    *err = "sequence longer than 510 residues on the d<=2 path";
    return -1;
  }
  if (p.sigma != 4 && p.sigma != 20) {
    *err = "alphabet size must be 4 or 20";
    return -1;
  }
  if (p.differences == 1) {
    const size_t smem = vk_layout(p.zrows, p.sigma, p.lmax, p.indels, true).total;
    const uint64_t blocks = ((p.w_count + VK_WB - 1) / VK_WB + VK_WARPS - 1) / VK_WARPS;
    if (p.sigma == 20)
      return p.indels ? launch_one(variant1_kernel<20, true>, p, smem, blocks, sm_count, st, err)
                      : launch_one(variant1_kernel<20, false>, p, smem, blocks, sm_count, st, err);
    return p.indels ? launch_one(variant1_kernel<4, true>, p, smem, blocks, sm_count, st, err)
                    : launch_one(variant1_kernel<4, false>, p, smem, blocks, sm_count, st, err);
  }
  const size_t smem = vk_layout(p.zrows, p.sigma, p.lmax, false, false).total;
  const uint64_t ctas = (p.w_count * p.split + VK_WARPS - 1) / VK_WARPS;
  return p.sigma == 20 ? launch_one(variant2_kernel<20>, p, smem, ctas, sm_count, st, err)
                       : launch_one(variant2_kernel<4>, p, smem, ctas, sm_count, st, err);
}

}  // namespace cb
