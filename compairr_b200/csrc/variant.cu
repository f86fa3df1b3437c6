// variant.cu — K3/K4 for d = 1 and d = 2: on-the-fly variant enumeration by incremental XOR,
// Bloom prefilter(s), table probe, exact verify, score, matrix accumulation, pair append
// (replaces generate_variants_1/_2, variants.cc:270-400; bloom_get, bloompat.h:55-58;
// find_variant_matches, overlap.cc:168-251; check_variant, variants.cc:166-240).
//
// Structure of both kernels (one warp works on one seed at a time):
//
//   enumerate  every lane decodes VK_U candidates per step from the seed's index spaces and XORs
//              their hashes together from shared-memory Zobrist values
//   filter     VK_U independent 8-byte loads per lane from the PARITY FILTERS (common.cuh): the
//              filter is chosen by the parity of the candidate's free position, which makes the
//              word address the same for all candidates at positions of that parity — the 32
//              loads of a warp step fall into one or two sectors and are served by L1 after the
//              first touch, instead of 32 random L2 sectors
//   queue      survivors are compacted (ballot + prefix popcount) into a per-warp ring in shared
//              memory; 32 at a time they are appended, coalesced, to the global candidate queue
//              that the table kernel (K4, own launch) drains
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
constexpr int VK_QCAP = 64;  // ring entries per warp
constexpr int VK_U = 4;      // probes per lane per step: 4 independent filter loads in flight per lane
constexpr int VK_WB = 8;     // seeds per warp batch (d = 1)
#ifndef VK_D1_CTAS
#define VK_D1_CTAS 3          // resident CTAs per SM the d = 1 kernel is compiled for
#endif

// Per-warp shared-memory block, addressed from ONE base pointer to keep the register footprint of
// the enumeration loop small:
//   [0, 1024)     survivor ring: hv[64] u64 | var[64] u32 | seed[64] u32
//   [1024, ...)   seed scratch: zo[lpad] (+ pre[lpad], sm[lpad], sp[lpad] with indels), u64 each
constexpr uint32_t VK_Q_BYTES = VK_QCAP * 16;

struct WarpCtx {
  unsigned char* wb;  // per-warp block
  uint32_t head, count;  // ring state
  uint32_t lane;
};

__device__ __forceinline__ uint64_t* q_hv(unsigned char* wb) { return reinterpret_cast<uint64_t*>(wb); }
__device__ __forceinline__ uint32_t* q_var(unsigned char* wb) { return reinterpret_cast<uint32_t*>(wb + VK_QCAP * 8); }
__device__ __forceinline__ uint32_t* q_seed(unsigned char* wb) { return reinterpret_cast<uint32_t*>(wb + VK_QCAP * 12); }

// mkvar() builds the 31-bit variant descriptor; it runs for survivors only (a fraction of a percent
// of the candidates), so the enumeration loop itself never packs one.
template <typename MkVar>
__device__ __forceinline__ void ring_push(WarpCtx& c, bool pass, uint64_t hv, MkVar mkvar, uint32_t seed) {
  const unsigned m = __ballot_sync(FULL, pass);
  if (m == 0) return;
  if (pass) {
    const uint32_t e = (c.head + c.count + __popc(m & ((1u << c.lane) - 1))) & (VK_QCAP - 1);
    q_hv(c.wb)[e] = hv;
    q_var(c.wb)[e] = mkvar();
    q_seed(c.wb)[e] = seed;
  }
  c.count += __popc(m);
  __syncwarp();
}

// Hand n <= 32 queued candidates to the table stage: one cursor bump per warp, coalesced stores
// into the global candidate queue.  The table stage is a separate kernel (table_kernel) — thread
// per candidate, whole GPU's worth of parallelism behind its dependent loads — so the enumeration
// kernel contains no call, no verify code and no matrix atomics, and its registers are its own.
// If the queue is full the entries are dropped and the cursor shows it: the host redoes that
// chunk of seeds in smaller pieces (the enumeration kernel has no other side effect).
__device__ __forceinline__ void ring_drain(const ProbeParams& P, WarpCtx& c, uint32_t n) {
  unsigned long long pos = 0;
  if (c.lane == 0) pos = atomicAdd(P.counters + CTR_GQ, (unsigned long long)n);
  pos = __shfl_sync(FULL, pos, 0);
  if (c.lane < n && pos + n <= P.gq_cap) {
    const uint32_t e = (c.head + c.lane) & (VK_QCAP - 1);
    P.gq_hv[pos + c.lane] = q_hv(c.wb)[e];
    P.gq_vs[pos + c.lane] = make_uint2(q_var(c.wb)[e], q_seed(c.wb)[e]);
  }
  __syncwarp();
  c.head = (c.head + n) & (VK_QCAP - 1);
  c.count -= n;
}

// One lane-step's verdicts into the pipeline.
template <typename MkVar>
__device__ __forceinline__ void submit(const ProbeParams& P, WarpCtx& c, bool pass, uint64_t hv,
                                       MkVar mkvar, uint32_t seed) {
  ring_push(c, pass, hv, mkvar, seed);
  if (c.count >= 32) ring_drain(P, c, 32);
}

__device__ __forceinline__ void finish(const ProbeParams& P, WarpCtx& c) {
  __syncwarp();
  while (c.count) ring_drain(P, c, c.count < 32 ? c.count : 32);
}

// VK_U filter lookups per lane: all loads first (VK_U words in flight per lane), then the tests.
// odd[u] = parity of candidate u's free position (which filter, common.cuh).
__device__ __forceinline__ void filter_step(const ProbeParams& P, const uint64_t (&hv)[VK_U],
                                            const bool (&odd)[VK_U], bool (&pass)[VK_U]) {
  if (!P.use_bloom) return;
  unsigned long long w[VK_U];
#pragma unroll
  for (int u = 0; u < VK_U; u++)  // unconditional: an inactive candidate's hash is a valid address too
    w[u] = __ldg(P.bloom + pfilter_word(hv[u], P.bloom_blocks, odd[u]));
#pragma unroll
  for (int u = 0; u < VK_U; u++)
    pass[u] = pass[u] & pattern_hit(w[u], odd[u] ? field_odd(hv[u]) : field_even(hv[u]));
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
  // d = 1 kernel only, after the scans (finalize_seed): everything a candidate needs, per "slot"
  // pp = position for substitutions (pp < L), L + position for insertions (pp in [L, 2L]):
  //   base2[pp]  hash of the variant minus the Zobrist value of its free residue
  //   word2[pp]  the parity-filter word that every residue at this slot is looked up in
  //   cmp2[pp]   the residue that must NOT be put there (substitution: the seed's own; insertion:
  //              the residue before it, variants.cc:341-353; 255 = none)
  // and for deletions (t < L) / the identical candidate (t = L, or 0 without indels):
  //   dh[t], wd[t]  variant hash and its filter word
  static constexpr uint32_t kScan = INDELS ? 4 : 1, kSlots = INDELS ? 2 : 1;
  __device__ __forceinline__ uint64_t* base2() const { return zo + kScan * lpad; }
  __device__ __forceinline__ uint64_t* word2() const { return zo + (kScan + kSlots) * lpad; }
  __device__ __forceinline__ uint64_t* dh() const { return zo + (kScan + 2 * kSlots) * lpad; }
  __device__ __forceinline__ uint64_t* wd() const { return zo + (kScan + 2 * kSlots + 1) * lpad; }
  __device__ __forceinline__ uint8_t* cmp2() const { return reinterpret_cast<uint8_t*>(zo + (kScan + 2 * kSlots + 2) * lpad); }
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
    bool pass[VK_U], odd[VK_U];
#pragma unroll
    for (int u = 0; u < VK_U; u++) {
      const uint32_t idx = base + u * 32 + c.lane;
      pass[u] = idx < T;
      hv[u] = h;
      odd[u] = true;
      var[u] = pack_var(VK_IDENTICAL, 0, 0, 0, 0);
      if (pass[u] && idx >= 1) {
        uint32_t t = idx - 1;
        if (t < nsub) {
          const uint32_t pos = t / S1, rp = t - pos * S1;
          const uint32_t r = sub_residue(rp, sres[pos]);
          hv[u] = h ^ s.zo[pos] ^ z[pos * SIGMA + r];
          odd[u] = pos & 1;
          var[u] = pack_var(VK_SUBSTITUTION, pos, r, 0, 0);
        } else if (INDELS) {
          t -= nsub;
          if (t < L) {  // deletion of residue t: only at the start of a run, only if L > 1
            pass[u] = (L > 1) && (t == 0 || sres[t] != sres[t - 1]);
            hv[u] = vjh ^ s.pre()[t] ^ s.sm()[t + 1];
            odd[u] = t & 1;  // no free residue: either filter, alternate for balance
            var[u] = pack_var(VK_DELETION, t, 0, 0, 0);
          } else {  // insertion of residue r before seed position pos
            t -= L;
            const uint32_t pos = t / SIGMA, r = t - pos * SIGMA;
            pass[u] = (pos == 0) || (r != sres[pos - 1]);
            hv[u] = vjh ^ s.pre()[pos] ^ z[pos * SIGMA + r] ^ s.sp()[pos];
            odd[u] = pos & 1;  // the inserted residue sits at position pos of the variant
            var[u] = pack_var(VK_INSERTION, pos, r, 0, 0);
          }
        }
      }
    }
    filter_step(P, hv, odd, pass);
#pragma unroll
    for (int u = 0; u < VK_U; u++) {
      const uint32_t v = var[u];
      submit(P, c, pass[u], hv[u], [v] { return v; }, slocal);
    }
  }
}

// d = 1.  Everything a candidate needs is laid out per slot first (SeedScratch); the filter words
// are fetched once per slot, because a slot's word does not depend on the residue placed there
// (parity filters, common.cuh), and the enumeration loops then run on shared memory and registers:
//   * the two words of ALL substitution slots (even / odd positions) depend only on the seed's
//     hash: they are fetched when the batch is staged (variant1_kernel) and arrive as ws_even/odd;
//   * the words of the insertion and deletion slots need the scans: their loads are ISSUED before
//     the substitution loop and CONSUMED after it, so their latency hides behind ~45 % of the
//     seed's work instead of stalling the warp.
template <int SIGMA, bool INDELS, int U>
__device__ __forceinline__ void enumerate_slots(const ProbeParams& P, WarpCtx& c, const uint64_t* __restrict__ z,
                                                const SeedScratch<SIGMA, INDELS>& s, uint32_t L,
                                                uint32_t q0, uint32_t q1, bool subs,
                                                unsigned long long ws_even, unsigned long long ws_odd,
                                                uint32_t slocal) {
  // candidates q in [q0, q1) of the slot space: slot pp = q / SIGMA, residue r = q % SIGMA
  for (uint32_t base = q0; base < q1; base += 32 * U) {
    uint64_t hv[U];
    uint32_t code[U];
    bool pass[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint32_t q = base + u * 32 + c.lane;
      const bool in = q < q1;
      const uint32_t qq = in ? q : q0;
      const uint32_t pp = qq / SIGMA, r = qq - pp * SIGMA;
      const uint32_t pos = subs ? pp : pp - L;
      const uint64_t b = s.base2()[pp];
      const uint32_t cmp = s.cmp2()[pp];
      const unsigned long long w = subs ? ((pos & 1) ? ws_odd : ws_even) : s.word2()[pp];
      hv[u] = b ^ z[pos * SIGMA + r];
      code[u] = qq;
      pass[u] = in & (r != cmp) & pattern_hit(w, (pos & 1) ? field_odd(hv[u]) : field_even(hv[u]));
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint32_t qq = code[u];
      submit(P, c, pass[u], hv[u], [qq, L] {
        const uint32_t pp = qq / SIGMA, r = qq - pp * SIGMA;
        return pp < L ? pack_var(VK_SUBSTITUTION, pp, r, 0, 0) : pack_var(VK_INSERTION, pp - L, r, 0, 0);
      }, slocal);
    }
  }
}

// Two candidates per lane per step: the loop has no global load left to overlap, and a small body
// matters more — with four (and a separate tail loop) a quarter of all stall samples were
// instruction-cache misses (profiles/r01_h_*).
#ifndef VK_U1
#define VK_U1 2
#endif
template <int SIGMA, bool INDELS>
__device__ __forceinline__ void enumerate_range(const ProbeParams& P, WarpCtx& c, const uint64_t* __restrict__ z,
                                                const SeedScratch<SIGMA, INDELS>& s, uint32_t L,
                                                uint32_t q0, uint32_t q1, bool subs,
                                                unsigned long long ws_even, unsigned long long ws_odd,
                                                uint32_t slocal) {
  enumerate_slots<SIGMA, INDELS, VK_U1>(P, c, z, s, L, q0, q1, subs, ws_even, ws_odd, slocal);
}

template <int SIGMA, bool INDELS>
__device__ __forceinline__ void seed_d1(const ProbeParams& P, WarpCtx& c, const uint64_t* __restrict__ z,
                                        const uint8_t* sres, const SeedScratch<SIGMA, INDELS>& s,
                                        uint32_t L, uint64_t h, uint64_t vjh,
                                        unsigned long long ws_even, unsigned long long ws_odd,
                                        uint32_t slocal) {
  const bool filt = P.use_bloom;
  const uint32_t lane = c.lane;
  // substitution slots
  for (uint32_t pp = lane; pp < L; pp += 32) {
    s.base2()[pp] = h ^ s.zo[pp];
    s.cmp2()[pp] = sres[pp];
  }
  // insertion slots (pp = L + pos) and deletions: bases now, words in flight.  The first 32 of
  // each are prefetched into registers; longer seeds fetch the rest when the words are stored.
  unsigned long long wi0 = ~0ull, wd0 = ~0ull;
  if (INDELS) {
    for (uint32_t pos = lane; pos <= L; pos += 32) {
      const uint64_t b = vjh ^ s.pre()[pos] ^ s.sp()[pos];
      s.base2()[L + pos] = b;
      s.cmp2()[L + pos] = (uint8_t)(pos == 0 ? 255u : sres[pos - 1]);
      // the inserted residue sits at position pos: it cannot change the field that picks the word
      if (pos < 32 && filt) wi0 = __ldg(P.bloom + pfilter_word(b, P.bloom_blocks, pos & 1));
    }
    for (uint32_t t = lane; t < L; t += 32) {
      const uint64_t hv = vjh ^ s.pre()[t] ^ s.sm()[t + 1];
      s.dh()[t] = hv;
      if (t < 32 && filt) wd0 = __ldg(P.bloom + pfilter_word(hv, P.bloom_blocks, true));
    }
  }
  __syncwarp();
  enumerate_range<SIGMA, INDELS>(P, c, z, s, L, 0, L * SIGMA, true, ws_even, ws_odd, slocal);
  if (INDELS) {
    for (uint32_t pos = lane; pos <= L; pos += 32)
      s.word2()[L + pos] = pos < 32 ? wi0
                           : (filt ? __ldg(P.bloom + pfilter_word(s.base2()[L + pos], P.bloom_blocks, pos & 1)) : ~0ull);
    for (uint32_t t = lane; t < L; t += 32)
      s.wd()[t] = t < 32 ? wd0 : (filt ? __ldg(P.bloom + pfilter_word(s.dh()[t], P.bloom_blocks, true)) : ~0ull);
    __syncwarp();
    enumerate_range<SIGMA, INDELS>(P, c, z, s, L, L * SIGMA, (2 * L + 1) * SIGMA, false, ws_even, ws_odd, slocal);
  }
  // deletions: one per run of equal residues, only if L > 1 (variants.cc:301-325); candidate
  // t = nd - 1 is the identical one (filter E, whose word for the seed's own hash is ws_odd)
  const uint32_t nd = INDELS ? L + 1 : 1;
  for (uint32_t t0 = 0; t0 < nd; t0 += 32) {
    const uint32_t t = t0 + lane;
    const bool in = t < nd;
    const bool is_del = INDELS && t < L;
    const uint32_t tt = is_del ? t : 0u;
    const bool valid = in && (!is_del || (L > 1 && (tt == 0 || sres[tt] != sres[tt - 1])));
    const uint64_t hv = is_del ? s.dh()[tt] : h;
    const unsigned long long w = is_del ? s.wd()[tt] : ws_odd;
    const bool pass = valid & pattern_hit(w, field_odd(hv));
    submit(P, c, pass, hv, [is_del, tt] {
      return is_del ? pack_var(VK_DELETION, tt, 0, 0, 0) : pack_var(VK_IDENTICAL, 0, 0, 0, 0);
    }, slocal);
  }
}

// Phase B: double substitutions i < j (variants.cc:357-400).  Outer (i, v) warp-uniform, lanes over
// the slots (j > i, r) of the second substitution: SIGMA candidates per j with the seed's own
// residue masked, hash = base2 ^ Z(j, s[j]) ^ Z(j, r).  The second substitution cannot change the
// filter word: for odd j all candidates read filter E's word of base2, for even j filter O's — two
// words per outer iteration, fetched one iteration AHEAD so that their latency never stalls the
// warp.
template <int SIGMA, bool INDELS>
__device__ __forceinline__ void phase_b(const ProbeParams& P, WarpCtx& c, const uint64_t* __restrict__ z,
                                        const uint8_t* sres, const SeedScratch<SIGMA, INDELS>& s,
                                        uint32_t L, uint64_t h, uint32_t slocal, uint32_t part,
                                        uint32_t split) {
  constexpr uint32_t S1 = SIGMA - 1;
  const uint32_t nouter = S1 * L;
  const bool filt = P.use_bloom;
  auto outer = [&](uint32_t o, uint32_t& i, uint32_t& v, uint64_t& b2, unsigned long long& we, unsigned long long& wo) {
    i = o / S1;
    v = sub_residue(o - i * S1, sres[i]);
    b2 = h ^ s.zo[i] ^ z[i * SIGMA + v];
    we = filt ? __ldg(P.bloom + pfilter_word(b2, P.bloom_blocks, true)) : ~0ull;
    wo = filt ? __ldg(P.bloom + pfilter_word(b2, P.bloom_blocks, false)) : ~0ull;
  };
  uint32_t i = 0, v = 0, ni = 0, nv = 0;
  uint64_t b2 = 0, nb2 = 0;
  unsigned long long we = 0, wo = 0, nwe = 0, nwo = 0;
  if (part < nouter) outer(part, ni, nv, nb2, nwe, nwo);
  for (uint32_t o = part; o < nouter; o += split) {
    i = ni; v = nv; b2 = nb2; we = nwe; wo = nwo;
    if (o + split < nouter) outer(o + split, ni, nv, nb2, nwe, nwo);
    const uint32_t ninner = SIGMA * (L - 1 - i);
    for (uint32_t tb = 0; tb < ninner; tb += 32) {
      const uint32_t t = tb + c.lane;
      const bool in = t < ninner;
      const uint32_t tt = in ? t : 0u;
      const uint32_t jj = tt / SIGMA, r = tt - jj * SIGMA;
      const uint32_t j = in ? i + 1 + jj : i;  // inactive lanes read a valid row
      const uint32_t cmp = sres[j];
      const uint64_t hv = b2 ^ s.zo[j] ^ z[j * SIGMA + r];
      const bool pass = in & (r != cmp) & pattern_hit((j & 1) ? we : wo, (j & 1) ? field_odd(hv) : field_even(hv));
      submit(P, c, pass, hv, [i, v, j, r] { return pack_var(VK_SUB_SUB, i, v, j, r); }, slocal);
    }
  }
}

// ---- shared-memory carve-up (host and device must agree) ------------------------------------------

struct VkLayout {
  uint32_t lpad;        // per-seed scratch entries (>= lmax + 2, multiple of 8)
  size_t z_u64;         // Zobrist rows
  size_t warp_bytes;    // per-warp block: survivor ring + seed scratch
  size_t blk_u64;       // staged seed batches, all warps: metas (4 u64 each) + hashes
  size_t res_per_warp;  // bytes of staged residues per warp
  size_t total;
};

__host__ __device__ inline VkLayout vk_layout(uint32_t zrows, uint32_t sigma, uint32_t lmax, bool indels,
                                              bool staged) {
  VkLayout l;
  l.lpad = (lmax + 2 + 7) & ~7u;
  l.z_u64 = (size_t)zrows * sigma;
  // scratch in u64 units of lpad: scans (4 or 1); d = 1 adds base2 + word2 (2 or 1 each), dh, wd, cmp2 (bytes, <= 1)
  const size_t units = (indels ? 4 : 1) + (staged ? (indels ? 4 : 2) + 2 + 1 : 0);
  l.warp_bytes = VK_Q_BYTES + (size_t)l.lpad * units * 8;
  l.blk_u64 = staged ? (size_t)VK_WARPS * VK_WB * 7 : 0;
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
  s.zo = reinterpret_cast<uint64_t*>(c.wb + VK_Q_BYTES);
  s.lpad = l.lpad;
  blk = reinterpret_cast<uint64_t*>(smem + l.z_u64 * 8 + VK_WARPS * l.warp_bytes);
  bytes = reinterpret_cast<uint8_t*>(blk + l.blk_u64);
  c.head = c.count = 0;
  c.lane = lane;
}

// ---- d = 1 -------------------------------------------------------------------------------------------
//
// Warps take batches of VK_WB consecutive seeds from a global dispenser and stage the batch's
// metadata, hashes and residues (contiguous in the arena) into per-warp shared memory with
// coalesced loads: the dependent global loads (dispenser -> metadata -> residues) are paid once
// per batch, not once per seed, and nothing waits on a CTA-wide barrier (a block-synchronous
// variant lost a third of its time at __syncthreads behind whichever warp was in the slow path,
// measured this round).

template <int SIGMA, bool INDELS>
__global__ void __launch_bounds__(VK_THREADS, VK_D1_CTAS) variant1_kernel(const __grid_constant__ ProbeParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const VkLayout lay = vk_layout(P.zrows, SIGMA, P.lmax, INDELS, true);
  uint64_t* z;
  SeedScratch<SIGMA, INDELS> sc;
  WarpCtx c;
  uint64_t* blk;
  uint8_t* bytes;
  carve_warp<SIGMA, INDELS>(lay, smem_raw, warp, lane, z, sc, c, blk, bytes);
  uint64_t* const my = blk + warp * (VK_WB * 7);                 // this warp's staging area
  SeqRec* const b_meta = reinterpret_cast<SeqRec*>(my);          // VK_WB records of 32 B
  uint64_t* const b_hash = my + VK_WB * 4;
  unsigned long long* const b_ws = reinterpret_cast<unsigned long long*>(my + VK_WB * 5);  // [seed][even, odd]
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
    // hashes, and with them the two filter words of each seed's substitution slots: filter O for
    // even positions, filter E for odd ones (the word index ignores the substituted residue)
    unsigned long long ws = ~0ull;
    if (lane < 2 * nb) {
      const uint64_t hk = __ldg(P.a.hash + P.a_first + first + (lane >> 1));
      if (lane & 1) b_hash[lane >> 1] = hk;
      if (P.use_bloom) ws = __ldg(P.bloom + pfilter_word(hk, P.bloom_blocks, lane & 1));
    }
    __syncwarp();
    const uint64_t res0 = b_meta[0].off_len & ((1ull << 40) - 1);
    const uint64_t last = b_meta[nb - 1].off_len;
    const uint32_t res_n = (uint32_t)((last & ((1ull << 40) - 1)) + (last >> 40) - res0);
    for (uint32_t i = lane; i < res_n; i += 32) b_res[i] = __ldg(P.a.res + res0 + i);
    if (lane < 2 * nb) b_ws[lane] = ws;
    __syncwarp();

    for (uint32_t k = 0; k < nb; k++) {
      const uint64_t off_len = b_meta[k].off_len;  // the enumeration needs only offset and length
      const uint32_t L = (uint32_t)(off_len >> 40);
      const uint64_t h = b_hash[k];
      const uint8_t* sres = b_res + (uint32_t)((off_len & ((1ull << 40) - 1)) - res0);
      __syncwarp();  // all lanes are done with the previous seed's scratch
      const uint64_t vjh = prepare_seed<SIGMA, INDELS>(z, sres, L, h, lane, sc);
      seed_d1<SIGMA, INDELS>(P, c, z, sres, sc, L, h, vjh, b_ws[2 * k], b_ws[2 * k + 1], (uint32_t)(first + k));
    }
  }
  finish(P, c);
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
  // candidates that passed the filter stage = entries of a queue that is consumed (a chunk that
  // overflowed is redone and counted then)
  if (P.count_bloom && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(P.counters + CTR_BLOOM_PASS, filled);
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
  // Shared-memory carve-out = what the resident CTAs need (+1 KB per CTA for the system), not the
  // driver's "room for the most CTAs" default: the rest of the 228 KB stays L1, which serves the
  // repeated filter words and the staged loads.  (With a single randomly probed filter the
  // carve-out had to stay under 100 KB — the rate of random 8-byte loads an SM sustains halves
  // beyond it, tools/bench_l2_random.cu; with the parity filters such loads are a few dozen per seed.)
  {
    const size_t need = (size_t)per_sm * (smem + 1024);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                         (int)std::min<size_t>(100, (need * 100 + 228 * 1024 - 1) / (228 * 1024)));
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
