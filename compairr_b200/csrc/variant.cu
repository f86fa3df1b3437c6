// variant.cu — K3 for d = 1 and d = 2: on-the-fly variant enumeration by incremental XOR and the
// class-filter test (replaces generate_variants_1/_2, variants.cc:270-400; bloom_get,
// bloompat.h:55-58), and K4, the table stage (find_variant_matches, overlap.cc:168-251;
// check_variant, variants.cc:166-240).
//
// Enumeration, round-2 design: ONE LANE PER SLOT, A LOOP OVER THE RESIDUES.
//
//   A "slot" is a place where a free residue goes: a substitution at position p, an insertion
//   before position p, or — for d = 2 — the second substitution (j, .) of a triple (i, v, j).  All
//   sigma candidates of a slot share
//       base2   the variant's hash minus the Zobrist value of the free residue,
//       word    their filter word (class filters, common.cuh: the word index does not depend on
//               the free residue),
//       zrow    the column of the TRANSPOSED Zobrist table (zT[r * ZP + pos]) their values sit in,
//   so a lane that owns a slot keeps all of that in registers and the inner loop over the residues
//   is one shared-memory load (lanes differ in pos, the residue is uniform), two XORs and the
//   pattern test per candidate — no decode, no per-candidate table lookups, no vote.  Passing
//   residues are collected in a per-lane bit mask; the (rare) survivors are extracted after the
//   loop and compacted into the warp's ring.  Round 1's loop decoded every candidate (q / sigma),
//   loaded base, compare residue and word from shared memory and voted per step: 50 warp
//   instructions per 32 candidates plus ~480 per seed of scans and slot tables.
//
//   d = 1   warps take batches of WB consecutive seeds; the batch's slots (L substitution + L+1
//           insertion slots per seed) form one flat list that the lanes walk 32 at a time, so
//           lanes stay busy whatever the lengths are.  The three XOR scans that the indel variants
//           need (prefix, and the two shifted suffixes; they replace the reference's serial walks,
//           variants.cc:311-324,341-353) are computed for the whole batch at once, one LANE per
//           (seed, scan), serially — 24 lanes x L steps instead of 3 x 5 shuffle rounds per seed.
//   d = 2   a warp takes a (seed, part) item; its slots are the (i < j, v) triples, 19 L (L-1) / 2
//           of them, again walked 32 at a time; the slot's word is fetched one pass ahead.
//
// Seeds longer than the fast kernels' tables (30 residues for ZP = 32, 94 for ZP = 96) go to the
// generic kernel below (any length up to 510), in a launch of its own.
#include <algorithm>

#include "device_utils.cuh"
#include "kernels.cuh"

namespace cb {

constexpr int VK_THREADS = 256;
constexpr int VK_WARPS = VK_THREADS / 32;
constexpr int VK_QCAP = 64;  // ring entries per warp
constexpr int VK_U = 4;      // generic kernel: probes per lane per step

// Per-warp survivor ring: hv[64] u64 | var[64] u32 | seed[64] u32
constexpr uint32_t VK_Q_BYTES = VK_QCAP * 16;

struct WarpCtx {
  unsigned char* wb;     // per-warp block (ring first)
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

// =====================================================================================================
// Fast kernels
// =====================================================================================================

// What a lane keeps for its slot (see the header).
struct SlotRegs {
  uint64_t base2;
  unsigned long long word;
  const uint64_t* zrow;  // &zT[pos]; residue r's value is zrow[r * ZP] (survivors only)
  const uint32_t* erow;  // &zE[pos]; residue r's part of the pattern field is erow[r * ZP]
  uint32_t fbase;        // the slot's part of the pattern field: pattern_field(base2, class of pos)
  uint32_t allowed;      // bit r: residue r is a candidate here (0: idle lane)
  uint32_t var;          // variant descriptor without the free residue
  uint32_t seed;         // seed number relative to a_first
};

// The inner loop: all SIGMA residues of every lane's slot.  RSHIFT: where the free residue goes in
// the descriptor (3 = res1: substitution / insertion, 8 = res2: second substitution).
//
// Two forms of the filter test:
//   one stage    all 3 + 3 pattern bits for every candidate; the (rare) survivors are extracted
//                after the loop.  14.75 warp instructions per 32 candidates.
//   two stages   the loop tests only the three bits of the LOW half of the word (9 instructions per
//                32 candidates); the ~3 % that pass are looked at again — high half, survivors into
//                the ring — in a loop over each lane's set bits, ~4 iterations per pass at 5 of 32
//                lanes active.
// Measured at C3 geometry (kernels incl. table stage): d = 2  7.9 ms two-stage vs 10.2 ms one-stage;
// d = 1 -i  17.2 ms two-stage vs 15.9 ms one-stage — in the larger d = 1 kernel the branchy second
// stage put 23 % of the stall samples on instruction-cache misses ("no instruction").  So d = 2 runs
// the two-stage form and d = 1 the one-stage form (CB_E1_TWO_STAGE for A/B builds).
#ifndef CB_E1_TWO_STAGE
#define CB_E1_TWO_STAGE 0
#endif
template <int SIGMA, int ZP, int RSHIFT, bool TWO_STAGE>
__device__ __forceinline__ void residue_loop(const ProbeParams& P, WarpCtx& c, const SlotRegs& R) {
  const uint32_t wlo = (uint32_t)R.word, whi = (uint32_t)(R.word >> 32);
  // the pattern field is linear in the hash: slot part ^ (position, residue) part (common.cuh)
  if (TWO_STAGE) {
    uint32_t ha = 0;
#pragma unroll
    for (int r = SIGMA - 1; r >= 0; r--)  // downwards: ha = 2 ha + bit leaves bit r for residue r
      ha = ha * 2u + (pattern_half_lo(wlo, R.fbase ^ R.erow[r * ZP]) & 1u);
    ha &= R.allowed;
    while (__any_sync(FULL, ha != 0)) {
      const bool live = ha != 0;
      const uint32_t r = live ? (uint32_t)__ffs((int)ha) - 1u : 0u;
      const bool pass = live && (pattern_half_hi(whi, R.fbase ^ R.erow[r * ZP]) & 1u);
      const uint64_t hv = R.base2 ^ R.zrow[r * ZP];
      const uint32_t var = R.var | (r << RSHIFT);
      submit(P, c, pass, hv, [var] { return var; }, R.seed);  // survivors: false positives + true matches
      ha &= ha - 1;
    }
    return;
  }
  uint32_t hits = 0;
#pragma unroll
  for (int r = 0; r < SIGMA; r++)
    if (pattern_hit_halves(wlo, whi, R.fbase ^ R.erow[r * ZP])) hits |= 1u << r;
  hits &= R.allowed;
  // survivors: a fraction of a percent of the candidates (false positives + true matches)
  while (__any_sync(FULL, hits != 0)) {
    const bool pass = hits != 0;
    const uint32_t r = pass ? (uint32_t)__ffs((int)hits) - 1u : 0u;
    const uint64_t hv = R.base2 ^ R.zrow[r * ZP];
    const uint32_t var = R.var | (r << RSHIFT);
    submit(P, c, pass, hv, [var] { return var; }, R.seed);
    hits &= hits - 1;
  }
}

// the word every residue at a free position of class cls is looked up in (h: the variant's hash
// with ANY residue there, or without one: the index is blind to that position)
__device__ __forceinline__ unsigned long long filter_word(const ProbeParams& P, uint64_t h, uint32_t cls) {
  return P.use_bloom ? __ldg(P.bloom + pfilter_word(h, P.bloom_blocks, cls)) : ~0ull;
}

// Transposed Zobrist table into shared memory: zT[r * ZP + p] = Z(p, r), rows beyond the table 0,
// and beside it zE[r * ZP + p] = the pattern-field part of Z(p, r) in the filter of p's class.
template <int SIGMA, int ZP>
__device__ __forceinline__ void stage_zt(const ProbeParams& P, uint64_t* zT, uint32_t* zE) {
  for (uint32_t i = threadIdx.x; i < SIGMA * ZP; i += VK_THREADS) {
    const uint32_t r = i / ZP, p = i - r * ZP;
    const uint64_t z = p < P.zrows ? P.ztab[p * SIGMA + r] : 0ull;
    zT[i] = z;
    zE[i] = pattern_field(z, pos_class(p));
  }
}
template <int SIGMA, int ZP>
constexpr size_t zt_bytes() { return (size_t)SIGMA * ZP * 12; }

// ---- d = 1 -------------------------------------------------------------------------------------------

template <int ZP>
struct E1Cfg {
  static constexpr int WB = ZP <= 32 ? 8 : 4;  // seeds per warp batch
  static constexpr uint32_t LMAX = ZP - 2;     // longest seed: insertion slot L uses row L, the suffix scan row q + 1
};

constexpr uint32_t LEN_SKIP = 0xffffffffu;  // a seed of the batch that this launch does not handle

// per-warp shared-memory block of the d = 1 kernel (host and device must agree)
template <int ZP, bool INDELS>
struct E1Layout {
  static constexpr int WB = E1Cfg<ZP>::WB;
  static constexpr size_t scan_u64 = INDELS ? (size_t)WB * 3 * ZP : 0;  // pre | sp | sm, [WB][ZP] each
  static constexpr size_t off_scan = VK_Q_BYTES;
  static constexpr size_t off_hash = off_scan + scan_u64 * 8;           // [WB] u64
  static constexpr size_t off_ws = off_hash + WB * 8;                   // [WB][4] u64: the seed's word in each class filter
  static constexpr size_t off_len = off_ws + WB * 32;                   // [WB] u32
  static constexpr size_t off_cum = off_len + WB * 4;                   // [WB + 1] u32, substitution slots
  static constexpr size_t off_icum = off_cum + (WB + 1) * 4;            // [WB + 1] u32, insertion slots
  static constexpr size_t off_dcum = off_icum + (WB + 1) * 4;           // [WB + 1] u32, deletion + identical items
  static constexpr size_t off_res = (off_dcum + (WB + 1) * 4 + 15) & ~(size_t)15;  // [WB][ZP] u8
  static constexpr size_t warp_bytes = (off_res + (size_t)WB * ZP + 15) & ~(size_t)15;
  static constexpr size_t total(int sigma) { return (size_t)sigma * ZP * 12 + VK_WARPS * warp_bytes; }
};

template <int SIGMA, bool INDELS, int ZP>
__global__ void __launch_bounds__(VK_THREADS, 3) enum1_kernel(const __grid_constant__ ProbeParams P) {
  using Lay = E1Layout<ZP, INDELS>;
  constexpr int WB = Lay::WB;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t* const zT = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* const zE = reinterpret_cast<uint32_t*>(smem_raw + (size_t)SIGMA * ZP * 8);
  WarpCtx c;
  c.wb = smem_raw + zt_bytes<SIGMA, ZP>() + warp * Lay::warp_bytes;
  c.head = c.count = 0;
  c.lane = lane;
  uint64_t* const scan = reinterpret_cast<uint64_t*>(c.wb + Lay::off_scan);
  uint64_t* const pre = scan;                       // pre[k][p]  = xor_{q<p}  Z(q,     s[q])
  uint64_t* const sp = scan + (size_t)WB * ZP;      // sp[k][p]   = xor_{q>=p} Z(q + 1, s[q])
  uint64_t* const sm = scan + (size_t)2 * WB * ZP;  // sm[k][p]   = xor_{q>=p} Z(q - 1, s[q])
  uint64_t* const b_hash = reinterpret_cast<uint64_t*>(c.wb + Lay::off_hash);
  unsigned long long* const b_ws = reinterpret_cast<unsigned long long*>(c.wb + Lay::off_ws);
  uint32_t* const b_len = reinterpret_cast<uint32_t*>(c.wb + Lay::off_len);
  uint32_t* const b_cum = reinterpret_cast<uint32_t*>(c.wb + Lay::off_cum);
  uint32_t* const b_icum = reinterpret_cast<uint32_t*>(c.wb + Lay::off_icum);
  uint32_t* const b_dcum = reinterpret_cast<uint32_t*>(c.wb + Lay::off_dcum);
  uint8_t* const b_res = c.wb + Lay::off_res;

  stage_zt<SIGMA, ZP>(P, zT, zE);
  __syncthreads();
  const uint64_t n_batches = (P.w_count + WB - 1) / WB;
  constexpr uint32_t ALL = (SIGMA >= 32) ? 0xffffffffu : ((1u << SIGMA) - 1u);

  for (;;) {
    unsigned long long b = 0;
    if (lane == 0) b = atomicAdd(P.counters + CTR_WORK, 1ull);
    b = __shfl_sync(FULL, b, 0);
    if (b >= n_batches) break;
    const uint64_t first = P.w_first + b * WB;  // relative to a_first
    const uint32_t nb = (uint32_t)((P.w_first + P.w_count - first < WB) ? P.w_first + P.w_count - first : WB);
    __syncwarp();  // previous batch fully consumed

    // ---- stage the batch: lengths, hashes, residues, each seed's own word in the four filters.
    // (Tried and dropped: running the dispenser two batches ahead and fetching the next batch's
    // records and hashes into registers while the current batch is worked on — the chain dispenser
    // -> records -> residues is 19 % of the stall samples — 16.1 vs 15.9 ms at C3 geometry: the six
    // extra live registers cost what the hidden latency gained; four deletion passes of loads in
    // flight instead of two: 17.5 ms, the unrolled copies grow the kernel by a quarter.)
    uint64_t my_off = 0;
    uint32_t my_len = LEN_SKIP;
    if (lane < nb) {
      const uint64_t off_len = __ldg(&P.a.meta[P.a_first + first + lane].off_len);
      const uint32_t L = (uint32_t)(off_len >> 40);
      my_off = off_len & ((1ull << 40) - 1);
      if (L >= P.len_lo && L <= P.len_hi) my_len = L;
    }
    if (lane < 4 * nb) {  // the seed's own word in each class filter = the word of all its substitution slots of that class
      const uint64_t hk = __ldg(P.a.hash + P.a_first + first + (lane >> 2));
      if ((lane & 3) == 0) b_hash[lane >> 2] = hk;
      b_ws[lane] = filter_word(P, hk, lane & 3);
    }
    {  // slot and item counts -> inclusive prefix sums over the WB seeds
      const bool live = my_len != LEN_SKIP;
      // L substitution slots, L + 1 insertion slots, L deletion candidates + the identical one: the
      // three lists differ by the number of live seeds before this one only
      uint32_t ns = live ? my_len : 0, nl = live ? 1 : 0;
#pragma unroll
      for (int o = 1; o < WB; o <<= 1) {
        const uint32_t xs = __shfl_up_sync(FULL, ns, o), xl = __shfl_up_sync(FULL, nl, o);
        if ((int)lane >= o) {
          ns += xs;
          nl += xl;
        }
      }
      if (lane < WB) {
        b_len[lane] = my_len;
        b_cum[lane + 1] = ns;
        b_icum[lane + 1] = INDELS ? ns + nl : 0;
        b_dcum[lane + 1] = INDELS ? ns + nl : nl;
      }
      if (lane == 0) b_cum[0] = b_icum[0] = b_dcum[0] = 0;
    }
#pragma unroll
    for (int k = 0; k < WB; k++) {  // residues: one row of ZP bytes per seed
      const uint32_t Lk = __shfl_sync(FULL, my_len, k);
      const uint64_t ok = __shfl_sync(FULL, my_off, k);
      if (Lk != LEN_SKIP)
        for (uint32_t p = lane; p < Lk; p += 32) b_res[k * ZP + p] = __ldg(P.a.res + ok + p);
    }
    __syncwarp();

    // ---- the three scans of every seed of the batch, one lane per (scan, seed) -------------------------
    if (INDELS) {
      const uint32_t which = lane / WB, k = lane % WB;  // 0 pre, 1 sp, 2 sm
      const uint32_t L = which < 3 ? b_len[k] : LEN_SKIP;
      if (L != LEN_SKIP) {
        uint64_t* const arr = scan + ((size_t)which * WB + k) * ZP;
        const uint8_t* const s = b_res + k * ZP;
        uint64_t x = 0;
        arr[which == 0 ? 0 : L] = 0ull;
        for (uint32_t t = 0; t < L; t++) {
          const uint32_t q = which == 0 ? t : L - 1 - t;
          const int zp = (int)q + (which == 0 ? 0 : which == 1 ? 1 : -1);
          x ^= zp >= 0 ? zT[s[q] * ZP + zp] : 0ull;
          arr[which == 0 ? q + 1 : q] = x;
        }
      }
      __syncwarp();
    }

    // ---- residue slots of the batch, 32 at a time: first all substitution slots, then all insertion
    // slots (two flat lists, so that a pass loads its slots through one code path) ----------------------------
    auto seed_of = [&](const uint32_t* cum, uint32_t g) {
      uint32_t k = 0;
#pragma unroll
      for (int j = 1; j < WB; j++) k += g >= cum[j];
      return k;
    };
    auto load_sub = [&](uint32_t g, uint32_t n) {
      SlotRegs R;
      const bool valid = g < n;
      if (!valid) g = 0;
      const uint32_t k = seed_of(b_cum, g), pos = g - b_cum[k];
      const uint32_t cmp = b_res[k * ZP + pos];
      R.base2 = b_hash[k] ^ zT[cmp * ZP + pos];
      R.word = b_ws[4 * k + pos_class(pos)];
      R.var = pack_var(VK_SUBSTITUTION, pos, 0, 0, 0);
      R.allowed = valid ? (ALL & ~(1u << cmp)) : 0u;
      R.zrow = zT + pos;
      R.erow = zE + pos;
      R.fbase = pattern_field(R.base2, pos_class(pos));
      R.seed = (uint32_t)first + k;
      return R;
    };
    auto load_ins = [&](uint32_t g, uint32_t n) {  // insertion before position pos: the new residue sits at position
      SlotRegs R;                                  // pos of the variant; not the residue before it, which would
      const bool valid = g < n;                    // repeat a variant (variants.cc:341-353)
      if (!valid) g = 0;
      const uint32_t k = seed_of(b_icum, g), pos = g - b_icum[k], L = b_len[k];
      const uint32_t cmp = pos ? b_res[k * ZP + pos - 1] : 31u;
      R.base2 = b_hash[k] ^ pre[k * ZP + L] ^ pre[k * ZP + pos] ^ sp[k * ZP + pos];
      R.word = filter_word(P, R.base2, pos_class(pos));
      R.var = pack_var(VK_INSERTION, pos, 0, 0, 0);
      R.allowed = valid ? (ALL & ~(1u << cmp)) : 0u;
      R.zrow = zT + pos;
      R.erow = zE + pos;
      R.fbase = pattern_field(R.base2, pos_class(pos));
      R.seed = (uint32_t)first + k;
      return R;
    };
    // one loop over both lists (the residue loop is inlined once: code size matters here, see residue_loop)
    const uint32_t n_sub = b_cum[WB], n_ins = INDELS ? b_icum[WB] : 0;
    const uint32_t p_sub = (n_sub + 31) / 32, p_all = p_sub + (n_ins + 31) / 32;
    auto load = [&](uint32_t t) {
      return (!INDELS || t < p_sub) ? load_sub(t * 32 + lane, n_sub) : load_ins((t - p_sub) * 32 + lane, n_ins);
    };
    if (p_all) {
      SlotRegs cur = load(0);
      for (uint32_t t = 0; t < p_all; t++) {
        SlotRegs nxt = cur;
        if (t + 1 < p_all) nxt = load(t + 1);  // the next pass's words are in flight during this one
        residue_loop<SIGMA, ZP, 3, CB_E1_TWO_STAGE != 0>(P, c, cur);
        cur = nxt;
      }
    }

    // ---- deletions (one per run of equal residues, only if L > 1, variants.cc:301-325) and the
    // identical candidate of every seed: one candidate per item ---------------------------------------------
    const uint32_t n_items = b_dcum[WB];
    constexpr int DU = 2;  // item passes in flight: their filter-word loads are issued together
    for (uint32_t g0 = 0; g0 < n_items; g0 += 32 * DU) {
      uint64_t hv[DU];
      unsigned long long w[DU];
      uint32_t desc[DU], ks[DU];  // descriptor | filter class << 29 | valid << 31
#pragma unroll
      for (int u = 0; u < DU; u++) {
        uint32_t g = g0 + u * 32 + lane;
        const bool in = g < n_items;
        if (!in) g = 0;
        uint32_t k = 0;
#pragma unroll
        for (int j = 1; j < WB; j++) k += g >= b_dcum[j];
        const uint32_t L = b_len[k], t = g - b_dcum[k];
        const bool is_del = INDELS && t < L;
        const uint64_t h = b_hash[k];
        hv[u] = h;
        bool valid = in;
        uint32_t cls = 0;
        w[u] = b_ws[4 * k];  // identical: any filter will do; the seed's word in filter 0 is at hand
        if (is_del) {        // no free residue: any filter, spread over the four
          hv[u] = h ^ pre[k * ZP + L] ^ pre[k * ZP + t] ^ sm[k * ZP + t + 1];
          valid = in && L > 1 && (t == 0 || b_res[k * ZP + t] != b_res[k * ZP + t - 1]);
          cls = pos_class(t);
          w[u] = filter_word(P, hv[u], cls);
        }
        desc[u] = (is_del ? pack_var(VK_DELETION, t, 0, 0, 0) : pack_var(VK_IDENTICAL, 0, 0, 0, 0)) | (cls << 29) | ((uint32_t)valid << 31);
        ks[u] = (uint32_t)first + k;
      }
#pragma unroll
      for (int u = 0; u < DU; u++) {
        if (g0 + u * 32 >= n_items) break;  // warp-uniform
        const bool pass = (desc[u] >> 31) & pattern_hit(w[u], pattern_field(hv[u], (desc[u] >> 29) & 3u));
        const uint32_t var = desc[u] & 0x1fffffffu;  // deletion / identical descriptors use bits 0..21 only
        submit(P, c, pass, hv[u], [var] { return var; }, ks[u]);
      }
    }
  }
  finish(P, c);
}

// ---- d = 2 -------------------------------------------------------------------------------------------

template <int ZP>
struct E2Layout {
  static constexpr uint32_t LMAX = ZP - 2;
  static constexpr uint32_t NPAIR = LMAX * (LMAX - 1) / 2;
  static constexpr size_t off_res = VK_Q_BYTES;                  // [ZP] u8 per warp
  static constexpr size_t warp_bytes = (off_res + ZP + 15) & ~(size_t)15;
  static constexpr size_t pair_bytes = ((size_t)NPAIR * 2 + 15) & ~(size_t)15;
  static constexpr size_t total(int sigma) { return (size_t)sigma * ZP * 12 + pair_bytes + VK_WARPS * warp_bytes; }
};

template <int SIGMA, int ZP>
__global__ void __launch_bounds__(VK_THREADS, 3) enum2_kernel(const __grid_constant__ ProbeParams P) {
  using Lay = E2Layout<ZP>;
  constexpr uint32_t S1 = SIGMA - 1;
  constexpr uint32_t ALL = (1u << SIGMA) - 1u;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t* const zT = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* const zE = reinterpret_cast<uint32_t*>(smem_raw + (size_t)SIGMA * ZP * 8);
  // pair e = j (j - 1) / 2 + i  (i < j): independent of the seed's length, a seed of length L owns e < L (L - 1) / 2
  uint16_t* const pairtab = reinterpret_cast<uint16_t*>(smem_raw + zt_bytes<SIGMA, ZP>());
  WarpCtx c;
  c.wb = smem_raw + zt_bytes<SIGMA, ZP>() + Lay::pair_bytes + warp * Lay::warp_bytes;
  c.head = c.count = 0;
  c.lane = lane;
  uint8_t* const sres = c.wb + Lay::off_res;

  stage_zt<SIGMA, ZP>(P, zT, zE);
  for (uint32_t j = 1 + threadIdx.x; j < Lay::LMAX; j += VK_THREADS)
    for (uint32_t i = 0; i < j; i++) pairtab[j * (j - 1) / 2 + i] = (uint16_t)(i | (j << 8));
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
    if (L < P.len_lo || L > P.len_hi) continue;
    const uint64_t h = __ldg(P.a.hash + sidx);
    __syncwarp();  // all lanes are done with the previous seed's residues
    for (uint32_t p = lane; p < L; p += 32) sres[p] = __ldg(P.a.res + (off_len & ((1ull << 40) - 1)) + p);
    __syncwarp();

    if (part == 0) {  // identical + single substitutions (the reference emits them with d = 2 too, variants.cc:410-427)
      const unsigned long long ws = filter_word(P, h, lane & 3);  // class of the lane = class of its position (p0 is a multiple of 32)
      for (uint32_t p0 = 0; p0 < L; p0 += 32) {
        const uint32_t pos = p0 + lane;
        const bool valid = pos < L;
        const uint32_t pc = valid ? pos : 0u;
        const uint32_t cmp = sres[pc];
        SlotRegs R;
        R.base2 = h ^ zT[cmp * ZP + pc];
        R.word = ws;
        R.allowed = valid ? (ALL & ~(1u << cmp)) : 0u;
        R.zrow = zT + pc;
        R.erow = zE + pc;
        R.fbase = pattern_field(R.base2, lane & 3);
        R.var = pack_var(VK_SUBSTITUTION, pc, 0, 0, 0);
        R.seed = slocal;
        residue_loop<SIGMA, ZP, 3, true>(P, c, R);
      }
      submit(P, c, lane == 0 && pattern_hit(ws, pattern_field(h, 0)), h, [] { return pack_var(VK_IDENTICAL, 0, 0, 0, 0); }, slocal);
    }

    // double substitutions i < j (variants.cc:357-400): slots x = (pair e, first residue v), lanes
    // over x, the loop over the second residue w.  The second substitution cannot change the filter
    // word: it is the word of h ^ Z(i,s[i]) ^ Z(i,v) in the filter of j's class.
    const uint32_t n_x = L * (L - 1) / 2 * S1;
    auto load_slot = [&](uint32_t x) {
      SlotRegs R;
      const bool valid = x < n_x;
      if (!valid) x = 0;
      const uint32_t e = x / S1, vp = x - e * S1;
      const uint32_t ij = pairtab[e];
      const uint32_t i = ij & 255u, j = ij >> 8;
      const uint32_t si = sres[i], sj = sres[j];
      const uint32_t v = sub_residue(vp, si);
      const uint64_t b2v = h ^ zT[si * ZP + i] ^ zT[v * ZP + i];
      R.base2 = b2v ^ zT[sj * ZP + j];
      R.word = filter_word(P, b2v, pos_class(j));
      R.allowed = valid ? (ALL & ~(1u << sj)) : 0u;
      R.zrow = zT + j;
      R.erow = zE + j;
      R.fbase = pattern_field(R.base2, pos_class(j));
      R.var = pack_var(VK_SUB_SUB, i, v, j, 0);
      R.seed = slocal;
      return R;
    };
    const uint32_t n_pass = (n_x + 31) / 32;
    if (part < n_pass) {
      SlotRegs cur = load_slot(part * 32 + lane);
      for (uint32_t t = part; t < n_pass; t += P.split) {
        SlotRegs nxt = cur;
        if (t + P.split < n_pass) nxt = load_slot((t + P.split) * 32 + lane);  // one pass ahead: its word is in flight
        residue_loop<SIGMA, ZP, 8, true>(P, c, cur);
        cur = nxt;
      }
    }
  }
  finish(P, c);
}

// =====================================================================================================
// Generic kernel: any seed length up to 510 (d = 1 with or without indels, d = 2).  One warp per
// (seed, part); every candidate is decoded from a flat index and looks its own filter word up.
// Slower per probe than the fast kernels; it exists so that a single long sequence does not
// change what the engine can do (the reference has no length limit).
// =====================================================================================================

__device__ __forceinline__ void filter_step(const ProbeParams& P, const uint64_t (&hv)[VK_U],
                                            const uint32_t (&cls)[VK_U], bool (&pass)[VK_U]) {
  if (!P.use_bloom) return;
  unsigned long long w[VK_U];
#pragma unroll
  for (int u = 0; u < VK_U; u++)  // unconditional: an inactive candidate's hash is a valid address too
    w[u] = __ldg(P.bloom + pfilter_word(hv[u], P.bloom_blocks, cls[u]));
#pragma unroll
  for (int u = 0; u < VK_U; u++) pass[u] = pass[u] & pattern_hit(w[u], pattern_field(hv[u], cls[u]));
}

// Per-warp scratch of the generic kernel for one seed.
struct SeedScratch {
  uint64_t* zo;    // Z(p, s[p]); with indels the scans follow at multiples of lpad:
  uint32_t lpad;   //   pre[p] = xor_{q<p} Z(q, s[q]),  sm[p] = xor_{q>=p} Z(q-1, s[q]),  sp[p] = xor_{q>=p} Z(q+1, s[q])
  __device__ __forceinline__ uint64_t* pre() const { return zo + lpad; }
  __device__ __forceinline__ uint64_t* sm() const { return zo + 2 * lpad; }
  __device__ __forceinline__ uint64_t* sp() const { return zo + 3 * lpad; }
};

template <bool ZG>
__device__ __forceinline__ uint64_t zval(const uint64_t* __restrict__ z, uint32_t i) { return ZG ? __ldg(z + i) : z[i]; }

// Fill zo[] (and the three scans) for the seed whose residues are at sres[0..L).  Returns VJ.
template <int SIGMA, bool INDELS, bool ZG>
__device__ __forceinline__ uint64_t prepare_seed(const uint64_t* __restrict__ z, const uint8_t* sres,
                                                 uint32_t L, uint64_t h, uint32_t lane, SeedScratch& s) {
  for (uint32_t p = lane; p < L; p += 32) s.zo[p] = zval<ZG>(z, p * SIGMA + sres[p]);
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
    uint64_t xm = (ok && q >= 1) ? zval<ZG>(z, (q - 1) * SIGMA + r) : 0ull;
    uint64_t xp = ok ? zval<ZG>(z, (q + 1) * SIGMA + r) : 0ull;
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
template <int SIGMA, bool INDELS, bool ZG>
__device__ __forceinline__ void phase_a(const ProbeParams& P, WarpCtx& c, const uint64_t* __restrict__ z,
                                        const uint8_t* sres, const SeedScratch& s,
                                        uint32_t L, uint64_t h, uint64_t vjh, uint32_t slocal) {
  constexpr uint32_t S1 = SIGMA - 1;
  const uint32_t nsub = S1 * L;
  const uint32_t T = 1 + nsub + (INDELS ? L + SIGMA * (L + 1) : 0);
  for (uint32_t base = 0; base < T; base += 32 * VK_U) {
    uint64_t hv[VK_U];
    uint32_t var[VK_U];
    bool pass[VK_U];
    uint32_t cls[VK_U];
#pragma unroll
    for (int u = 0; u < VK_U; u++) {
      const uint32_t idx = base + u * 32 + c.lane;
      pass[u] = idx < T;
      hv[u] = h;
      cls[u] = 0;
      var[u] = pack_var(VK_IDENTICAL, 0, 0, 0, 0);
      if (pass[u] && idx >= 1) {
        uint32_t t = idx - 1;
        if (t < nsub) {
          const uint32_t pos = t / S1, rp = t - pos * S1;
          const uint32_t r = sub_residue(rp, sres[pos]);
          hv[u] = h ^ s.zo[pos] ^ zval<ZG>(z, pos * SIGMA + r);
          cls[u] = pos_class(pos);
          var[u] = pack_var(VK_SUBSTITUTION, pos, r, 0, 0);
        } else if (INDELS) {
          t -= nsub;
          if (t < L) {  // deletion of residue t: only at the start of a run, only if L > 1
            pass[u] = (L > 1) && (t == 0 || sres[t] != sres[t - 1]);
            hv[u] = vjh ^ s.pre()[t] ^ s.sm()[t + 1];
            cls[u] = pos_class(t);  // no free residue: any filter, spread over the four
            var[u] = pack_var(VK_DELETION, t, 0, 0, 0);
          } else {  // insertion of residue r before seed position pos
            t -= L;
            const uint32_t pos = t / SIGMA, r = t - pos * SIGMA;
            pass[u] = (pos == 0) || (r != sres[pos - 1]);
            hv[u] = vjh ^ s.pre()[pos] ^ zval<ZG>(z, pos * SIGMA + r) ^ s.sp()[pos];
            cls[u] = pos_class(pos);  // the inserted residue sits at position pos of the variant
            var[u] = pack_var(VK_INSERTION, pos, r, 0, 0);
          }
        }
      }
    }
    filter_step(P, hv, cls, pass);
#pragma unroll
    for (int u = 0; u < VK_U; u++) {
      const uint32_t v = var[u];
      submit(P, c, pass[u], hv[u], [v] { return v; }, slocal);
    }
  }
}

// Phase B: double substitutions i < j (variants.cc:357-400).  Outer (i, v) warp-uniform, lanes over
// the slots (j > i, r) of the second substitution, one filter word per class of j per outer iteration.
template <int SIGMA, bool ZG>
__device__ __forceinline__ void phase_b(const ProbeParams& P, WarpCtx& c, const uint64_t* __restrict__ z,
                                        const uint8_t* sres, const SeedScratch& s,
                                        uint32_t L, uint64_t h, uint32_t slocal, uint32_t part,
                                        uint32_t split) {
  constexpr uint32_t S1 = SIGMA - 1;
  const uint32_t nouter = S1 * L;
  for (uint32_t o = part; o < nouter; o += split) {
    const uint32_t i = o / S1, v = sub_residue(o - i * S1, sres[i]);
    const uint64_t b2 = h ^ s.zo[i] ^ zval<ZG>(z, i * SIGMA + v);
    unsigned long long wc[CB_CLASSES];
#pragma unroll
    for (uint32_t q = 0; q < CB_CLASSES; q++) wc[q] = filter_word(P, b2, q);
    const uint32_t ninner = SIGMA * (L - 1 - i);
    for (uint32_t tb = 0; tb < ninner; tb += 32) {
      const uint32_t t = tb + c.lane;
      const bool in = t < ninner;
      const uint32_t tt = in ? t : 0u;
      const uint32_t jj = tt / SIGMA, r = tt - jj * SIGMA;
      const uint32_t j = in ? i + 1 + jj : i;  // inactive lanes read a valid row
      const uint32_t cmp = sres[j];
      const uint64_t hv = b2 ^ s.zo[j] ^ zval<ZG>(z, j * SIGMA + r);
      const uint32_t jc = pos_class(j);
      const unsigned long long w = jc == 0 ? wc[0] : jc == 1 ? wc[1] : jc == 2 ? wc[2] : wc[3];
      const bool pass = in & (r != cmp) & pattern_hit(w, pattern_field(hv, jc));
      submit(P, c, pass, hv, [i, v, j, r] { return pack_var(VK_SUB_SUB, i, v, j, r); }, slocal);
    }
  }
}

struct GenLayout {
  uint32_t lpad;      // per-seed scratch entries (>= lmax + 2, multiple of 8)
  size_t z_u64;       // Zobrist rows staged in shared memory (0 with ZG)
  size_t warp_bytes;  // survivor ring + seed scratch + residues
  size_t total;
};

static inline GenLayout gen_layout(uint32_t zrows, uint32_t sigma, uint32_t lmax, bool indels, bool zg) {
  GenLayout l;
  l.lpad = (lmax + 2 + 7) & ~7u;
  l.z_u64 = zg ? 0 : (size_t)zrows * sigma;
  l.warp_bytes = VK_Q_BYTES + (size_t)l.lpad * (indels ? 4 : 1) * 8 + l.lpad;
  l.total = l.z_u64 * 8 + VK_WARPS * l.warp_bytes;
  return l;
}

template <int SIGMA, bool INDELS, bool ZG>
__global__ void __launch_bounds__(VK_THREADS, 2) generic_kernel(const __grid_constant__ ProbeParams P, uint32_t lpad) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t z_u64 = ZG ? 0 : (size_t)P.zrows * SIGMA;
  const size_t warp_bytes = VK_Q_BYTES + (size_t)lpad * (INDELS ? 4 : 1) * 8 + lpad;
  uint64_t* zs = reinterpret_cast<uint64_t*>(smem_raw);
  WarpCtx c;
  c.wb = smem_raw + z_u64 * 8 + warp * warp_bytes;
  c.head = c.count = 0;
  c.lane = lane;
  SeedScratch sc;
  sc.zo = reinterpret_cast<uint64_t*>(c.wb + VK_Q_BYTES);
  sc.lpad = lpad;
  uint8_t* const sres = c.wb + VK_Q_BYTES + (size_t)lpad * (INDELS ? 4 : 1) * 8;
  if (!ZG) {
    for (uint32_t i = threadIdx.x; i < P.zrows * SIGMA; i += VK_THREADS) zs[i] = P.ztab[i];
    __syncthreads();
  }
  const uint64_t* const z = ZG ? P.ztab : zs;

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
    const uint64_t off_len = __ldg(&P.a.meta[sidx].off_len);
    const uint32_t L = (uint32_t)(off_len >> 40);
    if (L < P.len_lo || L > P.len_hi) continue;
    const uint64_t h = __ldg(P.a.hash + sidx);
    __syncwarp();
    for (uint32_t p = lane; p < L; p += 32) sres[p] = __ldg(P.a.res + (off_len & ((1ull << 40) - 1)) + p);
    __syncwarp();
    const uint64_t vjh = prepare_seed<SIGMA, INDELS, ZG>(z, sres, L, h, lane, sc);
    if (part == 0) phase_a<SIGMA, INDELS, ZG>(P, c, z, sres, sc, L, h, vjh, slocal);
    if (!INDELS && P.differences == 2) phase_b<SIGMA, ZG>(P, c, z, sres, sc, L, h, slocal, part, P.split);
  }
  finish(P, c);
}

// ---- K4: the table stage ----------------------------------------------------------------------------
//
// One thread per queued candidate (hash, variant, seed): probe chain, exact verify against the head
// of the occurrence list, score + matrix atomics + pair append per occurrence.  A queue that
// overflowed is not touched at all: the chunk is recorded and redone by the host.
__global__ void __launch_bounds__(256) table_kernel(const __grid_constant__ ProbeParams P, uint32_t chunk_id) {
  extern __shared__ __align__(16) unsigned char tile_raw[];
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
  double* const tile = matrix_tile_begin(P, tile_raw);
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
    nmatch += probe_chains(&P, act, hv, vs.x, P.a_first + vs.y, vs.y, tile);
  }
  matrix_tile_flush(P, tile);
  flush_counters(P, nmatch, 0);
}

void launch_table_stage(const ProbeParams& p, int sm_count, uint32_t chunk_id, cudaStream_t st) {
  const size_t smem = (size_t)p.tile_cells * sizeof(double);
  if (smem > 48 * 1024) cudaFuncSetAttribute(table_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  // with a tile fewer, longer-lived CTAs: every CTA flushes its whole tile once
  const int per_sm = smem ? (smem > 48 * 1024 ? 2 : 4) : 8;
  table_kernel<<<sm_count * per_sm, 256, smem, st>>>(p, chunk_id);
}

// ---- launch ------------------------------------------------------------------------------------------

template <typename K, typename... Extra>
static int launch_one(K kern, const ProbeParams& p, size_t smem, uint64_t work_ctas, int sm_count,
                      cudaStream_t st, const char** err, Extra... extra) {
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
  // repeated filter words and the staged loads.
  {
    const size_t need = (size_t)per_sm * (smem + 1024);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                         (int)std::min<size_t>(100, (need * 100 + 228 * 1024 - 1) / (228 * 1024)));
  }
  uint64_t grid = (uint64_t)sm_count * per_sm;  // persistent: whole waves of resident CTAs
  if (work_ctas < grid) grid = work_ctas;
  if (grid == 0) return 0;
  kern<<<(unsigned)grid, VK_THREADS, smem, st>>>(p, extra...);
  return 1;
}

template <int SIGMA, int ZP>
static int launch_fast(const ProbeParams& p, int sm_count, cudaStream_t st, const char** err) {
  if (p.differences == 1) {
    constexpr int WB = E1Cfg<ZP>::WB;
    const uint64_t ctas = ((p.w_count + WB - 1) / WB + VK_WARPS - 1) / VK_WARPS;
    return p.indels ? launch_one(enum1_kernel<SIGMA, true, ZP>, p, E1Layout<ZP, true>::total(SIGMA), ctas, sm_count, st, err)
                    : launch_one(enum1_kernel<SIGMA, false, ZP>, p, E1Layout<ZP, false>::total(SIGMA), ctas, sm_count, st, err);
  }
  const uint64_t ctas = (p.w_count * p.split + VK_WARPS - 1) / VK_WARPS;
  return launch_one(enum2_kernel<SIGMA, ZP>, p, E2Layout<ZP>::total(SIGMA), ctas, sm_count, st, err);
}

template <int SIGMA>
static int launch_generic(ProbeParams p, int sm_count, cudaStream_t st, const char** err) {
  const bool indels = p.indels && p.differences == 1;
  p.zrows = std::min(p.zrows, p.lmax + 2);  // rows the longest seed can touch (+1 for the shifted rows of the indel scans)
  bool zg = false;
  GenLayout l = gen_layout(p.zrows, SIGMA, p.lmax, indels, false);
  if (l.total > 160 * 1024) {  // the table itself stays in global memory / L1
    zg = true;
    l = gen_layout(p.zrows, SIGMA, p.lmax, indels, true);
  }
  const uint64_t ctas = (p.w_count * p.split + VK_WARPS - 1) / VK_WARPS;
  if (indels)
    return zg ? launch_one(generic_kernel<SIGMA, true, true>, p, l.total, ctas, sm_count, st, err, l.lpad)
              : launch_one(generic_kernel<SIGMA, true, false>, p, l.total, ctas, sm_count, st, err, l.lpad);
  return zg ? launch_one(generic_kernel<SIGMA, false, true>, p, l.total, ctas, sm_count, st, err, l.lpad)
            : launch_one(generic_kernel<SIGMA, false, false>, p, l.total, ctas, sm_count, st, err, l.lpad);
}

// Seeds are split by length between up to three launches (each kernel skips the seeds that are not
// its own): <= 30 residues the ZP = 32 kernels, <= 94 the ZP = 96 kernels, longer ones the generic
// kernel.  p.lmax = the longest seed of the set decides which launches are needed at all.
template <int SIGMA>
static int launch_by_length(const ProbeParams& p_in, int sm_count, cudaStream_t st, const char** err) {
  ProbeParams p = p_in;
  int launches = 0, l;
  const uint32_t f32 = E1Cfg<32>::LMAX, f96 = E1Cfg<96>::LMAX;
  uint32_t lo = 0;
  if (!p_in.force_generic) {
    p.len_lo = 0;
    p.len_hi = f32;
    if ((l = launch_fast<SIGMA, 32>(p, sm_count, st, err)) < 0) return l;
    launches += l;
    lo = f32 + 1;
    if (p_in.lmax >= lo) {
      p.len_lo = lo;
      p.len_hi = f96;
      cudaMemsetAsync(p.counters + CTR_WORK, 0, sizeof(unsigned long long), st);  // every launch walks all seeds: its own dispenser
      if ((l = launch_fast<SIGMA, 96>(p, sm_count, st, err)) < 0) return l;
      launches += l;
      lo = f96 + 1;
    }
  }
  if (p_in.lmax >= lo) {
    p.len_lo = lo;
    p.len_hi = 0xffffffffu;
    if (launches) cudaMemsetAsync(p.counters + CTR_WORK, 0, sizeof(unsigned long long), st);
    if ((l = launch_generic<SIGMA>(p, sm_count, st, err)) < 0) return l;
    launches += l;
  }
  return launches;
}

int launch_variant_kernels(const ProbeParams& p, int sm_count, cudaStream_t st, const char** err) {
  if (p.lmax + 1 > VAR_MAX_POS) {
    *err = "sequence longer than 510 residues on the d<=2 path (the variant descriptor holds 9-bit positions)";
    return -1;
  }
  if (p.sigma != 4 && p.sigma != 20) {
    *err = "alphabet size must be 4 or 20";
    return -1;
  }
  return p.sigma == 20 ? launch_by_length<20>(p, sm_count, st, err) : launch_by_length<4>(p, sm_count, st, err);
}

}  // namespace cb
