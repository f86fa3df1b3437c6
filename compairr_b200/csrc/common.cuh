// common.cuh — data layout and the host/device-shared pieces of the overlap hot path.
//
// Everything in here is plain integer arithmetic that both the kernels (kernels.cu) and the
// host side of the engine (engine.cu: table generation, probe counting) use.  The functions
// marked CB_HD are also compiled by g++ in tests/ (tests/csrc/hd_check.cpp) so that the variant
// decoding rules can be checked against the oracle without a GPU; that harness is a test of
// this header, not a product path.
//
// Reference semantics restated here (file:line in /root/reference/src):
//   variant enumeration rules     variants.cc:260-428
//   exact verification            variants.cc:166-240
//   score summands                overlap.cc:144-166
//   Zobrist hash                  zobrist.cc:74-88   (table VALUES are ours)
//   Bloom filter                  bloompat.h:40-58   (geometry is ours: class filters)
//   hash-table indexing           hashtable.h:36-46
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CB_HD __host__ __device__ __forceinline__
#else
#define CB_HD inline
#endif

namespace cb {

// ---- device data layout -------------------------------------------------------------------

// One 32-byte record per sequence = exactly one L2/DRAM sector, so a verify touches one sector
// for all of (offset, length, V, J, repertoire, count, next) instead of seven arrays (reference AoS
// seqinfo_s, db.cc:77-88, is 56 B).  `next` chains the occurrences of one distinct
// (sequence, V, J): identical set-B sequences share ONE table slot and hang off it as a list.
struct alignas(32) SeqRec {  // memory format
  uint64_t off_len;  // first residue in the arena (low 40 bits) | length (high 24 bits)
  uint64_t count;    // duplicate_count
  uint32_t v;
  uint32_t j;
  uint32_t rep;
  uint32_t next;     // next occurrence of the same (sequence, V, J) in set B, SEQ_NIL = end
};
static_assert(sizeof(SeqRec) == 32, "SeqRec must be one 32-byte sector");
constexpr uint32_t SEQ_NIL = 0xffffffffu;
constexpr uint32_t SEQ_MAX_LEN = (1u << 24) - 1;

struct SeqMeta {  // register format
  uint64_t off;
  uint64_t count;
  uint32_t len;
  uint32_t v;
  uint32_t j;
  uint32_t rep;
  uint32_t next;
};

CB_HD uint64_t pack_off_len(uint64_t off, uint32_t len) { return (off & ((1ull << 40) - 1)) | ((uint64_t)len << 40); }
CB_HD SeqMeta unpack_rec(uint64_t off_len, uint64_t count, uint32_t v, uint32_t j, uint32_t rep, uint32_t next) {
  SeqMeta m;
  m.off = off_len & ((1ull << 40) - 1);
  m.len = (uint32_t)(off_len >> 40);
  m.count = count;
  m.v = v;
  m.j = j;
  m.rep = rep;
  m.next = next;
  return m;
}

// Open-addressing slot, probed with one 128-bit load (reference keeps three arrays:
// hash_values / hash_data / hash_occupied bitmap, hashtable.h:22-29).  One slot per DISTINCT
// (sequence, V, J) of set B; idx is the head of its occurrence list.  "Empty" lives in the index
// word, so a stored hash may legitimately be any 64-bit value including 0.  The index word is
// (low 32 bits of the hash) << 32 | head index; sequence indices are < 2^32 - 1, so an occupied
// slot can never look like SLOT_EMPTY.
struct alignas(16) Slot {
  uint64_t hash;
  uint64_t idx;
};
static_assert(sizeof(Slot) == 16, "Slot must be 16 bytes");
constexpr uint64_t SLOT_EMPTY = ~0ull;

enum VariantKind : uint32_t {  // same numbering as mutation_kind_enum, variants.h:24-31
  VK_IDENTICAL = 0,
  VK_SUBSTITUTION = 1,
  VK_DELETION = 2,
  VK_INSERTION = 3,
  VK_SUB_SUB = 4
};

constexpr int MAXDIFF_HASH = 2;     // compairr.h:113
constexpr int BLOOM_K_HALF = 3;     // bits set in each 32-bit half of a 64-bit Bloom block

// ---- hashing --------------------------------------------------------------------------------

CB_HD uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// Zobrist value of residue r at position p.  A pure function of (seed, p, r) so the table can be
// extended to longer sequences without invalidating hashes already computed (the reference draws
// from glibc random(), zobrist.cc:52-63; results do not depend on the values).
//
// The values are STRUCTURED by the CLASS of the position, class = p mod 4: the 64-bit hash is four
// 16-bit fields, and a residue at a position of class c contributes 16 random bits to field c and
// nothing to the others.  So, h being the XOR of the values of a sequence (and of a 64-bit V/J
// term), field c of h changes only when a position of class c changes: the 48 bits of the other
// three fields are BLIND to class c.  All the single-residue variants of a seed at one position
// (and at every other position of the same class) share them — which is what lets ONE filter word
// answer for all of them (class filters below).
// Round 1 used two classes (position parity, two 32-bit fields).  A word index then sees only
// every other position: with -g (no V/J term) a set of 10^8 CDR3s has ~10^7 distinct values of it,
// the words of the short sequences saturate and 11 % of all candidates passed the filter (measured,
// C4 geometry).  Four classes let the index see three positions out of four.
// Two different sequences collide with probability 2^-16 at worst (they differ in one class
// only); every hash match is verified on the residues anyway.
constexpr uint32_t CB_CLASSES = 4;
CB_HD uint32_t pos_class(uint32_t p) { return p & 3u; }
// The 16-bit values of one position are non-zero and pairwise distinct (redrawn until they are), so
// a substitution always changes the hash and its pattern field; used on the host to fill the table
// (and by the tests), never per probe.
CB_HD uint64_t zobrist_gen(uint64_t seed, uint32_t p, uint32_t r) {
  uint32_t vals[32];
  const uint64_t base = splitmix64(seed ^ 0x5A0B1157ull) + ((uint64_t)p << 16);
  for (uint32_t q = 0; q <= (r & 31u); q++) {
    for (uint32_t attempt = 0;; attempt++) {
      const uint32_t x = (uint32_t)(splitmix64(base + (attempt << 8) + q) >> 48);
      bool fresh = x != 0;
      for (uint32_t k = 0; k < q && fresh; k++) fresh = vals[k] != x;
      if (fresh) {
        vals[q] = x;
        break;
      }
    }
  }
  return (uint64_t)vals[r & 31u] << (16 * pos_class(p));
}
// 32 bits of the hash that depend on every position, for the "possibly the same sequence" tag of
// a table slot.
CB_HD uint32_t slot_tag(uint64_t h) { return (uint32_t)(h >> 32) * 0x9E3779B1u ^ (uint32_t)h; }

// Contribution of the (V gene, J gene) pair: zobrist_v_base[v] ^ zobrist_d_base[j] in the
// reference (zobrist.cc:83-84).  Computed, not tabulated, so no table sized by #V + #J.
CB_HD uint64_t vj_hash(uint64_t seed, uint32_t v, uint32_t j) {
  return splitmix64(splitmix64(seed ^ 0x7E11C0DEull) ^ (((uint64_t)v << 32) | j));
}

// Home slot (hashtable.h:36-41 takes the upper half of the hash): the top bits of h * K, a
// multiplicative mix — every field of the hash reaches them.  Keys ordered by h * K touch the
// table in address order, what the partitioned build relies on (it sorts h * K and gets h back
// with the inverse multiplier: K is odd, the map is a bijection).
constexpr uint64_t CB_HOME_MUL = 0x9E3779B97F4A7C15ull;
constexpr uint64_t CB_HOME_INV = 0xF1DE83E19937733Dull;  // CB_HOME_MUL * CB_HOME_INV == 1 (mod 2^64)
static_assert(CB_HOME_MUL * CB_HOME_INV == 1ull, "inverse multiplier");
CB_HD uint64_t table_home(uint64_t h, uint64_t mask) {
#if defined(__CUDA_ARCH__)
  const int bits = __popcll(mask);
#else
  const int bits = __builtin_popcountll(mask);
#endif
  return ((h * CB_HOME_MUL) >> (64 - bits)) & mask;
}
constexpr int CB_PARTITION_TOP_BIT = 64;  // partition keys are bits [64 - p, 64) of h * CB_HOME_MUL

// Class filters (replace bloom_s, bloompat.h:26-58): FOUR blocked Bloom filters of `nblocks`
// 64-bit words each, laid out back to back; every set-B key is in all four.
//     filter c (words [c nblocks, (c + 1) nblocks)): word picked by the 48 bits of the hash that
//     are blind to class c, 3 + 3 bits picked by field c
// A variant may be looked up in any of them (no false negatives in all four).  The enumeration
// kernels use filter c for a variant whose free residue sits at a position of class c: the 19 (or
// 20) variants at that position — and those at every other position of that class — then read the
// SAME word: it is fetched once per slot, not once per candidate (variant.cu), where a single
// filter cost one random L2 sector per candidate.  Word choice by multiply-shift (any word count);
// normal polarity (1 = present; the reference's is inverted, an implementation detail).
CB_HD uint32_t mulhi32(uint32_t x, uint32_t n) {
#if defined(__CUDA_ARCH__)
  return __umulhi(x, n);
#else
  return (uint32_t)(((uint64_t)x * n) >> 32);
#endif
}
// 32 well-mixed bits of h that do not change when a position of class c changes
CB_HD uint32_t blind_field(uint64_t h, uint32_t c) {
  const uint64_t g = h & ~(0xFFFFull << (16 * c));
  uint32_t m = (uint32_t)g * 0x9E3779B1u + (uint32_t)(g >> 32) * 0x85EBCA77u;
  m ^= m >> 15;
  m *= 0x2C1B3C6Du;
  m ^= m >> 13;
  return m * 0x297A2D39u;
}
// word index of h in the filter serving a free position of class c
CB_HD uint64_t pfilter_word(uint64_t h, uint32_t nblocks, uint32_t c) {
  return (uint64_t)c * nblocks + mulhi32(blind_field(h, c), nblocks);
}
// The 32 bits the bit pattern of h in filter c is cut from: field c — the one field the index of
// that filter does NOT see, and the only one that differs between the candidates of a slot — spread
// over 32 bits by a GF(2)-linear map.  Taken from that field alone, because a seed that is itself in
// set B has its own key in the very word its substitution variants are looked up in: pattern bits
// drawn from the other fields would be the same as the key's, i.e. set, and the test would be down
// to the remaining ones (measured with such a pattern: 3 x the false positives).  Linear, so that
// the pattern field of a variant is the XOR of a per-slot part and a per-(position, residue) part:
// the enumeration loop reads the latter from a 32-bit table and never forms the 64-bit hash.
CB_HD uint32_t class_field(uint64_t h, uint32_t c) { return (uint32_t)(h >> (16 * c)) & 0xFFFFu; }
// Six taps: with three (x * 0x10001 ^ two shifts, half the instructions) the six windows overlap in
// the same input bits and the false-positive rate rose by a third (host simulation, 2e7 keys).
CB_HD uint32_t expand16(uint32_t x) { return x ^ (x << 3) ^ (x << 7) ^ (x << 11) ^ (x << 14) ^ (x << 16); }
CB_HD uint32_t pattern_field(uint64_t h, uint32_t c) { return expand16(class_field(h, c)); }
// Bits per key in each 32-bit half of a filter word: 3 (default) or 2 (compile-time knob for A/B
// runs of the enumeration kernels: four fewer instructions per candidate, ~2.5x the false positives).
#ifndef CB_PATTERN_HALF_BITS
#define CB_PATTERN_HALF_BITS 3
#endif
// Six 5-bit windows of the pattern field.
constexpr int CB_PAT_A0 = 0, CB_PAT_A1 = 16, CB_PAT_A2 = 5;     // low half of the word
constexpr int CB_PAT_B0 = 21, CB_PAT_B1 = 10, CB_PAT_B2 = 26;   // high half
CB_HD uint32_t bloom_pat_lo(uint32_t f) {
  uint32_t p = (1u << ((f >> CB_PAT_A0) & 31)) | (1u << ((f >> CB_PAT_A1) & 31));
  if (CB_PATTERN_HALF_BITS >= 3) p |= 1u << ((f >> CB_PAT_A2) & 31);
  return p;
}
CB_HD uint32_t bloom_pat_hi(uint32_t f) {
  uint32_t p = (1u << ((f >> CB_PAT_B0) & 31)) | (1u << ((f >> CB_PAT_B1) & 31));
  if (CB_PATTERN_HALF_BITS >= 3) p |= 1u << ((f >> CB_PAT_B2) & 31);
  return p;
}
CB_HD uint64_t bloom_pattern(uint32_t f) {
  return (uint64_t)bloom_pat_lo(f) | ((uint64_t)bloom_pat_hi(f) << 32);
}
// x >> (s mod 32): on the device one funnel shift in wrap mode, no separate "& 31"
CB_HD uint32_t shr_wrap(uint32_t x, uint32_t s) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(x, 0u, s);
#else
  return x >> (s & 31);
#endif
}
// Does the filter word (lo, hi) contain bloom_pattern(f)?  The same test as (w & pattern) == pattern,
// written as shifts of the word instead of a mask built from six variable shifts (tests/csrc/
// hd_check.cpp checks the equivalence): no branches.
// f >> k for a constant k.  (Tried: the high half of f * 2^(32-k), an IMAD.HI on the FMA pipe, to
// take the five constant shifts off the ALU pipe, which issues every other cycle and bounds the
// enumeration loop — no measurable change at C3 geometry, 15.9 vs 15.8 ms; plain shifts kept.)
template <int K>
CB_HD uint32_t shr_const(uint32_t f) { return f >> K; }
// bit 0 of the result: the pattern bits of f in the low (high) half of the word are all set
CB_HD uint32_t pattern_half_lo(uint32_t lo, uint32_t f) {
  uint32_t a = shr_wrap(lo, shr_const<CB_PAT_A0>(f)) & shr_wrap(lo, shr_const<CB_PAT_A1>(f));
  if (CB_PATTERN_HALF_BITS >= 3) a &= shr_wrap(lo, shr_const<CB_PAT_A2>(f));
  return a;
}
CB_HD uint32_t pattern_half_hi(uint32_t hi, uint32_t f) {
  uint32_t b = shr_wrap(hi, shr_const<CB_PAT_B0>(f)) & shr_wrap(hi, shr_const<CB_PAT_B1>(f));
  if (CB_PATTERN_HALF_BITS >= 3) b &= shr_wrap(hi, shr_const<CB_PAT_B2>(f));
  return b;
}
CB_HD bool pattern_hit_halves(uint32_t lo, uint32_t hi, uint32_t f) {
  return (pattern_half_lo(lo, f) & pattern_half_hi(hi, f) & 1u) != 0u;
}
CB_HD bool pattern_hit(unsigned long long w, uint32_t f) {
  return pattern_hit_halves((uint32_t)w, (uint32_t)(w >> 32), f);
}
// bit pattern of h in filter c
CB_HD uint64_t pfilter_pattern(uint64_t h, uint32_t c) { return bloom_pattern(pattern_field(h, c)); }

// ---- score summand (overlap.cc:144-166) ---------------------------------------------------------

CB_HD double score_of(int score, bool ignore_counts, uint64_t a, uint64_t b) {
  if (ignore_counts) return 1.0;
  switch (score) {
    case 0:  // product
    case 5:  // MH uses the sum of products
      return (double)a * (double)b;
    case 1:  // ratio
      return (double)a / (double)b;
    case 2:  // min
    case 6:  // Jaccard uses the sum of minima
      return (double)(a < b ? a : b);
    case 3:  // max
      return (double)(a > b ? a : b);
    default:  // 4: mean
      return ((double)a + (double)b) / 2;
  }
}

// ---- variant index space ------------------------------------------------------------------------
//
// For one seed of length L over an alphabet of S residues the engine walks these candidate
// index spaces (the reference materialises a var_s list, variants.cc:242-258; we never do):
//
//   identical      1 candidate                                              variants.cc:260-268
//   substitution   t in [0,(S-1)L): pos = t/(S-1), r' = t%(S-1),
//                  new residue r = r' + (r' >= seed[pos])                   variants.cc:280-293
//   deletion       p in [0,L), emitted iff L > 1 and (p == 0 or
//                  seed[p] != seed[p-1])   (one per run of equal residues)  variants.cc:301-325
//   insertion      (p,r) in [0,L] x [0,S), emitted iff p == 0 or
//                  r != seed[p-1]   (leftmost position of equal results)    variants.cc:329-353
//   sub_sub        i<j, two substitutions as above                          variants.cc:357-400
//
// Every emitted candidate is a distinct variant SEQUENCE of the seed, which is what makes each
// matching (seed, hit) pair count exactly once.

CB_HD uint32_t sub_residue(uint32_t rprime, uint32_t seed_residue) {
  return rprime + (rprime >= seed_residue ? 1u : 0u);
}

// Number of maximal runs of equal residues.
CB_HD uint32_t count_runs(const uint8_t* s, uint32_t len) {
  uint32_t runs = 0;
  for (uint32_t p = 0; p < len; p++)
    if (p == 0 || s[p] != s[p - 1]) runs++;
  return runs;
}

// Closed-form variant count = what generate_variants() emits (variants.cc:402-428).
CB_HD uint64_t probe_count(const uint8_t* s, uint32_t len, uint32_t sigma, int d, bool indels) {
  uint64_t L = len, S = sigma;
  uint64_t n = 1;
  if (d >= 1) {
    n += (S - 1) * L;
    if (indels) {
      if (L > 1) n += count_runs(s, len);
      n += S * (L + 1) - L;
    }
  }
  if (d >= 2) n += (S - 1) * (S - 1) * (L * (L - 1) / 2);
  return n;
}

// Exact verification: is `hit` exactly the seed with THIS edit applied? (check_variant,
// variants.cc:166-240).  Verifying the specific edit, not just "distance <= d", is what rejects
// a hash collision between two different variants of one seed and keeps the once-only count.
CB_HD bool verify_variant(const uint8_t* seed, uint32_t slen, const uint8_t* hit, uint32_t hlen,
                          uint32_t kind, uint32_t pos1, uint32_t r1, uint32_t pos2, uint32_t r2) {
  switch (kind) {
    case VK_IDENTICAL:
      if (hlen != slen) return false;
      for (uint32_t p = 0; p < slen; p++)
        if (seed[p] != hit[p]) return false;
      return true;
    case VK_SUBSTITUTION:
      if (hlen != slen || hit[pos1] != r1) return false;
      for (uint32_t p = 0; p < slen; p++)
        if (p != pos1 && seed[p] != hit[p]) return false;
      return true;
    case VK_DELETION:
      if (hlen + 1 != slen) return false;
      for (uint32_t p = 0; p < pos1; p++)
        if (seed[p] != hit[p]) return false;
      for (uint32_t p = pos1; p < hlen; p++)
        if (seed[p + 1] != hit[p]) return false;
      return true;
    case VK_INSERTION:
      if (hlen != slen + 1 || hit[pos1] != r1) return false;
      for (uint32_t p = 0; p < pos1; p++)
        if (seed[p] != hit[p]) return false;
      for (uint32_t p = pos1; p < slen; p++)
        if (seed[p] != hit[p + 1]) return false;
      return true;
    case VK_SUB_SUB:
      if (hlen != slen || hit[pos1] != r1 || hit[pos2] != r2) return false;
      for (uint32_t p = 0; p < slen; p++)
        if (p != pos1 && p != pos2 && seed[p] != hit[p]) return false;
      return true;
    default:
      return false;
  }
}

}  // namespace cb
