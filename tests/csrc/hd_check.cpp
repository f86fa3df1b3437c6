// hd_check.cpp — host-side check of the integer arithmetic in compairr_b200/csrc/common.cuh that
// the kernels and the engine share (CB_HD functions).  Compiled with g++ by
// tests/test_parity_fields.py; test infrastructure, not a product path.  No GPU involved.
//
// Checks, on random sequences:
//   1. the Zobrist values are structured by position parity: a substitution at an odd position
//      leaves field_even() of the hash unchanged, one at an even position leaves field_odd()
//      unchanged — for insertions (everything behind the insertion shifts by one) and for the
//      second substitution of a double substitution alike;
//   2. therefore pfilter_word() is the same for every residue at a slot, in the filter the
//      enumeration kernels pick for that slot (odd free position -> filter E, even -> filter O);
//   3. the two filters never overlap (E words in [0, n), O words in [n, 2n));
//   4. a key inserted into both filters is found by either lookup (no false negatives);
//   5. probe_count() equals a literal enumeration count of the reference's rules
//      (variants.cc:260-428) for small sequences.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <set>
#include <string>
#include <vector>

#include "../../compairr_b200/csrc/common.cuh"

using namespace cb;

static uint64_t rng_state = 88172645463325252ull;
static uint64_t rnd() {
  rng_state ^= rng_state << 13;
  rng_state ^= rng_state >> 7;
  rng_state ^= rng_state << 17;
  return rng_state;
}

static uint64_t hash_of(const std::vector<uint8_t>& s, uint64_t seed, uint64_t vj) {
  uint64_t h = vj;
  for (uint32_t p = 0; p < s.size(); p++) h ^= zobrist_gen(seed, p, s[p]);
  return h;
}

#define CHECK(cond)                                                         \
  do {                                                                      \
    if (!(cond)) {                                                          \
      fprintf(stderr, "hd_check: %s failed at line %d\n", #cond, __LINE__); \
      return 1;                                                             \
    }                                                                       \
  } while (0)

int main() {
  const uint64_t seed = 1;
  for (int sigma : {4, 20}) {
    for (int it = 0; it < 300; it++) {
      const uint32_t L = 1 + rnd() % 30;
      const uint32_t nblocks = 16 + (uint32_t)(rnd() % 1000003);
      std::vector<uint8_t> s(L);
      for (auto& x : s) x = (uint8_t)(rnd() % sigma);
      const uint64_t vj = vj_hash(seed, (uint32_t)(rnd() % 60), (uint32_t)(rnd() % 13));
      const uint64_t h = hash_of(s, seed, vj);
      std::vector<unsigned long long> filt(2 * (size_t)nblocks, 0);
      // 4: insert h into both filters, look it up both ways
      filt[pfilter_word(h, nblocks, true)] |= pfilter_pattern(h, true);
      filt[pfilter_word(h, nblocks, false)] |= pfilter_pattern(h, false);
      for (bool odd : {true, false}) {
        const uint64_t w = pfilter_word(h, nblocks, odd);
        CHECK(odd ? w < nblocks : (w >= nblocks && w < 2ull * nblocks));  // 3
        CHECK((filt[w] & pfilter_pattern(h, odd)) == pfilter_pattern(h, odd));
      }
      // 1 + 2: substitutions
      for (uint32_t p = 0; p < L; p++) {
        const bool odd = p & 1;
        for (int r = 0; r < sigma; r++) {
          std::vector<uint8_t> t = s;
          t[p] = (uint8_t)r;
          const uint64_t hv = hash_of(t, seed, vj);
          CHECK(hv == (h ^ zobrist_gen(seed, p, s[p]) ^ zobrist_gen(seed, p, (uint32_t)r)));
          CHECK(odd ? field_even(hv) == field_even(h) : field_odd(hv) == field_odd(h));
          CHECK(pfilter_word(hv, nblocks, odd) == pfilter_word(h, nblocks, odd));
          // 1 for double substitutions: a second one at j > p of either parity
          for (uint32_t j = p + 1; j < L && j < p + 4; j++) {
            std::vector<uint8_t> u = t;
            u[j] = (uint8_t)((s[j] + 1) % sigma);
            const uint64_t h2 = hash_of(u, seed, vj);
            CHECK(pfilter_word(h2, nblocks, j & 1) == pfilter_word(hv, nblocks, j & 1));
          }
        }
      }
      // 1 + 2: insertions before position p (p = L appends)
      for (uint32_t p = 0; p <= L; p++) {
        uint64_t first = 0;
        for (int r = 0; r < sigma; r++) {
          std::vector<uint8_t> t(s.begin(), s.begin() + p);
          t.push_back((uint8_t)r);
          t.insert(t.end(), s.begin() + p, s.end());
          const uint64_t w = pfilter_word(hash_of(t, seed, vj), nblocks, p & 1);
          if (r == 0) first = w;
          CHECK(w == first);
        }
      }
    }
  }
  // pattern_hit(w, f) is (w & bloom_pattern(f)) == bloom_pattern(f), for sparse and dense words
  for (int it = 0; it < 200000; it++) {
    const uint32_t f = (uint32_t)rnd();
    unsigned long long w = rnd();
    if (it % 3 == 0) w |= rnd() | rnd();
    if (it % 5 == 0) w |= bloom_pattern(f);
    CHECK(pattern_hit(w, f) == ((w & bloom_pattern(f)) == bloom_pattern(f)));
  }
  // 5: probe_count against a literal enumeration of distinct variant strings + the rules
  for (int it = 0; it < 200; it++) {
    const int sigma = (it & 1) ? 4 : 20;
    const uint32_t L = 1 + rnd() % 6;
    std::vector<uint8_t> s(L);
    for (auto& x : s) x = (uint8_t)(rnd() % (it % 3 == 0 ? 2 : sigma));  // low complexity: runs
    for (int d = 0; d <= 2; d++)
      for (int indels = 0; indels <= (d == 1 ? 1 : 0); indels++) {
        std::set<std::string> seen;
        uint64_t n = 1;
        seen.insert(std::string(s.begin(), s.end()));
        auto add = [&](const std::vector<uint8_t>& t) { n += seen.insert(std::string(t.begin(), t.end())).second ? 1 : 0; };
        if (d >= 1) {
          for (uint32_t p = 0; p < L; p++)
            for (int r = 0; r < sigma; r++)
              if (r != s[p]) {
                auto t = s;
                t[p] = (uint8_t)r;
                add(t);
              }
          if (indels) {
            if (L > 1)
              for (uint32_t p = 0; p < L; p++) {
                auto t = s;
                t.erase(t.begin() + p);
                add(t);
              }
            for (uint32_t p = 0; p <= L; p++)
              for (int r = 0; r < sigma; r++) {
                auto t = s;
                t.insert(t.begin() + p, (uint8_t)r);
                add(t);
              }
          }
        }
        if (d >= 2)
          for (uint32_t i = 0; i < L; i++)
            for (uint32_t j = i + 1; j < L; j++)
              for (int v = 0; v < sigma; v++)
                for (int w = 0; w < sigma; w++)
                  if (v != s[i] && w != s[j]) n++;  // all distinct from each other and from d <= 1
        CHECK(probe_count(s.data(), L, (uint32_t)sigma, d, indels != 0) == n);
      }
  }
  printf("hd_check ok\n");
  return 0;
}
