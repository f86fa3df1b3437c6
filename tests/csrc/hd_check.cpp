// hd_check.cpp — host-side check of the integer arithmetic in compairr_b200/csrc/common.cuh that
// the kernels and the engine share (CB_HD functions).  Compiled with g++ by
// tests/test_parity_fields.py; test infrastructure, not a product path.  No GPU involved.
//
// Checks, on random sequences:
//   1. the Zobrist values are structured by position class (p mod 4): a substitution at a position
//      of class c leaves blind_field(h, c) unchanged — for insertions (everything behind the
//      insertion shifts by one) and for the second substitution of a double substitution alike;
//   2. therefore pfilter_word() is the same for every residue at a slot, in the filter the
//      enumeration kernels pick for that slot (free position of class c -> filter c);
//   3. the four filters never overlap (filter c owns words [c n, (c + 1) n));
//   4. a key inserted into all four filters is found by every lookup (no false negatives);
//   5. probe_count() equals a literal enumeration count of the reference's rules
//      (variants.cc:260-428) for small sequences;
//   6. the home-slot multiplier is invertible (the partitioned build sorts h * K and recovers h),
//      table_home() stays in range and sees every class of positions.
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
      std::vector<unsigned long long> filt(CB_CLASSES * (size_t)nblocks, 0);
      // 4: insert h into all four filters, look it up in each
      for (uint32_t c = 0; c < CB_CLASSES; c++) filt[pfilter_word(h, nblocks, c)] |= pfilter_pattern(h, c);
      for (uint32_t c = 0; c < CB_CLASSES; c++) {
        const uint64_t w = pfilter_word(h, nblocks, c);
        CHECK(w >= (uint64_t)c * nblocks && w < (uint64_t)(c + 1) * nblocks);  // 3
        CHECK(pattern_hit(filt[w], pattern_field(h, c)));
      }
      // 6
      CHECK(h * CB_HOME_MUL * CB_HOME_INV == h);
      for (int bits : {3, 21, 28, 33}) CHECK(table_home(h, (1ull << bits) - 1) < (1ull << bits));
      // 1 + 2: substitutions
      for (uint32_t p = 0; p < L; p++) {
        const uint32_t c = pos_class(p);
        for (int r = 0; r < sigma; r++) {
          std::vector<uint8_t> t = s;
          t[p] = (uint8_t)r;
          const uint64_t hv = hash_of(t, seed, vj);
          CHECK(hv == (h ^ zobrist_gen(seed, p, s[p]) ^ zobrist_gen(seed, p, (uint32_t)r)));
          CHECK(blind_field(hv, c) == blind_field(h, c));
          CHECK(pfilter_word(hv, nblocks, c) == pfilter_word(h, nblocks, c));
          if (r != s[p]) {  // ... while the hash itself, its pattern field and (almost always) the home slot move
            CHECK(hv != h);
            CHECK(pattern_field(hv, c) != pattern_field(h, c));
            // the pattern field is linear: per-slot part ^ per-(position, residue) part
            CHECK(pattern_field(hv, c) == (pattern_field(h ^ zobrist_gen(seed, p, s[p]), c) ^ pattern_field(zobrist_gen(seed, p, (uint32_t)r), c)));
            for (uint32_t o = 0; o < CB_CLASSES; o++)
              if (o != c) CHECK(pattern_field(hv, o) == pattern_field(h, o));
          }
          // 1 for double substitutions: a second one at j > p of any class
          for (uint32_t j = p + 1; j < L && j < p + 6; j++) {
            std::vector<uint8_t> u = t;
            u[j] = (uint8_t)((s[j] + 1) % sigma);
            const uint64_t h2 = hash_of(u, seed, vj);
            CHECK(pfilter_word(h2, nblocks, pos_class(j)) == pfilter_word(hv, nblocks, pos_class(j)));
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
          const uint64_t w = pfilter_word(hash_of(t, seed, vj), nblocks, pos_class(p));
          if (r == 0) first = w;
          CHECK(w == first);
        }
      }
    }
  }
  // 6: the home slot depends on positions of every class (flip one residue per class, the slot
  // moves in the large majority of cases)
  {
    int moved[CB_CLASSES] = {0, 0, 0, 0};
    for (int it = 0; it < 400; it++) {
      std::vector<uint8_t> s(16);
      for (auto& x : s) x = (uint8_t)(rnd() % 20);
      const uint64_t h = hash_of(s, seed, 0);
      for (uint32_t p = 4; p < 8; p++) {
        auto t = s;
        t[p] = (uint8_t)((t[p] + 1) % 20);
        moved[pos_class(p)] += table_home(hash_of(t, seed, 0), (1ull << 24) - 1) != table_home(h, (1ull << 24) - 1);
      }
    }
    for (uint32_t c = 0; c < CB_CLASSES; c++) CHECK(moved[c] > 390);
  }
  // pattern_hit(w, f) is (w & bloom_pattern(f)) == bloom_pattern(f), for sparse and dense words
  for (int it = 0; it < 200000; it++) {
    const uint32_t f = (uint32_t)rnd();
    unsigned long long w = rnd();
    if (it % 3 == 0) w |= rnd() | rnd();
    if (it % 5 == 0) w |= bloom_pattern(f);
    CHECK(pattern_hit(w, f) == ((w & bloom_pattern(f)) == bloom_pattern(f)));
    CHECK(__builtin_popcountll(bloom_pattern(f)) <= 2 * CB_PATTERN_HALF_BITS && bloom_pattern(f) != 0);
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
