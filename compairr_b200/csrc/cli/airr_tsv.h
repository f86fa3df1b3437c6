// airr_tsv.h — AIRR TSV reader: file -> structure-of-arrays sequence set + the strings the
// writers need.  Behaviour-compatible with the reference's db_read()/parse_airr_tsv_*()
// (src/db.cc:172-901): same column names, same defaults, same error texts and exit codes, ids
// numbered in first-seen order.  Host-only; the GPU never sees strings.
#pragma once
#include <stdint.h>

#include <memory>
#include <new>
#include <utility>
#include <string>
#include <unordered_map>
#include <vector>

#include "options.h"

struct GeneTables {  // V and J gene names are shared by both sets (db.cc:119-125)
  std::vector<std::string> v_names, j_names;
  std::unordered_map<std::string, uint32_t> v_map, j_map;
};

// Column storage: a std::vector whose resize() does not zero the new elements.  The reader sizes
// the columns of a 10^8-line file once (4.5 GB) and its threads then fill them in parallel; with
// value-initialising resize() that was 1.5 s of serial memset and page faults on one core.
template <class T>
struct default_init_allocator : std::allocator<T> {
  template <class U>
  struct rebind {
    using other = default_init_allocator<U>;
  };
  using std::allocator<T>::allocator;
  template <class U, class... Args>
  void construct(U* p, Args&&... args) {
    if constexpr (sizeof...(Args) == 0)
      ::new ((void*)p) U;
    else
      ::new ((void*)p) U(std::forward<Args>(args)...);
  }
};
template <class T>
using RawVec = std::vector<T, default_init_allocator<T>>;

struct SeqDb {
  RawVec<uint8_t> residues;
  RawVec<uint64_t> offsets{0};
  RawVec<uint32_t> v, j, rep;
  RawVec<uint64_t> count;
  // sequence ids (only filled when needed: pairs / existence) and -k columns (tab-joined):
  // NUL-terminated strings in one arena each
  RawVec<char> id_arena, keep_arena;
  RawVec<uint64_t> id_off, keep_off;
  bool has_ids() const { return !id_off.empty(); }
  const char* seq_id(uint64_t i) const { return id_arena.data() + id_off[i]; }
  const char* keep(uint64_t i) const { return keep_arena.data() + keep_off[i]; }
  std::vector<std::string> rep_names;
  std::unordered_map<std::string, uint32_t> rep_map;
  unsigned longest = 0, shortest = ~0u;
  uint64_t total_count = 0, ignored_unknown = 0, ignored_empty = 0;
  uint64_t n() const { return v.size(); }
};

// Reads `filename` ("-" = stdin) into db; logs the summary block; exits with the reference's
// messages on malformed input.
void read_airr_tsv(const char* filename, const Options& o, bool require_sequence_id, bool want_ids,
                   const char* default_repertoire_id, GeneTables& genes, SeqDb& db);

// progress/timing lines in the reference's format (util.cc:32-70)
void progress_begin(const Options& o, const char* prompt);
void progress_end(const Options& o, const char* prompt);
