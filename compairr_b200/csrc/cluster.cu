// cluster.cu — `compairr -c`: single-linkage clusters of the d-neighbour graph of ONE set.
//
// Replaces, of src/cluster.cc: hash_insert_cluster + the insert loop (:57-69, :333-341), the
// network phase network_thread / process_variants / find_variant_matches / process_trad
// (:71-274) and the clustering phase process_seed + the seed loop (:277-300, :356-407) and
// the size sort (:411).  The network phase is the overlap hot path run as a self-comparison
// in pairs mode (same enumeration, filter, table and verify kernels, no matrix): every match
// (seed, hit, variant) comes back as a pair whose high half carries the variant descriptor.
// The host then lays each seed's hits out in the order the reference finds them — variant
// enumeration order (variants.cc:402-428), equal sequences in index order (probe-chain order of
// a table filled in index order) — because the reference's breadth-first walk, and so the ORDER
// of the rows it prints, depends on that order; the partition into clusters does not.
#include <algorithm>
#include <atomic>
#include <new>
#include <thread>
#include <vector>

#include "engine_internal.h"

using namespace cb;

namespace {

constexpr uint32_t NO_CLUSTER = 0xffffffffu;  // cluster.cc:24

// Position of a variant in the reference's enumeration: identical < substitutions (pos, residue)
// < deletions (pos) < insertions (pos, residue) < double substitutions (pos1, res1, pos2, res2)
// — the loop nests of generate_variants_0/_1/_2 (variants.cc:260-400).  The kind numbering is the
// reference's (variants.h:24-31), so it orders as the calls in generate_variants do.
// Descriptor layout: kind(3) | res1(5) | res2(5) | pos1(9) | pos2(9) (device_utils.cuh pack_var).
inline uint64_t enumeration_rank(uint32_t var) {
  const uint64_t kind = var & 7, r1 = (var >> 3) & 31, r2 = (var >> 8) & 31;
  const uint64_t pos1 = (var >> 13) & 511, pos2 = (var >> 22) & 511;
  return (kind << 28) | (pos1 << 19) | (r1 << 14) | (pos2 << 5) | r2;
}

struct Edge {
  uint64_t key;  // enumeration rank << 32 | hit
};

}  // namespace

extern "C" int cb_cluster(cb_ctx* c, cb_dset* s, uint32_t* order_out, uint32_t* cluster_no_out,
                          uint32_t* cluster_size_out, uint64_t* n_clusters_out, uint64_t* n_edges_out) {
  if (!c || !s || !order_out || !cluster_no_out || !cluster_size_out)
    return cb_fail(c, CB_ERR_INVALID, "cb_cluster: NULL argument");
  if (n_clusters_out) *n_clusters_out = 0;
  if (n_edges_out) *n_edges_out = 0;
  const uint64_t n = s->n;
  if (n == 0) return CB_OK;
  if (n >= NO_CLUSTER) return cb_fail(c, CB_ERR_LIMIT, "cb_cluster: more than 2^32-2 sequences");
  if (s->index_base != 0) return cb_fail(c, CB_ERR_INVALID, "cb_cluster: the set must have index_base 0");
  int rc = CB_OK;
  if (c->b != s || (!c->d_table && c->cfg.differences <= MAXDIFF_HASH)) {
    rc = cb_build_b(c, s);
    if (rc) return rc;
  }

  // ---- network: hits of every seed, in the reference's order -----------------------------------
  std::vector<uint64_t> start;  // n + 1
  std::vector<uint32_t> network;
  try {
    start.assign(n + 1, 0);
  } catch (...) {
    return cb_fail(c, CB_ERR_NOMEM, "cb_cluster: out of host memory");
  }
  const cb_config saved = c->cfg;
  c->cfg.want_pairs = 1;
  c->cfg.no_matrix = 1;
  c->cfg.mode = CB_MODE_MATRIX;
  c->network_mode = true;
  c->pending.clear();
  const uint64_t chunk = 1u << 20;
  const unsigned hw = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  std::vector<uint64_t> cnt;
  std::vector<Edge> edges;
  for (uint64_t first = 0; first < n && !rc; first += chunk) {
    const uint64_t m = std::min(chunk, n - first);
    rc = cb_run(c, s, first, m);
    if (rc) break;
    try {
      // counting sort of the chunk's pairs by seed, self hits dropped (cluster.cc:105 `seed != hit`)
      cnt.assign(m + 1, 0);
      for (const cb_pair& p : c->pending)
        if (p.a != (uint32_t)p.b) cnt[p.a - first + 1]++;
      for (uint64_t k = 0; k < m; k++) cnt[k + 1] += cnt[k];
      const uint64_t e = cnt[m];
      edges.resize(e);
      {
        std::vector<uint64_t> at(cnt.begin(), cnt.end() - 1);
        for (const cb_pair& p : c->pending)
          if (p.a != (uint32_t)p.b)
            edges[at[p.a - first]++].key = (enumeration_rank((uint32_t)(p.b >> 32)) << 32) | (uint32_t)p.b;
      }
      c->pending.clear();
      // each seed's hits in enumeration order (host threads over seed ranges)
      std::atomic<uint64_t> next{0};
      auto work = [&] {
        for (;;) {
          const uint64_t k0 = next.fetch_add(4096);
          if (k0 >= m) return;
          const uint64_t k1 = std::min(m, k0 + 4096);
          for (uint64_t k = k0; k < k1; k++)
            if (cnt[k + 1] - cnt[k] > 1)
              std::sort(edges.begin() + cnt[k], edges.begin() + cnt[k + 1],
                        [](const Edge& x, const Edge& y) { return x.key < y.key; });
        }
      };
      std::vector<std::thread> th;
      for (unsigned t = 1; t < hw && e > (1u << 16); t++) th.emplace_back(work);
      work();
      for (auto& t : th) t.join();
      const uint64_t base = network.size();
      network.resize(base + e);
      for (uint64_t k = 0; k < e; k++) network[base + k] = (uint32_t)edges[k].key;
      for (uint64_t k = 0; k < m; k++) start[first + k + 1] = base + cnt[k + 1];
    } catch (const std::bad_alloc&) {
      rc = cb_fail(c, CB_ERR_NOMEM, "cb_cluster: out of host memory for the network");
    }
  }
  c->cfg = saved;
  c->network_mode = false;
  c->pending.clear();
  if (rc) return rc;
  if (n_edges_out) *n_edges_out = network.size();

  // ---- clustering: the reference's breadth-first walk (cluster.cc:277-300, 356-407) -------------
  // order_out doubles as the queue: a cluster's members are appended in the order they are
  // reached, and scanned in that order for their own hits.
  std::vector<uint32_t> cid;
  struct Cl {
    uint32_t first_row, size;
  };
  std::vector<Cl> clusters;
  try {
    cid.assign(n, NO_CLUSTER);
    uint64_t rows = 0;
    for (uint64_t seed = 0; seed < n; seed++) {
      if (cid[seed] != NO_CLUSTER) continue;
      const uint32_t id = (uint32_t)clusters.size();
      const uint64_t row0 = rows;
      cid[seed] = id;
      order_out[rows++] = (uint32_t)seed;
      for (uint64_t scan = row0; scan < rows; scan++) {
        const uint32_t x = order_out[scan];
        for (uint64_t k = start[x]; k < start[x + 1]; k++) {
          const uint32_t hit = network[k];
          if (cid[hit] == NO_CLUSTER) {
            cid[hit] = id;
            order_out[rows++] = hit;
          }
        }
      }
      clusters.push_back({(uint32_t)row0, (uint32_t)(rows - row0)});
    }
    // largest first; equal sizes keep the order of their first members (qsort on glibc is a
    // stable merge sort, cluster.cc:41-55,411)
    std::stable_sort(clusters.begin(), clusters.end(), [](const Cl& x, const Cl& y) { return x.size > y.size; });
    std::vector<uint32_t> tmp(order_out, order_out + n);
    uint64_t row = 0;
    for (size_t k = 0; k < clusters.size(); k++)
      for (uint32_t t = 0; t < clusters[k].size; t++, row++) {
        order_out[row] = tmp[clusters[k].first_row + t];
        cluster_no_out[row] = (uint32_t)k + 1;
        cluster_size_out[row] = clusters[k].size;
      }
  } catch (const std::bad_alloc&) {
    return cb_fail(c, CB_ERR_NOMEM, "cb_cluster: out of host memory");
  }
  if (n_clusters_out) *n_clusters_out = clusters.size();
  return CB_OK;
}
