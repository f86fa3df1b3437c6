// cluster_cmd.cpp — see cluster_cmd.h.  Output formats follow src/cluster.cc:417-446 (cluster rows),
// :449-450 (log) and src/dedup.cc:27-59,169-173,185-195 (dedup rows, log).
#include "cluster_cmd.h"

#include <string.h>

#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "airr_tsv.h"
#include "compairr_b200.h"
#include "row_writer.h"

namespace {

struct OneSet {
  GeneTables genes;
  SeqDb db;
  cb_ctx* ctx = nullptr;
  cb_dset* dev = nullptr;
};

[[noreturn]] void engine_fatal(cb_ctx* c) {
  const std::string msg = c ? cb_last_error(c) : cb_global_error();
  if (c) cb_destroy(c);
  fatal(msg.c_str());
}

// COMPAIRR_B200_TRACE=1: wall-clock marks on stderr (as in overlap_cmd.cpp)
const auto g_t0 = std::chrono::steady_clock::now();
void mark(const char* what) {
  static const bool trace = getenv("COMPAIRR_B200_TRACE") != nullptr;
  if (trace)
    fprintf(stderr, "[trace] %8.3f s  %s\n",
            std::chrono::duration<double>(std::chrono::steady_clock::now() - g_t0).count(), what);
}

// db_read(d, filename, false, "1") + upload (both commands read one set, sequence ids optional).
// The CUDA context comes up on another thread while the host parses (a throw-away engine
// context leaves the device's primary context initialised).
void load(const Options& o, OneSet& s, bool want_ids) {
  std::thread warm([&] {
    cb_config w{};
    w.abi_version = CB_ABI_VERSION;
    w.device = o.device;
    w.alphabet_size = o.alphabet_size;
    w.n_reps_a = 1;
    w.no_matrix = 1;
    cb_ctx* c = nullptr;
    if (cb_create(&w, &c) == 0) cb_destroy(c);
  });
  struct Joiner {
    std::thread& t;
    ~Joiner() { if (t.joinable()) t.join(); }
  } guard{warm};
  read_airr_tsv(o.input1, o, false, want_ids, "1", s.genes, s.db);
  mark("set read");
  warm.join();
  mark("warm-up joined");
  cb_config cfg{};
  cfg.abi_version = CB_ABI_VERSION;
  cfg.device = o.device;
  cfg.alphabet_size = o.alphabet_size;
  cfg.differences = (int32_t)std::min<int64_t>(o.differences, 1 << 20);
  cfg.indels = o.indels;
  cfg.ignore_genes = o.ignore_genes;
  cfg.ignore_counts = o.ignore_counts;
  cfg.mode = CB_MODE_MATRIX;
  cfg.no_matrix = 1;
  cfg.n_reps_a = (uint32_t)std::max<size_t>(s.db.rep_names.size(), 1);
  cfg.seed = 1;
  if (cb_create(&cfg, &s.ctx)) engine_fatal(nullptr);
  cb_set h{};
  h.n = s.db.n();
  h.residues = s.db.residues.data();
  h.offsets = s.db.offsets.data();
  h.v_gene = s.db.v.data();
  h.j_gene = s.db.j.data();
  h.rep = s.db.rep.data();
  h.count = s.db.count.data();
  h.n_reps = (uint32_t)s.db.rep_names.size();
  if (cb_upload(s.ctx, &h, &s.dev)) engine_fatal(s.ctx);
  mark("set uploaded + hashed");
}

// The engine is not torn down: main() leaves with _exit() once the files are on disk, and
// destroying a context with multi-GB pools takes longer (0.2 - 0.6 s measured) than anything the
// operating system does to reclaim it.
void unload(OneSet& s) {
  if (getenv("COMPAIRR_B200_TEARDOWN")) {
    cb_free_set(s.ctx, s.dev);
    cb_destroy(s.ctx);
  }
}

void append_sequence(std::string& buf, const Options& o, const SeqDb& d, uint64_t i) {
  const char* alpha = o.nucleotides ? "acgt" : "ACDEFGHIKLMNPQRSTVWY";  // db.cc:73-74
  for (uint64_t p = d.offsets[i]; p < d.offsets[i + 1]; p++) buf += alpha[d.residues[p]];
}

void log_gene_counts(const GeneTables& g) {
  fprintf(g_log, "Unique V genes:    %lu\n", (unsigned long)g.v_names.size());
  fprintf(g_log, "Unique J genes:    %lu\n", (unsigned long)g.j_names.size());
}

}  // namespace

void cluster_command(const Options& o, FILE* outfile) {
  fprintf(g_log, "Immune receptor repertoire clustering\n\n");
  OneSet s;
  load(o, s, true);
  const SeqDb& d = s.db;
  const uint64_t n = d.n();
  fprintf(g_log, "\n");
  log_gene_counts(s.genes);
  fprintf(g_log, "\n");

  // upload + hash + table/Bloom build happened in load() / happens inside cb_cluster; the
  // reference's phase names are kept for the log
  progress_begin(o, "Hashing sequences:");
  progress_end(o, "Hashing sequences:");
  std::vector<uint32_t> order(n), no(n), size(n);
  uint64_t clusters = 0;
  progress_begin(o, "Building network: ");
  if (cb_cluster(s.ctx, s.dev, order.data(), no.data(), size.data(), &clusters, nullptr)) engine_fatal(s.ctx);
  progress_end(o, "Building network: ");
  mark("clusters computed");
  progress_begin(o, "Clustering:       ");  // done inside cb_cluster, as is the size sort
  progress_end(o, "Clustering:       ");
  progress_begin(o, "Sorting clusters: ");
  progress_end(o, "Sorting clusters: ");
  unload(s);

  progress_begin(o, "Writing clusters: ");
  fprintf(outfile, "#cluster_no\tcluster_size\trepertoire_id\tsequence_id\tduplicate_count\tv_call\tj_call\t%s\n",
          o.seq_header);
  write_rows_parallel(outfile, n, host_threads(o.threads, o.threads_given), [&](uint64_t k0, uint64_t k1, std::string& buf) {
    for (uint64_t k = k0; k < k1; k++) {
      const uint64_t a = order[k];
      append_u64(buf, no[k]);
      buf += '\t';
      append_u64(buf, size[k]);
      buf += '\t';
      buf += d.rep_names[d.rep[a]];
      buf += '\t';
      if (d.has_ids()) buf += d.seq_id(a);
      buf += '\t';
      append_u64(buf, d.count[a]);
      buf += '\t';
      buf += s.genes.v_names[d.v[a]];
      buf += '\t';
      buf += s.genes.j_names[d.j[a]];
      buf += '\t';
      append_sequence(buf, o, d, a);
      buf += '\n';
    }
  });
  progress_end(o, "Writing clusters: ");
  mark("clusters written");
  fprintf(g_log, "\n");
  fprintf(g_log, "Clusters:          %u\n", (unsigned)clusters);
}

void dedup_command(const Options& o, FILE* outfile) {
  OneSet s;
  load(o, s, false);
  const SeqDb& d = s.db;
  const uint64_t n = d.n();
  log_gene_counts(s.genes);

  fprintf(outfile, "repertoire_id\tduplicate_count");
  if (!o.ignore_genes) fprintf(outfile, "\tv_call\tj_call");
  fprintf(outfile, "\t%s\n", o.seq_header);

  std::vector<uint32_t> leader(n);
  std::vector<uint64_t> count(n);
  uint64_t merged = 0;
  progress_begin(o, "Deduplicating:    ");
  if (cb_dedup(s.ctx, s.dev, leader.data(), count.data(), &merged)) engine_fatal(s.ctx);
  progress_end(o, "Deduplicating:    ");
  mark("groups computed");
  unload(s);
  fprintf(g_log, "Duplicates merged: %lu\n", (unsigned long)merged);

  progress_begin(o, "Writing output:   ");
  write_rows_parallel(outfile, n, host_threads(o.threads, o.threads_given), [&](uint64_t i0, uint64_t i1, std::string& buf) {
    for (uint64_t i = i0; i < i1; i++) {
      if (leader[i] != i) continue;  // a group is reported once, at its first member
      buf += d.rep_names[d.rep[i]];
      buf += '\t';
      append_u64(buf, count[i]);
      if (!o.ignore_genes) {
        buf += '\t';
        buf += s.genes.v_names[d.v[i]];
        buf += '\t';
        buf += s.genes.j_names[d.j[i]];
      }
      buf += '\t';
      append_sequence(buf, o, d, i);
      buf += '\n';
    }
  });
  progress_end(o, "Writing output:   ");
  mark("output written");
  fprintf(g_log, "\n");
}
