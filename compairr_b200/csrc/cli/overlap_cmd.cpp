// overlap_cmd.cpp — see overlap_cmd.h.  Output formats follow src/overlap.cc:540-577 (values),
// :908-925 (pairs header), :455-507 (pairs rows), :944-1039 (matrix layouts).
#include "overlap_cmd.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "airr_tsv.h"
#include "compairr_b200.h"
#include "row_writer.h"

namespace {

struct RepStats {
  std::vector<uint64_t> size, count;
  std::vector<double> sq_count;
  std::vector<unsigned> order;  // display order: strcmp on the repertoire id (overlap.cc:130-142,659-667)
  uint64_t sum_size = 0, sum_count = 0;
};

// Per-repertoire sizes, total counts and sums of squared counts (overlap.cc:640-657, 733-750).  The
// reference adds (double)(c * c) in file order; here host threads add the same u64 products exactly
// (128-bit partial sums) and, as long as a repertoire's total stays below 2^53 — every partial sum
// of the serial loop is then exact too — the result is the same double, bit for bit.  A repertoire
// beyond that (counts in the hundreds of millions) is summed serially, as the reference does.
RepStats rep_stats(const SeqDb& d, int threads) {
  RepStats s;
  const size_t r = d.rep_names.size();
  s.size.assign(r, 0);
  s.count.assign(r, 0);
  s.sq_count.assign(r, 0.0);
  const uint64_t n = d.n();
  const unsigned nt = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)std::max(threads, 1), n / (1u << 20)));
  std::vector<std::vector<uint64_t>> size(nt, std::vector<uint64_t>(r, 0)), count(nt, std::vector<uint64_t>(r, 0));
  std::vector<std::vector<unsigned __int128>> sq(nt, std::vector<unsigned __int128>(r, 0));
  {
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; t++)
      pool.emplace_back([&, t] {
        for (uint64_t i = n * t / nt; i < n * (t + 1) / nt; i++) {
          const unsigned k = d.rep[i];
          const uint64_t c = d.count[i];
          size[t][k]++;
          count[t][k] += c;
          sq[t][k] += (unsigned __int128)(c * c);  // u64 product, as overlap.cc:654
        }
      });
    for (auto& th : pool) th.join();
  }
  bool exact = true;
  for (size_t k = 0; k < r; k++) {
    unsigned __int128 q = 0;
    for (unsigned t = 0; t < nt; t++) {
      s.size[k] += size[t][k];
      s.count[k] += count[t][k];
      q += sq[t][k];
    }
    if (q >= ((unsigned __int128)1 << 53)) exact = false;
    s.sq_count[k] = (double)(uint64_t)q;
  }
  if (!exact) {
    s.sq_count.assign(r, 0.0);
    for (uint64_t i = 0; i < n; i++) {
      const uint64_t c = d.count[i];
      s.sq_count[d.rep[i]] += (double)(c * c);
    }
  }
  s.order.resize(r);
  for (unsigned i = 0; i < r; i++) s.order[i] = i;
  std::sort(s.order.begin(), s.order.end(),
            [&](unsigned a, unsigned b) { return strcmp(d.rep_names[a].c_str(), d.rep_names[b].c_str()) < 0; });
  for (unsigned k = 0; k < r; k++) {
    s.sum_size += s.size[k];
    s.sum_count += s.count[k];
  }
  return s;
}

void log_rep_table(const SeqDb& d, const RepStats& s) {
  const size_t r = d.rep_names.size();
  const int w1 = std::max(1, (int)(1 + floor(log10((double)r))));
  const int w2 = std::max(9, (int)(1 + floor(log10((double)s.sum_size))));
  const int w3 = std::max(5, (int)(1 + floor(log10((double)s.sum_count))));
  fprintf(g_log, "Repertoires in set:\n");
  fprintf(g_log, "%*s %*s %*s %s\n", w1, "#", w2, "Sequences", w3, "Count", "Repertoire ID");
  for (unsigned i = 0; i < r; i++) {
    const unsigned k = s.order[i];
    fprintf(g_log, "%*u %*lu %*lu %s\n", w1, i + 1, w2, (unsigned long)s.size[k], w3, (unsigned long)s.count[k],
            d.rep_names[k].c_str());
  }
  fprintf(g_log, "\n");
}

cb_set as_cb_set(const SeqDb& d, uint64_t first, uint64_t n) {
  cb_set s{};
  s.n = n;
  s.residues = d.residues.data();
  s.offsets = d.offsets.data() + first;
  s.v_gene = d.v.data() + first;
  s.j_gene = d.j.data() + first;
  s.rep = d.rep.data() + first;
  s.count = d.count.data() + first;
  s.n_reps = (uint32_t)d.rep_names.size();
  s.index_base = first;
  return s;
}

[[noreturn]] void engine_fatal(cb_ctx* c) {
  const std::string msg = c ? cb_last_error(c) : cb_global_error();
  if (c) cb_destroy(c);
  fatal(msg.c_str());
}

int64_t hamming(const uint8_t* a, const uint8_t* b, int64_t n) {
  int64_t d = 0;
  for (int64_t i = 0; i < n; i++) d += a[i] != b[i];
  return d;
}

struct PairWriter {
  const Options& o;
  const SeqDb &d1, &d2;
  const GeneTables& g;
  FILE* f;

  void header() {
    fprintf(f, "#repertoire_id_1\tsequence_id_1\tduplicate_count_1\tv_call_1\tj_call_1\t%s_1", o.seq_header);
    for (auto& k : o.keep_names) fprintf(f, "\t%s_1", k.c_str());
    fprintf(f, "\trepertoire_id_2\tsequence_id_2\tduplicate_count_2\tv_call_2\tj_call_2\t%s_2", o.seq_header);
    for (auto& k : o.keep_names) fprintf(f, "\t%s_2", k.c_str());
    if (o.distance) fprintf(f, "\tdistance");
    fprintf(f, "\n");
  }
  void side(std::string& buf, const SeqDb& d, uint64_t i) const {
    const char* alpha = o.nucleotides ? "acgt" : "ACDEFGHIKLMNPQRSTVWY";  // db.cc:73-74
    buf += d.rep_names[d.rep[i]];
    buf += '\t';
    if (d.has_ids()) buf += d.seq_id(i);
    buf += '\t';
    append_u64(buf, d.count[i]);
    buf += '\t';
    buf += g.v_names[d.v[i]];
    buf += '\t';
    buf += g.j_names[d.j[i]];
    buf += '\t';
    for (uint64_t p = d.offsets[i]; p < d.offsets[i + 1]; p++) buf += alpha[d.residues[p]];
    if (!o.keep_names.empty()) {
      buf += '\t';
      buf += d.keep(i);
    }
  }
  // rows formatted on the host threads (-t, default: all), written in the order of p[] (row_writer.h)
  void write(const cb_pair* p, size_t n) const {
    write_rows_parallel(f, n, host_threads(o.threads, o.threads_given), [&](uint64_t k0, uint64_t k1, std::string& buf) {
      for (uint64_t k = k0; k < k1; k++) {
        const uint64_t a = p[k].a, b = p[k].b;
        side(buf, d1, a);
        buf += '\t';
        side(buf, d2, b);
        if (o.distance) {  // Hamming if equal length, else 1 (one indel), overlap.cc:492-502
          const int64_t l1 = (int64_t)(d1.offsets[a + 1] - d1.offsets[a]), l2 = (int64_t)(d2.offsets[b + 1] - d2.offsets[b]);
          int64_t dist = 1;
          if (l1 == l2) dist = hamming(d1.residues.data() + d1.offsets[a], d2.residues.data() + d2.offsets[b], l1);
          buf += '\t';
          append_u64(buf, (uint64_t)dist);
        }
        buf += '\n';
      }
    }, 1u << 13);
  }
};

}  // namespace

void overlap_command(const Options& o, FILE* outfile, FILE* pairsfile) {
  GeneTables genes;
  SeqDb d1, d2s;
  // COMPAIRR_B200_TRACE=1: wall-clock marks on stderr (where the time outside the logged phases goes)
  const bool trace = getenv("COMPAIRR_B200_TRACE") != nullptr;
  const auto t_start = std::chrono::steady_clock::now();
  auto mark = [&](const char* what) {
    if (trace)
      fprintf(stderr, "[trace] %8.3f s  %s\n",
              std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count(), what);
  };

  // the CUDA contexts come up (0.6 - 1.4 s) on another thread while the host parses its input: a
  // throw-away engine context per device leaves the device's primary context initialised
  std::thread warm;
  if (!getenv("COMPAIRR_B200_NO_WARM"))
    warm = std::thread([&] {
      const int nd = cb_device_count();
      for (int g = 0; g < o.gpus && o.device + g < nd; g++) {
        cb_config w{};
        w.abi_version = CB_ABI_VERSION;
        w.device = o.device + g;
        w.alphabet_size = o.alphabet_size;
        w.n_reps_a = 1;
        w.no_matrix = 1;
        cb_ctx* c = nullptr;
        if (cb_create(&w, &c) == 0) cb_destroy(c);
      }
    });
  struct Joiner {
    std::thread& t;
    ~Joiner() { if (t.joinable()) t.join(); }
  } warm_guard{warm};
  fprintf(g_log, "Immune receptor repertoire set 1\n\n");
  read_airr_tsv(o.input1, o, o.existence, o.existence || o.pairs, "1", genes, d1);
  fprintf(g_log, "\n");
  mark("set 1 read");
  const RepStats s1 = rep_stats(d1, host_threads(o.threads, o.threads_given));
  log_rep_table(d1, s1);
  mark("set 1 repertoire table");
  if (o.existence && d1.rep_names.size() > 1)
    fatal("Multiple repertoires are not allowed in the first file specified on the command line with the -x or --existence command.");

  fprintf(g_log, "Immune receptor repertoire set 2\n\n");
  const bool two_sets = o.input2 && strcmp(o.input1, o.input2);
  if (two_sets) {
    read_airr_tsv(o.input2, o, false, o.pairs != nullptr, "2", genes, d2s);
    fprintf(g_log, "\n");
  } else {
    fprintf(g_log, "Set 2 is identical to set 1\n\n");
  }
  const SeqDb& d2 = two_sets ? d2s : d1;
  const RepStats s2s = two_sets ? rep_stats(d2s, host_threads(o.threads, o.threads_given)) : RepStats();
  const RepStats& s2 = two_sets ? s2s : s1;
  if (two_sets) {
    if (d2.rep_names.empty()) fatal("Repertoire set missing repertoire_id.");
    log_rep_table(d2, s2);
  } else if (d2.rep_names.empty()) {
    fatal("Repertoire set is missing repertoire_id.");
  }
  mark("set 2 read + table");
  fprintf(g_log, "Unique V genes:    %lu\n", (unsigned long)genes.v_names.size());
  fprintf(g_log, "Unique J genes:    %lu\n", (unsigned long)genes.j_names.size());

  const uint64_t R1 = d1.rep_names.size(), R2 = d2.rep_names.size(), N1 = d1.n();
  const int ngpu = std::min<int>(o.gpus, std::max(1, cb_device_count() - o.device));
  if (ngpu != o.gpus) fprintf(g_log, "GPUs used:         %d (of %d requested)\n", ngpu, o.gpus);

  // ---- engine: one context per GPU joined in an NCCL communicator; every GPU gets all of set B
  // (each uploads 1/ngpu of it, all-gather over NVLink) and a shard of set A ----------------------
  cb_config cfg{};
  cfg.abi_version = CB_ABI_VERSION;
  cfg.alphabet_size = o.alphabet_size;
  cfg.differences = (int32_t)std::min<int64_t>(o.differences, 1 << 20);
  cfg.indels = o.indels;
  cfg.ignore_genes = o.ignore_genes;
  cfg.ignore_counts = o.ignore_counts;
  cfg.score = (int32_t)o.score;
  cfg.mode = o.existence ? CB_MODE_EXISTENCE : CB_MODE_MATRIX;
  cfg.no_matrix = o.no_matrix;
  cfg.want_pairs = o.pairs != nullptr;
  cfg.n_reps_a = (uint32_t)std::max<uint64_t>(R1, 1);
  cfg.seed = 1;

  if (warm.joinable()) warm.join();
  mark("warm-up joined");
  std::vector<cb_ctx*> ctx(ngpu, nullptr);
  for (int g = 0; g < ngpu; g++) {
    cfg.device = o.device + g;
    if (cb_create(&cfg, &ctx[g])) engine_fatal(nullptr);
  }
  if (ngpu > 1 && cb_comm_init_all(ctx.data(), ngpu)) engine_fatal(ctx[0]);
  mark("contexts created");
  // a failure on one GPU ends the command at once: the other ranks may be waiting in a collective
  auto rank_fatal = [&](cb_ctx* c) {
    fprintf(stderr, "\nError: %s\n", cb_last_error(c));
    cli_exit(1);
  };
  auto cols_of = [](const SeqDb& d, uint64_t first, uint64_t n) {
    cb_set_cols s{};
    s.n = n;
    s.residues = d.residues.data();
    s.offsets = {d.offsets.data() + first, 8, 0};
    s.v_gene = {d.v.data() + first, 4, 0};
    s.j_gene = {d.j.data() + first, 4, 0};
    s.rep = {d.rep.data() + first, 4, 0};
    s.count = {d.count.data() + first, 8, 0};
    s.n_reps = (uint32_t)d.rep_names.size();
    s.index_base = first;
    return s;
  };

  progress_begin(o, "Hashing sequences:");  // upload + hash (+ all-gather) + table/filter build + duplicate check
  {
    std::vector<std::thread> th;
    for (int g = 0; g < ngpu; g++)
      th.emplace_back([&, g] {
        uint64_t first = 0, count = d2.n();
        cb_shard_range(d2.n(), g, ngpu, &first, &count);
        cb_set_cols shard = cols_of(d2, first, count);
        shard.index_base = 0;  // of the whole set: hits are reported by their index in set 2
        if (cb_set_b_sharded(ctx[g], &shard, d2.n())) rank_fatal(ctx[g]);
      });
    for (auto& t : th) t.join();
  }
  progress_end(o, "Hashing sequences:");
  mark("set B uploaded + built");
  cb_dset* whole_a_dev = nullptr;  // set 1 on GPU 0, kept for the analysis when GPU 0 takes all of it
  if (o.differences <= 2) {
    if (two_sets) {  // duplicates in set 1 are only checked with two distinct sets (overlap.cc:846-852)
      const cb_set whole_a = as_cb_set(d1, 0, N1);
      uint64_t dup1 = 0;
      if (cb_upload(ctx[0], &whole_a, &whole_a_dev) || cb_count_dups(ctx[0], whole_a_dev, &dup1)) engine_fatal(ctx[0]);
      if (ngpu > 1) {
        cb_free_set(ctx[0], whole_a_dev);
        whole_a_dev = nullptr;
      }
      if (dup1) fprintf(g_log, "Warning: %lu duplicates detected in repertoire set 1\n", (unsigned long)dup1);
    }
    const uint64_t dup2 = cb_dups_b(ctx[0]);
    if (dup2) fprintf(g_log, "Warning: %lu duplicates detected in repertoire set 2\n", (unsigned long)dup2);
  }

  mark("duplicate checks");
  // ---- analysis --------------------------------------------------------------------------------
  std::vector<double> matrix;
  const uint64_t rows = o.existence ? N1 : R1;
  if (!o.no_matrix) matrix.assign(rows * R2, 0.0);
  PairWriter pw{o, d1, d2, genes, pairsfile};
  if (o.pairs) pw.header();

  progress_begin(o, "Analysing:        ");
  // shard boundaries balanced by expected probes (cost ~ L for d=1, L^2 for d=2)
  std::vector<uint64_t> bound(ngpu + 1, 0);
  {
    std::vector<double> pre(N1 + 1, 0.0);
    for (uint64_t i = 0; i < N1; i++) {
      const double L = (double)(d1.offsets[i + 1] - d1.offsets[i]);
      pre[i + 1] = pre[i] + (o.differences >= 2 ? L * L : L) + 1.0;
    }
    for (int g = 1; g < ngpu; g++)
      bound[g] = std::lower_bound(pre.begin(), pre.end(), pre[N1] * g / ngpu) - pre.begin();
    bound[ngpu] = N1;
  }
  const uint64_t chunk = o.pairs || o.existence ? (1u << 20) : N1 + 1;  // bound host memory for pairs / -x rows
  std::mutex pairs_mu;  // pairs of every GPU go to the file chunk by chunk (their order is unspecified, README.md:163)
  {
    std::vector<std::thread> th;
    for (int g = 0; g < ngpu; g++)
      th.emplace_back([&, g] {
        cb_ctx* c = ctx[g];
        const bool self_dev = !two_sets;  // set A is the resident set B
        cb_dset* da = nullptr;
        const bool reuse = whole_a_dev != nullptr;  // one GPU: the dup check left set 1 resident
        if (bound[g + 1] > bound[g]) {
          if (self_dev) {
            da = cb_resident_b(c);
          } else if (reuse) {
            da = whole_a_dev;
          } else {
            const cb_set shard = as_cb_set(d1, bound[g], bound[g + 1] - bound[g]);
            if (cb_upload(c, &shard, &da)) rank_fatal(c);
          }
        }
        const uint64_t base = (self_dev || reuse) ? bound[g] : 0;  // range inside the device set
        std::vector<cb_pair> pair_out;
        for (uint64_t at = bound[g]; at < bound[g + 1]; at += chunk) {
          const uint64_t n = std::min(chunk, bound[g + 1] - at);
          if (cb_run(c, da, base + (at - bound[g]), n)) rank_fatal(c);
          if (o.existence && !o.no_matrix && cb_get_matrix(c, matrix.data() + at * R2, n * R2)) rank_fatal(c);
          if (o.pairs) {
            uint64_t np = 0;
            cb_pairs_pending(c, &np);
            pair_out.resize(np);
            size_t got = 0;
            cb_drain_pairs(c, pair_out.data(), np, &got);
            std::lock_guard<std::mutex> lk(pairs_mu);
            pw.write(pair_out.data(), got);
          }
        }
        if (da && !self_dev && !reuse) cb_free_set(c, da);
        // -m: the sum of the partial matrices (sim_thread's merge, overlap.cc:510-527) is an NCCL
        // all-reduce; every rank joins, also one whose shard was empty
        if (!o.existence && !o.no_matrix && ngpu > 1 && cb_allreduce_matrix(c)) rank_fatal(c);
      });
    for (auto& t : th) t.join();
  }
  if (!o.existence && !o.no_matrix && cb_get_matrix(ctx[0], matrix.data(), matrix.size())) engine_fatal(ctx[0]);
  progress_end(o, "Analysing:        ");
  mark("analysis done");
  // The engines are not torn down: main() leaves with _exit() once the files are on disk, and
  // destroying contexts with multi-GB pools took 0.5 s (measured) for nothing the operating system
  // does not reclaim anyway.
  if (getenv("COMPAIRR_B200_TEARDOWN")) {
    for (int g = 0; g < ngpu; g++) {
      if (g == 0 && whole_a_dev) cb_free_set(ctx[0], whole_a_dev);
      cb_destroy(ctx[g]);
    }
    mark("contexts destroyed");
  }

  // ---- results (overlap.cc:540-577, 944-1039) ---------------------------------------------------
  auto value = [&](uint64_t s, unsigned t) -> double {
    const double x = matrix[R2 * s + t];
    if (o.score == SCORE_MH) {
      const double lx = s1.sq_count[s] / s1.count[s] / s1.count[s];
      const double ly = s2.sq_count[t] / s2.count[t] / s2.count[t];
      const double xy = 1.0 * s1.count[s] * s2.count[t];
      return (2.0 * x) / ((lx + ly) * xy);
    }
    if (o.score == SCORE_JACCARD) {
      const double sa = (double)s1.count[s], sb = (double)s2.count[t];
      return x / (sa + sb - x);
    }
    return x;
  };
  if (!o.no_matrix) {
    progress_begin(o, "Writing results:  ");
    const uint64_t nrows = o.existence ? N1 : R1;
    auto row_index = [&](uint64_t i) -> uint64_t { return o.existence ? i : s1.order[i]; };
    auto row_name = [&](uint64_t i) -> const char* {
      return o.existence ? d1.seq_id(i) : d1.rep_names[s1.order[i]].c_str();
    };
    if (o.alternative) {
      fprintf(outfile, o.existence ? "#sequence_id_1\trepertoire_id_2\tmatches\n" : "#repertoire_id_1\trepertoire_id_2\tmatches\n");
      for (uint64_t i = 0; i < nrows; i++)
        for (unsigned j = 0; j < R2; j++)
          fprintf(outfile, "%s\t%s\t%.10lg\n", row_name(i), d2.rep_names[s2.order[j]].c_str(),
                  value(row_index(i), s2.order[j]));
    } else {
      fprintf(outfile, "#");
      for (unsigned j = 0; j < R2; j++) fprintf(outfile, "\t%s", d2.rep_names[s2.order[j]].c_str());
      fprintf(outfile, "\n");
      for (uint64_t i = 0; i < nrows; i++) {
        fprintf(outfile, "%s", row_name(i));
        for (unsigned j = 0; j < R2; j++) fprintf(outfile, "\t%.10lg", value(row_index(i), s2.order[j]));
        fprintf(outfile, "\n");
      }
    }
    progress_end(o, "Writing results:  ");
  }
  fprintf(g_log, "\n");
  mark("results written");
}
