// airr_tsv.cpp — see airr_tsv.h.
#include "airr_tsv.h"

#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>

static std::chrono::steady_clock::time_point g_t0;

void progress_begin(const Options& o, const char* prompt) {
  if (o.log)
    fprintf(g_log, "%s", prompt);
  else
    fprintf(g_log, "%s %.0f%%", prompt, 0.0);
  fflush(g_log);
  g_t0 = std::chrono::steady_clock::now();
}

void progress_end(const Options& o, const char* prompt) {
  const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - g_t0).count();
  if (o.log)
    fprintf(g_log, " %.0f%% (%.9lfs)\n", 100.0, dt);
  else
    fprintf(g_log, "  \r%s %.0f%% (%.9lfs)\n", prompt, 100.0, dt);
  fflush(g_log);
}

namespace {

struct Columns {
  int repertoire_id = 0, sequence_id = 0, duplicate_count = 0, v_call = 0, j_call = 0;
  int junction = 0, junction_aa = 0, cdr3 = 0, cdr3_aa = 0;
  std::vector<int> keep;
};

signed char g_map[256];

void init_map(bool nt) {
  memset(g_map, -1, sizeof g_map);
  const char* alpha = nt ? "ACGT" : "ACDEFGHIKLMNPQRSTVWY";
  for (int i = 0; alpha[i]; i++) {
    g_map[(unsigned char)alpha[i]] = (signed char)i;
    g_map[(unsigned char)(alpha[i] | 0x20)] = (signed char)i;
  }
  if (nt) g_map['U'] = g_map['u'] = 3;  // db.cc:53-71
}

// split a line in place on tabs; returns field count
size_t split_tabs(char* line, std::vector<char*>& out) {
  out.clear();
  char* p = line;
  out.push_back(p);
  for (; *p; p++)
    if (*p == '\t') {
      *p = 0;
      out.push_back(p + 1);
    }
  return out.size();
}

void parse_header(char* line, const Options& o, bool require_sequence_id, Columns& c) {
  std::vector<char*> f;
  split_tabs(line, f);
  c.keep.assign(o.keep_names.size(), 0);
  for (size_t i = 0; i < f.size(); i++) {
    const int col = (int)i + 1;
    const char* t = f[i];
    if (!strcmp(t, "repertoire_id")) c.repertoire_id = col;
    else if (!strcmp(t, "sequence_id")) c.sequence_id = col;
    else if (!strcmp(t, "duplicate_count")) c.duplicate_count = col;
    else if (!strcmp(t, "v_call")) c.v_call = col;
    else if (!strcmp(t, "j_call")) c.j_call = col;
    else if (!strcmp(t, "junction")) c.junction = col;
    else if (!strcmp(t, "junction_aa")) c.junction_aa = col;
    else if (!strcmp(t, "cdr3")) c.cdr3 = col;
    else if (!strcmp(t, "cdr3_aa")) c.cdr3_aa = col;
    for (size_t k = 0; k < o.keep_names.size(); k++)
      if (o.keep_names[k] == t) c.keep[k] = col;
  }
  const int seqcol = o.cdr3 ? (o.nucleotides ? c.cdr3 : c.cdr3_aa) : (o.nucleotides ? c.junction : c.junction_aa);
  const bool missing = (require_sequence_id && !c.sequence_id) || (!o.ignore_counts && !c.duplicate_count) ||
                       (!o.ignore_genes && (!c.v_call || !c.j_call)) || !seqcol;
  if (missing) {  // db.cc:233-279
    fprintf(g_log, "\nMissing essential column(s) in header of AIRR TSV input file:");
    if (require_sequence_id && !c.sequence_id) fprintf(g_log, " sequence_id");
    if (!o.ignore_counts && !c.duplicate_count) fprintf(g_log, " duplicate_count");
    if (!o.ignore_genes) {
      if (!c.v_call) fprintf(g_log, " v_call");
      if (!c.j_call) fprintf(g_log, " j_call");
    }
    if (!seqcol) fprintf(g_log, " %s", o.seq_header);
    fprintf(g_log, "\n");
    exit(1);
  }
  bool any = false;
  for (int k : c.keep) any |= k < 1;
  if (any) {
    fprintf(g_log, "\nWarning: missing column(s) to keep in header:");
    for (size_t k = 0; k < c.keep.size(); k++)
      if (c.keep[k] < 1) fprintf(g_log, " %s", o.keep_names[k].c_str());
    fprintf(g_log, "\n");
  }
}

}  // namespace

void read_airr_tsv(const char* filename, const Options& o, bool require_sequence_id, bool want_ids,
                   const char* default_repertoire_id, GeneTables& genes, SeqDb& db) {
  FILE* fp = nullptr;
  if (!strcmp(filename, "-")) {
    const int fd = dup(STDIN_FILENO);
    fp = fd < 0 ? nullptr : fdopen(fd, "rb");
  } else {
    fp = fopen(filename, "rb");
  }
  if (!fp) {
    fprintf(g_log, "\nError: Unable to open input data file (%s).\n", filename);
    exit(1);
  }
  struct stat fs;
  if (fstat(fileno(fp), &fs)) {
    fprintf(g_log, "\nUnable to fstat on input file (%s)\n", filename);
    exit(1);
  }
  if (!S_ISREG(fs.st_mode)) fprintf(g_log, "Waiting for data from standard input...\n");
  init_map(o.nucleotides);

  progress_begin(o, "Reading sequences:");
  char* line = nullptr;
  size_t cap = 0;
  ssize_t len = getline(&line, &cap, fp);
  if (len < 0) fatal("Unable to read from the input file");
  uint64_t lineno = 0;
  bool in_header = true;
  Columns col;
  std::vector<char*> f;
  const char* const empty = "";
  while (len >= 0) {
    if (len > 0 && line[len - 1] == '\n') line[--len] = 0;
    if (len > 0 && line[len - 1] == '\r') line[--len] = 0;
    lineno++;
    if (in_header) {
      if (line[0] != '#' && line[0] != '@') {  // leading comment lines are skipped (db.cc:766-782)
        parse_header(line, o, require_sequence_id, col);
        in_header = false;
      }
    } else {
      const int nf = (int)split_tabs(line, f);
      auto field = [&](int c) -> const char* { return (c >= 1 && c <= nf) ? f[c - 1] : nullptr; };
      const char* seq = field(o.cdr3 ? (o.nucleotides ? col.cdr3 : col.cdr3_aa)
                                     : (o.nucleotides ? col.junction : col.junction_aa));
      // residues first, exactly in the reference's order of checks (db.cc:400-503)
      const size_t slen = seq ? strlen(seq) : 0;
      const size_t base = db.residues.size();
      bool ignore = false;
      for (size_t i = 0; i < slen; i++) {
        const unsigned char ch = (unsigned char)seq[i];
        const signed char m = g_map[ch];
        if (m >= 0) {
          db.residues.push_back((uint8_t)m);
        } else if (ch >= 32 && ch <= 126) {
          if (o.ignore_unknown) {
            ignore = true;
            db.ignored_unknown++;
          } else {
            fprintf(g_log, "\n\nError: Illegal character '%c' in sequence on line %lu. Use -u to ignore.\n", ch,
                    (unsigned long)lineno);
            exit(1);
          }
        } else {
          fprintf(g_log, "\n\nError: Illegal character (ascii no %d) in sequence on line %lu\n", ch,
                  (unsigned long)lineno);
          exit(1);
        }
      }
      const unsigned seqlen = (unsigned)(db.residues.size() - base);
      if (seqlen == 0) {
        if (o.ignore_empty) {
          ignore = true;
          db.ignored_empty++;
        } else {
          fprintf(g_log, "\n\nError: Empty sequence in sequence on line %lu. Use -e to ignore.\n",
                  (unsigned long)lineno);
          exit(1);
        }
      }
      if (ignore) {
        db.residues.resize(base);
      } else {
        if (seqlen > db.longest) db.longest = seqlen;
        if (seqlen < db.shortest) db.shortest = seqlen;
        // repertoire id (db.cc:505-520)
        const char* rid = field(col.repertoire_id);
        if (!rid) rid = default_repertoire_id;
        auto rit = db.rep_map.find(rid);
        uint32_t rno;
        if (rit == db.rep_map.end()) {
          rno = (uint32_t)db.rep_names.size();
          db.rep_names.emplace_back(rid);
          db.rep_map.emplace(rid, rno);
        } else {
          rno = rit->second;
        }
        // sequence id (db.cc:523-540)
        const char* sid = field(col.sequence_id);
        if (!(sid && *sid) && require_sequence_id) {
          fprintf(g_log, "\n\nError: missing or empty sequence_id value on line %lu\n", (unsigned long)lineno);
          exit(1);
        }
        // duplicate count (db.cc:543-572)
        const char* dc = field(col.duplicate_count);
        uint64_t cnt = 1;
        if (dc && *dc) {
          char* end = nullptr;
          const long v = strtol(dc, &end, 10);
          if (end && *end == 0 && v >= 1) {
            cnt = (uint64_t)v;
          } else {
            fprintf(g_log, "\n\nError: Illegal duplicate_count on line %lu: %s\n", (unsigned long)lineno, dc);
            exit(1);
          }
        } else if (!o.ignore_counts) {
          fprintf(g_log, "\n\nError: missing or empty duplicate_count on line %lu\n", (unsigned long)lineno);
          exit(1);
        }
        db.total_count += cnt;
        // genes (db.cc:577-631)
        const char* vc = field(col.v_call);
        const char* jc = field(col.j_call);
        if (!o.ignore_genes && !(vc && *vc)) {
          fprintf(g_log, "\n\nError: missing or empty v_call value on line %lu\n", (unsigned long)lineno);
          exit(1);
        }
        if (!vc) vc = empty;
        auto intern = [](std::unordered_map<std::string, uint32_t>& map, std::vector<std::string>& names,
                         const char* s) {
          auto it = map.find(s);
          if (it != map.end()) return it->second;
          const uint32_t no = (uint32_t)names.size();
          names.emplace_back(s);
          map.emplace(s, no);
          return no;
        };
        const uint32_t vno = intern(genes.v_map, genes.v_names, vc);
        if (!o.ignore_genes && !(jc && *jc)) {
          fprintf(g_log, "\n\nError: missing or empty j_call value on line %lu\n", (unsigned long)lineno);
          exit(1);
        }
        if (!jc) jc = empty;
        const uint32_t jno = intern(genes.j_map, genes.j_names, jc);
        if (!(seq && *seq)) {
          fprintf(g_log, "\n\nError: missing or empty %s value on line %lu\n", o.seq_header, (unsigned long)lineno);
          exit(1);
        }
        db.offsets.push_back(db.residues.size());
        db.v.push_back(vno);
        db.j.push_back(jno);
        db.rep.push_back(rno);
        db.count.push_back(cnt);
        if (want_ids) db.seq_id.emplace_back(sid ? sid : empty);
        if (!o.keep_names.empty()) {
          std::string k;
          for (size_t x = 0; x < col.keep.size(); x++) {
            if (x) k.push_back('\t');
            const char* kv = field(col.keep[x]);
            if (kv) k += kv;
          }
          db.keep.push_back(std::move(k));
        }
      }
    }
    len = getline(&line, &cap, fp);
  }
  progress_end(o, "Reading sequences:");
  free(line);
  fclose(fp);

  if (db.ignored_unknown) fprintf(g_log, "%lu sequences with unknown symbols ignored.\n", (unsigned long)db.ignored_unknown);
  if (db.ignored_empty) fprintf(g_log, "%lu empty sequences ignored.\n", (unsigned long)db.ignored_empty);
  fprintf(g_log, "Repertoires:       %lu\nSequences:         %lu\nResidues:          %lu\n",
          (unsigned long)db.rep_names.size(), (unsigned long)db.n(), (unsigned long)db.residues.size());
  if (db.n() > 0)
    fprintf(g_log, "Shortest:          %u\nLongest:           %u\nAverage length:    %.1lf\n", db.shortest, db.longest,
            1.0 * db.residues.size() / db.n());
  else
    fprintf(g_log, "Shortest:          -\nLongest:           -\nAverage length:    -\n");
  fprintf(g_log, "Total dupl. count: %lu\n", (unsigned long)db.total_count);
}
