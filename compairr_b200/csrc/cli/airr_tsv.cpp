// airr_tsv.cpp — see airr_tsv.h.
#include "airr_tsv.h"

#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <string_view>
#include <thread>

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

// The whole input in memory: a private mapping for regular files, a buffer for pipes.
struct FileData {
  const char* p = nullptr;
  size_t n = 0;
  void* map = nullptr;
  std::string owned;
  ~FileData() {
    if (map) munmap(map, n);
  }
};

using sv = std::string_view;

// split [b, e) on tabs into at most f.size() fields; returns the number of fields on the line
// (counting stops once every wanted column has been seen)
inline int split_tabs(const char* b, const char* e, std::vector<sv>& f) {
  int nf = 0;
  const int want = (int)f.size();
  while (nf < want) {
    const char* t = (const char*)memchr(b, '\t', (size_t)(e - b));
    if (!t) {
      f[nf++] = sv(b, (size_t)(e - b));
      return nf;
    }
    f[nf++] = sv(b, (size_t)(t - b));
    b = t + 1;
  }
  return nf;
}

void parse_header(const std::string& line, const Options& o, bool require_sequence_id, Columns& c) {
  std::vector<std::string> f;
  {
    size_t b = 0;
    for (;;) {
      const size_t t = line.find('\t', b);
      f.emplace_back(line.substr(b, t == std::string::npos ? std::string::npos : t - b));
      if (t == std::string::npos) break;
      b = t + 1;
    }
  }
  c.keep.assign(o.keep_names.size(), 0);
  for (size_t i = 0; i < f.size(); i++) {
    const int col = (int)i + 1;
    const char* t = f[i].c_str();
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
    cli_exit(1);
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

// First-seen-order string interning with a one-entry cache (consecutive lines mostly repeat the
// repertoire and often the genes).  Views point into the file data or at static strings.
struct Interner {
  std::unordered_map<sv, uint32_t> map;
  std::vector<sv> names;
  sv last;
  uint32_t last_no = 0;
  bool have_last = false;
  uint32_t get(sv s) {
    if (have_last && s == last) return last_no;
    auto it = map.find(s);
    uint32_t no;
    if (it != map.end()) {
      no = it->second;
    } else {
      no = (uint32_t)names.size();
      names.push_back(s);
      map.emplace(s, no);
    }
    last = s;
    last_no = no;
    have_last = true;
    return no;
  }
};

enum ErrKind {
  ERR_NONE = 0, ERR_CHAR_PRINTABLE, ERR_CHAR_OTHER, ERR_EMPTY_SEQ, ERR_NO_SEQID, ERR_BAD_COUNT, ERR_NO_COUNT,
  ERR_NO_V, ERR_NO_J, ERR_NO_SEQ
};

// What one reader thread produced from its range of lines: the same columns as SeqDb with
// thread-local id numbering, merged in file order afterwards.
struct Part {
  RawVec<uint8_t> residues;
  RawVec<uint32_t> len, v, j, rep;
  RawVec<uint64_t> count;
  RawVec<char> id_arena, keep_arena;
  RawVec<uint64_t> id_off, keep_off;
  Interner reps, vs, js;
  unsigned longest = 0, shortest = ~0u;
  uint64_t total_count = 0, ignored_unknown = 0, ignored_empty = 0, lines = 0;
  ErrKind err = ERR_NONE;
  uint64_t err_line = 0;  // 1-based inside this part
  int err_char = 0;
  std::string err_text;
};

struct ParseCtx {
  const Options* o;
  const Columns* col;
  bool require_sequence_id, want_ids;
  sv default_rep;
  int seqcol, maxcol;
};

void parse_range(const char* b, const char* e, const ParseCtx& cx, Part& out) {
  const Options& o = *cx.o;
  const Columns& col = *cx.col;
  std::vector<sv> f((size_t)cx.maxcol);
  const bool want_keep = !o.keep_names.empty();
  // capacity guesses from the byte count keep reallocation out of the loop
  const size_t guess = (size_t)(e - b) / 48 + 16;
  out.len.reserve(guess);
  out.v.reserve(guess);
  out.j.reserve(guess);
  out.rep.reserve(guess);
  out.count.reserve(guess);
  out.residues.reserve(guess * 16);
  auto fail = [&](ErrKind k, int ch = 0, sv text = sv()) {
    out.err = k;
    out.err_line = out.lines;
    out.err_char = ch;
    out.err_text.assign(text);
  };
  while (b < e) {
    const char* nl = (const char*)memchr(b, '\n', (size_t)(e - b));
    const char* le = nl ? nl : e;
    const char* next = nl ? nl + 1 : e;
    if (le > b && le[-1] == '\r') le--;
    out.lines++;
    const int nf = split_tabs(b, le, f);
    b = next;
    auto has = [&](int c) { return c >= 1 && c <= nf; };
    // residues first, exactly in the reference's order of checks (db.cc:400-503)
    const bool has_seq = has(cx.seqcol);
    const sv seq = has_seq ? f[cx.seqcol - 1] : sv();
    const size_t base = out.residues.size();
    bool ignore = false;
    {  // the common case first: every symbol legal, translated in one branch-free sweep
      out.residues.resize(base + seq.size());
      uint8_t* const dst = out.residues.data() + base;
      signed char bad = 0;
      for (size_t i = 0; i < seq.size(); i++) {
        const signed char m = g_map[(unsigned char)seq[i]];
        dst[i] = (uint8_t)m;
        bad |= m;
      }
      if (bad < 0) {  // an illegal symbol somewhere: redo the sequence symbol by symbol (db.cc:400-440)
        out.residues.resize(base);
        for (size_t i = 0; i < seq.size(); i++) {
          const unsigned char ch = (unsigned char)seq[i];
          const signed char m = g_map[ch];
          if (m >= 0) {
            out.residues.push_back((uint8_t)m);
          } else if (ch >= 32 && ch <= 126) {
            if (o.ignore_unknown) {
              ignore = true;
              out.ignored_unknown++;
            } else {
              return fail(ERR_CHAR_PRINTABLE, ch);
            }
          } else {
            return fail(ERR_CHAR_OTHER, ch);
          }
        }
      }
    }
    const unsigned seqlen = (unsigned)(out.residues.size() - base);
    if (seqlen == 0) {
      if (o.ignore_empty) {
        ignore = true;
        out.ignored_empty++;
      } else {
        return fail(ERR_EMPTY_SEQ);
      }
    }
    if (ignore) {
      out.residues.resize(base);
      continue;
    }
    if (seqlen > out.longest) out.longest = seqlen;
    if (seqlen < out.shortest) out.shortest = seqlen;
    // repertoire id (db.cc:505-520)
    const uint32_t rno = out.reps.get(has(col.repertoire_id) ? f[col.repertoire_id - 1] : cx.default_rep);
    // sequence id (db.cc:523-540)
    const bool has_sid = has(col.sequence_id);
    const sv sid = has_sid ? f[col.sequence_id - 1] : sv();
    if (sid.empty() && cx.require_sequence_id) return fail(ERR_NO_SEQID);
    // duplicate count (db.cc:543-572): strtol over the whole field, value >= 1
    uint64_t cnt = 1;
    const sv dc = has(col.duplicate_count) ? f[col.duplicate_count - 1] : sv();
    if (!dc.empty()) {
      bool ok = false;
      if (dc.size() <= 18 && dc[0] >= '0' && dc[0] <= '9') {  // plain digits: the common case
        uint64_t v = 0;
        ok = true;
        for (char ch : dc) {
          if (ch < '0' || ch > '9') {
            ok = false;
            break;
          }
          v = v * 10 + (uint64_t)(ch - '0');
        }
        if (ok && v >= 1) cnt = v; else ok = false;
      }
      if (!ok) {  // anything else goes through strtol itself (signs, blanks, overflow)
        const std::string z(dc);
        if (strlen(z.c_str()) == z.size()) {
          char* end = nullptr;
          const long v = strtol(z.c_str(), &end, 10);
          if (end && *end == 0 && v >= 1) {
            cnt = (uint64_t)v;
            ok = true;
          }
        }
      }
      if (!ok) return fail(ERR_BAD_COUNT, 0, dc);
    } else if (!o.ignore_counts) {
      return fail(ERR_NO_COUNT);
    }
    out.total_count += cnt;
    // genes (db.cc:577-631)
    const sv vc = has(col.v_call) ? f[col.v_call - 1] : sv();
    const sv jc = has(col.j_call) ? f[col.j_call - 1] : sv();
    if (!o.ignore_genes && vc.empty()) return fail(ERR_NO_V);
    const uint32_t vno = out.vs.get(vc);
    if (!o.ignore_genes && jc.empty()) return fail(ERR_NO_J);
    const uint32_t jno = out.js.get(jc);
    if (seq.empty()) return fail(ERR_NO_SEQ);
    out.len.push_back(seqlen);
    out.v.push_back(vno);
    out.j.push_back(jno);
    out.rep.push_back(rno);
    out.count.push_back(cnt);
    if (cx.want_ids) {
      out.id_off.push_back(out.id_arena.size());
      out.id_arena.insert(out.id_arena.end(), sid.begin(), sid.end());
      out.id_arena.push_back(0);
    }
    if (want_keep) {
      out.keep_off.push_back(out.keep_arena.size());
      for (size_t x = 0; x < col.keep.size(); x++) {
        if (x) out.keep_arena.push_back('\t');
        if (has(col.keep[x])) {
          const sv kv = f[col.keep[x] - 1];
          out.keep_arena.insert(out.keep_arena.end(), kv.begin(), kv.end());
        }
      }
      out.keep_arena.push_back(0);
    }
  }
}

[[noreturn]] void report(const Part& p, uint64_t lineno, const Options& o) {
  const unsigned long ln = (unsigned long)lineno;
  switch (p.err) {
    case ERR_CHAR_PRINTABLE:
      fprintf(g_log, "\n\nError: Illegal character '%c' in sequence on line %lu. Use -u to ignore.\n", p.err_char, ln);
      break;
    case ERR_CHAR_OTHER:
      fprintf(g_log, "\n\nError: Illegal character (ascii no %d) in sequence on line %lu\n", p.err_char, ln);
      break;
    case ERR_EMPTY_SEQ:
      fprintf(g_log, "\n\nError: Empty sequence in sequence on line %lu. Use -e to ignore.\n", ln);
      break;
    case ERR_NO_SEQID:
      fprintf(g_log, "\n\nError: missing or empty sequence_id value on line %lu\n", ln);
      break;
    case ERR_BAD_COUNT:
      fprintf(g_log, "\n\nError: Illegal duplicate_count on line %lu: %s\n", ln, p.err_text.c_str());
      break;
    case ERR_NO_COUNT:
      fprintf(g_log, "\n\nError: missing or empty duplicate_count on line %lu\n", ln);
      break;
    case ERR_NO_V:
      fprintf(g_log, "\n\nError: missing or empty v_call value on line %lu\n", ln);
      break;
    case ERR_NO_J:
      fprintf(g_log, "\n\nError: missing or empty j_call value on line %lu\n", ln);
      break;
    default:
      fprintf(g_log, "\n\nError: missing or empty %s value on line %lu\n", o.seq_header, ln);
      break;
  }
  cli_exit(1);
}

// run fn(t) for t in [0, n) on n threads (inline when n == 1)
template <class F>
void parallel(unsigned n, F fn) {
  if (n <= 1) {
    fn(0u);
    return;
  }
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < n; t++) pool.emplace_back(fn, t);
  for (auto& th : pool) th.join();
}

}  // namespace

// The reference reads with getline/strsep/std::map on one thread (db.cc:708-901), ~0.4 us per
// line: 40 s for a 10^8-line file, far more than the whole GPU phase.  Here the
// file is mapped, cut into one range of whole lines per host thread, parsed into thread-local
// columns with thread-local id numbering, and merged in file order — so sequence order (the -x
// row order and the index space of pairs) and the first-seen numbering of repertoires and genes
// are exactly the serial reader's, and the earliest malformed line is the one reported.
void read_airr_tsv(const char* filename, const Options& o, bool require_sequence_id, bool want_ids,
                   const char* default_repertoire_id, GeneTables& genes, SeqDb& db) {
  int fd = -1;
  if (!strcmp(filename, "-"))
    fd = dup(STDIN_FILENO);
  else
    fd = open(filename, O_RDONLY);
  if (fd < 0) {
    fprintf(g_log, "\nError: Unable to open input data file (%s).\n", filename);
    cli_exit(1);
  }
  struct stat fs;
  if (fstat(fd, &fs)) {
    fprintf(g_log, "\nUnable to fstat on input file (%s)\n", filename);
    cli_exit(1);
  }
  if (!S_ISREG(fs.st_mode)) fprintf(g_log, "Waiting for data from standard input...\n");
  init_map(o.nucleotides);

  progress_begin(o, "Reading sequences:");
  auto data = std::make_shared<FileData>();
  if (S_ISREG(fs.st_mode) && fs.st_size > 0) {
    void* m = mmap(nullptr, (size_t)fs.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m != MAP_FAILED) {
      data->map = m;
      data->p = (const char*)m;
      data->n = (size_t)fs.st_size;
      madvise(m, data->n, MADV_WILLNEED);
    }
  }
  if (!data->map) {  // pipes, empty files, or a mapping that failed
    char buf[1 << 16];
    for (;;) {
      const ssize_t r = read(fd, buf, sizeof buf);
      if (r < 0) fatal("Unable to read from the input file");
      if (r == 0) break;
      data->owned.append(buf, (size_t)r);
    }
    data->p = data->owned.data();
    data->n = data->owned.size();
  }
  close(fd);
  if (data->n == 0) fatal("Unable to read from the input file");

  // leading comment lines are skipped, the first other line is the header (db.cc:766-782)
  const char* p = data->p;
  const char* const end = p + data->n;
  uint64_t lineno = 0;
  Columns col;
  bool have_header = false;
  while (p < end && !have_header) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
    const char* le = nl ? nl : end;
    lineno++;
    if (*p != '#' && *p != '@') {
      const char* he = (le > p && le[-1] == '\r') ? le - 1 : le;
      std::string header(p, (size_t)(he - p));
      header.resize(strlen(header.c_str()));
      parse_header(header, o, require_sequence_id, col);
      have_header = true;
    }
    p = nl ? nl + 1 : end;
  }

  ParseCtx cx;
  cx.o = &o;
  cx.col = &col;
  cx.require_sequence_id = require_sequence_id;
  cx.want_ids = want_ids;
  cx.default_rep = sv(default_repertoire_id);
  cx.seqcol = o.cdr3 ? (o.nucleotides ? col.cdr3 : col.cdr3_aa) : (o.nucleotides ? col.junction : col.junction_aa);
  cx.maxcol = std::max({col.repertoire_id, col.sequence_id, col.duplicate_count, col.v_call, col.j_call, cx.seqcol, 1});
  for (int k : col.keep) cx.maxcol = std::max(cx.maxcol, k);

  // one range of whole lines per thread
  const size_t body = (size_t)(end - p);
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 1;
  unsigned n_thr = (o.threads_given || o.threads > 1) ? (unsigned)o.threads : std::min(hw, 32u);
  size_t min_bytes = 1 << 20;  // per thread; COMPAIRR_B200_READ_MIN_BYTES lets tests split small files
  if (const char* e = getenv("COMPAIRR_B200_READ_MIN_BYTES")) min_bytes = std::max<size_t>(1, strtoull(e, nullptr, 10));
  n_thr = (unsigned)std::max<size_t>(1, std::min<size_t>(n_thr, body / min_bytes));
  std::vector<const char*> cut(n_thr + 1, end);
  cut[0] = p;
  for (unsigned t = 1; t < n_thr; t++) {
    const char* c = p + body / n_thr * t;
    if (c < cut[t - 1]) c = cut[t - 1];
    const char* nl = (const char*)memchr(c, '\n', (size_t)(end - c));
    cut[t] = nl ? nl + 1 : end;
  }
  const bool trace = getenv("COMPAIRR_B200_TRACE") != nullptr;
  const auto t_parse0 = std::chrono::steady_clock::now();
  auto since = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_parse0).count(); };
  std::vector<Part> parts(n_thr);
  parallel(n_thr, [&](unsigned t) { parse_range(cut[t], cut[t + 1], cx, parts[t]); });
  if (trace) fprintf(stderr, "[trace]   reader: %u threads parsed in %.3f s\n", n_thr, since());
  for (unsigned t = 0; t < n_thr; t++) {  // the earliest malformed line, numbered from the top of the file
    if (parts[t].err != ERR_NONE) report(parts[t], lineno + parts[t].err_line, o);
    lineno += parts[t].lines;
  }

  // merge: global ids in file order of first appearance, then every thread copies its columns
  // into place
  auto intern = [](std::unordered_map<std::string, uint32_t>& map, std::vector<std::string>& names, sv s) {
    std::string key(s);
    auto it = map.find(key);
    if (it != map.end()) return it->second;
    const uint32_t no = (uint32_t)names.size();
    names.push_back(key);
    map.emplace(std::move(key), no);
    return no;
  };
  std::vector<std::vector<uint32_t>> rmap(n_thr), vmap(n_thr), jmap(n_thr);
  std::vector<uint64_t> seq0(n_thr + 1, db.n()), res0(n_thr + 1, db.residues.size());
  std::vector<uint64_t> id0(n_thr + 1, db.id_arena.size()), keep0(n_thr + 1, db.keep_arena.size());
  for (unsigned t = 0; t < n_thr; t++) {
    Part& pt = parts[t];
    for (sv s : pt.reps.names) rmap[t].push_back(intern(db.rep_map, db.rep_names, s));
    for (sv s : pt.vs.names) vmap[t].push_back(intern(genes.v_map, genes.v_names, s));
    for (sv s : pt.js.names) jmap[t].push_back(intern(genes.j_map, genes.j_names, s));
    seq0[t + 1] = seq0[t] + pt.len.size();
    res0[t + 1] = res0[t] + pt.residues.size();
    id0[t + 1] = id0[t] + pt.id_arena.size();
    keep0[t + 1] = keep0[t] + pt.keep_arena.size();
    if (!pt.len.empty()) {
      db.longest = std::max(db.longest, pt.longest);
      db.shortest = std::min(db.shortest, pt.shortest);
    }
    db.total_count += pt.total_count;
    db.ignored_unknown += pt.ignored_unknown;
    db.ignored_empty += pt.ignored_empty;
  }
  const uint64_t n_total = seq0[n_thr];
  db.residues.resize(res0[n_thr]);
  db.offsets.resize(n_total + 1);
  db.v.resize(n_total);
  db.j.resize(n_total);
  db.rep.resize(n_total);
  db.count.resize(n_total);
  if (want_ids) {
    db.id_arena.resize(id0[n_thr]);
    db.id_off.resize(n_total);
  }
  if (!o.keep_names.empty()) {
    db.keep_arena.resize(keep0[n_thr]);
    db.keep_off.resize(n_total);
  }
  parallel(n_thr, [&](unsigned t) {
    Part& pt = parts[t];
    const uint64_t s0 = seq0[t], cnt = pt.len.size();
    if (!pt.residues.empty()) memcpy(db.residues.data() + res0[t], pt.residues.data(), pt.residues.size());
    uint64_t off = res0[t];
    for (uint64_t i = 0; i < cnt; i++) {
      off += pt.len[i];
      db.offsets[s0 + i + 1] = off;
      db.v[s0 + i] = vmap[t][pt.v[i]];
      db.j[s0 + i] = jmap[t][pt.j[i]];
      db.rep[s0 + i] = rmap[t][pt.rep[i]];
      db.count[s0 + i] = pt.count[i];
    }
    if (want_ids && cnt) {
      memcpy(db.id_arena.data() + id0[t], pt.id_arena.data(), pt.id_arena.size());
      for (uint64_t i = 0; i < cnt; i++) db.id_off[s0 + i] = id0[t] + pt.id_off[i];
    }
    if (!o.keep_names.empty() && cnt) {
      memcpy(db.keep_arena.data() + keep0[t], pt.keep_arena.data(), pt.keep_arena.size());
      for (uint64_t i = 0; i < cnt; i++) db.keep_off[s0 + i] = keep0[t] + pt.keep_off[i];
    }
    Part().residues.swap(pt.residues);  // give the memory back early
  });
  parts.clear();
  if (trace) fprintf(stderr, "[trace]   reader: merged at %.3f s\n", since());
  progress_end(o, "Reading sequences:");

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
