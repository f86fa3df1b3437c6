// options.cpp — see options.h.  Written against the behaviour of src/compairr.cc:292-706.
#include "options.h"

#include <unistd.h>

#include <getopt.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <time.h>

FILE* g_log = stderr;

static const char* const kScoreNames[SCORE_END] = {"Product", "Ratio", "Min", "Max", "Mean", "MH", "Jaccard"};
static const char* const kScoreDescr[SCORE_END] = {
    "Sum of products of counts", "Sum of ratios of counts", "Sum of minimum of counts",
    "Sum of maximum of counts",  "Sum of mean of counts",   "Morisita-Horn index",
    "Jaccard index"};

void cli_exit(int code) {
  fflush(nullptr);
  _exit(code);
}

[[noreturn]] void fatal(const char* msg) {
  fprintf(stderr, "\nError: %s\n", msg);
  cli_exit(1);
}

void show_header() {
  fprintf(g_log, "CompAIRR-B200 1.13.0 - Comparison of Adaptive Immune Receptor Repertoires (GPU overlap engine)\n");
  fprintf(g_log, "https://github.com/uio-bmi/compairr\n\n");
}

void show_usage() {
  static const char* const lines[] = {
      "Usage: compairr [OPTIONS] TSVFILE1 [TSVFILE2]", "", "Commands:",
      " -h, --help                  display this help and exit",
      " -v, --version               display version information",
      " -m, --matrix                compute overlap matrix between two sets",
      " -x, --existence             check existence of sequences in repertoires",
      " -c, --cluster               cluster sequences in one repertoire",
      " -z, --deduplicate           deduplicate sequences in repertoires", "",
      "General options:",
      " -d, --differences INTEGER   number of differences accepted (0*)",
      " -i, --indels                allow insertions or deletions when d=1",
      " -f, --ignore-counts         ignore duplicate_count information",
      " -g, --ignore-genes          ignore V and J gene information",
      " -n, --nucleotides           compare nucleotides, not amino acids",
      " -s, --score STRING          MH, Jaccard, product*, ratio, min, max, or mean",
      " -t, --threads INTEGER       number of host threads to use (1*-256)",
      " -u, --ignore-unknown        ignore sequences with unknown symbols",
      " -e, --ignore-empty          ignore empty sequences", "", "Input/output options:",
      " -a, --alternative           output results in three-column format, not matrix",
      "     --cdr3                  use the cdr3(_aa) column instead of junction(_aa)",
      "     --distance              include sequence distance in pairs file",
      " -k, --keep-columns STRING   comma-separated columns to copy to pairs file",
      " -l, --log FILENAME          log to file (stderr*)",
      " -o, --output FILENAME       output results to file (stdout*)",
      "     --no-matrix             do not keep or output any matrix",
      " -p, --pairs FILENAME        output matching pairs to file (none*)", "", "GPU options:",
      "     --gpus INTEGER          number of GPUs to shard set 1 over (1*)",
      "     --device INTEGER        first CUDA device to use (0*)", "",
      "                             * default value", ""};
  for (const char* l : lines) fprintf(stderr, "%s\n", l);
}

void show_time(const char* prompt) {
  char buf[100];
  const time_t now = time(nullptr);
  const size_t n = strftime(buf, sizeof buf, "%a %b %d %T %Z %Y", localtime(&now));
  fprintf(g_log, "%s%s\n", prompt, n > 0 ? buf : "?");
}

void show_args(const Options& o) {
  if (o.matrix) fprintf(g_log, "Command:           Overlap (-m)\n");
  if (o.cluster) fprintf(g_log, "Command:           Cluster (-c)\n");
  if (o.existence) fprintf(g_log, "Command:           Existence (-x)\n");
  if (o.deduplicate) fprintf(g_log, "Command:           Deduplicate (--deduplicate)\n");
  if (o.matrix) {
    fprintf(g_log, "Repertoire set 1:  %s\n", o.input1);
    fprintf(g_log, "Repertoire set 2:  %s\n", o.input2 ? o.input2 : "(same as set 1)");
  } else {
    fprintf(g_log, "Repertoire:        %s\n", o.input1);
    if (o.existence) fprintf(g_log, "Repertoire set:    %s\n", o.input2);
  }
  auto yn = [](bool b) { return b ? "Yes" : "No"; };
  fprintf(g_log, "Nucleotides (n):   %s\n", yn(o.nucleotides));
  fprintf(g_log, "Differences (d):   %ld\n", (long)o.differences);
  fprintf(g_log, "Indels (i):        %s\n", yn(o.indels));
  fprintf(g_log, "Ignore counts (f): %s\n", yn(o.ignore_counts));
  fprintf(g_log, "Ignore genes (g):  %s\n", yn(o.ignore_genes));
  fprintf(g_log, "Ign. unknown (u):  %s\n", yn(o.ignore_unknown));
  fprintf(g_log, "Ignore empty (e):  %s\n", yn(o.ignore_empty));
  fprintf(g_log, "Use cdr3 column:   %s\n", yn(o.cdr3));
  fprintf(g_log, "Threads (t):       %ld\n", (long)o.threads);
  fprintf(g_log, "GPUs:              %d (first device %d)\n", o.gpus, o.device);
  fprintf(g_log, "Output file (o):   %s\n", o.no_matrix ? "(none)" : o.output);
  if (o.matrix || o.existence) {
    fprintf(g_log, "Output format (a): %s\n", o.alternative ? "Column" : "Matrix");
    fprintf(g_log, "Score (s):         %s\n", kScoreDescr[o.score]);
    fprintf(g_log, "Pairs file (p):    %s\n", o.pairs ? o.pairs : "(none)");
    fprintf(g_log, "Keep columns:      %s\n", o.keep_columns ? o.keep_columns : "");
  }
  fprintf(g_log, "Log file (l):      %s\n", o.log ? o.log : "(stderr)");
}

static int64_t parse_long(const char* s, const char* what) {
  char* end = nullptr;
  const int64_t v = strtol(s, &end, 10);
  if (*end) {
    fprintf(stderr, "\nInvalid numeric argument for option %s\n", what);
    cli_exit(1);
  }
  return v;
}

// comma-separated list of [A-Za-z0-9_]+ (parse_keep_columns, compairr.cc:111-171)
static bool parse_keep(const char* s, std::vector<std::string>& out) {
  std::string cur;
  for (const char* p = s;; p++) {
    const char c = *p;
    if (c == ',' || c == 0) {
      if (cur.empty()) return false;
      out.push_back(cur);
      cur.clear();
      if (c == 0) return true;
    } else if ((c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || (c >= '0' && c <= '9') || c == '_') {
      cur.push_back(c);
    } else {
      return false;
    }
  }
}

void parse_args(int argc, char** argv, Options& o) {
  static const char short_opts[] = "acd:efghik:l:mno:p:s:t:uvxz";
  enum { L_CDR3 = 1000, L_DISTANCE, L_NO_MATRIX, L_GPUS, L_DEVICE };
  static const struct option long_opts[] = {
      {"alternative", no_argument, nullptr, 'a'},      {"cdr3", no_argument, nullptr, L_CDR3},
      {"cluster", no_argument, nullptr, 'c'},          {"differences", required_argument, nullptr, 'd'},
      {"distance", no_argument, nullptr, L_DISTANCE},  {"ignore-empty", no_argument, nullptr, 'e'},
      {"ignore-counts", no_argument, nullptr, 'f'},    {"ignore-genes", no_argument, nullptr, 'g'},
      {"help", no_argument, nullptr, 'h'},             {"indels", no_argument, nullptr, 'i'},
      {"keep-columns", required_argument, nullptr, 'k'}, {"log", required_argument, nullptr, 'l'},
      {"matrix", no_argument, nullptr, 'm'},           {"nucleotides", no_argument, nullptr, 'n'},
      {"no-matrix", no_argument, nullptr, L_NO_MATRIX}, {"output", required_argument, nullptr, 'o'},
      {"pairs", required_argument, nullptr, 'p'},      {"score", required_argument, nullptr, 's'},
      {"summands", required_argument, nullptr, 's'},   {"threads", required_argument, nullptr, 't'},
      {"ignore-unknown", no_argument, nullptr, 'u'},   {"version", no_argument, nullptr, 'v'},
      {"existence", no_argument, nullptr, 'x'},        {"deduplicate", no_argument, nullptr, 'z'},
      {"gpus", required_argument, nullptr, L_GPUS},    {"device", required_argument, nullptr, L_DEVICE},
      {nullptr, 0, nullptr, 0}};
  bool used[26] = {false};
  opterr = 1;
  int c;
  while ((c = getopt_long(argc, argv, short_opts, long_opts, nullptr)) != -1) {
    if (c >= 'a' && c <= 'z') {
      if (used[c - 'a']) {  // every option at most once (compairr.cc:403-423)
        const char* lname = "";
        for (const struct option* lo = long_opts; lo->name; lo++)
          if (lo->val == c) {
            lname = lo->name;
            break;
          }
        fprintf(stderr, "Error: Option -%c or --%s specified more than once.\n", c, lname);
        cli_exit(1);
      }
      used[c - 'a'] = true;
    }
    switch (c) {
      case 'a': o.alternative = true; break;
      case 'c': o.cluster = true; break;
      case 'd': o.differences = parse_long(optarg, "-d or --differences"); break;
      case 'e': o.ignore_empty = true; break;
      case 'f': o.ignore_counts = true; break;
      case 'g': o.ignore_genes = true; break;
      case 'h': o.help = true; break;
      case 'i': o.indels = true; break;
      case 'k': o.keep_columns = optarg; break;
      case 'l': o.log = optarg; break;
      case 'm': o.matrix = true; break;
      case 'n': o.nucleotides = true; break;
      case 'o': o.output = optarg; break;
      case 'p': o.pairs = optarg; break;
      case 's': o.score_string = optarg; break;
      case 't':
        o.threads = parse_long(optarg, "-t or --threads");
        o.threads_given = true;
        break;
      case 'u': o.ignore_unknown = true; break;
      case 'v': o.version = true; break;
      case 'x': o.existence = true; break;
      case 'z': o.deduplicate = true; break;
      case L_CDR3: o.cdr3 = true; break;
      case L_DISTANCE: o.distance = true; break;
      case L_NO_MATRIX: o.no_matrix = true; break;
      case L_GPUS: o.gpus = (int)parse_long(optarg, "--gpus"); break;
      case L_DEVICE: o.device = (int)parse_long(optarg, "--device"); break;
      default:
        show_header();
        show_usage();
        cli_exit(1);
    }
  }

  const int cmds = o.help + o.version + o.matrix + o.cluster + o.existence + o.deduplicate;
  if (cmds == 0)
    fatal("Please specify a command (--help, --version, --matrix, --existence, --cluster, or --deduplicate)");
  if (cmds > 1)
    fatal("Please specify just one command (--help, --version, --matrix, --existence, --cluster, or --deduplicate)");

  const int nfiles = argc - optind;
  if (o.help || o.version) {
    if (nfiles != 0) fatal("Incorrect number of arguments");
  } else if (o.matrix) {
    if (nfiles == 2) {
      o.input1 = argv[optind];
      o.input2 = argv[optind + 1];
    } else if (nfiles == 1) {
      o.input1 = argv[optind];
    } else {
      fatal("Incorrect number of arguments. One or two input files must be specified.");
    }
  } else if (o.existence) {
    if (nfiles != 2) fatal("Incorrect number of arguments. Two input files must be specified.");
    o.input1 = argv[optind];
    o.input2 = argv[optind + 1];
  } else {
    if (nfiles != 1) fatal("Incorrect number of arguments. One input file must be specified.");
    o.input1 = argv[optind];
  }

  if (o.deduplicate) {
    if (o.differences != 0) fatal("Option -d or --differences must be 0 for deduplication.");
    if (o.indels) fatal("Option -i or --indels is not allowed for deduplication.");
  }
  if (o.keep_columns) {
    if (!o.pairs) fatal("Option --keep-columns only allowed with --pairs options.");
    if (!parse_keep(o.keep_columns, o.keep_names))
      fatal("Illegal list of columns with --keep-columns option. It must be a comma-separated list of column names. Allowed symbols: A-Z, a-z, _, and 0-9.");
  }
  if (o.threads < 1 || o.threads > 256) {
    fprintf(stderr, "\nError: Illegal number of threads specified with -t or --threads, must be in the range 1 to %u.\n", 256u);
    cli_exit(1);
  }
  if (o.differences < 0) fatal("Differences specified with -d or -differences cannot be negative.");
  if (o.indels && o.differences != 1) fatal("Indels are only allowed when d=1");
  if (o.cluster) {
    if (o.pairs) fatal("Option -p or --pairs is not allowed with -c or --cluster");
    if (o.alternative) fatal("Option -a or --alternative is not allowed with -c or --cluster");
    if (o.score_string) fatal("Option -s or --score is not allowed with -c or --cluster");
  }
  if (o.score_string) {
    o.score = -1;
    for (int i = 0; i < SCORE_END; i++)
      if (strcasecmp(o.score_string, kScoreNames[i]) == 0) {
        o.score = i;
        break;
      }
    if (o.score < 0) fatal("Argument to -s or --score must be MH, Jaccard, product, ratio, min, max or mean");
  }
  if (!o.matrix) {
    if (o.score == SCORE_MH) fatal("The Morisita-Horn index is only allowed when computing repertoire overlap");
    if (o.score == SCORE_JACCARD) fatal("The Jaccard index is only allowed when computing repertoire overlap");
  }
  if (o.differences > 0) {
    if (o.score == SCORE_MH) fatal("The Morisita-Horn index is not defined when d>0");
    if (o.score == SCORE_JACCARD) fatal("The Jaccard index is not defined when d>0");
  }
  if (o.gpus < 1 || o.gpus > 64) fatal("Option --gpus must be in the range 1 to 64.");
  if (o.device < 0) fatal("Option --device cannot be negative.");
  o.alphabet_size = o.nucleotides ? 4 : 20;
  o.seq_header = o.cdr3 ? (o.nucleotides ? "cdr3" : "cdr3_aa") : (o.nucleotides ? "junction" : "junction_aa");
}
