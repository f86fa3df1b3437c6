// options.h — command line of the CompAIRR-compatible front end.
// Same grammar, defaults, validation order and error texts as the reference's args_init()
// (src/compairr.cc:292-706); two long-only extensions select GPUs (--gpus N, --device K).
#pragma once
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

struct Options {
  bool alternative = false, cdr3 = false, cluster = false, deduplicate = false, distance = false;
  bool existence = false, help = false, ignore_counts = false, ignore_empty = false;
  bool ignore_genes = false, ignore_unknown = false, indels = false, matrix = false;
  bool nucleotides = false, no_matrix = false, version = false;
  const char* keep_columns = nullptr;
  const char* log = nullptr;
  const char* output = "-";
  const char* pairs = nullptr;
  const char* score_string = nullptr;
  int64_t differences = 0;
  int64_t score = 0;  // reference enum order: product ratio min max mean mh jaccard
  int64_t threads = 1;
  bool threads_given = false;  // an explicit -t (also -t 1) sizes the host-side reader and writers
  const char* input1 = nullptr;
  const char* input2 = nullptr;
  const char* seq_header = "junction_aa";
  int alphabet_size = 20;
  std::vector<std::string> keep_names;
  // extensions
  int gpus = 1;
  int device = 0;
};

enum { SCORE_PRODUCT, SCORE_RATIO, SCORE_MIN, SCORE_MAX, SCORE_MEAN, SCORE_MH, SCORE_JACCARD, SCORE_END };

extern FILE* g_log;  // stderr or the -l file

[[noreturn]] // flush the C streams and leave at once: no exit handlers (a CUDA context may be coming up on
// another thread, and the CUDA runtime's own teardown takes seconds)
[[noreturn]] void cli_exit(int code);
[[noreturn]] void fatal(const char* msg);  // "\nError: <msg>\n" on stderr, exit 1 (util.cc:84-88)
void parse_args(int argc, char** argv, Options& o);
void show_header();
void show_usage();
void show_args(const Options& o);
void show_time(const char* prompt);
