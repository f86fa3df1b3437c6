// main.cpp — compairr_b200: a CompAIRR-compatible command line (-m / -x / -c / -z) on top of
// the GPU engine (main(), src/compairr.cc:743-798).
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "options.h"
#include "cluster_cmd.h"
#include "overlap_cmd.h"

static FILE* open_output(const char* name) {  // "-" = stdout (util.cc:157-170)
  if (!strcmp(name, "-")) {
    const int fd = dup(STDOUT_FILENO);
    return fd < 0 ? nullptr : fdopen(fd, "w");
  }
  return fopen(name, "w");
}

int main(int argc, char** argv) {
  Options o;
  parse_args(argc, argv, o);
  if (o.log) {
    g_log = open_output(o.log);
    if (!g_log) fatal("Unable to open log file for writing.");
  }
  FILE* outfile = open_output(o.output);
  if (!outfile) fatal("Unable to open output file for writing.");
  FILE* pairsfile = nullptr;
  if (o.pairs) {
    pairsfile = open_output(o.pairs);
    if (!pairsfile) fatal("Unable to open pairs file for writing.");
  }
  show_header();
  if (o.version || o.help) {
    if (o.help) show_usage();
    return 0;
  }
  show_time("Start time:        ");
  show_args(o);
  fprintf(g_log, "\n");
  if (o.matrix || o.existence)
    overlap_command(o, outfile, pairsfile);
  else if (o.deduplicate)
    dedup_command(o, outfile);
  else
    cluster_command(o, outfile);
  show_time("End time:          ");
  // a full disk or a closed pipe must not look like success (the rows go through stdio buffers:
  // the error may only surface at the flush inside fclose)
  auto close_checked = [](FILE* f, const char* what) {
    const bool bad = ferror(f) != 0;
    if (fclose(f) != 0 || bad) {
      fprintf(stderr, "\nError: Unable to write to the %s file.\n", what);
      fflush(nullptr);
      _exit(1);
    }
  };
  if (pairsfile) close_checked(pairsfile, "pairs");
  close_checked(outfile, "output");
  if (g_log != stderr) close_checked(g_log, "log");
  // Everything the user asked for is on disk.  Leave without running the CUDA runtime's exit
  // handlers: tearing down the primary context and its memory pools takes seconds (measured: 4 s
  // of a 5.4 s command) and frees nothing the operating system does not free anyway.
  fflush(nullptr);
  _exit(0);
}
