// overlap_cmd.h — the -m / -x command: what the reference's overlap() (src/overlap.cc:607-1079)
// does around the hot path — read both sets, log the repertoire tables, hand the arrays to the
// GPU engine through the C ABI, write the matrix and the pairs file.
#pragma once
#include "options.h"

void overlap_command(const Options& o, FILE* outfile, FILE* pairsfile);
