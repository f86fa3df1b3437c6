// cluster_cmd.h — the `-c` (cluster) and `-z` (deduplicate) commands of the CompAIRR-compatible
// front end: cluster() (src/cluster.cc:302-475) and dedup() (src/dedup.cc:139-215).  The host
// reads the one input set and writes the rows; grouping (dedup) and network + clustering
// (cluster) are done by the engine (cb_dedup / cb_cluster).
#pragma once
#include <stdio.h>

#include "options.h"

void cluster_command(const Options& o, FILE* outfile);
void dedup_command(const Options& o, FILE* outfile);
