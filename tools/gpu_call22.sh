#!/bin/bash
# Round-2 GPU call 22 (1 GPU): filter word ranges of 64 / 96 MiB (3 / 2 passes per class instead of 4).
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_c22_build_ab.txt; : > $O
for k in 48 64 96 200; do
  TAG=part_mib_$k COMPAIRR_B200_FILTER_PART_MIB=$k timeout 300 python tools/build_ab.py 0 2>&1 | tee -a $O | tail -1
done
