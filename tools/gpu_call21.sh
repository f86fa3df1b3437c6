#!/bin/bash
# Round-2 GPU call 21 (1 GPU): filter passes beside the build with fewer CTAs per SM.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_c21_build_ab.txt; : > $O
for k in 8 5 4 3 2; do
  TAG=filter_ctas_$k COMPAIRR_B200_FILTER_CTAS=$k timeout 300 python tools/build_ab.py 0 2>&1 | tee -a $O | tail -1
done
TAG=filter_ctas_4_no_overlap COMPAIRR_B200_FILTER_CTAS=4 COMPAIRR_B200_FILTER_OVERLAP=0 timeout 300 python tools/build_ab.py 0 2>&1 | tee -a $O | tail -1
