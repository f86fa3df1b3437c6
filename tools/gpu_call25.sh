#!/bin/bash
# Round-2 GPU call 25 (1 GPU): records of the keys set aside prefetched to L2 in pass 1 of the tile kernel.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_c25_build_ab.txt; : > $O
TAG=prefetch timeout 300 python tools/build_ab.py 0 2>&1 | tee -a $O | tail -1
TAG=prefetch_t384x3 COMPAIRR_B200_LIB=$PWD/_scratch/lib_t384x3.so timeout 300 python tools/build_ab.py 0 2>&1 | tee -a $O | tail -1
