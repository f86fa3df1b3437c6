#!/bin/bash
# Round-2 GPU call 17 (1 GPU): the tiled set-B build — its tests, then build time at 10^8 keys against the swept and
# direct routes and three CTA shapes of the tile kernel.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_redo_paths.py tests/test_gpu_cluster_dedup.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02_c17_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c17_pytest.txt
tail -5 gpurun_out/r02_c17_pytest.txt
TAG=t512x2 timeout 600 python tools/build_ab.py 0 64 8 2>&1 | tee gpurun_out/r02_c17_build_ab.txt | tail -4
for v in t256x3 t384x3 t1024x1; do
  TAG=$v COMPAIRR_B200_LIB=$PWD/_scratch/lib_$v.so timeout 300 python tools/build_ab.py 0 2>&1 | tee -a gpurun_out/r02_c17_build_ab.txt | tail -1
done
