#!/bin/bash
# Round-2 GPU call 18 (1 GPU): tiled build, filters beside it on the side stream, unrolled filter passes — A/B at 10^8
# keys, then the tests of the build routes.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_c18_build_ab.txt; : > $O
TAG=default timeout 600 python tools/build_ab.py 0 64 2>&1 | tee -a $O | tail -2
TAG=no_overlap COMPAIRR_B200_FILTER_OVERLAP=0 timeout 300 python tools/build_ab.py 0 2>&1 | tee -a $O | tail -1
TAG=no_overlap_unroll1 COMPAIRR_B200_FILTER_OVERLAP=0 COMPAIRR_B200_FILTER_UNROLL=1 timeout 300 python tools/build_ab.py 0 2>&1 | tee -a $O | tail -1
TAG=overlap_part24 COMPAIRR_B200_FILTER_PART_MIB=24 timeout 300 python tools/build_ab.py 0 2>&1 | tee -a $O | tail -1
for v in t384x3 t512x3; do
  TAG=$v COMPAIRR_B200_LIB=$PWD/_scratch/lib_$v.so timeout 300 python tools/build_ab.py 0 2>&1 | tee -a $O | tail -1
done
timeout 900 python -m pytest tests/test_gpu_redo_paths.py tests/test_gpu_cluster_dedup.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02_c18_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c18_pytest.txt
tail -5 gpurun_out/r02_c18_pytest.txt
