#!/bin/bash
# Round-2 GPU call 15 (1 GPU): word-wide verify A/B + parity subset, CLI at C3 size with the parallel repertoire statistics.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_redo_paths.py tests/test_gpu_lengths.py tests/test_gpu_cluster_dedup.py -x -q 2>&1 | tail -3
TAG=main timeout 600 python tools/bigcase.py both 0 > gpurun_out/r02_c15_bigcase.txt 2>&1
TAG=verify_bytes COMPAIRR_B200_LIB=$PWD/_scratch/lib_vbytes.so timeout 600 python tools/bigcase.py both 0 >> gpurun_out/r02_c15_bigcase.txt 2>&1
cat gpurun_out/r02_c15_bigcase.txt
timeout 900 python tools/cli_trace.py 1000 1000 > gpurun_out/r02_c15_cli_c3.txt 2>&1; tail -22 gpurun_out/r02_c15_cli_c3.txt
