#!/bin/bash
# Round-2 GPU call 30 (1 GPU): the hub-sequence test of the tiled build (+ the other build-route tests).
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_redo_paths.py -m gpu -x -q -k "tiled or partitioned" > gpurun_out/r02_c30_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c30_pytest.txt; tail -5 gpurun_out/r02_c30_pytest.txt
