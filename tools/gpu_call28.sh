#!/bin/bash
# Round-2 GPU call 28 (2 GPUs): tiled build on two devices of one process.
set -u
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_multi.py -q -k "tiled or matrix_and_pairs" > gpurun_out/r02_c28_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c28_pytest.txt; tail -4 gpurun_out/r02_c28_pytest.txt
