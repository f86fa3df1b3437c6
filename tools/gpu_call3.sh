#!/bin/bash
# Round-2 GPU call 3 (1 GPU): new enumeration kernels — parity suite, rates at C3 geometry.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c3_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c3_pytest.txt
tail -12 gpurun_out/r02_c3_pytest.txt
timeout 600 python tools/bigcase.py both 24 > gpurun_out/r02_c3_bigcase.txt 2>&1
cat gpurun_out/r02_c3_bigcase.txt | tail -3
CB_FLAGS=16 timeout 600 python tools/bigcase.py d1 24 > gpurun_out/r02_c3_bigcase_generic.txt 2>&1
tail -2 gpurun_out/r02_c3_bigcase_generic.txt
