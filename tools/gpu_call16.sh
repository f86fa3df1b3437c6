#!/bin/bash
# Round-2 GPU call 16 (1 GPU): final verification of HEAD: GPU suite, smoke, default bench, reference arm.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c16_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c16_pytest.txt
tail -3 gpurun_out/r02_c16_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 1200 python bench.py > gpurun_out/r02_c16_bench_n1.txt 2> gpurun_out/r02_c16_bench_n1.err ) 2>&1 | grep real
echo "bench n1 rc=$?"; tail -c 800 gpurun_out/r02_c16_bench_n1.txt
( time timeout 900 python bench.py --impl reference > gpurun_out/r02_c16_bench_ref.txt 2>&1 ) 2>&1 | grep real; tail -c 400 gpurun_out/r02_c16_bench_ref.txt
