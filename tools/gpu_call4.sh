#!/bin/bash
# Round-2 GPU call 4 (1 GPU): class filters + new enumeration kernels: parity suite, rates, profiles, bench.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c4_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c4_pytest.txt
tail -12 gpurun_out/r02_c4_pytest.txt
timeout 600 python tools/bigcase.py both 0 > gpurun_out/r02_c4_bigcase.txt 2>&1
tail -3 gpurun_out/r02_c4_bigcase.txt
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:enum1_kernel -s 1 -c 1 -o gpurun_out/r02_enum1 -f python tools/bigcase.py d1 0 > gpurun_out/r02_c4_ncu_e1.log 2>&1
timeout 600 $NCU -k regex:enum2_kernel -s 1 -c 1 -o gpurun_out/r02_enum2 -f python tools/bigcase.py d2 0 > gpurun_out/r02_c4_ncu_e2.log 2>&1
timeout 600 $NCU -k regex:build_kernel -c 1 -o gpurun_out/r02_build4 -f python tools/bigcase.py d1 0 > gpurun_out/r02_c4_ncu_build.log 2>&1
BENCH_DEBUG=1 timeout 900 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline > gpurun_out/r02_c4_bench_n1.txt 2> gpurun_out/r02_c4_bench_n1.err
echo "bench n1 rc=$?"; tail -c 2500 gpurun_out/r02_c4_bench_n1.txt; tail -3 gpurun_out/r02_c4_bench_n1.err
