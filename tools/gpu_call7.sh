#!/bin/bash
# Round-2 GPU call 7 (1 GPU): two-stage filter test + FMA-pipe shifts: parity suite, A/B, bench with C5, traces.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c7_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c7_pytest.txt
tail -4 gpurun_out/r02_c7_pytest.txt
TAG=main timeout 600 python tools/bigcase.py both 0 > gpurun_out/r02_c7_bigcase.txt 2>&1
TAG=shifts_on_alu COMPAIRR_B200_LIB=$PWD/_scratch/lib_shf.so timeout 600 python tools/bigcase.py both 0 >> gpurun_out/r02_c7_bigcase.txt 2>&1
TAG=main24 timeout 600 python tools/bigcase.py both 24 >> gpurun_out/r02_c7_bigcase.txt 2>&1
cat gpurun_out/r02_c7_bigcase.txt
timeout 600 python tools/tile_ab.py 500000 > gpurun_out/r02_c7_tile_ab.txt 2>&1; cat gpurun_out/r02_c7_tile_ab.txt
timeout 900 python tools/cli_trace.py 10 1000 > gpurun_out/r02_c7_cli_trace.txt 2>&1; cat gpurun_out/r02_c7_cli_trace.txt
timeout 900 python tools/cli_trace.py 100 100 --ref > gpurun_out/r02_c7_cli_trace_1e7.txt 2>&1; tail -40 gpurun_out/r02_c7_cli_trace_1e7.txt
BENCH_DEBUG=1 timeout 900 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline --skip-strong > gpurun_out/r02_c7_bench_n1.txt 2> gpurun_out/r02_c7_bench_n1.err
echo "bench n1 rc=$?"; tail -c 3500 gpurun_out/r02_c7_bench_n1.txt; tail -3 gpurun_out/r02_c7_bench_n1.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:enum1_kernel -s 1 -c 1 -o gpurun_out/r02_enum1 -f python tools/bigcase.py d1 0 > gpurun_out/r02_c7_ncu_e1.log 2>&1
