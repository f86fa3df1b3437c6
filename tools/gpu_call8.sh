#!/bin/bash
# Round-2 GPU call 8 (1 GPU): single-stage loop restored (one inlined copy), pinned staging, A/B, full bench.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c8_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c8_pytest.txt
tail -4 gpurun_out/r02_c8_pytest.txt
TAG=main timeout 600 python tools/bigcase.py both 0 > gpurun_out/r02_c8_bigcase.txt 2>&1
TAG=shifts_on_alu COMPAIRR_B200_LIB=$PWD/_scratch/lib_shf.so timeout 600 python tools/bigcase.py both 0 >> gpurun_out/r02_c8_bigcase.txt 2>&1
for mib in 32 64 96; do TAG=part$mib COMPAIRR_B200_FILTER_PART_MIB=$mib BENCH_DEBUG=1 timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --skip-strong --skip-d2 --skip-c5 --skip-parity 2>&1 | grep "^\[step\]" | tail -1 | sed "s/^/part_mib=$mib /" >> gpurun_out/r02_c8_bigcase.txt; done
cat gpurun_out/r02_c8_bigcase.txt
timeout 900 python tools/cli_trace.py 10 1000 > gpurun_out/r02_c8_cli_trace.txt 2>&1; tail -22 gpurun_out/r02_c8_cli_trace.txt
BENCH_DEBUG=1 timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_c8_bench_n1.txt 2> gpurun_out/r02_c8_bench_n1.err
echo "bench n1 rc=$?"; tail -c 6000 gpurun_out/r02_c8_bench_n1.txt; tail -3 gpurun_out/r02_c8_bench_n1.err
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_c8_bench_ref.txt 2>&1; tail -c 1500 gpurun_out/r02_c8_bench_ref.txt
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:enum1_kernel -s 1 -c 1 -o gpurun_out/r02_enum1 -f python tools/bigcase.py d1 0 > gpurun_out/r02_c8_ncu_e1.log 2>&1
