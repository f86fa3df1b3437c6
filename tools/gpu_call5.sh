#!/bin/bash
# Round-2 GPU call 5 (1 GPU): class filters v2 (pattern from the free field), L2-blocked filter build, matrix tile.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c5_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c5_pytest.txt
tail -12 gpurun_out/r02_c5_pytest.txt
for bpk in 16 24; do timeout 600 python tools/bigcase.py both $bpk >> gpurun_out/r02_c5_bigcase.txt 2>&1; done
CB_FLAGS=32 TAG=filters_in_build timeout 600 python tools/bigcase.py d1 16 >> gpurun_out/r02_c5_bigcase.txt 2>&1
cat gpurun_out/r02_c5_bigcase.txt
timeout 600 python tools/tile_ab.py 1000000 > gpurun_out/r02_c5_tile_ab.txt 2>&1; cat gpurun_out/r02_c5_tile_ab.txt
timeout 600 python tools/perf_probe.py c2 > gpurun_out/r02_c5_c2.txt 2>&1; cat gpurun_out/r02_c5_c2.txt
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:enum1_kernel -s 1 -c 1 -o gpurun_out/r02_enum1 -f python tools/bigcase.py d1 0 > gpurun_out/r02_c5_ncu_e1.log 2>&1
timeout 600 $NCU -k regex:enum2_kernel -s 1 -c 1 -o gpurun_out/r02_enum2 -f python tools/bigcase.py d2 0 > gpurun_out/r02_c5_ncu_e2.log 2>&1
timeout 600 $NCU -k regex:"build_kernel|filter_kernel" -c 3 -o gpurun_out/r02_build5 -f python tools/bigcase.py d1 0 > gpurun_out/r02_c5_ncu_build.log 2>&1
timeout 600 $NCU -k regex:table_kernel -s 1 -c 1 -o gpurun_out/r02_table5 -f python tools/bigcase.py d1 0 > gpurun_out/r02_c5_ncu_table.log 2>&1
BENCH_DEBUG=1 timeout 900 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline --skip-strong > gpurun_out/r02_c5_bench_n1.txt 2> gpurun_out/r02_c5_bench_n1.err
echo "bench n1 rc=$?"; tail -c 2500 gpurun_out/r02_c5_bench_n1.txt; tail -3 gpurun_out/r02_c5_bench_n1.err
