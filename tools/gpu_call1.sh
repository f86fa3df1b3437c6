#!/bin/bash
# Round-2 GPU call 1: full GPU test-suite, ncu captures of the kernels that had no summary yet
# (build, table, hash, brute) and of the enumeration kernel at the bench's 24 bits/key, bench baseline.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_c1_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c1_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c1_pytest.txt
tail -5 gpurun_out/r02_c1_pytest.txt
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:build_kernel -c 1 -o gpurun_out/r02_build -f python tools/bigcase.py d1 24 > gpurun_out/r02_c1_ncu_build.log 2>&1
timeout 600 $NCU -k regex:table_kernel -s 1 -c 1 -o gpurun_out/r02_table -f python tools/bigcase.py d1 24 > gpurun_out/r02_c1_ncu_table.log 2>&1
timeout 600 $NCU -k regex:hash_kernel -s 3 -c 1 -o gpurun_out/r02_hash -f python tools/bigcase.py d1 24 > gpurun_out/r02_c1_ncu_hash.log 2>&1
timeout 600 $NCU -k regex:variant1_kernel -s 1 -c 1 -o gpurun_out/r02_variant1_24 -f python tools/bigcase.py d1 24 > gpurun_out/r02_c1_ncu_v1.log 2>&1
timeout 600 $NCU -k regex:variant2_kernel -s 1 -c 1 -o gpurun_out/r02_variant2_24 -f python tools/bigcase.py d2 24 > gpurun_out/r02_c1_ncu_v2.log 2>&1
CB_FLAGS=4 timeout 600 $NCU -k regex:brute_kernel -c 1 -o gpurun_out/r02_brute -f python tools/brute_rate.py 100 > gpurun_out/r02_c1_ncu_brute.log 2>&1
timeout 600 python tools/bigcase.py both 24 > gpurun_out/r02_c1_bigcase.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_c1_bench.txt 2>&1
tail -2 gpurun_out/r02_c1_bigcase.txt
tail -c 1500 gpurun_out/r02_c1_bench.txt
ls -la gpurun_out/*.ncu-rep
