#!/bin/bash
# Round-2 GPU call 11 (1 GPU): final state: parity suite, smoke, bench, profiles (launch list, enumeration kernels, build, traffic).
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c11_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c11_pytest.txt
tail -3 gpurun_out/r02_c11_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
TAG=main timeout 600 python tools/bigcase.py both 0 > gpurun_out/r02_c11_bigcase.txt 2>&1; cat gpurun_out/r02_c11_bigcase.txt
BENCH_DEBUG=1 timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_c11_bench_n1.txt 2> gpurun_out/r02_c11_bench_n1.err
echo "bench n1 rc=$?"; tail -c 1500 gpurun_out/r02_c11_bench_n1.txt
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:enum1_kernel -c 1 -o gpurun_out/r02_enum1_first -f python tools/bigcase.py d1 0 > gpurun_out/r02_c11_ncu_e1a.log 2>&1
timeout 600 $NCU -k regex:enum1_kernel -s 1 -c 1 -o gpurun_out/r02_enum1 -f python tools/bigcase.py d1 0 > gpurun_out/r02_c11_ncu_e1.log 2>&1
timeout 600 $NCU -k regex:enum2_kernel -s 1 -c 1 -o gpurun_out/r02_enum2 -f python tools/bigcase.py d2 0 > gpurun_out/r02_c11_ncu_e2.log 2>&1
timeout 600 $NCU -k regex:build_kernel -c 1 -o gpurun_out/r02_build_final -f python tools/bigcase.py d1 0 > gpurun_out/r02_c11_ncu_build.log 2>&1
timeout 600 $NCU -k regex:hamming_tc -c 1 -o gpurun_out/r02_tc -f python tools/brute_rate.py 100 > gpurun_out/r02_c11_ncu_tc.log 2>&1
timeout 900 bash tools/launch_list.sh
ls -la gpurun_out/r02_launches.csv
