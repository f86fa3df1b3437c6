#!/bin/bash
# Round-2 GPU call 20 (1 GPU): two-pass tile kernel — build time at 10^8 keys, launch list, the build-route tests.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_c20_build_ab.txt; : > $O
TAG=default timeout 600 python tools/build_ab.py 0 64 2>&1 | tee -a $O | tail -2
TAG=t384x3 COMPAIRR_B200_LIB=$PWD/_scratch/lib_t384x3.so timeout 300 python tools/build_ab.py 0 2>&1 | tee -a $O | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c20_build_launches.csv \
  python tools/build_ab.py 0 > gpurun_out/r02_c20_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r02_c20_build_launches.csv gpurun_out/r02_c20_build_launches.txt "python tools/build_ab.py 0  (three set-B builds at 10^8 keys + one run of 10^6 seeds)" | head -12
timeout 900 python -m pytest tests/test_gpu_redo_paths.py tests/test_gpu_cluster_dedup.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02_c20_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c20_pytest.txt
tail -4 gpurun_out/r02_c20_pytest.txt
