#!/bin/bash
# Round-2 GPU call 9 (2 GPUs): multi-GPU tests, bench N=2, d=1 two-stage A/B.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/r02_c9_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c9_pytest.txt
tail -6 gpurun_out/r02_c9_pytest.txt
TAG=main timeout 600 python tools/bigcase.py both 0 > gpurun_out/r02_c9_bigcase.txt 2>&1
TAG=e1_two_stage COMPAIRR_B200_LIB=$PWD/_scratch/lib_e1ts.so timeout 600 python tools/bigcase.py d1 0 >> gpurun_out/r02_c9_bigcase.txt 2>&1
cat gpurun_out/r02_c9_bigcase.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_c9_bench_n2.txt 2> gpurun_out/r02_c9_bench_n2.err
echo "bench n2 rc=$?"; tail -c 5000 gpurun_out/r02_c9_bench_n2.txt; tail -3 gpurun_out/r02_c9_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_c9_bench_ref_n2.txt 2>&1; tail -c 600 gpurun_out/r02_c9_bench_ref_n2.txt
