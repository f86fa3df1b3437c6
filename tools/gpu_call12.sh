#!/bin/bash
# Round-2 GPU call 12 (8 GPUs): multi-GPU tests with 4 ranks, bench at N=8 (all sections) with phase timing.
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_c12_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r02_c12_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c12_pytest.txt; tail -3 gpurun_out/r02_c12_pytest.txt
BENCH_DEBUG=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_c12_bench_n8.txt 2> gpurun_out/r02_c12_bench_n8.err
echo "bench n8 rc=$?"; grep "e2e rank 0" gpurun_out/r02_c12_bench_n8.err | tail -3; grep "e2e rank 7" gpurun_out/r02_c12_bench_n8.err | tail -2; tail -c 4000 gpurun_out/r02_c12_bench_n8.txt; tail -5 gpurun_out/r02_c12_bench_n8.err
