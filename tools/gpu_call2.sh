#!/bin/bash
# Round-2 GPU call 2 (2 GPUs): redo-path + multi-GPU tests, bench at N=1 (all sections) and N=2.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/r02_c2_smi.txt 2>&1
df -h /dev/shm >> gpurun_out/r02_c2_smi.txt; nproc >> gpurun_out/r02_c2_smi.txt
timeout 900 python -m pytest tests/test_gpu_redo_paths.py tests/test_gpu_multi.py -q -x > gpurun_out/r02_c2_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c2_pytest.txt
tail -8 gpurun_out/r02_c2_pytest.txt
BENCH_DEBUG=1 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_c2_bench_n1.txt 2> gpurun_out/r02_c2_bench_n1.err
echo "bench n1 rc=$?"; tail -c 3000 gpurun_out/r02_c2_bench_n1.txt; tail -5 gpurun_out/r02_c2_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_c2_bench_n2.txt 2> gpurun_out/r02_c2_bench_n2.err
echo "bench n2 rc=$?"; tail -c 3000 gpurun_out/r02_c2_bench_n2.txt; tail -5 gpurun_out/r02_c2_bench_n2.err
