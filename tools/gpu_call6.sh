#!/bin/bash
# Round-2 GPU call 6 (2 GPUs): multi-GPU tests, bench N=1 (full) and N=2, tile crossover, CLI trace.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_redo_paths.py tests/test_gpu_lengths.py -q -x > gpurun_out/r02_c6_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c6_pytest.txt
tail -6 gpurun_out/r02_c6_pytest.txt
timeout 600 python tools/bigcase.py both 0 > gpurun_out/r02_c6_bigcase.txt 2>&1; tail -2 gpurun_out/r02_c6_bigcase.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_c6_bench_n2.txt 2> gpurun_out/r02_c6_bench_n2.err
echo "bench n2 rc=$?"; tail -c 2500 gpurun_out/r02_c6_bench_n2.txt; tail -3 gpurun_out/r02_c6_bench_n2.err
timeout 600 python tools/tile_ab.py 500000 > gpurun_out/r02_c6_tile_ab.txt 2>&1; cat gpurun_out/r02_c6_tile_ab.txt
timeout 900 python tools/cli_trace.py 10 1000 > gpurun_out/r02_c6_cli_trace.txt 2>&1; cat gpurun_out/r02_c6_cli_trace.txt
timeout 900 python tools/cli_trace.py 100 100 --ref > gpurun_out/r02_c6_cli_trace_1e7.txt 2>&1; tail -30 gpurun_out/r02_c6_cli_trace_1e7.txt
