#!/bin/bash
# Round-2 GPU call 10 (2 GPUs): multi-GPU tests again, e2e phase timing at N=2, d=1 A/B (prefetch, DU, two-stage).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r02_c10_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c10_pytest.txt
tail -6 gpurun_out/r02_c10_pytest.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_lengths.py -q -x 2>&1 | tail -2
TAG=main timeout 600 python tools/bigcase.py both 0 > gpurun_out/r02_c10_bigcase.txt 2>&1
TAG=du2 COMPAIRR_B200_LIB=$PWD/_scratch/lib_du2.so timeout 600 python tools/bigcase.py d1 0 >> gpurun_out/r02_c10_bigcase.txt 2>&1
TAG=e1_two_stage COMPAIRR_B200_LIB=$PWD/_scratch/lib_e1ts.so timeout 600 python tools/bigcase.py d1 0 >> gpurun_out/r02_c10_bigcase.txt 2>&1
cat gpurun_out/r02_c10_bigcase.txt
BENCH_DEBUG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 4 --warmup 3 --skip-strong --skip-d2 --skip-c5 --skip-parity > gpurun_out/r02_c10_bench_n2.txt 2> gpurun_out/r02_c10_bench_n2.err
echo "bench n2 rc=$?"; grep "e2e rank" gpurun_out/r02_c10_bench_n2.err | tail -8; python -c "
import json
d=json.loads([x for x in open('gpurun_out/r02_c10_bench_n2.txt') if x.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e'])"
