#!/bin/bash
# Round-2 GPU call 24 (2 GPUs): e2e with the filter passes on their own side stream (N=1 short bench), the multi-GPU tests
# and the bench at N=2 with the tiled build.
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --skip-strong --skip-c5 --skip-cpu-baseline --skip-parity --skip-d2 > gpurun_out/r02_c24_bench_n1_short.txt 2> gpurun_out/r02_c24_bench_n1_short.err
python - <<'P'
import json
d=json.loads([x for x in open('gpurun_out/r02_c24_bench_n1_short.txt') if x.startswith('{')][-1])
print("N=1 value", d["value"]/1e9, d["ms_per_step"], "e2e", d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"])
P
timeout 600 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r02_c24_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c24_pytest.txt; tail -3 gpurun_out/r02_c24_pytest.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_c24_bench_n2.txt 2> gpurun_out/r02_c24_bench_n2.err
echo "bench n2 rc=$?"
python - <<'P'
import json
d=json.loads([x for x in open('gpurun_out/r02_c24_bench_n2.txt') if x.startswith('{')][-1])
print("N=2 value", d["value"]/1e9, d["ms_per_step"], "e2e", d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], "strong", d["strong"]["value"]/1e9, d["strong"]["ms_per_step"], "parity", d["parity_checked"]["matrix_identical"])
P
