#!/bin/bash
# Round-2 GPU call 27 (4 GPUs): bench at N=4 (all sections) with the tiled build.
set -u
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r02_c27_bench_n4.txt 2> gpurun_out/r02_c27_bench_n4.err
echo "bench n4 rc=$?"
python - <<'P'
import json
d=json.loads([x for x in open('gpurun_out/r02_c27_bench_n4.txt') if x.startswith('{')][-1])
print("N=4 value", d["value"]/1e9, d["ms_per_step"], "e2e", d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], "strong", d["strong"]["value"]/1e9, d["strong"]["ms_per_step"], "d2", d["d2"]["value"]/1e9, "parity", d["parity_checked"]["matrix_identical"])
P
tail -3 gpurun_out/r02_c27_bench_n4.err
