#!/bin/bash
# Round-2 GPU call 26 (1 GPU): verification of HEAD: GPU suite, smoke, default bench, reference arm.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c26_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c26_pytest.txt
tail -3 gpurun_out/r02_c26_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 1200 python bench.py > gpurun_out/r02_c26_bench_n1.txt 2> gpurun_out/r02_c26_bench_n1.err ) 2>&1 | grep real
echo "bench n1 rc=$?"
python - <<'P'
import json
d=json.loads([x for x in open('gpurun_out/r02_c26_bench_n1.txt') if x.startswith('{')][-1])
print("N=1 value", d["value"]/1e9, d["ms_per_step"], "e2e", d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], "d2", d["d2"]["value"]/1e9, d["d2"]["ms_per_step"], "strong", d["strong"]["value"]/1e9, d["strong"]["ms_per_step"])
print("cli", d["cli_wall"]["ours_wall_s"], d["cli_wall"]["reference_wall_s"], "cpu", d["cpu_baseline"]["value"]/1e9, "parity", d["parity_checked"]["matrix_identical"], d["parity_checked"]["full_size_b"]["matrix_identical"])
P
( time timeout 900 python bench.py --impl reference > gpurun_out/r02_c26_bench_ref.txt 2>&1 ) 2>&1 | grep real; tail -c 300 gpurun_out/r02_c26_bench_ref.txt
