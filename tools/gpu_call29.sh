#!/bin/bash
# Round-2 GPU call 29 (1 GPU): word-wide hash loads + conditional link reset: full GPU suite, smoke, short bench.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c29_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c29_pytest.txt
tail -3 gpurun_out/r02_c29_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --skip-strong --skip-c5 --skip-cpu-baseline --skip-parity > gpurun_out/r02_c29_bench_n1_short.txt 2> gpurun_out/r02_c29_bench_n1_short.err
python - <<'P'
import json
d=json.loads([x for x in open('gpurun_out/r02_c29_bench_n1_short.txt') if x.startswith('{')][-1])
print("N=1 value", d["value"]/1e9, d["ms_per_step"], "e2e", d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], "d2", d["d2"]["value"]/1e9, d["d2"]["ms_per_step"])
print({k: v for k, v in d.items() if k.startswith("ms_") or k in ("phases", "stats")})
P
