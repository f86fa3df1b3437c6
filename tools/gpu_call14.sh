#!/bin/bash
# Round-2 GPU call 14 (1 GPU): reader changes: goldens + full suite, bench, the whole C3 problem through the CLI vs the reference.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c14_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c14_pytest.txt
tail -3 gpurun_out/r02_c14_pytest.txt
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_c14_bench_n1.txt 2> gpurun_out/r02_c14_bench_n1.err
echo "bench n1 rc=$?"; tail -c 1200 gpurun_out/r02_c14_bench_n1.txt
timeout 1500 python tools/cli_trace.py 1000 1000 --ref > gpurun_out/r02_c14_cli_c3.txt 2>&1; cat gpurun_out/r02_c14_cli_c3.txt
