#!/bin/bash
# Round-2 GPU call 13 (1 GPU): L2 fetch granularity A/B on the kernels and on the build.
set -u
mkdir -p gpurun_out
: > gpurun_out/r02_c13.txt
for g in 64 32 128; do
  TAG=l2fetch$g COMPAIRR_B200_L2_FETCH_BYTES=$g timeout 600 python tools/bigcase.py both 0 >> gpurun_out/r02_c13.txt 2>&1
  COMPAIRR_B200_L2_FETCH_BYTES=$g BENCH_DEBUG=1 timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --skip-strong --skip-d2 --skip-c5 --skip-parity 2>&1 | grep -E "^\[step\]|^\[e2e" | tail -2 | sed "s/^/l2fetch=$g /" >> gpurun_out/r02_c13.txt
done
cat gpurun_out/r02_c13.txt
