#!/bin/bash
# ncu launch list of the bench command (per-launch device times: compare SHARES, not absolutes).
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-strong --skip-c5 --skip-parity > gpurun_out/r02_bench_under_ncu.log 2>&1
