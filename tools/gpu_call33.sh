#!/bin/bash
# Round-2 GPU call 33 (1 GPU): compute-sanitizer racecheck (shared-memory hazards) over the tiled build.
set -u
mkdir -p gpurun_out
timeout 100 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/memcheck_tiled.py > gpurun_out/r02_c33_racecheck.txt 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_c33_racecheck.txt
grep -c "hazard" gpurun_out/r02_c33_racecheck.txt; tail -12 gpurun_out/r02_c33_racecheck.txt
