#!/bin/bash
# Round-2 GPU call 34 (1 GPU): compute-sanitizer racecheck over every probe path at small size.
set -u
mkdir -p gpurun_out
timeout 80 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/memcheck_paths.py > gpurun_out/r02_c34_racecheck.txt 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_c34_racecheck.txt
grep -c "hazard" gpurun_out/r02_c34_racecheck.txt; grep -i "hazard\|Race reported" gpurun_out/r02_c34_racecheck.txt | sort | uniq -c | sort -rn | head -20; tail -4 gpurun_out/r02_c34_racecheck.txt
