#!/bin/bash
# Round-2 GPU call 31 (1 GPU): compute-sanitizer memcheck over hash + tiled build + a short run.
set -u
mkdir -p gpurun_out
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/memcheck_tiled.py > gpurun_out/r02_c31_memcheck.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_c31_memcheck.txt
tail -6 gpurun_out/r02_c31_memcheck.txt
