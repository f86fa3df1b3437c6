#!/bin/bash
# Round-2 GPU call 32 (1 GPU): compute-sanitizer memcheck over every probe path at small size.
set -u
mkdir -p gpurun_out
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/memcheck_paths.py > gpurun_out/r02_c32_memcheck.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_c32_memcheck.txt
tail -22 gpurun_out/r02_c32_memcheck.txt
