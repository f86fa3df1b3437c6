#!/bin/bash
# Round-2 GPU call 19 (1 GPU): where the set-B build's time goes — launch list of three builds at 10^8 keys, full captures
# of the tile kernel and one filter pass.
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c19_build_launches.csv \
  python tools/build_ab.py 0 > gpurun_out/r02_c19_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r02_c19_build_launches.csv gpurun_out/r02_c19_build_launches.txt "python tools/build_ab.py 0  (three set-B builds at 10^8 keys + one run of 10^6 seeds)" | head -30
timeout 600 ncu --set full --clock-control none --import-source on -k regex:build_tile_kernel -s 1 -c 1 -f -o gpurun_out/r02_c19_tile \
  python tools/build_ab.py 0 > gpurun_out/r02_c19_under_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:filter_kernel -s 21 -c 1 -f -o gpurun_out/r02_c19_filter \
  python tools/build_ab.py 0 > gpurun_out/r02_c19_under_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
