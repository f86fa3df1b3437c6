#!/bin/bash
# Round-2 GPU call 23 (1 GPU): verification of HEAD with the tiled build: GPU suite, smoke, default bench, the bench's launch
# list and a full capture of the tile kernel.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c23_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c23_pytest.txt
tail -3 gpurun_out/r02_c23_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 1200 python bench.py > gpurun_out/r02_c23_bench_n1.txt 2> gpurun_out/r02_c23_bench_n1.err ) 2>&1 | grep real
echo "bench n1 rc=$?"; tail -c 300 gpurun_out/r02_c23_bench_n1.txt
bash tools/launch_list.sh
python tools/launch_summary.py gpurun_out/r02_launches.csv gpurun_out/r02_c23_bench_launches.txt "python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-strong --skip-c5 --skip-parity (first 600 launches)" | head -24
timeout 600 ncu --set full --clock-control none --import-source on -k regex:build_tile_kernel -s 1 -c 1 -f -o gpurun_out/r02_c23_tile \
  python tools/build_ab.py 0 > gpurun_out/r02_c23_under_ncu2.log 2>&1
ls -la gpurun_out/r02_c23_tile.ncu-rep
