#!/bin/sh
# Build a variant of libcompairr_b200.so with extra -D flags into _scratch/ for A/B runs:
#   tools/build_variant.sh u4 -DVK_U1=4            # d=1 loop: four candidates per lane per step
#   tools/build_variant.sh k2 -DCB_PATTERN_HALF_BITS=2
#   tools/build_variant.sh cta4 -DVK_D1_CTAS=4
# then:  COMPAIRR_B200_LIB=$PWD/_scratch/lib_u4.so python tools/prof_big.py 24 1
# (_scratch/ is git-ignored but travels to the GPU box with gpurun.)
set -e
name=$1; shift
cd "$(dirname "$0")/../compairr_b200/csrc"
mkdir -p ../../_scratch/obj_$name
for f in kernels variant engine upload brute hamming_tc cluster comm; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC \
    -ccbin /usr/bin/g++ --expt-relaxed-constexpr "$@" -c -o ../../_scratch/obj_$name/$f.o $f.cu &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../_scratch/lib_$name.so \
  ../../_scratch/obj_$name/*.o -ccbin /usr/bin/g++ -cudart shared -lnccl -lpthread
rm -rf ../../_scratch/obj_$name
echo "built _scratch/lib_$name.so"
