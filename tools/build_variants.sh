#!/bin/bash
# Build tuning variants of libsparc_b200.so into build/variants/ (selected at run time with
# SPARC_B200_LIB=...).  build/ is git-ignored but travels to the GPU box with gpurun.
set -e
cd "$(dirname "$0")/.."
SRC=sparc_gym_b200/csrc/sparc_b200.cu
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC"
mkdir -p build/variants
for pc in "8 8" "12 4" "10 6" "6 10" "12 8"; do
  set -- $pc
  nvcc $FLAGS -DRING_PROD=$1 -DRING_CONS=$2 -o build/variants/lib_ring_p$1_c$2.so $SRC &
done
wait
ls -la build/variants
