#!/bin/bash
# Build tuning variants of libsparc_b200.so into build/variants/ (selected at run time with
# SPARC_B200_LIB=...).  build/ is git-ignored but travels to the GPU box with gpurun.
set -e
cd "$(dirname "$0")/.."
SRC=sparc_gym_b200/csrc/sparc_b200.cu
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC"
mkdir -p build/variants
for mb in 1 5 6 8 10 12; do
  nvcc $FLAGS -DDYN_MINBLOCKS=$mb -o build/variants/lib_dyn_b$mb.so $SRC &
done
wait
ls build/variants
