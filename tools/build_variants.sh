#!/bin/bash
# Build tuning variants of libsparc_b200.so into build/variants/ (selected at run time with
# SPARC_B200_LIB=...).  build/ is git-ignored but travels to the GPU box with gpurun.
set -e
cd "$(dirname "$0")/.."
SRC=sparc_gym_b200/csrc/sparc_b200.cu
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC"
mkdir -p build/variants
for mb in 1 4 6 8; do
  for hint in cs wb; do
    h=".$hint"; [ "$hint" = wb ] && h=""
    nvcc $FLAGS -DSPARC_MINBLOCKS=$mb -DSPARC_STORE_HINT="\"$h\"" -o build/variants/lib_mb${mb}_${hint}.so $SRC &
  done
done
wait
ls -la build/variants
