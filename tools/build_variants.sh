#!/bin/bash
# Build tuning variants of libsparc_b200.so into build/variants/ (selected at run time with
# SPARC_B200_LIB=...).  build/ is git-ignored but travels to the GPU box with gpurun.
# Compile-time knobs: SPARC_MINBLOCKS (k_env), ENC_THREADS / ENC_MINBLOCKS (k_encode), DYN_MINBLOCKS (k_dyn),
# RING_PROD / RING_CONS / RING_TMA_STORE / RING_PROFILE (k_env_ring).
set -e
cd "$(dirname "$0")/.."
SRC=sparc_gym_b200/csrc/sparc_b200.cu
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC"
mkdir -p build/variants
for cfg in "$@"; do      # e.g.  tools/build_variants.sh "enc256 -DENC_THREADS=256 -DENC_MINBLOCKS=4" "ringtma -DRING_TMA_STORE"
  set -- $cfg; name=$1; shift
  nvcc $FLAGS "$@" -o build/variants/lib_$name.so $SRC &
done
wait
ls build/variants
