#!/bin/bash
# validator prefetches (in-tree) vs without (build/variants/lib_noprefetch.so)
b() { env "$@" python bench.py --steps 500 --warmup 20 --no-e2e --no-cpu-baseline --no-extra $EXTRA 2>gpurun_out/err.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); r=j['roofline']; print('$* $EXTRA', j['ms_per_step'], r['kernel_ms'], j['gpu_launches'], j['value'], r['algorithmic_bytes_per_env_step'], r['whole_step']['frac'])" || tail -3 gpurun_out/err.log; }
OLD=SPARC_B200_LIB=$PWD/build/variants/lib_noprefetch.so
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -2
for rep in 1 2; do
for X in "" "--obs-format compact" "--obs-update delta" "--no-obs"; do
EXTRA="$X" b A=new
EXTRA="$X" b $OLD
done; done
