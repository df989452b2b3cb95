#!/bin/bash
# A/B: k_val launched before k_encode (in-tree build) vs after it (build/variants/lib_valafter.so)
b() { env "$@" python bench.py --steps 500 --warmup 20 --no-e2e --no-cpu-baseline --no-extra $EXTRA 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); r=j['roofline']; print('$* $EXTRA', j['ms_per_step'], r['kernel_ms'], j['gpu_launches'], j['value'], r['frac'], r['whole_step']['frac'])"; }
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -2
OLD=SPARC_B200_LIB=$PWD/build/variants/lib_valafter.so
for rep in 1 2; do
EXTRA="--obs-format compact" b A=new
EXTRA="--obs-format compact" b $OLD
EXTRA="" b A=new
EXTRA="" b $OLD
done
EXTRA="--obs-format compact" b SPARC_VAL_GRID=64
EXTRA="--obs-format compact" b SPARC_VAL_GRID=148
EXTRA="" b SPARC_VAL_GRID=64
EXTRA="--workload config3" b A=new
EXTRA="--workload config3" b $OLD
EXTRA="--no-graph" b A=new
EXTRA="--no-graph" b $OLD
