#!/bin/bash
# A/B of library builds on one box: default (in-tree) vs build/variants/*.so; prints ms/step, encoder ms, launches, env-steps/s, fractions
b() { env "$@" python bench.py --steps 500 --warmup 20 --no-e2e --no-cpu-baseline --no-extra $EXTRA 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); r=j['roofline']; print('$* $EXTRA', j['ms_per_step'], r['kernel_ms'], j['gpu_launches'], j['value'], r['frac'], r['whole_step']['frac'])"; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
OLD=SPARC_B200_LIB=$PWD/build/variants/libsparc_b200_shifted_rows.so
LUT=SPARC_B200_LIB=$PWD/build/variants/lib_lut.so
for rep in 1 2; do
EXTRA="" b A=new
EXTRA="" b $OLD
EXTRA="--obs-format compact" b A=new
EXTRA="--obs-format compact" b $OLD
EXTRA="--obs-format compact" b $LUT
done
EXTRA="--workload config3" b A=new
EXTRA="--workload config3" b $OLD
EXTRA="--workload config2" b A=new
EXTRA="--workload config2" b $OLD
