#!/bin/bash
run() { env "$@" python bench.py --steps 500 --warmup 20 --no-e2e --no-cpu-baseline --no-extra $EXTRA 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$* $EXTRA', j['ms_per_step'])"; }
for lib in "" mb1 mb3 mb4; do
  L=""; [ -n "$lib" ] && L="SPARC_B200_LIB=$PWD/build/variants/lib_$lib.so"
  EXTRA="--workload config2" run A=1 $L
  EXTRA="--workload config3 --envs-per-gpu 4096" run A=1 $L
done
