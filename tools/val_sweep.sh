#!/bin/bash
run() { env "$@" python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline $EXTRA 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$* $EXTRA', j['ms_per_step'], j['roofline']['kernel_ms'])"; }
run A=1
for t in 4 8 16; do run SPARC_B200_LIB=$PWD/build/variants/lib_skip1.so SPARC_ENC_TILE=$t; done
