#!/bin/bash
run() { env "$@" python bench.py --steps 200 --warmup 10 --no-e2e --no-cpu-baseline $EXTRA 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$* $EXTRA', j['ms_per_step'], j['roofline']['whole_step']['frac'] if 'whole_step' in j['roofline'] else '')"; }
for g in 148 296 592 1184; do run SPARC_VAL_GRID=$g; done
EXTRA="--no-obs" run A=1
EXTRA="--workload config3" run A=1
