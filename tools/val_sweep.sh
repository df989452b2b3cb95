#!/bin/bash
run() { env "$@" python bench.py --steps 200 --warmup 10 --no-e2e --no-cpu-baseline $EXTRA 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$* $EXTRA', j['ms_per_step'])"; }
run A=1
EXTRA="--no-obs" run A=1
