#!/bin/bash
run() { env "$@" python bench.py --steps 200 --warmup 10 --no-e2e --no-cpu-baseline $EXTRA 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$* $EXTRA', j['ms_per_step'], j['gpu_launches'])"; }
for nc in 1 2 3 4 8; do run SPARC_SPLIT=$nc; done
