#!/bin/bash
# phase split of k_env: full step vs dynamics only (--no-obs) vs no validation
run() { python bench.py --workload config4 --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --pool 4096 "$@" 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$*', j['ms_per_step'], j['roofline']['frac'])"; }
run
run --no-obs
run --validate never
run --no-obs --validate never
run --autoreset disabled
