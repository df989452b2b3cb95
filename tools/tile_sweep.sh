#!/bin/bash
# bench.py over tile sizes with the default library (SPARC_TILE = envs per CTA)
W=${1:-config4}; shift
for t in "$@"; do
  SPARC_TILE=$t python bench.py --workload $W --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --pool 512 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$W tile=$t', j['value'], j['roofline']['frac'], j['ms_per_step'])"
done
