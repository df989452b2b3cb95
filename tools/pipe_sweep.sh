#!/bin/bash
W=${1:-config4}
for pipe in 0 1; do for t in 32; do
  SPARC_RING=$pipe SPARC_TILE=$t python bench.py --workload $W --steps 40 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$W ring=$pipe tile=$t', j['value'], j['roofline']['frac'], j['ms_per_step'])"
done; done
