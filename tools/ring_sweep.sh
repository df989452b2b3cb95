#!/bin/bash
# k_dyn build variants (min resident CTAs -> register cap)
W=${1:-config4}
for lib in build/variants/lib_dyn_*.so; do
  SPARC_B200_LIB=$PWD/$lib python bench.py --workload $W --steps 40 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$W $lib', j['value'], j['roofline']['frac'], j['ms_per_step'])"
  SPARC_B200_LIB=$PWD/$lib python bench.py --workload $W --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --no-obs 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('   dynamics only', j['ms_per_step'])"
done
