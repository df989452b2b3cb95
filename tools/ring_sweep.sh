#!/bin/bash
W=${1:-config4}
for lib in build/variants/lib_ring_*.so; do
  SPARC_B200_LIB=$PWD/$lib SPARC_RING=1 python bench.py --workload $W --steps 40 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$W $lib', j['value'], j['roofline']['frac'], j['ms_per_step'])"
done
