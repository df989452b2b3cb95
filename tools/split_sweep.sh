#!/bin/bash
# launch strategies: fused k_env vs the k_dyn || k_encode pipeline (chunks x encoder tile)
W=${1:-config4}
run() { env "$@" python bench.py --workload $W --steps 40 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$W $*', j['value'], j['roofline']['frac'], j['ms_per_step'])"; }
run SPARC_SPLIT=0 SPARC_TILE=32
for nc in 1 2 4 8; do for t in 4 8 16 32; do run SPARC_SPLIT=$nc SPARC_ENC_TILE=$t; done; done
