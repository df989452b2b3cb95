#!/bin/bash
# where should the fused kernel hand over to the split step?  (config-3 pool)
run() { env "$@" python bench.py --workload config3 --steps 300 --warmup 10 --no-e2e --no-cpu-baseline $EXTRA 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$* $EXTRA', j['ms_per_step'])"; }
for n in 4096 8192 16384 32768; do for sp in 0 1; do EXTRA="--envs-per-gpu $n" run SPARC_SPLIT=$sp; done; done
for n in 4096 8192; do for t in 4 8 16; do EXTRA="--envs-per-gpu $n" run SPARC_SPLIT=1 SPARC_ENC_TILE=$t; done; done
