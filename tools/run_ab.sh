#!/bin/bash
# k_val_g8 (eight lanes per snapshot) vs k_val: GPU tests, then step times in the three observation modes
b() { env "$@" python bench.py --steps 500 --warmup 20 --no-e2e --no-cpu-baseline --no-extra $EXTRA 2>gpurun_out/err.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); r=j['roofline']; print('$* $EXTRA', j['ms_per_step'], r['kernel_ms'], j['gpu_launches'], j['value'], r['algorithmic_bytes_per_env_step'], r['whole_step']['frac'])" || tail -3 gpurun_out/err.log; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for g8 in 0 1; do
EXTRA="" b SPARC_VAL_G8=$g8
EXTRA="--obs-format compact" b SPARC_VAL_G8=$g8
EXTRA="--obs-update delta" b SPARC_VAL_G8=$g8
EXTRA="--no-obs" b SPARC_VAL_G8=$g8
done
for g in 32 64 296; do EXTRA="--obs-update delta" b SPARC_VAL_GRID_DELTA=$g; done
for g in 16 64 148; do EXTRA="--obs-format compact" b SPARC_VAL_GRID=$g; done
for g in 16 64; do EXTRA="" b SPARC_VAL_GRID=$g; done
EXTRA="--obs-update delta --validate never" b A=1
python tools/hard_tail.py 2>&1 | tail -2
