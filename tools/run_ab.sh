#!/bin/bash
# delta observation mode: GPU tests, then step times next to the full rewrite (both formats, with and without validation)
b() { env "$@" python bench.py --steps 500 --warmup 20 --no-e2e --no-cpu-baseline --no-extra $EXTRA 2>gpurun_out/err.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); r=j['roofline']; print('$* $EXTRA', j['ms_per_step'], r['kernel_ms'], j['gpu_launches'], j['value'], r['algorithmic_bytes_per_env_step'], r['whole_step']['frac'])" || tail -3 gpurun_out/err.log; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
EXTRA="" b A=1
EXTRA="--obs-update delta" b A=1
EXTRA="--obs-update delta --validate never" b A=1
EXTRA="--obs-update delta --obs-format compact" b A=1
EXTRA="--obs-update delta --workload config3" b A=1
EXTRA="--obs-update delta --workload config2" b A=1
EXTRA="--workload config2" b A=1
EXTRA="--obs-update delta" b SPARC_VAL_GRID=64
EXTRA="--obs-update delta" b SPARC_VAL_GRID=148
EXTRA="--obs-update delta" b SPARC_VAL_GRID=296
python bench.py --steps 200 --warmup 10 --obs-update delta --no-extra --cpu-seconds 3 > gpurun_out/bench_config4_delta.json 2>gpurun_out/err2.log; tail -c 600 gpurun_out/bench_config4_delta.json
