#!/bin/bash
# A/B of library builds / run-time knobs on ONE box (same clocks, same neighbours): every argument is one arm,
# written as "ENV=VALUE ... -- bench.py flags".  The in-tree build is used unless an arm sets SPARC_B200_LIB
# (build variants with tools/build_variants.sh).  Prints ms/step, dominant-kernel ms, launches, env-steps/s,
# algorithmic bytes per env-step and the whole-step fraction of the HBM copy bandwidth.
#
#   gpurun -- 'bash tools/run_ab.sh "A=1 --" "SPARC_B200_LIB=$PWD/build/variants/lib_x.so --" \
#              "A=1 -- --obs-format compact" "SPARC_VAL_GRID=64 -- --obs-update delta" | tee gpurun_out/ab.log'
#
# (round 1's A/B logs are in profiles/r1_ab_*.txt)
for arm in "$@"; do
  envs="${arm%%--*}"; flags="${arm#*--}"
  env $envs python bench.py --steps 500 --warmup 20 --no-e2e --no-cpu-baseline --no-extra $flags 2>gpurun_out/err.log |
    python -c "import sys,json; j=json.loads(sys.stdin.read()); r=j['roofline']; print('$arm |', j['ms_per_step'], r['kernel_ms'], j['gpu_launches'], j['value'], r['algorithmic_bytes_per_env_step'], r['whole_step']['frac'])" ||
    tail -3 gpurun_out/err.log
done
