#!/bin/bash
# round-2 evidence with the final kernels (second half of the round: validator warp in k_env, one-board legality,
# k_env_host / k_info host calls, early-return step_host): bench lines, host-call probes, the k_env cycle profile,
# ncu launch lists + full captures of the hot kernels (run under gpurun, one GPU).  Files -> gpurun_out/, copied to profiles/.
exec > gpurun_out/final_prof_r2b.log 2>&1
set -x
B="python bench.py --no-cpu-baseline --no-extra"
python bench.py --steps 200 --warmup 10 --stats-every 16 > gpurun_out/r2_bench_config4.json 2> gpurun_out/bench_config4.err; echo bench=$?
$B --workload config3 --steps 200 --warmup 10 > gpurun_out/r2_bench_config3.json 2>/dev/null
$B --workload config2 --steps 500 --warmup 20 > gpurun_out/r2_bench_config2.json 2>/dev/null
python bench.py --workload config1 > gpurun_out/r2_bench_config1_dropin.json 2>/dev/null
$B --obs-format compact --steps 200 --warmup 10 > gpurun_out/r2_bench_config4_compact.json 2>/dev/null
$B --obs-update delta --steps 200 --warmup 10 > gpurun_out/r2_bench_config4_delta.json 2>/dev/null
$B --obs-update delta --obs-format compact --steps 200 --warmup 10 > gpurun_out/r2_bench_config4_delta_compact.json 2>/dev/null
$B --obs-update delta --validate never --steps 200 --warmup 10 --no-e2e > gpurun_out/r2_bench_config4_delta_novalidate.json 2>/dev/null
$B --validate never --steps 200 --warmup 10 --no-e2e > gpurun_out/r2_bench_config4_novalidate.json 2>/dev/null
for f in gpurun_out/r2_bench_config*.json; do python -c "
import json,sys; j=json.loads(open('$f').read().strip().splitlines()[-1]); r=j.get('roofline') or {}; print('$f', j['value'], j['ms_per_step'], r.get('frac'), (r.get('whole_step') or {}).get('frac'), (j.get('e2e') or {}).get('value'), (j.get('cpu_baseline') or {}).get('value'), j.get('stats_collective'))"; done
python tools/hostcall_probe.py 4096 1024 256 > gpurun_out/r2_hostcall_config2.txt 2>&1
python tools/probe.py "workload=config2" "workload=config2 validate=never" "workload=config2 obs=0" "workload=config2 envs=4" "workload=config2 envs=4 obs=0" "workload=config2 envs=1024" "workload=config2 envs=16384" "workload=config4" "workload=config4 SPARC_SPLIT=0" "workload=config4 SPARC_SPLIT=0 SPARC_TILE=8" "format=compact" "update=delta" "update=delta format=compact" "workload=config3" "workload=config3 update=delta" > gpurun_out/r2_probe_final.txt 2>&1
SPARC_B200_LIB=$PWD/build/variants/lib_envprof.so python tools/envprof.py > gpurun_out/r2_envprof_config2.txt 2>&1
python tools/hard_tail.py --steps 400 > gpurun_out/r2_hard_tail.txt 2>&1; cat gpurun_out/r2_hard_tail.txt
SPARC_B200_LIB=$PWD/build/variants/lib_valprof.so python tools/valprof.py --obs-update delta > gpurun_out/r2_valprof.txt 2>&1; cat gpurun_out/r2_valprof.txt
# steady state: 200 warm-up steps, then profile a few steps (3 launches per step: k_dyn, k_encode, k_val)
CMD="python bench.py --steps 8 --warmup 200 --no-e2e --no-cpu-baseline --no-graph --no-extra"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 60 --csv --log-file gpurun_out/r2_launches_steady_state.csv $CMD > gpurun_out/ncu1.log 2>&1; echo ncu1=$?
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_encode|k_dyn|k_val' -s 606 -c 6 -f -o gpurun_out/r2_prof $CMD > gpurun_out/ncu2.log 2>&1; echo ncu2=$?
CMDD="python bench.py --steps 30 --warmup 200 --no-e2e --no-cpu-baseline --no-graph --no-extra --obs-update delta"
$CMDD > gpurun_out/plain4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 60 --csv --log-file gpurun_out/r2_launches_delta.csv $CMDD > gpurun_out/ncu3.log 2>&1; echo ncu3=$?
$CMDD > gpurun_out/plain5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_post|k_dyn' -s 404 -c 4 -f -o gpurun_out/r2_prof_delta $CMDD > gpurun_out/ncu4.log 2>&1; echo ncu4=$?
CMDC="python bench.py --steps 8 --warmup 200 --no-e2e --no-cpu-baseline --no-graph --no-extra --obs-format compact"
$CMDC > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_encode' -s 202 -c 2 -f -o gpurun_out/r2_prof_compact $CMDC > gpurun_out/ncu5.log 2>&1; echo ncu5=$?
CMD2="python bench.py --workload config2 --steps 8 --warmup 50 --no-e2e --no-cpu-baseline --no-graph --no-extra"
$CMD2 > gpurun_out/plain6.log 2>&1 && ncu --set full --cache-control none --clock-control none --import-source on -k regex:k_env -s 52 -c 3 -f -o gpurun_out/r2_prof_config2 $CMD2 > gpurun_out/ncu6.log 2>&1; echo ncu6=$?
ls -la gpurun_out/*.ncu-rep
