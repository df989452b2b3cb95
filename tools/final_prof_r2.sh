#!/bin/bash
# round-2 evidence: bench lines, tails, ncu launch lists + full captures of the hot kernels (run under gpurun, one GPU)
exec > gpurun_out/final_prof_r2.log 2>&1
set -x
B="python bench.py --no-cpu-baseline --no-extra"
python bench.py --steps 200 --warmup 10 --stats-every 16 > gpurun_out/r2_bench_config4.json 2> gpurun_out/bench_config4.err; echo bench=$?
$B --workload config3 --steps 200 --warmup 10 > gpurun_out/r2_bench_config3.json 2>/dev/null
$B --obs-format compact --steps 200 --warmup 10 > gpurun_out/r2_bench_config4_compact.json 2>/dev/null
$B --obs-update delta --steps 200 --warmup 10 > gpurun_out/r2_bench_config4_delta.json 2>/dev/null
$B --obs-update delta --obs-format compact --steps 200 --warmup 10 > gpurun_out/r2_bench_config4_delta_compact.json 2>/dev/null
$B --obs-update delta --validate never --steps 200 --warmup 10 --no-e2e > gpurun_out/r2_bench_config4_delta_novalidate.json 2>/dev/null
$B --validate never --steps 200 --warmup 10 --no-e2e > gpurun_out/r2_bench_config4_novalidate.json 2>/dev/null
for f in gpurun_out/r2_bench_config*.json; do python -c "
import json,sys; j=json.loads(open('$f').read().strip().splitlines()[-1]); r=j.get('roofline') or {}; print('$f', j['value'], j['ms_per_step'], r.get('frac'), (r.get('whole_step') or {}).get('frac'), (j.get('e2e') or {}).get('value'), (j.get('cpu_baseline') or {}).get('value'), j.get('stats_collective'))"; done
python tools/hard_tail.py --steps 400 > gpurun_out/r2_hard_tail.txt 2>&1; cat gpurun_out/r2_hard_tail.txt
SPARC_B200_LIB=$PWD/build/variants/lib_valprof.so python tools/valprof.py --obs-update delta > gpurun_out/r2_valprof.txt 2>&1; cat gpurun_out/r2_valprof.txt
# steady state: 200 warm-up steps, then profile a few steps (3 launches per step: k_dyn, k_encode, k_val)
CMD="python bench.py --steps 8 --warmup 200 --no-e2e --no-cpu-baseline --no-graph --no-extra"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 60 --csv --log-file gpurun_out/r2_launches_steady_state.csv $CMD > gpurun_out/ncu1.log 2>&1; echo ncu1=$?
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_encode|k_dyn|k_val' -s 606 -c 6 -o gpurun_out/r2_prof $CMD > gpurun_out/ncu2.log 2>&1; echo ncu2=$?
CMDD="python bench.py --steps 30 --warmup 200 --no-e2e --no-cpu-baseline --no-graph --no-extra --obs-update delta"
$CMDD > gpurun_out/plain4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 60 --csv --log-file gpurun_out/r2_launches_delta.csv $CMDD > gpurun_out/ncu3.log 2>&1; echo ncu3=$?
$CMDD > gpurun_out/plain5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_post|k_dyn' -s 404 -c 4 -o gpurun_out/r2_prof_delta $CMDD > gpurun_out/ncu4.log 2>&1; echo ncu4=$?
CMDC="python bench.py --steps 8 --warmup 200 --no-e2e --no-cpu-baseline --no-graph --no-extra --obs-format compact"
$CMDC > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_encode' -s 202 -c 2 -o gpurun_out/r2_prof_compact $CMDC > gpurun_out/ncu5.log 2>&1; echo ncu5=$?
ls -la gpurun_out/*.ncu-rep
