#!/bin/bash
# round-1 evidence: parity, bench lines, ncu launch list + full capture of the hot kernels
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -2 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo smoke=$?; tail -1 gpurun_out/smoke.log
python bench.py --steps 200 --warmup 10 --cpu-seconds 10 > gpurun_out/bench_config4.json 2> gpurun_out/bench_config4.err; echo bench=$?
python bench.py --workload config3 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_config3.json 2>/dev/null
python bench.py --workload config2 --steps 500 --warmup 10 --no-cpu-baseline > gpurun_out/bench_config2.json 2>/dev/null
python bench.py --obs-format compact --steps 200 --warmup 10 --no-cpu-baseline --no-extra > gpurun_out/bench_config4_compact.json 2>/dev/null
python bench.py --obs-update delta --steps 200 --warmup 10 --no-cpu-baseline --no-extra > gpurun_out/bench_config4_delta.json 2>/dev/null
python bench.py --obs-update delta --validate never --steps 200 --warmup 10 --no-cpu-baseline --no-extra --no-e2e > gpurun_out/bench_config4_delta_novalidate.json 2>/dev/null
python bench.py --impl reference --steps 1 --cpu-seconds 10 > gpurun_out/bench_reference_arm.json 2>/dev/null
for f in gpurun_out/bench_config*.json; do python -c "
import json,sys; j=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', j['value'], j['ms_per_step'], j['roofline']['frac'], j['roofline']['whole_step']['frac'], (j.get('e2e') or {}).get('value'), (j.get('cpu_baseline') or {}).get('value'))"; done
python tools/hard_tail.py > gpurun_out/hard_tail.txt 2>&1; cat gpurun_out/hard_tail.txt
# steady state: 200 warm-up steps, then profile a few steps (3 launches per step: k_dyn, k_val, k_encode)
CMD="python bench.py --steps 8 --warmup 200 --no-e2e --no-cpu-baseline --no-graph --no-extra"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 40 --csv --log-file gpurun_out/r1_launches.csv $CMD > gpurun_out/ncu1.log 2>&1; echo ncu1=$?
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_encode|k_dyn|k_val' -s 606 -c 6 -o gpurun_out/r1_prof $CMD > gpurun_out/ncu2.log 2>&1; echo ncu2=$?
CMDC="python bench.py --steps 8 --warmup 200 --no-e2e --no-cpu-baseline --no-graph --no-extra --obs-format compact"
$CMDC > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_encode' -s 202 -c 2 -o gpurun_out/r1_prof_compact $CMDC > gpurun_out/ncu3.log 2>&1; echo ncu3=$?
CMDD="python bench.py --steps 8 --warmup 200 --no-e2e --no-cpu-baseline --no-graph --no-extra --obs-update delta"
$CMDD > gpurun_out/plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_val|k_dyn|k_encode_list' -s 606 -c 3 -o gpurun_out/r1_prof_delta $CMDD > gpurun_out/ncu4.log 2>&1; echo ncu4=$?
# the CTA-per-entry k_encode_list of the first delta version, for comparison
SPARC_B200_LIB=$PWD/build/variants/lib_list_cta.so python bench.py --obs-update delta --validate never --steps 200 --warmup 10 --no-cpu-baseline --no-extra --no-e2e 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('list_cta delta novalidate', j['ms_per_step'], j['value'])"
