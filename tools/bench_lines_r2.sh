#!/bin/bash
# the round's bench lines only (no ncu): gpurun -- bash tools/bench_lines_r2.sh ; files -> gpurun_out/r2_bench_*.json
exec > gpurun_out/bench_lines_r2.log 2>&1
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
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null
