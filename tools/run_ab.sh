b() { env "$@" python bench.py --steps 500 --warmup 20 --no-e2e --no-cpu-baseline --no-extra $EXTRA 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); r=j['roofline']; print('$* $EXTRA', j['ms_per_step'], r['kernel_ms'], j['gpu_launches'], j['value'], r['frac'], r['whole_step']['frac'])"; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
EXTRA="" b A=1
EXTRA="--obs-format compact" b A=1
EXTRA="--workload config3" b A=1
EXTRA="--workload config2" b A=1
EXTRA="--no-obs" b A=1
python tools/hard_tail.py 2>&1 | tail -2
