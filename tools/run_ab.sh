b() { env "$@" python bench.py --steps 500 --warmup 20 --no-e2e --no-cpu-baseline --no-extra $EXTRA 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); r=j['roofline']; print('$* $EXTRA', j['ms_per_step'], r['kernel_ms'], j['gpu_launches'], j['value'], r['frac'], r['whole_step']['frac'])"; }
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -2
V=$PWD/build/variants
for rep in 1 2; do
EXTRA="--no-obs --validate never" b A=default
EXTRA="--no-obs --validate never" b SPARC_B200_LIB=$V/lib_rows1.so
EXTRA="--no-obs --validate never" b SPARC_B200_LIB=$V/lib_nostats.so
done
EXTRA="" b A=default
EXTRA="" b SPARC_B200_LIB=$V/lib_rows1.so
EXTRA="" b SPARC_B200_LIB=$V/lib_nostats.so
EXTRA="--obs-format compact" b A=default
EXTRA="--obs-format compact" b SPARC_B200_LIB=$V/lib_rows1.so
EXTRA="--workload config3" b A=default
EXTRA="--workload config3" b SPARC_B200_LIB=$V/lib_rows1.so
EXTRA="--workload config2" b A=default
EXTRA="--workload config2" b SPARC_B200_LIB=$V/lib_rows1.so
