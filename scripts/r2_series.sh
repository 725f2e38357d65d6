# materialised-path mode through the record pass: GPU tests, series / series19 with and without it, launch list
set -x
F=gpurun_out/${1:-r2s}; rm -rf $F; mkdir -p $F
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "series or run_strategies or run_strategy" > $F/tests.log 2>&1; tail -3 $F/tests.log
for w in series series19; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > $F/bench_${w}_split.json 2> $F/err_${w}_split.txt; echo rc $?
  LETF_NO_SPLIT=1 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > $F/bench_${w}_fused.json 2> $F/err_${w}_fused.txt; echo rc $?
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $F/series19_launches.csv python bench.py --workload series19 --steps 1 --warmup 3 --no-cpu > $F/ncu.log 2>&1
F=$F python - <<'PY'
import json,glob,os
for f in sorted(glob.glob(os.environ['F']+'/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), '%.4e'%d['value'], 'ms %.2f'%d['ms_per_step'], d['roofline']['kernel'], 'kernel_ms %.2f'%d['roofline']['kernel_ms'], 'GB/s %.0f'%d['roofline']['achieved'], 'e2e %.3e'%d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
grep -i "letf" $F/series19_launches.csv | awk -F'","' '{print $5, $NF}' | tail -6
