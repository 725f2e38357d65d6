# generic split pipeline: GPU tests, then the sub5 workload with and without it, and a launch list of the split run
set -x
F=gpurun_out/${1:-r2g}; rm -rf $F; mkdir -p $F
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "generic_split or typed_slot or split" > $F/tests.log 2>&1; tail -3 $F/tests.log
python bench.py --workload sub5 --steps 3 --warmup 3 --no-cpu > $F/bench_sub5_split.json 2> $F/err_split.txt; echo rc $?
LETF_NO_ROUTED_SPLIT=1 python bench.py --workload sub5 --steps 3 --warmup 3 --no-cpu > $F/bench_sub5_generic.json 2> $F/err_generic.txt; echo rc $?
LETF_NO_GENERIC_SPLIT=1 python bench.py --workload sub5 --steps 3 --warmup 3 --no-cpu > $F/bench_sub5_fused.json 2> $F/err_fused.txt; echo rc $?
python bench.py --workload sweep19 --steps 3 --warmup 3 --no-cpu > $F/bench_sweep19.json 2> $F/err_sweep19.txt; echo rc $?
python bench.py --workload c4 --trials 303104 --steps 3 --warmup 3 --no-cpu > $F/bench_c4.json 2> $F/err_c4.txt; echo rc $?
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $F/sub5_launches.csv python bench.py --workload sub5 --trials 113664 --steps 1 --warmup 1 --no-cpu > $F/ncu.log 2>&1
F=$F python - <<'PY'
import json,glob,os
for f in sorted(glob.glob(os.environ['F']+'/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), '%.4e'%d['value'], 'ms %.2f'%d['ms_per_step'], d['config']['kernel']['name'], d['config']['kernel']['regs'], 'e2e %.3e'%d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
grep -i "letf_" $F/sub5_launches.csv | awk -F'","' '{print $5, $NF}' | tail -8
