set -x
F=gpurun_out/r2g8; rm -rf $F; mkdir -p $F
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "generic_split or record_pass" > $F/tests.log 2>&1; tail -3 $F/tests.log
LETF_NO_ROUTED_SPLIT=1 python bench.py --workload sub5 --steps 3 --warmup 3 --no-cpu > $F/bench_sub5_generic8.json 2> $F/err1.txt; echo rc $?
LETF_NO_ROUTED_SPLIT=1 LETF_KERNEL=f32_rules_all python bench.py --workload sub5 --steps 3 --warmup 3 --no-cpu > $F/bench_sub5_generic32.json 2> $F/err2.txt; echo rc $?
F=$F python - <<'PY'
import json,glob,os
for f in sorted(glob.glob(os.environ['F']+'/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), '%.4e'%d['value'], 'ms %.2f'%d['ms_per_step'], d['config']['kernel']['name'], d['config']['kernel']['regs'], 'e2e %.3e'%d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
