# round-2 full pass on one GPU: all gpu tests, smoke, default bench (c3), reference arm, other workloads
set -x
F=gpurun_out/${1:-r2b}; rm -rf $F; mkdir -p $F
python -m pytest tests -m gpu -x -q -s > $F/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 $F/pytest_gpu.log; grep -E "joint_pair|f32 tape|c2 FAST" $F/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $F/smoke.log 2>&1; echo "smoke rc $?"; tail -1 $F/smoke.log
python bench.py > $F/bench_default.json 2> $F/bench_default.err; echo "default rc $?"; tail -3 $F/bench_default.err
python bench.py --impl reference > $F/bench_reference.json 2> $F/bench_reference.err; echo "ref rc $?"
for w in c1 c1x c2; do python bench.py --workload $w --steps 5 --warmup 3 > $F/bench_${w}.json 2> $F/bench_${w}.err; echo "$w rc $?"; tail -2 $F/bench_${w}.err; done
for w in c4 sweep19; do python bench.py --workload $w --trials 303104 --steps 3 --warmup 3 --no-cpu > $F/bench_${w}.json 2> $F/bench_${w}.err; echo "$w rc $?"; done
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob(os.environ.get('F','gpurun_out/r2b')+'/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), '%.3e'%d['value'], 'ms %.2f'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], 'frac', (d.get('roofline') or {}).get('frac'), (d.get('cpu_baseline') or {}).get('value'), (d.get('cpu_baseline_numpy') or {}))
    except Exception as e: print(f, 'ERR', e)
PY
