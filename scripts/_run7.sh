python -m pytest tests -m gpu -x -q -k "series or run_strategies" 2>&1 | tail -15
for w in series series19; do python bench.py --workload $w --steps 3 --warmup 3 2> gpurun_out/bench_$w.err | tee gpurun_out/bench_$w.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['config']['kernel'])"; tail -2 gpurun_out/bench_$w.err; done
