python -m pytest tests -m gpu -x -q -k "run_strategies" 2>&1 | tail -3
python bench.py --workload series --steps 3 --warmup 3 2> gpurun_out/bench_series.err | tee gpurun_out/bench_series.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], d['roofline'], d['config']['kernel'])"
tail -3 gpurun_out/bench_series.err
