python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "edge_sizes or partition or public_api or dropin or mc_runner" 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c3', d['value'], d['e2e']['value'], d['e2e']['per_trial_route'])"
python bench.py --workload sweep19 --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('sweep19', d['value'], d['e2e']['value'], d['e2e']['per_trial_route'])"
