F=gpurun_out/r2sub6; rm -rf $F; mkdir -p $F
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "generic_split or record_pass or joint_pair or chunk_invariant or partition" > $F/tests.log 2>&1; tail -2 $F/tests.log
python bench.py --workload sma2 --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('sma2', d['value'], d['ms_per_step'], d['config']['kernel']['name'], d['e2e']['value'])"
LETF_KERNEL=f32_rules_sub19 python bench.py --workload sma2 --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('sma2 via sub19', d['value'], d['ms_per_step'], d['config']['kernel']['name'])"
