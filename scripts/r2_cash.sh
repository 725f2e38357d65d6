F=gpurun_out/r2cash; rm -rf $F; mkdir -p $F
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "record_pass or series or generic_split or split" > $F/tests.log 2>&1; tail -2 $F/tests.log
for w in sweep19 c4; do python bench.py --workload $w --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', d['value'], d['roofline']['kernel_ms'], d['e2e']['value'])"; done
