# split strategy pipeline: parity tests, c4 / sweep19 benches (split and, for A/B, fused), ncu of both kernels
set -x
F=gpurun_out/${1:-r2c}; rm -rf $F; mkdir -p $F
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "split_pipeline or typed_slot or edge_sizes or benchmark_horizons" > $F/pytest.log 2>&1; echo "pytest rc $?"; tail -4 $F/pytest.log
for w in c4 sweep19; do
  python bench.py --workload $w --trials 303104 --steps 3 --warmup 3 --no-cpu > $F/bench_${w}.json 2> $F/bench_${w}.err; echo "$w rc $?"; tail -2 $F/bench_${w}.err
  LETF_NO_SPLIT=1 python bench.py --workload $w --trials 303104 --steps 2 --warmup 3 --no-cpu > $F/bench_${w}_fused.json 2> $F/bench_${w}_fused.err; echo "$w fused rc $?"
done
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob(os.environ['F']+'/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), '%.3e'%d['value'], 'ms %.2f'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], 'frac %.3f'%d['roofline']['frac'], d['config']['kernel'])
    except Exception as e: print(f, 'ERR', e)
PY
for w in sweep19 c4; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"letf_path_kernel|letf_jpair_kernel" -s 6 -c 2 -f -o $F/prof_${w} python bench.py --workload $w --trials 113664 --steps 1 --warmup 3 --no-cpu > $F/ncu_${w}.log 2>&1; echo "ncu $w rc $?"
done
./scripts/pipe_peaks $F/pipe_peaks.json > $F/pipe_peaks.txt 2>&1; echo "pipe rc $?"; cat $F/pipe_peaks.txt
