# A/B of kernel variants named in $VARIANTS ("workload:kernel" pairs; kernel "-" = default selection)
set -x
F=gpurun_out/${1:-r2f}; rm -rf $F; mkdir -p $F
for v in $VARIANTS; do
  w=${v%%:*}; k=${v##*:}
  if [ "$k" = "-" ]; then python bench.py --workload $w --trials ${TRIALS:-303104} --steps 3 --warmup 3 --no-cpu > $F/bench_${w}_default.json 2> $F/err_${w}_default.txt
  else LETF_KERNEL=$k python bench.py --workload $w --trials ${TRIALS:-303104} --steps 3 --warmup 3 --no-cpu > $F/bench_${w}_${k}.json 2> $F/err_${w}_${k}.txt; fi
  echo "$v rc $?"
done
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob(os.environ['F']+'/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), '%.4e'%d['value'], 'ms %.2f'%d['ms_per_step'], 'kernel_ms %.2f'%d['roofline']['kernel_ms'], d['config']['kernel']['name'], d['config']['kernel']['regs'])
    except Exception as e: print(f, 'ERR', e)
PY
./scripts/pipe_peaks $F/pipe_peaks.json > $F/pipe_peaks.txt 2>&1; cat $F/pipe_peaks.txt
