# 8-GPU lines of the final build: c3, c4 at its literal size (10M trials over 8 GPUs), the default sweep, five of the nineteen,
# and the exact cross-rank percentile check over NCCL
F=gpurun_out/${1:-r2n8}; rm -rf $F; mkdir -p $F
N=${NG:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 "$@"; }
for w in c3 c4 sweep19 sub5; do
  run bench.py --gpus $N --workload $w --steps 3 --warmup 3 --no-cpu > $F/bench_${w}_${N}gpu.json 2> $F/err_${w}_${N}gpu.txt; echo "$w rc $?"
done
run scripts/check_exact_percentiles_nccl.py > $F/exact_percentiles.txt 2>&1; echo "percentiles rc $?"; tail -1 $F/exact_percentiles.txt
F=$F python - <<'PY'
import json,glob,os
for f in sorted(glob.glob(os.environ['F']+'/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), '%.4e'%d['value'], 'ms %.2f'%d['ms_per_step'], 'e2e %.4e'%d['e2e']['value'], d.get('summary_check'), d['n_gpus'])
    except Exception as e: print(f, 'ERR', e)
PY
