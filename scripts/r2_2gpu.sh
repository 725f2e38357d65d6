set -x
F=gpurun_out/${1:-r2g}; rm -rf $F; mkdir -p $F
N=${NG:-2}
for w in c3 c2 sweep19; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $w --steps 3 --warmup 3 $( [ $w = sweep19 ] && echo --trials 227328 ) > $F/bench_${w}_${N}gpu.json 2> $F/err_${w}_${N}gpu.txt; echo "$w rc $?"; tail -2 $F/err_${w}_${N}gpu.txt
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload c5 --trials 500000 --steps 2 --warmup 3 > $F/bench_c5_${N}gpu.json 2> $F/err_c5_${N}gpu.txt; echo "c5 rc $?"; tail -2 $F/err_c5_${N}gpu.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --impl reference --steps 1 --warmup 0 > $F/bench_ref_${N}gpu.json 2> $F/err_ref_${N}gpu.txt; echo "ref rc $?"
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob(os.environ['F']+'/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), '%.4e'%d['value'], 'ms %.2f'%d['ms_per_step'], 'e2e %.4e'%d['e2e']['value'], d.get('summary_check'), d['n_gpus'])
    except Exception as e: print(f, 'ERR', e)
PY
