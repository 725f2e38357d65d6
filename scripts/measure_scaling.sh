set -x
F=gpurun_out/final2; mkdir -p $F
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > $F/bench_c2_${N}gpu.json 2> $F/bench_c2_${N}gpu.err; echo "c2 N=$N rc $?"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --workload c4 --steps 2 --warmup 3 --no-cpu > $F/bench_c4_8gpu.json 2> $F/bench_c4_8gpu.err; echo "c4 N=8 rc $?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --workload sweep19 --steps 2 --warmup 3 --no-cpu > $F/bench_sweep19_8gpu.json 2> $F/bench_sweep19_8gpu.err; echo "sweep19 N=8 rc $?"
