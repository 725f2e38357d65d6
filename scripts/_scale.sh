set -x
mkdir -p gpurun_out/final
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/final/bench_c2_${N}gpu.json 2> gpurun_out/final/bench_c2_${N}gpu.err; echo "c2 N=$N rc $?"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --workload c4 --steps 2 --warmup 3 --no-cpu > gpurun_out/final/bench_c4_8gpu.json 2> gpurun_out/final/bench_c4_8gpu.err; echo "c4 N=8 rc $?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --workload c3 --steps 2 --warmup 3 --no-cpu > gpurun_out/final/bench_c3_8gpu.json 2> gpurun_out/final/bench_c3_8gpu.err; echo "c3 N=8 rc $?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --workload c5 --steps 1 --warmup 3 --no-cpu > gpurun_out/final/bench_c5_8gpu.json 2> gpurun_out/final/bench_c5_8gpu.err; echo "c5 N=8 rc $?"
for f in gpurun_out/final/bench_c*gpu.json; do echo $f; cut -c1-160 $f; done
