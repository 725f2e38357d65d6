set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s11.log 2>&1; echo "pytest rc $?" 
for w in c3 c4; do python bench.py --workload $w --trials 303104 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_${w}_s11.json 2> gpurun_out/bench_${w}_s11.err; done
python bench.py --workload sweep19 --trials 303104 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_sweep19_s11.json 2> gpurun_out/bench_sweep19_s11.err
tail -3 gpurun_out/pytest_gpu_s11.log; cat gpurun_out/bench_c3_s11.json gpurun_out/bench_c4_s11.json gpurun_out/bench_sweep19_s11.json
