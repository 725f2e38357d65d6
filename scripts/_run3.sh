python bench.py --workload sweep19 --trials 303104 --steps 2 --warmup 3 --no-cpu 2> gpurun_out/bench_sweep19_s6.err | tee gpurun_out/bench_sweep19_s6.json | cut -c1-200
