python -m pytest tests -m gpu -x -q -k "bootstrap or boot" > gpurun_out/pytest_gpu_s9.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_gpu_s9.log
python bench.py --workload boot --steps 2 --warmup 3 --no-cpu 2> gpurun_out/bench_boot_s9.err | tee gpurun_out/bench_boot_s9.json | cut -c1-300
