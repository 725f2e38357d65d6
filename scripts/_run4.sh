python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s7.log 2>&1; echo "pytest rc $?"; tail -15 gpurun_out/pytest_gpu_s7.log
python bench.py --workload boot --steps 2 --warmup 3 --no-cpu 2> gpurun_out/bench_boot_s7.err | tee gpurun_out/bench_boot_s7.json | cut -c1-2500
tail -5 gpurun_out/bench_boot_s7.err
