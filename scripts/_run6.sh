python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s14.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/pytest_gpu_s14.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | cut -c1-200
