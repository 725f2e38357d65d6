python -m pytest tests -m gpu -x -q -k "sampled_tax or bootstrap" > gpurun_out/pytest_gpu_s13.log 2>&1; echo "pytest rc $?"; tail -15 gpurun_out/pytest_gpu_s13.log
