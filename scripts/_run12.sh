mkdir -p gpurun_out/final
timeout 600 ncu --set full --clock-control none --import-source on -k regex:letf_path_kernel -s 3 -c 1 -f -o gpurun_out/final/prof_c3b python bench.py --workload c3 --trials 75776 --steps 1 --warmup 3 --no-cpu > gpurun_out/final/ncu_c3b.log 2>&1; echo rc $?
