F=gpurun_out/r2c1x; rm -rf $F; mkdir -p $F
timeout 600 ncu --set full --clock-control none --import-source on -k regex:letf_path_kernel -s 3 -c 1 -f -o $F/prof_c1x python bench.py --workload c1x --trials 20000000 --steps 1 --warmup 3 --no-cpu > $F/ncu.log 2>&1; echo rc $?
python scripts/ncu_metrics.py $F/prof_c1x.ncu-rep 0 > $F/c1x_summary.csv; cat $F/c1x_summary.csv
python scripts/ncu_sass_mix.py $F/prof_c1x.ncu-rep $((20000000*252)) --blocks > $F/c1x_sass_mix.txt; head -40 $F/c1x_sass_mix.txt
