# ncu capture of the dynamic-slot rule kernel (generic pipeline) on the five-strategy list
F=gpurun_out/r2gncu; rm -rf $F; mkdir -p $F
LETF_NO_ROUTED_SPLIT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:letf_path_kernel -s 3 -c 1 -f -o $F/prof_all8 python bench.py --workload sub5 --trials 75776 --steps 1 --warmup 3 --no-cpu > $F/ncu.log 2>&1; echo rc $?
python scripts/ncu_metrics.py $F/prof_all8.ncu-rep 0 > $F/all8_summary.csv; cat $F/all8_summary.csv
python scripts/ncu_sass_mix.py $F/prof_all8.ncu-rep $((75776*7560)) --blocks > $F/all8_sass_mix.txt; head -30 $F/all8_sass_mix.txt
