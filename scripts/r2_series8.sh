# record pass of the materialised-path mode: duration and DRAM traffic per launch (ncu), then the series19 line
F=gpurun_out/${1:-r2s8}; rm -rf $F; mkdir -p $F
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:letf_series_records_kernel -c 4 --csv --log-file $F/rec.csv python bench.py --workload series19 --steps 1 --warmup 3 --no-cpu > $F/ncu.log 2>&1
grep letf_series $F/rec.csv | awk -F'","' '{print $(NF-2), $NF}' | tail -3
python bench.py --workload series19 --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('series19', d['value'], d['roofline']['kernel_ms'])"
python bench.py --workload series --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('series', d['value'], d['roofline']['kernel_ms'])"
