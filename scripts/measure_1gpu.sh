# round-2 end-of-round measurement set on one B200 (all from one gpurun call): gpu tests, smoke, the default bench line
# (c3) + reference arm, the other workloads at their literal sizes, the launch list of the default command and one
# `ncu --set full` capture per kernel of the benched shapes.  Collected by scripts/collect_profiles.py <dir> <tag> r2 ...
set -x
F=gpurun_out/${1:-final_r2}; rm -rf $F; mkdir -p $F
python -m pytest tests -m gpu -x -q -s > $F/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -2 $F/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $F/smoke.log 2>&1; echo "smoke rc $?"; tail -2 $F/smoke.log
python bench.py > $F/bench_default.json 2> $F/bench_default.err; echo "default rc $?"
python bench.py --impl reference > $F/bench_reference.json 2> $F/bench_reference.err; echo "ref rc $?"
for w in c1 c1x c2; do python bench.py --workload $w --steps 5 --warmup 3 > $F/bench_${w}.json 2> $F/bench_${w}.err; echo "$w rc $?"; done
for w in c4 sweep19 sub5 sma2 boot; do python bench.py --workload $w --steps 3 --warmup 3 --no-cpu > $F/bench_${w}.json 2> $F/bench_${w}.err; echo "$w rc $?"; done
python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu > $F/bench_c5.json 2> $F/bench_c5.err; echo "c5 rc $?"
for w in series series19; do python bench.py --workload $w --steps 3 --warmup 3 > $F/bench_${w}.json 2> $F/bench_${w}.err; echo "$w rc $?"; done
LETF_NO_ROUTED_SPLIT=1 python bench.py --workload sub5 --steps 3 --warmup 3 --no-cpu > $F/bench_sub5generic.json 2> $F/bench_sub5generic.err; echo "sub5 generic pipeline rc $?"
LETF_NO_GENERIC_SPLIT=1 python bench.py --workload sub5 --steps 3 --warmup 3 --no-cpu > $F/bench_sub5fused.json 2> $F/bench_sub5fused.err; echo "sub5 fused rc $?"
LETF_NO_SPLIT=1 python bench.py --workload series19 --steps 3 --warmup 3 --no-cpu > $F/bench_series19fused.json 2> $F/bench_series19fused.err; echo "series19 fused rc $?"
LETF_BENCH_JOINT_NU=5.3,4.2 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu > $F/bench_c3nu.json 2> $F/bench_c3nu.err; echo "c3 calibrated nu rc $?"
LETF_BENCH_MODEL=vix_te python bench.py --workload sweep19 --steps 3 --warmup 3 --no-cpu > $F/bench_sweep19te.json 2> $F/bench_sweep19te.err; echo "sweep19 calibrated te rc $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $F/launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu > $F/ncu_launches.log 2>&1; echo "launch list rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:letf_jpair_kernel -s 3 -c 1 -f -o $F/prof_c3 python bench.py --workload c3 --trials 113664 --steps 1 --warmup 3 --no-cpu > $F/ncu_c3.log 2>&1; echo "ncu c3 rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:letf_path_kernel -s 3 -c 1 -f -o $F/prof_c2 python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu > $F/ncu_c2.log 2>&1; echo "ncu c2 rc $?"
for w in sweep19 c4; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"letf_path_kernel|letf_jpair_kernel" -s 6 -c 2 -f -o $F/prof_${w} python bench.py --workload $w --trials 113664 --steps 1 --warmup 3 --no-cpu > $F/ncu_${w}.log 2>&1; echo "ncu $w rc $?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"letf_series_records_kernel|letf_path_kernel" -s 6 -c 2 -f -o $F/prof_series19 python bench.py --workload series19 --steps 1 --warmup 3 --no-cpu > $F/ncu_series19.log 2>&1; echo "ncu series19 rc $?"
./scripts/pipe_peaks $F/pipe_peaks.json > $F/pipe_peaks.txt 2>&1; echo "pipe rc $?"
F=$F python - <<'PY'
import json,glob,os
for f in sorted(glob.glob(os.environ['F']+'/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), '%.3e'%d['value'], 'ms %.2f'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], 'frac', (d.get('roofline') or {}).get('frac'), (d.get('config') or {}).get('kernel'))
    except Exception as e: print(f, 'ERR', e)
PY
