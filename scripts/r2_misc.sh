set -x
F=gpurun_out/${1:-r2j}; rm -rf $F; mkdir -p $F
python scripts/e2e_breakdown.py c3 > $F/e2e_c3.txt 2>&1; cat $F/e2e_c3.txt
python scripts/e2e_breakdown.py c2 > $F/e2e_c2.txt 2>&1; cat $F/e2e_c2.txt
for w in series19 series; do python bench.py --workload $w --steps 3 --warmup 3 > $F/bench_${w}.json 2> $F/bench_${w}.err; echo "$w rc $?"; done
TRIALS=1000000 VARIANTS="c3:- c3:f32_joint_t_pair_c3_b64" F=$F/ab bash scripts/r2_ab.sh $(basename $F)/ab
TRIALS=1250000 VARIANTS="c4:- c4:f32_joint_t_prod_c4_b64 c4:f32_joint_t_prod_c4_r128" F=$F/ab4 bash scripts/r2_ab.sh $(basename $F)/ab4
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob(os.environ['F']+'/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), '%.4e'%d['value'], 'ms %.2f'%d['ms_per_step'], d['roofline'].get('achieved'), d['roofline'].get('frac'))
    except Exception as e: print(f, 'ERR', e)
PY
