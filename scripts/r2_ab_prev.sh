for rep in 1 2; do
for d in . _ab_prev; do
  for w in c4 sweep19; do (cd $d; python bench.py --workload $w --trials 227328 --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$d', '$w', d['value'], d['roofline']['kernel_ms'])"); done
done
done
