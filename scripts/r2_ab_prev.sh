# same-box A/B of the working tree against a build of the previous commit in ./_ab_prev ("workload[:trials]" list in $WL)
for rep in 1 2; do
for d in . _ab_prev; do
  for wt in ${WL:-c4:227328 sweep19:227328}; do w=${wt%%:*}; t=${wt##*:}; [ "$t" = "$w" ] && t=0
    (cd $d; python bench.py --workload $w $( [ $t != 0 ] && echo --trials $t ) --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$d', '$w', d['value'], d['roofline']['kernel_ms'])"); done
done
done
