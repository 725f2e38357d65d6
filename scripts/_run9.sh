python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in c2 c3 c4 sweep19 boot; do t=303104; [ $w = c2 ] && t=100000; python bench.py --workload $w --trials $t --steps 2 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w value %.4e'%d['value'], 'e2e %.3e'%d['e2e']['value'], d['config']['kernel'])"; done
python bench.py --workload series19 --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('series19 value %.4e'%d['value'], d['roofline']['achieved'], d['config']['kernel'])"
