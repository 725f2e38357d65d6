for i in 1 2 3; do for w in c3 boot; do python bench.py --workload $w --trials 303104 --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w value %.4e'%d['value'])"; done; done
