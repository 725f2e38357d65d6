for w in c3 c4; do python bench.py --workload $w --trials 303104 --steps 2 --warmup 3 --no-cpu 2>/dev/null | cut -c1-140; done
