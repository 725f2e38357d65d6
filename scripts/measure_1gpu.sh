set -x
rm -rf gpurun_out/final2; mkdir -p gpurun_out/final2
F=gpurun_out/final2
python -m pytest tests -m gpu -x -q > $F/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -2 $F/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $F/smoke.log 2>&1; echo "smoke rc $?"; tail -1 $F/smoke.log
python bench.py > $F/bench_default.json 2> $F/bench_default.err; echo "default rc $?"
python bench.py --impl reference > $F/bench_reference.json 2> $F/bench_reference.err; echo "ref rc $?"
for w in c3 c4 sweep19 boot; do python bench.py --workload $w --trials 303104 --steps 3 --warmup 3 --no-cpu > $F/bench_${w}.json 2> $F/bench_${w}.err; echo "$w rc $?"; done
python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu > $F/bench_c5.json 2> $F/bench_c5.err; echo "c5 rc $?"
for w in series series19; do python bench.py --workload $w --steps 3 --warmup 3 > $F/bench_${w}.json 2> $F/bench_${w}.err; echo "$w rc $?"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $F/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu > $F/ncu_launches.log 2>&1; echo "launch list rc $?"
for w in c2 c3 c4 boot sweep19; do
  n=75776; [ $w = c4 ] && n=50560; [ $w = boot ] && n=50560; [ $w = c2 ] && n=100000; [ $w = sweep19 ] && n=37888
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:letf_path_kernel -s 3 -c 1 -f -o $F/prof_${w} python bench.py --workload $w --trials $n --steps 1 --warmup 3 --no-cpu > $F/ncu_${w}.log 2>&1
  echo "ncu $w rc $?"
done
