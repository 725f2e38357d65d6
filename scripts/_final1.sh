set -x
mkdir -p gpurun_out/final
python bench.py > gpurun_out/final/bench_default.json 2> gpurun_out/final/bench_default.err; echo "default rc $?"
python bench.py --impl reference > gpurun_out/final/bench_reference.json 2> gpurun_out/final/bench_reference.err; echo "ref rc $?"
for w in c3 c4 sweep19 boot; do python bench.py --workload $w --trials 303104 --steps 2 --warmup 3 --no-cpu > gpurun_out/final/bench_${w}.json 2> gpurun_out/final/bench_${w}.err; echo "$w rc $?"; done
python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu > gpurun_out/final/bench_c5.json 2> gpurun_out/final/bench_c5.err; echo "c5 rc $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/final/ncu_launches.log 2>&1; echo "launch list rc $?"
for w in c2 c3 c4 boot; do
  n=75776; [ $w = c4 ] && n=50560; [ $w = boot ] && n=56832; [ $w = c2 ] && n=100000
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:letf_path_kernel -s 3 -c 1 -f -o gpurun_out/final/prof_${w} python bench.py --workload $w --trials $n --steps 1 --warmup 3 --no-cpu > gpurun_out/final/ncu_${w}.log 2>&1
  echo "ncu $w rc $?"
done
cat gpurun_out/final/bench_default.json | cut -c1-400
