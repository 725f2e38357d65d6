set -x
for w in ${WL:-c3}; do
  n=75776; [ $w = c4 ] && n=50560
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:letf_path_kernel -s 3 -c 1 -f -o gpurun_out/prof_${w}_${TAG} python bench.py --workload $w --trials $n --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_${w}_${TAG}.log 2>&1
  echo "ncu $w rc $?"
done
