# round-2 c3 measurement pass: parity test of the pair kernel, c3 bench (pair kernel and, for A/B, the generic lean build), ncu
set -x
F=gpurun_out/${1:-r2a}; rm -rf $F; mkdir -p $F
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "joint_pair or edge_sizes or partition" > $F/pytest_pair.log 2>&1; echo "pytest rc $?"; tail -5 $F/pytest_pair.log
python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu > $F/bench_c3.json 2> $F/bench_c3.err; echo "c3 rc $?"; cat $F/bench_c3.json | cut -c1-400
LETF_NO_JPAIR=1 python bench.py --workload c3 --trials 303104 --steps 3 --warmup 3 --no-cpu > $F/bench_c3_generic.json 2> $F/bench_c3_generic.err; echo "c3 generic rc $?"; cut -c1-200 $F/bench_c3_generic.json
python bench.py --workload c3 --trials 303104 --steps 3 --warmup 3 --no-cpu > $F/bench_c3_303k.json 2> $F/bench_c3_303k.err; cut -c1-200 $F/bench_c3_303k.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:letf_jpair_kernel -s 3 -c 1 -f -o $F/prof_c3 python bench.py --workload c3 --trials 113664 --steps 1 --warmup 3 --no-cpu > $F/ncu_c3.log 2>&1; echo "ncu rc $?"
