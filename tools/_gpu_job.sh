set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_final1.json 2> gpurun_out/bench_final1.err; tail -c 300 gpurun_out/bench_final1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json
B="python bench.py --steps 1 --warmup 1 --perms 500 --no-cpu --no-rooted"
ncu --metrics gpu__time_duration.sum --clock-control none -k "$(cat tools/own_kernels.regex)" --csv --log-file gpurun_out/launches_r02.csv $B > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log | cut -c1-300
python __graft_entry__.py --smoke 2>&1 | tail -3
