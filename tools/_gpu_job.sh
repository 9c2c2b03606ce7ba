set -x
python bench.py --steps 3 --warmup 2 --no-mapper --no-cpu > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -3 gpurun_out/bench_full.err
