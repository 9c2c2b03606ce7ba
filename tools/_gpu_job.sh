set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_s2.json 2> gpurun_out/bench_s2.err; tail -c 2500 gpurun_out/bench_s2.json; tail -5 gpurun_out/bench_s2.err
