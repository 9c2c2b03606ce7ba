# the job tools/gpu.sh sends by default: GPU parity suite, smoke, one bench line
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python __graft_entry__.py --smoke 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 400 gpurun_out/bench.json
