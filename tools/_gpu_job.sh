set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_n8d.json 2> gpurun_out/bench_n8d.err; tail -c 300 gpurun_out/bench_n8d.json; tail -3 gpurun_out/bench_n8d.err
