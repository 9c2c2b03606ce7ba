set -x
python -m pytest tests/test_conn_gpu.py -m gpu -x -q 2>&1 | tail -5
P="python tools/conn_probe.py --genes 2048 --perms 256"
SCTDA_CONN_LEGACY=1 $P --scores --save gpurun_out/legacy.npz
$P --scores --check gpurun_out/legacy.npz
$P --order pca --scores --check gpurun_out/legacy.npz
SCTDA_CONN64_WARPS=24 $P --order pca --check gpurun_out/legacy.npz
SCTDA_CONN64_WARPS=16 $P --order pca --check gpurun_out/legacy.npz
