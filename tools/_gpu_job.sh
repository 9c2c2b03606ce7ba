set -x
P="python tools/conn_probe.py --genes 2048 --perms 256"
SCTDA_CONN_LEGACY=1 $P --scores --save gpurun_out/legacy.npz
$P --order pca --scores --check gpurun_out/legacy.npz
$P --scores --check gpurun_out/legacy.npz
