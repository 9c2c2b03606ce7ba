set -x
python -m pytest tests/test_conn_gpu.py tests/test_end_to_end_gpu.py -m gpu -x -q 2>&1 | tail -4
P="python tools/conn_probe.py --genes 4096 --perms 1184 --order pca --reps 3 --scores"
cp build/base/libbase.so sctda_b200/csrc/libsctda_b200.so
$P --save /tmp/base.npz
cp build/base/libnew.so sctda_b200/csrc/libsctda_b200.so
$P --check /tmp/base.npz
SCTDA_CONNW_WARPS=20 $P
SCTDA_CONNW_WARPS=28 $P
P="python tools/conn_probe.py --genes 1024 --perms 1184 --order pca --reps 2"
ncu --set full --clock-control none --import-source on -k regex:conn_perm_wide -s 1 -c 1 -o gpurun_out/prof_w6 -f $P > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
