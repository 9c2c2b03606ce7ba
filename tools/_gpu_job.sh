P="python tools/conn_probe.py --genes 8192 --perms 1184 --order pca --reps 2 --scores"
cp sctda_b200/csrc/libsctda_b200.so /tmp/cur.so
cp build/base/libnorowpf.so sctda_b200/csrc/libsctda_b200.so
$P --save /tmp/base.npz
cp /tmp/cur.so sctda_b200/csrc/libsctda_b200.so
$P --check /tmp/base.npz
