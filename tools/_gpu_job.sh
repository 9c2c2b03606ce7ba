set -x
P="python tools/conn_probe.py --genes 4096 --perms 1184 --order pca --reps 2"
export SCTDA_CONNW_PF=2
cp sctda_b200/csrc/libsctda_b200.so /tmp/cur.so
cp build/base/libg6.so sctda_b200/csrc/libsctda_b200.so
SCTDA_CONNW_WARPS=20 $P
SCTDA_CONNW_WARPS=16 $P
cp build/base/libg8.so sctda_b200/csrc/libsctda_b200.so
SCTDA_CONNW_WARPS=16 $P
SCTDA_CONNW_WARPS=20 $P
cp /tmp/cur.so sctda_b200/csrc/libsctda_b200.so
