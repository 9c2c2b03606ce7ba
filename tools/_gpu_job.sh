M="python tools/mapper_run.py 100000 2000 30 3 euclidean 2"
cp sctda_b200/csrc/libsctda_b200.so /tmp/cur.so
$M | tail -1 | cut -c1-60
for v in c4s2 c3s2 c2s4; do cp build/base/lib$v.so sctda_b200/csrc/libsctda_b200.so; echo $v; $M | tail -1 | cut -c1-60; done
cp /tmp/cur.so sctda_b200/csrc/libsctda_b200.so
