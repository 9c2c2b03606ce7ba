set -x
python -m pytest tests -m gpu -q 2>&1 | tail -12
P="python tools/conn_probe.py --genes 1024 --perms 1184 --order pca --reps 2"
ncu --set full --clock-control none --import-source on -k regex:conn_perm_wide -s 1 -c 1 -o gpurun_out/prof_w4 -f $P > gpurun_out/ncu.log 2>&1
