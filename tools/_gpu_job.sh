set -x
CMD="python tools/conn_probe.py --genes 4096 --perms 128 --order pca --reps 2"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conn_perm64 -s 1 -c 1 -o gpurun_out/prof_t64c -f $CMD > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/plain.log
