M="python tools/mapper_run.py 100000 2000 30 3 euclidean 1"
SCTDA_PIPELINE=0 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 2 -c 1 -o gpurun_out/prof_gemm_r02 -f $M > gpurun_out/ncu4.log 2>&1
tail -2 gpurun_out/ncu4.log | cut -c1-200
