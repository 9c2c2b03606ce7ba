python tools/lpt_probe.py 8 2>&1 | grep "chain=True"
python tools/mapper_run.py 100000 2000 30 3 euclidean 3 | tail -2 | grep -o '"gemm_ms": [0-9.]*\|"total_ms": [0-9.]*' | paste - -
SCTDA_PIPELINE=0 python tools/lpt_probe.py 8 2>&1 | grep "chain=True: slowest"
