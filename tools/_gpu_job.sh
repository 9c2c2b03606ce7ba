python tools/gemm_probe.py 8192 2000 | sed -n 3p
python tools/mapper_run.py 100000 2000 30 3 euclidean 3 | tail -2 | grep -o '"gemm_ms": [0-9.]*\|"total_ms": [0-9.]*' | paste - -
python tools/mapper_run.py 20000 5000 30 3 correlation 3 | tail -1 | grep -o '"gemm_ms": [0-9.]*\|"total_ms": [0-9.]*' | paste - -
python tools/lpt_probe.py 8 2>&1 | grep "chain=True"
python -m pytest tests/test_mapper_gpu.py -m gpu -x -q 2>&1 | tail -2
