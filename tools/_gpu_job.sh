python tools/mapper_run.py 100000 2000 30 3 euclidean 3 | tail -2 | grep -o '"gemm_ms": [0-9.]*\|"total_ms": [0-9.]*' | paste - -
python tools/mapper_run.py 20000 5000 30 3 correlation 3 | tail -2 | grep -o '"gemm_ms": [0-9.]*\|"total_ms": [0-9.]*' | paste - -
python -m pytest tests/test_mapper_gpu.py tests/test_end_to_end_gpu.py -m gpu -x -q 2>&1 | tail -3
