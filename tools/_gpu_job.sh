python -m pytest tests/test_conn_gpu.py -m gpu -x -q -k "sharded or skeleton" 2>&1 | tail -15
