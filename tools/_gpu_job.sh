cp sctda_b200/csrc/libsctda_b200.so /tmp/cur.so
cp build/base/libcgsplit.so sctda_b200/csrc/libsctda_b200.so
echo cg; python tools/mapper_run.py 20000 5000 30 3 correlation 3 fp16x3 | tail -1 | grep -o '"gemm_ms": [0-9.]*\|"total_ms": [0-9.]*' | paste - -
python tools/mapper_run.py 100000 2000 30 3 euclidean 3 fp16x3 | tail -1 | grep -o '"gemm_ms": [0-9.]*\|"total_ms": [0-9.]*' | paste - -
cp /tmp/cur.so sctda_b200/csrc/libsctda_b200.so
echo ca; python tools/mapper_run.py 20000 5000 30 3 correlation 3 fp16x3 | tail -1 | grep -o '"gemm_ms": [0-9.]*\|"total_ms": [0-9.]*' | paste - -
python tools/mapper_run.py 100000 2000 30 3 euclidean 3 fp16x3 | tail -1 | grep -o '"gemm_ms": [0-9.]*\|"total_ms": [0-9.]*' | paste - -
python -m pytest tests/test_mapper_gpu.py -m gpu -x -q -k "split" 2>&1 | tail -2
