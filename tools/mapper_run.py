"""Runs only the Mapper build at a given size (profiling helper): python tools/mapper_run.py N G r gain metric [reps] [fp64|fp16x3]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sctda_b200 import _lib, mapper, synth  # noqa: E402

N, G, r, gain, metric = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), sys.argv[5]
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 2
dev = torch.device('cuda', 0)
ctx = _lib.context(0)
ctx.set_kernel_timing(True)
if len(sys.argv) > 7:
    ctx.set_precision(sys.argv[7])
X = synth.expression_genes_torch(G, N, seed=synth.SEED + 11, device=dev).T.contiguous()
lens = mapper.pca_lens_device(X)
for _ in range(reps):
    prof = {}
    res = mapper.build_graph_device(ctx, X, lens, r, gain, metric=metric, profile=prof)
    print(json.dumps(prof))
