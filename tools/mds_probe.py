"""MDS lens at scale (profiling helper): distance matrix + SMACOF on the device.
python tools/mds_probe.py [N=20000] [G=5000] [metric=correlation] [max_iter=300]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sctda_b200 import _lib, mapper, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
metric = sys.argv[3] if len(sys.argv) > 3 else 'correlation'
max_iter = int(sys.argv[4]) if len(sys.argv) > 4 else 300
dev = torch.device('cuda', 0)
ctx = _lib.context(0)
ctx.set_kernel_timing(True)
X = synth.expression_genes_torch(G, N, seed=synth.SEED + 11, device=dev).T.contiguous()
torch.cuda.synchronize()
t0 = time.perf_counter()
D = mapper.pairwise_distances(ctx, X, metric)
torch.cuda.synchronize()
t1 = time.perf_counter()
print('pairwise %s distances %d x %d x %d: %.1f ms, %.2f TFLOP/s' % (metric, N, N, G, 1e3 * (t1 - t0), 2.0 * N * N * G / (t1 - t0) / 1e12))
t_a = time.perf_counter()
pos, stress, it = mapper.smacof_device(ctx, D, max_iter=5, random_state=0, return_n_iter=True)
torch.cuda.synchronize()
print('5-iteration run: %.1f ms' % (1e3 * (time.perf_counter() - t_a)))
t1 = time.perf_counter()
pos, stress, it = mapper.smacof_device(ctx, D, max_iter=max_iter, random_state=0, return_n_iter=True)
torch.cuda.synchronize()
t2 = time.perf_counter()
ms = ctx.last_kernel_ms()
print('SMACOF: %d iterations, stress %.6g, %.1f ms total, last pass kernel %.3f ms = %.0f GB/s of dissimilarities'
      % (it, stress, 1e3 * (t2 - t1), ms, 8.0 * N * N / ms / 1e6))
