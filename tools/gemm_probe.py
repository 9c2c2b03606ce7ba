"""FP64 DMMA contraction alone (profiling helper): out = Xn[rows] Xn[rows]^T for n rows of G features.
python tools/gemm_probe.py [n=8192] [G=5000]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sctda_b200 import _lib, mapper  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
G = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
dev = torch.device('cuda', 0)
ctx = _lib.context(0)
ctx.set_kernel_timing(True)
g = torch.Generator(device=dev)
g.manual_seed(1)
X = torch.randn((n, G), dtype=torch.float64, device=dev, generator=g)
Xn, _ = mapper.row_normalize(ctx, X, 'euclidean')
rows = torch.randperm(n, device=dev, generator=g).to(torch.int32)
for _ in range(3):
    out = mapper.gram_gather(ctx, Xn, rows, rows)
    ms = ctx.last_kernel_ms()
    print('gemm_kernel %d x %d x %d: %.3f ms, %.2f TFLOP/s' % (n, n, G, ms, 2.0 * n * n * G / ms / 1e9))
ref = (X[rows.long()[:64]] @ X[rows.long()[:64]].T)
print('max rel err vs torch on a 64x64 corner: %.2e' % float(((out[:64, :64] - ref).abs() / ref.abs().clamp_min(1e-300)).max()))
