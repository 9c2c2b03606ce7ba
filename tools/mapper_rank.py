"""What one rank of an 8-GPU Mapper build does, on one GPU (profiling helper): the LPT patch shard of
`rank` out of `world` for the 100k x 2k configuration, clustered by sctda_bin_cluster.
python tools/mapper_rank.py [rank=0] [world=8]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy  # noqa: E402
import torch  # noqa: E402

from sctda_b200 import _lib, mapper, parallel, synth  # noqa: E402

rank = int(sys.argv[1]) if len(sys.argv) > 1 else 0
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
N, G, r, gain = 100000, 2000, 30, 3.0
dev = torch.device('cuda', 0)
ctx = _lib.context(0)
ctx.set_kernel_timing(True)
X = synth.expression_genes_torch(G, N, seed=synth.SEED + 11, device=dev).T.contiguous()
lens = mapper.pca_lens_device(X)
res = mapper.build_graph_device(ctx, X, lens, r, gain, metric='euclidean')
bin_off = res.bin_off.cpu().numpy().astype(numpy.int64)
members = res.members.cpu().numpy()
parts = parallel.assign_patches_lpt(bin_off[1:] - bin_off[:-1], world)
so, sm = parallel.sub_csr(bin_off, members, parts[rank])
Xn, _ = mapper.row_normalize(ctx, X, 'euclidean')
sizes = so[1:] - so[:-1]
print('rank %d of %d: %d patches, largest %d, sum m^2 = %.3g' % (rank, world, len(sizes), sizes.max(), float((sizes * sizes).sum())))
for _ in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mapper.cluster_sub_csr(ctx, Xn, so, sm)
    torch.cuda.synchronize()
    print('cluster_sub_csr %.1f ms (gemm %.1f ms)' % (1e3 * (time.perf_counter() - t0), ctx.last_kernel_ms()))
