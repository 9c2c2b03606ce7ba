"""The patch dealer's cost model on one GPU: every rank's shard of the 8-GPU 100k x 2k build is clustered in turn
(sctda_bin_cluster on the rank's sub-CSR), with and without the MST-chain term of parallel.patch_cost.
python tools/lpt_probe.py [world=8]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy  # noqa: E402
import torch  # noqa: E402

from sctda_b200 import _lib, mapper, parallel, synth  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N, G, r, gain = 100000, 2000, 30, 3.0
dev = torch.device('cuda', 0)
ctx = _lib.context(0)
ctx.set_kernel_timing(True)
X = synth.expression_genes_torch(G, N, seed=synth.SEED + 11, device=dev).T.contiguous()
lens = mapper.pca_lens_device(X)
res = mapper.build_graph_device(ctx, X, lens, r, gain, metric='euclidean')
bin_off = res.bin_off.cpu().numpy().astype(numpy.int64)
members = res.members.cpu().numpy()
sizes = bin_off[1:] - bin_off[:-1]
print('patches above %d cells:' % parallel.CHAIN_MIN, sorted(sizes[sizes > parallel.CHAIN_MIN].tolist(), reverse=True))
Xn, _ = mapper.row_normalize(ctx, X, 'euclidean')
for chain in (False, True):
    parts = parallel.assign_patches_lpt(sizes, world, chain=chain)
    worst = 0.0
    for rank in range(world):
        so, sm = parallel.sub_csr(bin_off, members, parts[rank])
        ts = []
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            mapper.cluster_sub_csr(ctx, Xn, so, sm)
            torch.cuda.synchronize()
            ts.append(1e3 * (time.perf_counter() - t0))
        s = so[1:] - so[:-1]
        print('chain=%s rank %d: %d patches, largest %d, sum m^2 %.3g, %.1f ms (gemm %.1f)'
              % (chain, rank, len(s), s.max(), float((s * s).sum()), min(ts), ctx.last_kernel_ms()))
        worst = max(worst, min(ts))
    print('chain=%s: slowest rank %.1f ms' % (chain, worst))
