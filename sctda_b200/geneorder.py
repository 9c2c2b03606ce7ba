"""Tile layout of the shared-permutation kernel: an order of the genes that puts genes with
similar supports (sets of expressing cells) next to each other, so that more 32-byte sectors of
the transposed table hold only zeros (csrc/conn_tile64.cu never requests those).  The order
changes nothing that is computed; results are written at the original gene positions.

Not part of the reference (its per-gene loop, /root/reference/scTDA/main.py:1068-1071, has no
data layout to speak of); a planning step like the CSR construction in graph.DeviceGraph.
"""
import numpy
import torch


def gene_order(Y, method='pca', sample=2048, dims=8, seed=0):
    """Y [G, cells] device table.  Returns an int32 device permutation of the genes.
    'density': ascending fraction of non-zero cells.
    'pca': recursive principal-direction bisection of the binary supports (sketched on `sample`
    cells, embedded in the top `dims` principal components), leaves of four genes."""
    G = int(Y.shape[0])
    dev = Y.device
    if method == 'density':
        return torch.argsort((Y != 0).sum(dim=1)).to(torch.int32).contiguous()
    if method != 'pca':
        raise ValueError(method)
    gen = torch.Generator(device='cpu')
    gen.manual_seed(seed)
    ncell = int(Y.shape[1])
    cols = torch.randperm(ncell, generator=gen)[:min(sample, ncell)].to(dev)
    B = (Y.index_select(1, cols) != 0).to(torch.float32)
    B = B / B.norm(dim=1, keepdim=True).clamp_min(1e-9)
    # top principal directions of the sketch through its small Gram matrix
    w, v = torch.linalg.eigh(B.T @ B)
    E = (B @ v[:, -dims:]).cpu().numpy().astype(numpy.float64)
    order = numpy.arange(G)
    segs = [(0, G)]
    while segs:
        nxt = []
        for lo, hi in segs:
            if hi - lo <= 4:
                continue
            idx = order[lo:hi]
            X = E[idx] - E[idx].mean(axis=0)
            _, vec = numpy.linalg.eigh(X.T @ X)
            o = numpy.argsort(X @ vec[:, -1], kind='stable')
            order[lo:hi] = idx[o]
            half = ((hi - lo) // 2 + 3) // 4 * 4
            half = max(4, min(half, hi - lo - 1))
            nxt.append((lo, lo + half))
            nxt.append((lo + half, hi))
        segs = nxt
    return torch.from_numpy(order.astype(numpy.int32)).to(dev)
