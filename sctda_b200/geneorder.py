"""Tile layout of the shared-permutation kernel: an order of the genes that puts genes with
similar supports (sets of expressing cells) next to each other, so that more 32-byte sectors of
the transposed table hold only zeros (csrc/conn_tile64.cu never requests those) and more whole
rows of a 64-gene tile are empty (dropped before the gather).  The order changes nothing that is
computed: results are written at the original gene positions.

Not part of the reference (its per-gene loop, /root/reference/scTDA/main.py:1068-1071, has no
data layout to speak of); a planning step like the CSR construction in graph.DeviceGraph, made
once per table.  torch is used as array plumbing on the device (a sketch, small batched
eigen-decompositions, sorts); nothing here is on the timed path of the permutation test.
"""
import torch


def gene_order(Y, method='pca', sample=2048, dims=8, seed=0, leaf=4):
    """Y [G, cells] device table (float64).  Returns an int32 device permutation of the genes.

    'density': ascending number of non-zero cells.
    'pca': recursive principal-direction bisection of the binary supports: the supports are sketched
    on `sample` random cells and embedded in the top `dims` principal components of the sketch;
    every segment of genes is then sorted along its own principal direction and cut in two
    (multiples of `leaf` genes, the genes of one 32-byte sector) until segments hold `leaf` genes."""
    G = int(Y.shape[0])
    dev = Y.device
    if G <= leaf or method == 'none':
        return torch.arange(G, dtype=torch.int32, device=dev)
    if method == 'density':
        return torch.argsort((Y != 0).sum(dim=1), stable=True).to(torch.int32).contiguous()
    if method != 'pca':
        raise ValueError(method)
    gen = torch.Generator(device='cpu')
    gen.manual_seed(seed)
    ncell = int(Y.shape[1])
    cols = torch.randperm(ncell, generator=gen)[:min(sample, ncell)].to(dev)
    B = (Y.index_select(1, cols) != 0).to(torch.float32)
    B = B / B.norm(dim=1, keepdim=True).clamp_min(1e-9)
    dims = min(dims, int(B.shape[1]))
    _, v = torch.linalg.eigh((B.T @ B).double())
    E = (B.double() @ v[:, -dims:])                                   # [G, dims] embedding of the supports
    order = torch.arange(G, device=dev)
    starts = torch.zeros(1, dtype=torch.int64, device=dev)
    sizes = torch.full((1,), G, dtype=torch.int64, device=dev)
    while int(sizes.max().item()) > leaf:
        nseg = int(sizes.numel())
        seg = torch.repeat_interleave(torch.arange(nseg, device=dev), sizes)
        X = E.index_select(0, order)
        mean = torch.zeros((nseg, dims), dtype=X.dtype, device=dev).index_add_(0, seg, X) / sizes.unsqueeze(1)
        Xc = X - mean.index_select(0, seg)
        cov = torch.zeros((nseg, dims, dims), dtype=X.dtype, device=dev).index_add_(0, seg, Xc.unsqueeze(2) * Xc.unsqueeze(1))
        vec = torch.linalg.eigh(cov)[1][:, :, -1]                     # principal direction of every segment
        proj = (Xc * vec.index_select(0, seg)).sum(dim=1)
        idx = torch.argsort(proj, stable=True)
        idx = idx.index_select(0, torch.argsort(seg.index_select(0, idx), stable=True))    # segments stay in place
        order = order.index_select(0, idx)
        # cut every segment above `leaf` genes in two multiples of `leaf`
        big = sizes > leaf
        half = torch.div(torch.div(sizes, 2, rounding_mode='floor') + (leaf - 1), leaf, rounding_mode='floor') * leaf
        half = torch.minimum(torch.clamp(half, min=leaf), sizes - 1)
        left = torch.where(big, half, sizes)
        right = torch.where(big, sizes - half, torch.zeros_like(sizes))
        ns = torch.stack([starts, starts + left], dim=1).reshape(-1)
        nz = torch.stack([left, right], dim=1).reshape(-1)
        keep = nz > 0
        starts, sizes = ns[keep], nz[keep]
    return order.to(torch.int32).contiguous()


def sector_density(Y, order=None, group=4):
    """Fraction of (group of `group` consecutive genes, cell) pairs that hold a non-zero: the share of
    32-byte sectors the kernel has to fetch for this layout."""
    nz = (Y != 0)
    if order is not None:
        nz = nz.index_select(0, order.long())
    G = (int(nz.shape[0]) // group) * group
    if G == 0:
        return 1.0
    return float(nz[:G].view(G // group, group, -1).any(dim=1).float().mean().item())
