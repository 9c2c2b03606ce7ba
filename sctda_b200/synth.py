"""Synthetic inputs for the hot path.

Nothing here is reference code: the reference ships no data generator.  The
generator follows the recipe SURVEY.md §8(d) fixes so that every bench / test
run sees the same cells x genes matrix for a given (seed, shape):

    latent pseudo-time t ~ U(0,1), branch b in {0,1,2};
    per-gene mean mu_g(t,b) = A_g * exp(-(t-tau_g)^2 / (2 sigma_g^2)) * [b in mask_g]
    counts ~ NegBin(mean mu, dispersion 0.5); values = log2(1 + TPM)

(the log2(1+TPM) transform is what Preprocess.save writes,
/root/reference/scTDA/main.py:692,726).
"""
import numpy

SEED = 20170501


def expression_matrix(n_cells, n_genes, seed=SEED, dtype=numpy.float64):
    """cells x genes matrix of log2(1+TPM) values (C-contiguous float64) and
    the latent (t, branch) per cell."""
    rng = numpy.random.RandomState(seed)
    t = rng.uniform(0.0, 1.0, n_cells)
    branch = rng.randint(0, 3, n_cells)
    # after t > 0.5 the three branches diverge; before that all cells share branch 0
    branch = numpy.where(t > 0.5, branch, 0)
    amp = rng.lognormal(3.0, 1.0, n_genes)
    tau = rng.uniform(0.0, 1.0, n_genes)
    sig = rng.uniform(0.05, 0.3, n_genes)
    mask = rng.randint(1, 8, n_genes)  # bit b set => gene active on branch b
    out = numpy.empty((n_cells, n_genes), dtype=dtype)
    disp = 0.5
    chunk = max(1, (1 << 24) // max(n_genes, 1))
    for lo in range(0, n_cells, chunk):
        hi = min(n_cells, lo + chunk)
        tt = t[lo:hi, None]
        active = ((mask[None, :] >> branch[lo:hi, None]) & 1).astype(numpy.float64)
        mu = amp[None, :] * numpy.exp(-(tt - tau[None, :]) ** 2 / (2.0 * sig[None, :] ** 2)) * active
        # sparsify: scRNA-seq dropout (60-80 % zeros overall)
        mu *= 0.15
        lam = rng.gamma(shape=1.0 / disp, scale=mu * disp + 1e-300)
        counts = rng.poisson(lam).astype(numpy.float64)
        tot = counts.sum(axis=1, keepdims=True)
        tot[tot == 0.0] = 1.0
        out[lo:hi] = numpy.log2(1.0 + 1e6 * counts / tot)
    return out, t, branch


def cover_graph(n_cells, n_nodes, memberships_per_cell=4, chords=None, seed=SEED):
    """Synthetic Mapper-like graph for the connectivity benchmark:
    every cell lies in exactly `memberships_per_cell` distinct nodes, node sizes are
    equal up to rounding, adjacency = ring lattice of degree 8 plus `chords` random
    chords.  Returns (members: list of int arrays per node, edges: (E,2) int array, i<j)."""
    rng = numpy.random.RandomState(seed + 1)
    c = memberships_per_cell
    node_of = numpy.empty((n_cells, c), dtype=numpy.int64)
    # c independent random balanced assignments; repair the (rare) duplicate node per cell
    base = numpy.arange(n_cells) % n_nodes
    for k in range(c):
        node_of[:, k] = base[rng.permutation(n_cells)]
    for _ in range(64):
        srt = numpy.sort(node_of, axis=1)
        dup = numpy.nonzero((srt[:, 1:] == srt[:, :-1]).any(axis=1))[0]
        if dup.size == 0:
            break
        node_of[dup, rng.randint(0, c, dup.size)] = rng.randint(0, n_nodes, dup.size)
    order = numpy.argsort(node_of.ravel(), kind='stable')
    cells = numpy.repeat(numpy.arange(n_cells), c)[order]
    counts = numpy.bincount(node_of.ravel(), minlength=n_nodes)
    offs = numpy.concatenate([[0], numpy.cumsum(counts)])
    members = [numpy.sort(cells[offs[k]:offs[k + 1]]) for k in range(n_nodes)]
    es = set()
    for d in (1, 2, 3, 4):
        for i in range(n_nodes):
            j = (i + d) % n_nodes
            if i != j:
                es.add((min(i, j), max(i, j)))
    if chords is None:
        chords = n_nodes
    while chords > 0 and n_nodes > 2:
        i, j = rng.randint(0, n_nodes, 2)
        if i != j and (min(i, j), max(i, j)) not in es:
            es.add((min(i, j), max(i, j)))
            chords -= 1
    edges = numpy.array(sorted(es), dtype=numpy.int64).reshape(-1, 2)
    return members, edges


def connectivity_case(n_cells, n_nodes, n_genes, memberships_per_cell=4, seed=SEED, dtype=numpy.float64):
    """Inputs of the connectivity permutation test at a given scale:
    returns (members, adj_csr, Y) with Y [n_genes, n_cells] gene-major log2(1+TPM) values,
    every cell a member of the (connected) graph, adjacency weights 1.0."""
    import scipy.sparse
    members, edges = cover_graph(n_cells, n_nodes, memberships_per_cell=memberships_per_cell, seed=seed)
    i = numpy.concatenate([edges[:, 0], edges[:, 1]])
    j = numpy.concatenate([edges[:, 1], edges[:, 0]])
    adj = scipy.sparse.csr_matrix((numpy.ones(i.size), (i, j)), shape=(n_nodes, n_nodes))
    X, _, _ = expression_matrix(n_cells, n_genes, seed=seed, dtype=dtype)
    return members, adj, numpy.ascontiguousarray(X.T)


def _gene_params(rng, n_genes):
    amp = rng.lognormal(3.0, 1.0, n_genes)
    tau = rng.uniform(0.0, 1.0, n_genes)
    sig = rng.uniform(0.05, 0.3, n_genes)
    mask = rng.randint(1, 8, n_genes)
    return amp, tau, sig, mask


def _cell_latents(n_cells, seed):
    rng = numpy.random.RandomState(seed)
    t = rng.uniform(0.0, 1.0, n_cells)
    branch = numpy.where(t > 0.5, rng.randint(0, 3, n_cells), 0)
    return t, branch


def expression_genes_numpy(n_genes, n_cells, seed=SEED, gene_offset=0):
    """Gene-major [n_genes, n_cells] float64 table of the same trajectory model as
    expression_matrix, generated gene block by gene block (no per-cell TPM normalisation:
    values = log2(1 + 50 * count)), 60-80 % zeros.  Host generator for the CPU arm."""
    t, branch = _cell_latents(n_cells, seed)
    out = numpy.empty((n_genes, n_cells), dtype=numpy.float64)
    blk = 256
    for lo in range(0, n_genes, blk):
        hi = min(n_genes, lo + blk)
        rng = numpy.random.RandomState((seed + 7919 * ((gene_offset + lo) // blk + 1)) % (2 ** 31))
        amp, tau, sig, mask = _gene_params(rng, hi - lo)
        active = ((mask[:, None] >> branch[None, :]) & 1).astype(numpy.float64)
        mu = 0.15 * amp[:, None] * numpy.exp(-(t[None, :] - tau[:, None]) ** 2 / (2.0 * sig[:, None] ** 2)) * active
        lam = rng.gamma(shape=2.0, scale=0.5 * mu + 1e-300)
        out[lo:hi] = numpy.log2(1.0 + 50.0 * rng.poisson(lam))
    return out


def expression_genes_torch(n_genes, n_cells, seed=SEED, device='cuda', gene_offset=0):
    """Device generator of the same model (bench.py): gene-major float64 [n_genes, n_cells].
    Gene g's parameters depend only on (seed, gene_offset + g), so shards of a gene range are
    consistent whatever the number of ranks; the Poisson draws come from torch's generator."""
    import torch
    t, branch = _cell_latents(n_cells, seed)
    dev = torch.device(device)
    td = torch.from_numpy(t).to(dev)
    bd = torch.from_numpy(branch.astype(numpy.int64)).to(dev)
    out = torch.empty((n_genes, n_cells), dtype=torch.float64, device=dev)
    gen = torch.Generator(device=dev)
    blk = 256
    lo = 0
    while lo < n_genes:
        g0 = gene_offset + lo
        hi = min(n_genes, lo + blk - (g0 % blk))
        bidx = g0 // blk
        rng = numpy.random.RandomState((seed + 7919 * (bidx + 1)) % (2 ** 31))
        amp, tau, sig, mask = _gene_params(rng, blk)
        sl = slice(g0 % blk, g0 % blk + (hi - lo))
        amp, tau, sig = [torch.from_numpy(x[sl]).to(dev) for x in (amp, tau, sig)]
        mk = torch.from_numpy(mask[sl].astype(numpy.int64)).to(dev)
        active = ((mk[:, None] >> bd[None, :]) & 1).to(torch.float64)
        mu = 0.15 * amp[:, None] * torch.exp(-(td[None, :] - tau[:, None]) ** 2 / (2.0 * sig[:, None] ** 2)) * active
        gen.manual_seed(int(seed + 104729 * (bidx + 1) + (g0 % blk)))
        lam = torch._standard_gamma(torch.full_like(mu, 2.0), generator=gen) * (0.5 * mu)
        out[lo:hi] = torch.log2(1.0 + 50.0 * torch.poisson(lam, generator=gen))
        lo = hi
    return out


def permutation_block_torch(n, S, seed=SEED, device='cuda'):
    """[n, S] int32: n uniformly random permutations of 0..S-1 (argsort of uniform keys)."""
    import torch
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    out = torch.empty((n, S), dtype=torch.int32, device=dev)
    step = max(1, (1 << 24) // max(S, 1))
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        out[lo:hi] = torch.rand((hi - lo, S), generator=gen, device=dev).argsort(dim=1).to(torch.int32)
    return out
