"""One-box multi-GPU sharding of the two paths: one process per GPU,
torch.distributed for the plumbing (NCCL on the GPUs, gloo in the CPU tests).

  (b) permutation test: genes are independent -> contiguous gene shards, no collective on the
      data path, one all_gather of the per-gene integer counts at the end;
  (a) Mapper build: cover patches are independent -> patches dealt to ranks by descending m^2
      (longest-processing-time first), every rank clusters its patches, one all_gather of the
      per-patch labels; cover, node lists and nerve are replicated integer work.

The compute itself is passed in as a callable, so the sharding logic is testable on CPU ranks.
"""
import numpy
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n, rank, world_size):
    """Contiguous shard [lo, hi) of n units for `rank` (sizes differ by at most one)."""
    return (n * rank) // world_size, (n * (rank + 1)) // world_size


def _comm_device():
    return torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' else torch.device('cpu')


def all_gather_varlen(t):
    """all_gather of 1-D tensors of different lengths; returns the list of per-rank tensors (on
    the device of `t`)."""
    rank, ws = world()
    if ws == 1:
        return [t]
    cdev = _comm_device()
    n = torch.tensor([t.numel()], dtype=torch.int64, device=cdev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=cdev) for _ in range(ws)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    m = max(max(sizes), 1)
    buf = torch.zeros(m, dtype=t.dtype, device=cdev)
    buf[:t.numel()] = t.to(cdev)
    out = [torch.zeros(m, dtype=t.dtype, device=cdev) for _ in range(ws)]
    dist.all_gather(out, buf)
    return [o[:s].to(t.device) for o, s in zip(out, sizes)]


def gene_sharded_counts(n_genes, compute):
    """Permutation test over `n_genes` genes sharded by rank.  compute(lo, hi) -> int32 tensor of
    the exceed counts of genes [lo, hi).  Returns the full [n_genes] int32 tensor on every rank."""
    rank, ws = world()
    lo, hi = shard_range(n_genes, rank, ws)
    mine = compute(lo, hi).to(torch.int32).reshape(-1)
    assert mine.numel() == hi - lo
    return torch.cat(all_gather_varlen(mine))


def assign_patches_lpt(sizes, world_size):
    """Deal patches to ranks by descending m^2 (the Gram tile and the MST are O(m^2)), each to the
    currently least loaded rank.  Deterministic.  Returns a list of ascending patch-index arrays."""
    sizes = numpy.asarray(sizes, dtype=numpy.int64)
    order = numpy.lexsort((numpy.arange(sizes.size), -sizes))        # descending size, ties by index
    load = numpy.zeros(world_size, dtype=numpy.float64)
    owner = numpy.full(sizes.size, -1, dtype=numpy.int64)
    for b in order:
        if sizes[b] == 0:
            owner[b] = b % world_size
            continue
        r = int(numpy.argmin(load))
        owner[b] = r
        load[r] += float(sizes[b]) ** 2
    return [numpy.nonzero(owner == r)[0] for r in range(world_size)]


def sub_csr(bin_off, members, patches):
    """CSR restricted to `patches` (ascending indices): (sub_off, sub_members)."""
    bin_off = numpy.asarray(bin_off, dtype=numpy.int64)
    sizes = bin_off[1:] - bin_off[:-1]
    sub_sizes = sizes[patches]
    sub_off = numpy.zeros(len(patches) + 1, dtype=numpy.int64)
    numpy.cumsum(sub_sizes, out=sub_off[1:])
    idx = numpy.concatenate([numpy.arange(bin_off[b], bin_off[b + 1]) for b in patches]) if len(patches) else \
        numpy.zeros(0, dtype=numpy.int64)
    return sub_off, numpy.asarray(members)[idx]


def patch_sharded_labels(bin_off, members, compute):
    """Per-patch clustering sharded over the ranks.  bin_off/members: the (replicated) cover CSR
    as host arrays.  compute(sub_off, sub_members) -> (labels int32 [len(sub_members)], K int32
    [n_sub_patches]) clusters a sub-CSR.  Returns (labels [Mtot], K [B]) as host int32 arrays,
    identical on every rank."""
    rank, ws = world()
    bin_off = numpy.asarray(bin_off, dtype=numpy.int64)
    B = bin_off.size - 1
    sizes = bin_off[1:] - bin_off[:-1]
    parts = assign_patches_lpt(sizes, ws)
    sub_off, sub_members = sub_csr(bin_off, members, parts[rank])
    lab, K = compute(sub_off, sub_members)
    lab = torch.as_tensor(numpy.asarray(lab, dtype=numpy.int32))
    K = torch.as_tensor(numpy.asarray(K, dtype=numpy.int32))
    labs = all_gather_varlen(lab)
    Ks = all_gather_varlen(K)
    labels = numpy.zeros(int(bin_off[-1]), dtype=numpy.int32)
    bin_K = numpy.zeros(B, dtype=numpy.int32)
    for r in range(ws):
        lr = labs[r].cpu().numpy()
        kr = Ks[r].cpu().numpy()
        pos = 0
        for i, b in enumerate(parts[r]):
            m = int(sizes[b])
            labels[bin_off[b]:bin_off[b + 1]] = lr[pos:pos + m]
            bin_K[b] = kr[i]
            pos += m
    return labels, bin_K
