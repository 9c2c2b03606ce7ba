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


def tile_shard_positions(n, rank, world_size, tile=128):
    """Positions (ascending) of a planned gene order that `rank` works on: the 128-gene tiles of the order are
    dealt round-robin, so that every rank gets the same mix of sparse and dense tiles (a contiguous range of the
    order would hand one rank all the dense genes) while each tile stays as homogeneous as the plan made it."""
    import numpy
    pos = numpy.arange(n, dtype=numpy.int64)
    return pos[(pos // tile) % world_size == rank]


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


# Cost model of the patch dealer, in units of one entry of a patch Gram tile (measured on a B200, 100k x 2k build,
# tools/lpt_probe.py: 7.4e-8 ms per entry for the contraction; 2.4e-3 ms per dependent step of the MST of a patch
# above CHAIN_MIN cells).  The MSTs of a rank's large patches run side by side, one cluster of CTAs each, after the
# contraction: a rank pays the chain of its LARGEST patch once, on top of its tiles.
CHAIN_MIN = 4096
CHAIN_ENTRIES_PER_STEP = 3.2e4


def assign_patches_lpt(sizes, world_size, chain=True):
    """Deal patches to ranks by descending size, each to the currently least loaded rank; load = the m^2 entries of
    the rank's Gram tiles plus the MST chain of its largest patch (the first one it is dealt).  Deterministic.
    Returns a list of ascending patch-index arrays."""
    sizes = numpy.asarray(sizes, dtype=numpy.int64)
    order = numpy.lexsort((numpy.arange(sizes.size), -sizes))        # descending size, ties by index
    load = numpy.zeros(world_size, dtype=numpy.float64)
    fresh = numpy.ones(world_size, dtype=bool)
    owner = numpy.full(sizes.size, -1, dtype=numpy.int64)
    for b in order:
        if sizes[b] == 0:
            owner[b] = b % world_size
            continue
        r = int(numpy.argmin(load))
        owner[b] = r
        load[r] += float(sizes[b]) ** 2
        if chain and fresh[r] and sizes[b] > CHAIN_MIN:
            load[r] += CHAIN_ENTRIES_PER_STEP * float(sizes[b])
        fresh[r] = False
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


# ----------------------------------------------------------------------------------------------
# NCCL communicator for the library's own collectives (include/sctda_b200.h: sctda_allgather_csr,
# sctda_gather_results): created HERE, by the host, with ctypes on the process's libnccl.so.2;
# torch.distributed only carries the 128-byte unique id from rank 0 to the others.
# ----------------------------------------------------------------------------------------------
_nccl_comm = {}


def nccl_comm(device):
    """The ncclComm_t (as ctypes.c_void_p) of this rank for the default process group, created once per
    process.  Needs an initialised torch.distributed group (any backend) and one GPU per rank."""
    import ctypes
    if 'comm' in _nccl_comm:
        return _nccl_comm['comm']
    rank, ws = world()

    class UniqueId(ctypes.Structure):
        _fields_ = [('internal', ctypes.c_char * 128)]
    lib = ctypes.CDLL('libnccl.so.2', mode=ctypes.RTLD_GLOBAL)          # the copy torch has already loaded
    lib.ncclGetUniqueId.argtypes = [ctypes.POINTER(UniqueId)]
    lib.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, UniqueId, ctypes.c_int]
    uid = UniqueId()
    if rank == 0:
        rc = lib.ncclGetUniqueId(ctypes.byref(uid))
        if rc != 0:
            raise RuntimeError('ncclGetUniqueId failed (%d)' % rc)
    cdev = _comm_device()
    t = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).clone().to(cdev)
    dist.broadcast(t, src=0)
    ctypes.memmove(ctypes.byref(uid), bytes(t.cpu().numpy().tobytes()), 128)
    torch.cuda.set_device(device)
    comm = ctypes.c_void_p()
    rc = lib.ncclCommInitRank(ctypes.byref(comm), int(ws), uid, int(rank))
    if rc != 0:
        raise RuntimeError('ncclCommInitRank failed (%d)' % rc)
    _nccl_comm['comm'] = comm
    _nccl_comm['lib'] = lib
    return comm


def patch_shard_plan(sizes, rank, world_size):
    """Host plan of the patch-sharded build from the patch sizes [B] alone (a few KB): patches dealt by
    descending m^2 (assign_patches_lpt).  Returns a dict of small int32 arrays:
      mine [nb]       ascending patch ids of this rank,   sub_off [nb+1]  CSR offsets of its sub-CSR,
      owner [B]       rank that clusters each patch,      frag_off [B]    offset of the patch inside its owner's
      frag_idx [B]    index of the patch among its owner's patches,        label fragment,
      max_frag, max_nb  the largest fragment / patch count over the ranks."""
    sizes = numpy.asarray(sizes, dtype=numpy.int64)
    parts = assign_patches_lpt(sizes, world_size)
    B = sizes.size
    owner = numpy.zeros(B, dtype=numpy.int32)
    frag_off = numpy.zeros(B, dtype=numpy.int32)
    frag_idx = numpy.zeros(B, dtype=numpy.int32)
    max_frag, max_nb = 0, 0
    for r, p in enumerate(parts):
        owner[p] = r
        off = numpy.zeros(p.size + 1, dtype=numpy.int64)
        numpy.cumsum(sizes[p], out=off[1:])
        frag_off[p] = off[:-1]
        frag_idx[p] = numpy.arange(p.size)
        max_frag = max(max_frag, int(off[-1]))
        max_nb = max(max_nb, int(p.size))
        if r == rank:
            mine, sub_off = p.astype(numpy.int32), off.astype(numpy.int32)
    return {'mine': mine, 'sub_off': sub_off, 'owner': owner, 'frag_off': frag_off, 'frag_idx': frag_idx,
            'max_frag': max_frag, 'max_nb': max_nb}


def gather_counts_device(ctx, mine, max_n):
    """The final gather of the permutation test on device buffers (sctda_gather_results, NCCL): `mine` int32
    device tensor of this rank's per-gene counts, max_n the largest shard.  Returns the [world, max_n] int32
    device tensor (rank-major, shards padded with zeros) on every rank; one rank: `mine` itself."""
    import ctypes
    from . import _lib
    rank, ws = world()
    if ws == 1:
        return mine
    dev = mine.device
    out = torch.empty((ws, int(max_n)), dtype=torch.int32, device=dev)
    ctx.check(ctx.lib.sctda_gather_results(ctx.handle, nccl_comm(dev.index), ws, _lib.ptr(mine), int(mine.numel()), int(max_n),
                                           _lib.ptr(out), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out

