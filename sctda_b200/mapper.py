"""Host-side mirror of the Mapper seam of scTDA: sakmapper.apply_lens / sakmapper.mapper_graph as
scTDA calls them (/root/reference/scTDA/main.py:744-753, 768-771, "ref:") and
scTDA.TopologicalRepresentation (ref:731-778), with the cover, the per-patch clustering, the node
lists and the nerve computed by the CUDA kernels behind include/sctda_b200.h.

The Mapper algorithm itself is not in the reference tree (sakmapper is an absent third-party
dependency); what is built here is the algorithm restated in oracle/mapper_oracle.py, with the
call-site contract of the reference: mapper_graph returns (G, all_clusters, patches), G's nodes
0..V-1 enumerate all_clusters, TopologicalRepresentation.save writes name.json + name.gexf.
There is no CPU fallback for the graph build.
"""
import ctypes

import numpy
import torch

from . import _lib
from . import formats


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


_dummy = {}


def _nonnull(t):
    """Pointer of a device tensor; an empty tensor maps to a one-element dummy buffer."""
    if t.numel():
        return ctypes.c_void_p(t.data_ptr())
    if t.device not in _dummy:
        _dummy[t.device] = torch.zeros(16, dtype=torch.int64, device=t.device)
    return ctypes.c_void_p(_dummy[t.device].data_ptr())


class _Buf(object):
    """A device array of n elements whose pointer is valid even when n == 0 (an empty torch
    tensor reports a NULL data_ptr, which the C-ABI rejects as a missing argument)."""

    def __init__(self, shape, dtype, dev):
        n = shape[0] if isinstance(shape, tuple) else shape
        full = (max(n, 1),) + (shape[1:] if isinstance(shape, tuple) else ())
        self.base = torch.empty(full, dtype=dtype, device=dev)
        self.view = self.base[:n]
        self.ptr = ctypes.c_void_p(self.base.data_ptr())


def _padded_device_table(X, dev):
    """float64 device copy of a cells x genes table with an even leading dimension (the
    contraction kernel stages 16-byte chunks)."""
    if isinstance(X, torch.Tensor):
        Xd = X.to(dev, dtype=torch.float64)
    else:
        Xd = torch.from_numpy(numpy.ascontiguousarray(numpy.asarray(X, dtype=numpy.float64))).to(dev)
    n, g = Xd.shape
    if g % 2 == 0 and Xd.stride(1) == 1 and Xd.stride(0) % 2 == 0 and Xd.data_ptr() % 16 == 0:
        return Xd
    out = torch.zeros((n, g + (g % 2)), dtype=torch.float64, device=dev)
    out[:, :g] = Xd
    return out[:, :g]


def row_normalize(ctx, Xd, metric):
    """sctda_row_normalize_f64: (Xn [N, G] view of an even-ld buffer, sq [N])."""
    dev = Xd.device
    n, g = Xd.shape
    ld = (g + 15) // 16 * 16          # zero-padded to a multiple of 16 columns: TMA bulk-copy staging
    buf = torch.empty((n, ld), dtype=torch.float64, device=dev)
    sq = torch.empty(n, dtype=torch.float64, device=dev)
    mode = {'euclidean': 0, 'correlation': 1}[metric]
    ctx.check(ctx.lib.sctda_row_normalize_f64(ctx.handle, n, g, _lib.ptr(Xd), int(Xd.stride(0)), mode, _lib.ptr(buf), ld,
                                              _lib.ptr(sq), _stream(dev)))
    return buf[:, :g], sq


def pairwise_distances(ctx, Xd, metric, row_block=None):
    """The N x N dissimilarity of apply_lens(lens='mds', metric=...) (ref:746-751) on the device.
    One call over all rows computes the upper triangle of tiles and mirrors it (bit-identical:
    products commute); `row_block` computes it row block by row block instead (the sharding unit
    of SURVEY.md §8e: independent row blocks against all columns)."""
    dev = Xd.device
    Xn, sq = row_normalize(ctx, Xd, metric)
    n, g = Xn.shape
    D = torch.empty((n, n), dtype=torch.float64, device=dev)
    code = {'euclidean': 0, 'correlation': 1}[metric]
    if row_block is None or ctx.get_precision() != 'fp64':
        row_block = max(n, 1)         # (the split-precision path splits the operands once per call)
    for r0 in range(0, n, row_block):
        r1 = min(n, r0 + row_block)
        ctx.check(ctx.lib.sctda_pairdist_rows_f64(ctx.handle, n, g, _lib.ptr(Xn), int(Xn.stride(0)), _lib.ptr(sq), r0, r1,
                                                  code, _lib.ptr(D[r0:r1]), n, _stream(dev)))
    return D


def smacof_device(ctx, D, n_components=2, init=None, n_init=1, max_iter=300, eps=1e-6, random_state=None,
                  return_n_iter=False):
    """Metric MDS by SMACOF on a device dissimilarity matrix D [N, N] (symmetric, float64), following
    sklearn.manifold.smacof / _smacof_single step by step: uniform random start from `random_state`
    (or `init`), Guttman transform, stress of the new configuration, stop when
    (old_stress - stress) / (sum of squared distances / 2) < eps, best of n_init restarts.
    Every iteration is one sctda_smacof_pass over D; returns (positions [N, 2] device tensor, stress)."""
    import sklearn.utils
    if n_components != 2:
        raise ValueError('the lens is two-dimensional')
    dev = D.device
    n = int(D.shape[0])
    rs = sklearn.utils.check_random_state(random_state)
    if init is not None:
        n_init = 1
    stats = torch.zeros(2, dtype=torch.float64, device=dev)
    best = None
    for _ in range(int(n_init)):
        X0 = numpy.asarray(init, dtype=numpy.float64) if init is not None else rs.uniform(size=n * 2).reshape((n, 2))
        cur = torch.from_numpy(numpy.ascontiguousarray(X0)).to(dev)
        nxt = torch.empty_like(cur)
        old = None
        n_iter = 0
        # pass p on X_p yields stress(X_p) and X_{p+1}; sklearn's iteration it produces X_{it+1} and tests its stress
        for p in range(int(max_iter) + 1):
            ctx.check(ctx.lib.sctda_smacof_pass(ctx.handle, n, _lib.ptr(D), int(D.stride(0)), _lib.ptr(cur), _lib.ptr(nxt),
                                                _lib.ptr(stats), _stream(dev)))
            if p >= 1:
                stress, ssd = [float(v) for v in stats.cpu()]
                n_iter = p
                if old is not None and (old - stress) / (ssd / 2.0) < eps:
                    break
                old = stress
                if p == int(max_iter):
                    break
            cur, nxt = nxt, cur
        if best is None or stress < best[1]:
            best = (cur.clone(), stress, n_iter)
    return (best[0], best[1], best[2]) if return_n_iter else (best[0], best[1])


def gram_gather(ctx, Xn, rows_a, rows_b):
    dev = Xn.device
    ra = torch.as_tensor(rows_a, dtype=torch.int32, device=dev).contiguous()
    rb = torch.as_tensor(rows_b, dtype=torch.int32, device=dev).contiguous()
    out = torch.empty((ra.numel(), rb.numel()), dtype=torch.float64, device=dev)
    ctx.check(ctx.lib.sctda_gram_gather_f64(ctx.handle, int(Xn.shape[1]), _lib.ptr(Xn), int(Xn.stride(0)), ra.numel(),
                                            _lib.ptr(ra), rb.numel(), _lib.ptr(rb), _lib.ptr(out), rb.numel(), _stream(dev)))
    return out


# ----------------------------------------------------------------------------------------------
# cover bounds: 2(r+1) order statistics of the lens (host reference form and device form)
# ----------------------------------------------------------------------------------------------
def cover_bounds(lens, resolution, gain, equalize=True):
    """Open interval bounds per axis: fenceposts = numpy.percentile(axis, k*100/r) (equalize) or
    evenly spaced between min and max, each interval widened by gain/2 of its width on both sides
    (SURVEY.md appendix C.2)."""
    lens = numpy.asarray(lens, dtype=numpy.float64)
    r = int(resolution)
    out = []
    for ax in (0, 1):
        v = lens[:, ax]
        if equalize:
            posts = numpy.array([numpy.percentile(v, k * 100.0 / r) for k in range(r + 1)])
        else:
            posts = numpy.linspace(v.min(), v.max(), r + 1)
        w = posts[1:] - posts[:-1]
        out += [posts[:-1] - (0.5 * gain) * w, posts[1:] + (0.5 * gain) * w]
    return out


def cover_bounds_device(lens_d, resolution, gain, equalize=True):
    """cover_bounds on the device, without a host round trip: torch.sort (library plumbing) for the
    order statistics, then numpy.percentile's 'linear' rule operation by operation in float64
    ((n-1)*q, floor, gamma, a + (b-a)*gamma, switched to b - (b-a)*(1-gamma) for gamma >= 0.5), so the
    fenceposts are the bits numpy produces (the GPU parity tests compare the patches exactly)."""
    r = int(resolution)
    dev = lens_d.device
    n = int(lens_d.shape[0])
    out = []
    k = torch.arange(r + 1, dtype=torch.float64, device=dev)
    for ax in (0, 1):
        v = lens_d[:, ax].contiguous()
        if equalize:
            srt = torch.sort(v).values
            q = ((k * 100.0) / float(r)) / 100.0
            virt = float(n - 1) * q
            prev = torch.floor(virt)
            above = virt >= float(n - 1)
            pi = torch.where(above, torch.full_like(prev, n - 1), prev).to(torch.int64).clamp_(0, n - 1)
            ni = torch.where(above, torch.full_like(prev, n - 1), prev + 1).to(torch.int64).clamp_(0, n - 1)
            gamma = virt - torch.where(above, torch.full_like(prev, -1.0), prev)
            a, b = srt[pi], srt[ni]
            diff = b - a
            lo_half = a + diff * gamma
            hi_half = b - diff * (1.0 - gamma)
            posts = torch.where(gamma >= 0.5, hi_half, lo_half)
        else:
            start, stop = v.min(), v.max()
            step = (stop - start) / float(r)
            posts = k * step + start                       # numpy.linspace: arange * step + start, last = stop
            posts[r] = stop
        w = posts[1:] - posts[:-1]
        out += [(posts[:-1] - (0.5 * gain) * w).contiguous(), (posts[1:] + (0.5 * gain) * w).contiguous()]
    return out


class MapperDeviceResult(object):
    """Device-resident result of one graph build (all int32 tensors)."""
    __slots__ = ('r', 'bin_off', 'members', 'labels', 'bin_K', 'bin_db', 'V', 'node_off', 'node_members', 'node_bin',
                 'E', 'edges')


def build_graph_device(ctx, Xd, lens, resolution, gain, equalize=True, stat='db', max_K=5, metric='euclidean',
                       bounds=None, profile=None, linkage='single'):
    """normalise -> cover -> per-patch clustering -> nodes -> nerve, everything on the device.
    Xd: [N, G] float64 device tensor (even leading dimension); lens: [N, 2] host or device;
    linkage: 'single' (default), 'average', 'ward' or 'complete' (sctda_set_linkage)."""
    if linkage != 'single':
        with ctx.linkage(linkage):
            return build_graph_device(ctx, Xd, lens, resolution, gain, equalize, stat, max_K, metric, bounds, profile)
    dev = Xd.device
    N, G = Xd.shape
    r = int(resolution)
    if stat not in ('db', 'histgap'):
        raise NotImplementedError("statistics=%r: the device path serves 'db' and 'histgap' (the gap statistic is randomised)" % (stat,))
    if not isinstance(lens, torch.Tensor):
        lens = numpy.array(lens, dtype=numpy.float64)
    lens_d = torch.as_tensor(lens, dtype=torch.float64).to(dev).contiguous()
    st = _stream(dev)
    marks = []

    def mark(name):
        if profile is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(torch.cuda.current_stream(dev))
            marks.append((name, e))
    mark('start')
    if bounds is None:
        b = cover_bounds_device(lens_d, r, gain, equalize)
    else:
        b = [torch.from_numpy(numpy.ascontiguousarray(x)).to(dev) for x in bounds]
    mark('bounds')
    Xn, _ = row_normalize(ctx, Xd, metric)
    mark('normalize')
    res = MapperDeviceResult()
    res.r = r
    B = r * r
    res.bin_off = torch.empty(B + 1, dtype=torch.int32, device=dev)
    ctx.check(ctx.lib.sctda_cover_count(ctx.handle, N, _lib.ptr(lens_d), 2, r, _lib.ptr(b[0]), _lib.ptr(b[1]),
                                        _lib.ptr(b[2]), _lib.ptr(b[3]), _lib.ptr(res.bin_off), st))
    mtot = int(res.bin_off[B].item())
    _b_members = _Buf(mtot, torch.int32, dev)
    res.members = _b_members.view
    ctx.check(ctx.lib.sctda_cover_fill(ctx.handle, N, r, _lib.ptr(res.bin_off), _b_members.ptr, st))
    mark('cover')
    _b_labels = _Buf(mtot, torch.int32, dev)
    res.labels = _b_labels.view
    res.bin_K = torch.empty(B, dtype=torch.int32, device=dev)
    res.bin_db = torch.empty((B, 8), dtype=torch.float64, device=dev)
    ctx.check(ctx.lib.sctda_bin_cluster(ctx.handle, G, _lib.ptr(Xn), int(Xn.stride(0)), B, _lib.ptr(res.bin_off),
                                        _b_members.ptr, int(max_K), 1 if stat == 'histgap' else 0,
                                        _b_labels.ptr, _lib.ptr(res.bin_K), _lib.ptr(res.bin_db), st))
    if profile is not None:
        profile['gemm_ms'] = ctx.last_kernel_ms() if ctx.timing_enabled else None
    mark('cluster')
    node_base = torch.empty(B + 1, dtype=torch.int32, device=dev)
    ctx.check(ctx.lib.sctda_nodes_count(ctx.handle, B, _lib.ptr(res.bin_K), _lib.ptr(node_base), st))
    V = int(node_base[B].item())
    res.V = V
    res.node_off = torch.empty(V + 1, dtype=torch.int32, device=dev)
    _b_node_members = _Buf(mtot, torch.int32, dev)
    res.node_members = _b_node_members.view
    _b_node_bin = _Buf(V, torch.int32, dev)
    res.node_bin = _b_node_bin.view
    ctx.check(ctx.lib.sctda_nodes_fill(ctx.handle, B, _lib.ptr(res.bin_off), _b_members.ptr,
                                       _b_labels.ptr, _lib.ptr(node_base), V,
                                       _lib.ptr(res.node_off), _b_node_members.ptr,
                                       _b_node_bin.ptr, st))
    mark('nodes')
    ecount = torch.zeros(1, dtype=torch.int64, device=dev)
    ctx.check(ctx.lib.sctda_nerve_count(ctx.handle, N, V, _lib.ptr(res.node_off), _b_node_members.ptr,
                                        mtot, _lib.ptr(ecount), st))
    E = int(ecount.item())
    res.E = E
    _b_edges = _Buf((E, 2), torch.int32, dev)
    res.edges = _b_edges.view
    if E:
        ctx.check(ctx.lib.sctda_nerve_fill(ctx.handle, V, _b_edges.ptr, st))
    mark('nerve')
    if profile is not None:
        torch.cuda.current_stream(dev).synchronize()
        for (_, e0), (name, e1) in zip(marks[:-1], marks[1:]):
            profile[name + '_ms'] = e0.elapsed_time(e1)
        profile['total_ms'] = marks[0][1].elapsed_time(marks[-1][1])
        profile['Mtot'] = mtot
        profile['V'] = V
        profile['E'] = E
    return res


def cluster_sub_csr(ctx, Xn, sub_off, sub_members, max_K=5):
    """sctda_bin_cluster on a sub-CSR given as host arrays; returns host (labels, K)."""
    dev = Xn.device
    nb = len(sub_off) - 1
    off_d = torch.as_tensor(numpy.asarray(sub_off, dtype=numpy.int32)).to(dev)
    mem_d = torch.as_tensor(numpy.asarray(sub_members, dtype=numpy.int32)).to(dev)
    m = int(mem_d.numel())
    labels = torch.zeros(max(m, 1), dtype=torch.int32, device=dev)
    K = torch.zeros(max(nb, 1), dtype=torch.int32, device=dev)
    if nb:
        ctx.check(ctx.lib.sctda_bin_cluster(ctx.handle, int(Xn.shape[1]), _lib.ptr(Xn), int(Xn.stride(0)), nb, _lib.ptr(off_d),
                                            ctypes.c_void_p(mem_d.data_ptr()), int(max_K), 0, _lib.ptr(labels), _lib.ptr(K), None,
                                            _stream(dev)))
    return labels[:m].cpu().numpy(), K[:nb].cpu().numpy()


def build_graph_distributed(ctx, Xd, lens, resolution, gain, equalize=True, max_K=5, metric='euclidean', profile=None,
                            exchange='device'):
    """The graph build with the per-patch clustering sharded over the ranks of the default
    process group: X and the lens are replicated, patches are dealt by
    descending m^2, cover / nodes / nerve are replicated.  Returns the same MapperDeviceResult on every rank.

    exchange='device' (default on GPUs): only the patch sizes (a few KB) visit the host, to plan the shards; the
    sub-CSR of this rank's patches is cut on the device (sctda_csr_select), and the label fragments meet
    through ONE pair of ncclAllGather calls on device buffers followed by a placement kernel
    (sctda_allgather_csr, communicator from parallel.nccl_comm).
    exchange='host': the round-1 path through host arrays and torch.distributed.all_gather
    (parallel.patch_sharded_labels); kept for the gloo tests of the host logic."""
    from . import parallel
    import time as _time
    dev = Xd.device
    N, G = Xd.shape
    r = int(resolution)
    B = r * r
    if not isinstance(lens, torch.Tensor):
        lens = numpy.array(lens, dtype=numpy.float64)
    lens_d = torch.as_tensor(lens, dtype=torch.float64).to(dev).contiguous()
    st = _stream(dev)
    if profile is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream(dev))
    b = cover_bounds_device(lens_d, r, gain, equalize)
    Xn, _ = row_normalize(ctx, Xd, metric)
    res = MapperDeviceResult()
    res.r = r
    res.bin_off = torch.empty(B + 1, dtype=torch.int32, device=dev)
    ctx.check(ctx.lib.sctda_cover_count(ctx.handle, N, _lib.ptr(lens_d), 2, r, _lib.ptr(b[0]), _lib.ptr(b[1]),
                                        _lib.ptr(b[2]), _lib.ptr(b[3]), _lib.ptr(res.bin_off), st))
    bin_off = res.bin_off.cpu().numpy()                       # B + 1 integers: the plan needs the patch sizes
    _t = [_time.perf_counter()]
    mtot = int(bin_off[B])
    _b_members = _Buf(mtot, torch.int32, dev)
    res.members = _b_members.view
    ctx.check(ctx.lib.sctda_cover_fill(ctx.handle, N, r, _lib.ptr(res.bin_off), _b_members.ptr, st))
    if exchange == 'host':
        members_h = res.members.cpu().numpy()
        _t.append(_time.perf_counter())

        def _compute(so, sm):
            out = cluster_sub_csr(ctx, Xn, so, sm, max_K)
            _t.append(_time.perf_counter())
            return out
        labels, bin_K = parallel.patch_sharded_labels(bin_off, members_h, _compute)
        _t.append(_time.perf_counter())
        res.labels = torch.as_tensor(labels).to(dev)
        res.bin_K = torch.as_tensor(bin_K).to(dev)
    else:
        rank, ws = parallel.world()
        sizes = bin_off[1:].astype(numpy.int64) - bin_off[:-1]
        plan = parallel.patch_shard_plan(sizes, rank, ws)
        nb = int(plan['mine'].size)
        frag_len = int(plan['sub_off'][-1])
        small = numpy.concatenate([plan['mine'], plan['sub_off'], plan['owner'], plan['frag_off'], plan['frag_idx']])
        small_d = torch.from_numpy(small.astype(numpy.int32)).to(dev, non_blocking=True)      # one small upload
        o = numpy.cumsum([0, nb, nb + 1, B, B, B])
        mine_d, sub_off_d, owner_d, frag_off_d, frag_idx_d = [small_d[o[i]:o[i + 1]] for i in range(5)]
        _t.append(_time.perf_counter())
        sub_members = _Buf(frag_len, torch.int32, dev)
        ctx.check(ctx.lib.sctda_csr_select(ctx.handle, nb, _nonnull(mine_d), _lib.ptr(res.bin_off), _b_members.ptr,
                                           _nonnull(sub_off_d), sub_members.ptr, st))
        frag_labels = _Buf(frag_len, torch.int32, dev)
        frag_K = _Buf(nb, torch.int32, dev)
        if nb:
            ctx.check(ctx.lib.sctda_bin_cluster(ctx.handle, G, _lib.ptr(Xn), int(Xn.stride(0)), nb, _nonnull(sub_off_d),
                                                sub_members.ptr, int(max_K), 0, frag_labels.ptr, frag_K.ptr, None, st))
        _t.append(_time.perf_counter())
        _b_labels = _Buf(mtot, torch.int32, dev)
        res.labels = _b_labels.view
        res.bin_K = torch.empty(B, dtype=torch.int32, device=dev)
        comm = parallel.nccl_comm(dev.index) if ws > 1 else None
        ctx.check(ctx.lib.sctda_allgather_csr(ctx.handle, comm, ws, frag_labels.ptr, frag_len, int(plan['max_frag']),
                                              frag_K.ptr, nb, int(plan['max_nb']), B, _lib.ptr(res.bin_off),
                                              _lib.ptr(owner_d), _lib.ptr(frag_off_d), _lib.ptr(frag_idx_d),
                                              _b_labels.ptr, _lib.ptr(res.bin_K), st))
        _t.append(_time.perf_counter())
    res.bin_db = None
    _finish_nodes_and_nerve(ctx, res, N, mtot)
    if profile is not None:
        e1.record(torch.cuda.current_stream(dev))
        torch.cuda.current_stream(dev).synchronize()
        _t.append(_time.perf_counter())
        profile.update({'host_ms': {'plan_and_upload' if exchange != 'host' else 'cover_to_host': 1e3 * (_t[1] - _t[0]),
                                    'cluster_shard': 1e3 * (_t[2] - _t[1]),
                                    'exchange_and_merge': 1e3 * (_t[3] - _t[2]), 'nodes_nerve': 1e3 * (_t[4] - _t[3])},
                        'exchange': exchange, 'total_ms': e0.elapsed_time(e1), 'Mtot': mtot, 'V': res.V, 'E': res.E,
                        'gemm_ms': ctx.last_kernel_ms() if ctx.timing_enabled else None})
    return res


def _finish_nodes_and_nerve(ctx, res, N, mtot):
    dev = res.bin_off.device
    st = _stream(dev)
    B = res.r * res.r
    node_base = torch.empty(B + 1, dtype=torch.int32, device=dev)
    ctx.check(ctx.lib.sctda_nodes_count(ctx.handle, B, _lib.ptr(res.bin_K), _lib.ptr(node_base), st))
    V = int(node_base[B].item())
    res.V = V
    res.node_off = torch.empty(V + 1, dtype=torch.int32, device=dev)
    _b_node_members = _Buf(mtot, torch.int32, dev)
    res.node_members = _b_node_members.view
    _b_node_bin = _Buf(V, torch.int32, dev)
    res.node_bin = _b_node_bin.view
    ctx.check(ctx.lib.sctda_nodes_fill(ctx.handle, B, _lib.ptr(res.bin_off), _nonnull(res.members),
                                       _nonnull(res.labels), _lib.ptr(node_base), V,
                                       _lib.ptr(res.node_off), _b_node_members.ptr,
                                       _b_node_bin.ptr, st))
    ecount = torch.zeros(1, dtype=torch.int64, device=dev)
    ctx.check(ctx.lib.sctda_nerve_count(ctx.handle, N, V, _lib.ptr(res.node_off), _b_node_members.ptr,
                                        mtot, _lib.ptr(ecount), st))
    res.E = int(ecount.item())
    _b_edges = _Buf((res.E, 2), torch.int32, dev)
    res.edges = _b_edges.view
    if res.E:
        ctx.check(ctx.lib.sctda_nerve_fill(ctx.handle, V, _b_edges.ptr, st))


# ----------------------------------------------------------------------------------------------
# the sakmapper-shaped API
# ----------------------------------------------------------------------------------------------
def _values(df):
    return df.values if hasattr(df, 'values') else numpy.asarray(df)


def apply_lens(df, lens='pca', metric='correlation', dissimilarity=None, device=0, **kwargs):
    """sakmapper.apply_lens as called at ref:745-753.  Returns a DataFrame (N x 2, index of df).
    'pca': the two leading principal components, computed on the device (sklearn.decomposition.PCA
    sign convention); 'mds': the dissimilarity matrix in `metric` is computed on the device
    (sctda_pairdist_rows_f64) and embedded on the device by SMACOF (smacof_device: sklearn.manifold.MDS's
    algorithm and keyword arguments n_init / max_iter / eps / random_state / init);
    dissimilarity='precomputed' takes df itself as the matrix;
    'neighbor': sklearn.manifold.SpectralEmbedding on the host."""
    import pandas
    X = numpy.asarray(_values(df), dtype=numpy.float64)
    index = df.index if hasattr(df, 'index') else None
    dev = torch.device('cuda', device)
    if lens == 'pca':
        if kwargs.get('n_components', 2) != 2:
            raise ValueError('the lens is two-dimensional')
        Xd = torch.from_numpy(numpy.ascontiguousarray(X)).to(dev)
        out = pca_lens_device(Xd).cpu().numpy()
    elif lens == 'mds':
        ctx = _lib.context(device)
        if dissimilarity == 'precomputed':
            D = torch.from_numpy(numpy.ascontiguousarray(0.5 * (X + X.T))).to(dev)      # sklearn's check_symmetric
        else:
            D = pairwise_distances(ctx, _padded_device_table(X, dev), metric)
        kw = {k: kwargs[k] for k in ('n_components', 'n_init', 'max_iter', 'eps', 'random_state', 'init') if k in kwargs}
        unknown = set(kwargs) - set(kw) - {'n_jobs', 'verbose'}
        if unknown:
            raise TypeError('apply_lens(lens="mds"): unsupported MDS arguments %s' % sorted(unknown))
        out = smacof_device(ctx, D, **kw)[0].cpu().numpy()
    elif lens == 'neighbor':
        import sklearn.manifold
        kw = dict(kwargs)
        kw.setdefault('n_components', 2)
        out = sklearn.manifold.SpectralEmbedding(**kw).fit_transform(X)
    else:
        raise ValueError("lens must be 'pca', 'mds' or 'neighbor'")
    return pandas.DataFrame(out, index=index)


def pca_lens_device(Xd, ctx=None, tol=1e-12, max_iter=1000, return_info=False):
    """Two leading principal component scores of the rows of Xd (device, float64), with
    sklearn.decomposition.PCA's sign rule (largest |loading| of each component positive):
    sctda_pca_lens_f64 (csrc/pca.cu: centred Gram matrix by the library's FP64 contraction, two leading
    eigenpairs by block subspace iteration, no library eigen-solver)."""
    dev = Xd.device
    if ctx is None:
        ctx = _lib.context(dev.index if dev.index is not None else torch.cuda.current_device())
    Xd = Xd if (Xd.dtype == torch.float64 and Xd.stride(1) == 1) else Xd.to(torch.float64).contiguous()
    n, g = Xd.shape
    lens = torch.empty((n, 2), dtype=torch.float64, device=dev)
    info = torch.zeros(6, dtype=torch.float64, device=dev)
    ctx.check(ctx.lib.sctda_pca_lens_f64(ctx.handle, int(n), int(g), _lib.ptr(Xd), int(Xd.stride(0)), _lib.ptr(lens), 2,
                                         _lib.ptr(info), int(max_iter), float(tol), _stream(dev)))
    if return_info:
        v = info.cpu().numpy()
        return lens, {'theta': [float(v[0]), float(v[1])], 'residual': [float(v[2]), float(v[3])],
                      'iterations': int(v[4]), 'status': int(v[5])}
    return lens


def mapper_graph(df, lens_data, resolution=10, gain=0.5, equalize=True, clust='agglomerative', stat='db', max_K=5,
                 metric='euclidean', device=0, verbose=True, precision='fp64', linkage='single'):
    """sakmapper.mapper_graph as called at ref:768-771 -> (G, all_clusters, patches):
    G networkx.Graph with nodes 0..V-1 and weight-1.0 edges, all_clusters[k] the ascending row
    positions of node k, patches {(i, j): row positions of cover patch (i, j)}.
    `metric` (extension): distances used by the per-patch clustering.
    `precision` (extension): 'fp64' (default, graph identical to the oracle) or 'fp16x3' (opt-in
    tcgen05 contraction, distances within 1e-5; see include/sctda_b200.h sctda_set_precision).
    `linkage` (extension): the linkage of clust='agglomerative'.  sakmapper passes the patch to
    sklearn.cluster.AgglomerativeClustering and its linkage is NOT pinned by the reference's call sites
    (SURVEY.md appendix C.2): 'single' (default: the minimum spanning tree, north_star (2)), 'average', 'ward'
    (sklearn's own default) or 'complete'.  clust='kmeans' and stat='gap' are randomised procedures and are
    not offered."""
    import networkx
    if clust != 'agglomerative':
        raise NotImplementedError("clust=%r: the device path offers clust='agglomerative' (linkage='single', 'average', "
                                  "'ward' or 'complete'); KMeans is randomised and not offered" % (clust,))
    ctx = _lib.context(device)
    dev = torch.device('cuda', device)
    Xd = _padded_device_table(_values(df), dev)
    lens = numpy.asarray(_values(lens_data), dtype=numpy.float64)[:, :2]
    with ctx.precision(precision):
        res = build_graph_device(ctx, Xd, lens, resolution, gain, equalize=equalize, stat=stat, max_K=max_K, metric=metric,
                                 linkage=linkage)
    bin_off = res.bin_off.cpu().numpy()
    members = res.members.cpu().numpy()
    node_off = res.node_off.cpu().numpy()
    node_members = res.node_members.cpu().numpy()
    edges = res.edges.cpu().numpy()
    r = res.r
    patches = {}
    for b in range(r * r):
        patches[(b // r, b % r)] = members[bin_off[b]:bin_off[b + 1]]
    all_clusters = [node_members[node_off[k]:node_off[k + 1]] for k in range(res.V)]
    if verbose:
        print('total of %d patches required clustering' % int((bin_off[1:] > bin_off[:-1]).sum()))
        print('this implies %d nodes in the mapper graph' % res.V)
    G = networkx.Graph()
    G.add_nodes_from(range(res.V))
    G.add_edges_from(((int(i), int(j)) for i, j in edges), weight=1.0)
    return G, all_clusters, patches


class TopologicalRepresentation(object):
    """Builds a topological representation of the data with the Mapper algorithm (ref:731-778)."""

    def __init__(self, table, lens='mds', metric='correlation', precomputed=False, device=0, **kwargs):
        """ref:735-756 (the scatter plot of ref:754-756 is not produced)."""
        import pandas
        self.device = device
        self.metric = metric
        self.df = pandas.read_table(table + '.mapper.tsv')                       # ref:743
        if lens == 'neighbor':
            self.lens_data_mds = apply_lens(self.df, lens=lens, device=device, **kwargs)
        elif lens == 'mds':
            if precomputed:
                self.lens_data_mds = apply_lens(self.df, lens=lens, metric=metric, dissimilarity='precomputed',
                                                device=device, **kwargs)
            else:
                self.lens_data_mds = apply_lens(self.df, lens=lens, metric=metric, device=device, **kwargs)
        else:
            self.lens_data_mds = apply_lens(self.df, lens=lens, device=device, **kwargs)

    def save(self, name, resolution, gain, equalize=True, cluster='agglomerative', statistics='db', max_K=5, *,
             cluster_metric='euclidean', precision='fp64', linkage='single'):
        """ref:758-778: writes name.json and name.gexf, returns the patches.  The positional signature is
        the reference's; `cluster_metric`, `precision` and `linkage` are keyword-only extensions (the
        opt-in contraction precision and the linkage of cluster='agglomerative', see mapper_graph)."""
        G, all_clusters, patches = mapper_graph(self.df, lens_data=self.lens_data_mds, resolution=resolution, gain=gain,
                                                equalize=equalize, clust=cluster, stat=statistics, max_K=max_K,
                                                metric=cluster_metric, device=self.device, precision=precision,
                                                linkage=linkage)
        formats.write_mapper_json(name + '.json', all_clusters)                  # ref:772-776
        formats.write_gexf(name + '.gexf', G.number_of_nodes(), sorted(G.edges()))   # ref:777
        return patches
