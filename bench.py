#!/usr/bin/env python
"""bench.py — connectivity permutation test throughput (gene·perms/s), BASELINE.json metric.

  python bench.py [--gpus N] [--steps K] [--warmup W]          own arm (CUDA, through the C-ABI)
  python bench.py --impl reference ...                          the reference's algorithm on the
                                                                host cores (CPU oracle port)

Workload (BASELINE.json configs[3], SURVEY.md §8d "C4"): 2,000-node Mapper graph over
S = 20,000 cells (every cell in 4 nodes, ring lattice of degree 8 + 2,000 chords),
20,000 genes x 10,000 permutations, genes sharded over the ranks with no data-path
collective (strong scaling: the total is fixed).  One step = one pass of the permutation
test over the rank's gene shard x all permutations.  The permutation block [n, S] is an
INPUT shared by all genes (SURVEY.md §7 H1 option i: 2e8 fresh per-gene permutations are
16 TB of indices; the per-gene reference stream is exercised by the parity tests).

JSON line: see the task contract.  `value` = device-resident throughput, the final gather of
the per-gene counts of all ranks included; `e2e` = the same through graph.conn_perm_test_host
with pinned HOST buffers (H2D of table + permutations, the tile layout and D2H of the counts
inside the timed region); `roofline` = the dominant kernel (conn_perm64_kernel) against the
measured HBM copy bandwidth with SURVEY.md §8d's algorithmic bytes (4*S + 8*S/n per
gene·perm), its duration averaged over all timed steps; `cpu_baseline` = the oracle port on
a bounded sample of the same workload.

`mapper`: the Mapper build of BASELINE.json's metric, 100,000 cells x 2,000 genes (configs[2]) at
EVERY N including 1 (cells/s; patches sharded over the ranks for N > 1), with the FP64
contraction's roofline in both conventions (SURVEY 8d's algorithmic flops without symmetry
credit, and the flops of the tiles actually executed); configs[1] (20k x 5k) rides along at N = 1.

`rooted`: BASELINE.json configs[4] at every N: RootedGraph on a 50,000-cell Mapper graph through the
reference-facing classes, the all-gene table (connectivity test + centroid / dispersion) with the genes
sharded over the ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy  # noqa: E402

METRIC = 'connectivity test gene*perms/s'
UNIT = 'gene*perms/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='own', choices=['own', 'reference'])
    ap.add_argument('--genes', type=int, default=20000)
    ap.add_argument('--perms', type=int, default=10000)
    ap.add_argument('--cells', type=int, default=20000)
    ap.add_argument('--nodes', type=int, default=2000)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-mapper', action='store_true')
    ap.add_argument('--no-rooted', action='store_true')
    ap.add_argument('--mapper-cells', type=int, default=20000)
    ap.add_argument('--mapper-genes', type=int, default=5000)
    ap.add_argument('--mapper-resolution', type=int, default=30)
    ap.add_argument('--mapper-gain', type=float, default=3.0)
    ap.add_argument('--mapper-metric', default='correlation')
    ap.add_argument('--no-cpu', action='store_true')
    return ap.parse_args()


def workload_config(a, n_gpus):
    return {'workload': 'connectivity permutation test: %d-node Mapper graph x %d genes x %d permutations, '
                        'S=%d cells, 4 nodes per cell; genes sharded over %d GPU(s)'
                        % (a.nodes, a.genes, a.perms, a.cells, n_gpus),
            'nodes': a.nodes, 'cells': a.cells, 'genes': a.genes, 'perms': a.perms,
            'permutations': 'one shared [n,S] int32 block (input)', 'l2': 'inputs_exceed_l2',
            'final_gather': 'counts of all ranks gathered inside the timed region (sctda_gather_results: ncclAllGather on device buffers)',
            'parallelism': 'genes/%d' % n_gpus}


# ---------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md)
# ---------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.stop = False
        self.t = None

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], stdout=subprocess.PIPE,
                                     stderr=subprocess.DEVNULL, timeout=5).stdout.decode().strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(numpy.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                'samples': len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port (reference algorithm, ref:966-992) on a bounded sample
# ---------------------------------------------------------------------------------------------
_W = {}


def _one_blas_thread():
    """One BLAS / OpenMP thread in this process: the reference is a per-gene loop that one process
    per core runs gene-parallel; a multithreaded numpy.dot inside every worker oversubscribes the
    cores (round 1: 4x slower at N = 1 than under torchrun, which exports OMP_NUM_THREADS=1)."""
    try:
        import threadpoolctl
        _W['tp'] = threadpoolctl.threadpool_limits(limits=1)
    except Exception:
        pass


def _cpu_init(members, adj_dense, Y, perms):
    from oracle import conn_oracle as co
    V = len(members)
    dic = {str(k): [int(c) for c in members[k]] for k in range(V)}
    st = co.GraphState([str(k) for k in range(V)], adj_dense, dic, {'g%d' % i: Y[i] for i in range(Y.shape[0])})
    _W['st'], _W['perms'], _W['co'] = st, perms, co


def _cpu_gene(i):
    co, st = _W['co'], _W['st']
    with numpy.errstate(invalid='ignore', divide='ignore'):
        return float(co.connectivity_pvalue(st, 'g%d' % i, n=_W['perms'].shape[0], perms=_W['perms']))


class CpuArm(object):
    """The oracle port (oracle/conn_oracle.py, the reference's per-gene loop ref:966-992) on the bench
    graph with `procs` single-threaded worker processes; sample(n_genes) times n_genes genes x all
    permutations of the block."""

    def __init__(self, a, n_perms, procs, max_genes):
        import multiprocessing as mp
        from sctda_b200 import synth
        members, adj, _ = synth.connectivity_case(a.cells, a.nodes, 1, seed=synth.SEED)
        self.Y = synth.expression_genes_numpy(max_genes, a.cells, seed=synth.SEED)
        rng = numpy.random.RandomState(7)
        perms = numpy.stack([rng.permutation(a.cells) for _ in range(n_perms)]).astype(numpy.int32)
        _one_blas_thread()
        _cpu_init(members, adj.toarray(), self.Y, perms)               # inherited by the forked workers
        self.n_perms, self.procs = n_perms, procs
        self.pool = mp.get_context('fork').Pool(procs, initializer=_one_blas_thread) if procs > 1 else None
        if self.pool:
            self.pool.map(_cpu_gene, range(min(procs, max_genes)), chunksize=1)      # spin-up, untimed

    def sample(self, n_genes, offset=0):
        idx = [(offset + i) % self.Y.shape[0] for i in range(n_genes)]
        t0 = time.perf_counter()
        if self.pool:
            self.pool.map(_cpu_gene, idx, chunksize=1)
        else:
            for i in idx:
                _cpu_gene(i)
        dt = time.perf_counter() - t0
        return n_genes * self.n_perms / dt, dt

    def close(self):
        if self.pool:
            self.pool.close()
            self.pool.join()


def mapper_cpu_sample(budget_s=20.0, N=100000, G=2000, r=30, gain=3.0, metric='euclidean', max_patch=5000):
    """The Mapper oracle (oracle/mapper_oracle.py) on a BOUNDED sample of the bench workload: the cover
    of the 100k x 2k table is cut as the device path cuts it (PCA lens on the host), then patches
    (those up to `max_patch` cells, in random order) are clustered one by one until `budget_s` is
    spent.  The per-patch clustering is O(m^2 G); the build rate of the whole table is extrapolated
    by the m^2 share of the sample: cells/s = N * (sum m^2 of the sample / sum m^2 of all patches) / seconds."""
    from oracle import mapper_oracle as mo
    from sctda_b200 import synth
    t_all = time.perf_counter()
    X = synth.expression_genes_numpy(G, N, seed=synth.SEED + 11).T.copy()
    Xc = X - X.mean(axis=0)
    w, v = numpy.linalg.eigh(Xc.T @ Xc)
    lens = Xc @ v[:, [-1, -2]]
    del Xc
    Xn, _ = mo.row_normalize(X, metric)
    bins, _ = mo.cover(lens, r, gain, True)
    sizes = numpy.array([c.size for c in bins], dtype=numpy.float64)
    rng = numpy.random.RandomState(3)
    done_m2, n_done, t_spent = 0.0, 0, 0.0
    for b in rng.permutation(len(bins)):
        cells = bins[int(b)]
        if cells.size < 2 or cells.size > max_patch:
            continue
        t0 = time.perf_counter()
        mo.cluster_bin(mo.gram(Xn, cells, cells), max_K=5)
        t_spent += time.perf_counter() - t0
        done_m2 += float(cells.size) ** 2
        n_done += 1
        if t_spent > budget_s:
            break
    total_m2 = float((sizes ** 2).sum())
    est = N * (done_m2 / total_m2) / t_spent
    return {'value': est, 'unit': 'cells/s', 'cores': os.cpu_count(), 'kind': 'port',
            'workload': 'Mapper build: %d cells x %d genes, %s metric, PCA lens, resolution %d, gain %g, max_K 5' % (N, G, metric, r, gain),
            'sample': '%d of %d patches of that cover (patches up to %d cells, random order; %.3g of sum m^2) clustered by '
                      'oracle/mapper_oracle.py in %.1f s (Gram matrices by an OpenMP fma loop, MST / selection single-threaded '
                      'numpy); whole-build rate extrapolated by the m^2 share; setup (table, lens, cover) %.0f s untimed'
                      % (n_done, len(bins), max_patch, done_m2 / total_m2, t_spent, time.perf_counter() - t_all - t_spent)}


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_perms = 2000                                     # >= 2,000 permutations per gene: get_gene's Python loops are
    n_genes = cores                                    # amortised as in the real workload (10,000), one gene per core and step
    arm = CpuArm(a, n_perms, cores, max(cores, 8) * 4)
    for i in range(max(0, min(a.warmup, 1))):          # one warm-up step is enough for a CPU loop (the pool is already spun up)
        arm.sample(n_genes, offset=i * n_genes)
    times = []
    for i in range(a.steps):
        v, dt = arm.sample(n_genes, offset=(i + 1) * n_genes)
        times.append(dt)
    arm.close()
    total = n_genes * n_perms * a.steps / sum(times)
    sample = ('%d genes x %d permutations of the same graph per step, one single-threaded process per core '
              '(threadpoolctl: 1 BLAS thread each), oracle/conn_oracle.py (numpy restatement of ref:966-992)'
              % (n_genes, n_perms))
    mapper_leg = None
    if not a.no_mapper:
        try:
            mapper_leg = mapper_cpu_sample()
        except Exception as e:                          # the conn line must not be lost to the secondary leg
            mapper_leg = {'error': repr(e)}
    line = {'metric': METRIC, 'value': total, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': 1e3 * sum(times) / a.steps, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'impl': 'reference',
            'config': workload_config(a, a.gpus),
            'cpu_baseline': {'value': total, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample,
                             'blas_threads_per_worker': 1},
            'e2e': {'value': total, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0, 'mapper': mapper_leg}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# Mapper build (BASELINE.json configs[1]: 20k cells x 5k genes, correlation metric, PCA lens,
# resolution 30, gain 3, one B200) — reported inside the same JSON line under "mapper"
# ---------------------------------------------------------------------------------------------
def fp64_probe_tflops(dev):
    """cuBLAS DGEMM yard-stick for the FP64 tensor pipe (MEASURED_PEAKS.json has no FP64 figure):
    torch.matmul on 6144^3 float64, best of 5, CUDA events."""
    import torch
    n = 6144
    A = torch.randn((n, n), dtype=torch.float64, device=dev)
    B = torch.randn((n, n), dtype=torch.float64, device=dev)
    best = 0.0
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(A, B)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def executed_tile_flops(sizes, G, tile):
    """Flops of the tile x tile output tiles gemm_kernel executes for symmetric patch Gram matrices (tile =
    sctda_gemm_tile_fp64()): the upper triangle of tiles only (T(T+1)/2 tiles of a patch of ceil(m/tile) = T tile
    rows), every tile counted in full (edge tiles compute their padding)."""
    T = (numpy.asarray(sizes, dtype=numpy.int64) + tile - 1) // tile
    return 2.0 * G * float(tile * tile) * float((T * (T + 1) // 2).sum())


def mapper_build_record(a, ctx, dev, N, G, metric, r, gain, steps, warmup, rank, world, peak, with_e2e):
    """Times `steps` Mapper builds (normalise -> cover -> per-patch clustering -> nodes -> nerve, inputs
    resident; patches sharded over the ranks when world > 1) and returns the record."""
    import torch
    import torch.distributed as dist
    from sctda_b200 import mapper, synth
    X = synth.expression_genes_torch(G, N, seed=synth.SEED + 11, device=dev).T.contiguous()      # cells x genes
    torch.cuda.synchronize()
    mapper.pca_lens_device(X, ctx=ctx)                       # warm-up (workspace allocation)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lens, lens_info = mapper.pca_lens_device(X, ctx=ctx, return_info=True)      # own kernels (csrc/pca.cu), every rank
    torch.cuda.synchronize()
    lens_ms = 1e3 * (time.perf_counter() - t0)
    if world > 1:
        dist.broadcast(lens, src=0)                         # every rank must cut the same cover

    def build(prof):
        if world > 1:
            return mapper.build_graph_distributed(ctx, X, lens, r, gain, metric=metric, profile=prof)
        return mapper.build_graph_device(ctx, X, lens, r, gain, metric=metric, profile=prof)

    def timed(n):
        times, gemm, prof, res = [], [], {}, None
        for _ in range(n):
            prof = {}
            if world > 1:
                dist.barrier()
            res = build(prof)
            t = torch.tensor([prof['total_ms']], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times.append(float(t.item()))
            gemm.append(prof['gemm_ms'])
        return times, gemm, prof, res
    for _ in range(max(1, warmup)):
        build({})
    times, gemm, prof, res = timed(steps)
    bin_off = res.bin_off.cpu().numpy().astype(numpy.int64)
    m = bin_off[1:] - bin_off[:-1]
    sum_m2 = float((m * m).sum())
    flops = 2.0 * G * sum_m2 / world                          # this rank's share (LPT balances m^2)
    flops_exec = executed_tile_flops(m, G, ctx.gemm_tile_fp64()) / world
    # the build has one host synchronisation (patch sizes come back to plan the batches): on a busy
    # host single builds are delayed by tens of ms, so the figure is the MEDIAN build; every build is listed
    ms = float(numpy.median(times))
    gms = float(numpy.median(gemm))
    out = {'metric': 'Mapper build cells/s', 'value': N / (ms * 1e-3), 'unit': 'cells/s', 'ms_per_build': ms,
           'ms_builds': [round(x, 2) for x in times], 'ms_per_build_mean': float(numpy.mean(times)), 'n_gpus': world,
           'config': {'workload': 'Mapper build: %d cells x %d genes, %s metric, PCA lens, resolution %d, gain %g, '
                                  'max_K 5, %d B200%s' % (N, G, metric, r, gain, world,
                                                          ', patches sharded by descending m^2' if world > 1 else ''),
                      'cover_multiplicity': float(prof['Mtot']) / N, 'nodes': prof['V'], 'edges': prof['E'],
                      'largest_patch': int(m.max()), 'sum_m2': sum_m2},
           'stages_ms': {k[:-3]: round(float(v), 3) for k, v in prof.items()
                         if k.endswith('_ms') and k not in ('total_ms', 'gemm_ms', 'host_ms') and isinstance(v, (int, float))},
           'pca_lens_ms': lens_ms, 'pca_lens': lens_info, 'value_with_lens': N / ((ms + lens_ms) * 1e-3),
           'host_ms': prof.get('host_ms'),
           'roofline': {'bound': 'tensor', 'achieved': flops / (gms * 1e-3) / 1e12, 'peak': peak, 'unit': 'TFLOP/s',
                        'frac': flops / (gms * 1e-3) / 1e12 / peak,
                        'achieved_executed': flops_exec / (gms * 1e-3) / 1e12,
                        'frac_executed': flops_exec / (gms * 1e-3) / 1e12 / peak,
                        'traffic': None, 'kernel': 'gemm_kernel (FP64 DMMA)', 'kernel_ms': gms, 'flops': flops,
                        'flops_executed': flops_exec,
                        'conventions': '`frac`: SURVEY.md 8d\'s algorithmic 2*G*sum m_b^2, no symmetry credit (can exceed 1); '
                                       '`frac_executed`: the output tiles (sctda_gemm_tile_fp64 squared) the kernel runs (upper triangle of every '
                                       'patch Gram tile, mirrored on the way out) -- the fraction of the tensor pipe '
                                       'actually used',
                        'peak_source': 'torch.matmul float64 6144^3 (cuBLAS DGEMM) probe in this run'}}
    # the opt-in split-precision contraction (sctda_set_precision(ctx, 1): tcgen05 kind::f16 on (hi, lo)
    # half-precision operand pairs, distances within 1e-5) on the same build; the headline above stays
    # the FP64 build whose graph is identical to the oracle's
    with ctx.precision('fp16x3'):
        build({})
        t2, g2, prof2, res2 = timed(max(2, steps // 2))
    ms2, gms2 = float(numpy.median(t2)), float(numpy.median(g2))
    same = bool(res2.V == res.V and res2.E == res.E and res2.node_members.shape == res.node_members.shape and
                (res2.node_members == res.node_members).all().item() and (res2.edges == res.edges).all().item())
    out['opt_in_fp16x3'] = {'value': N / (ms2 * 1e-3), 'unit': 'cells/s', 'ms_per_build': ms2, 'ms_builds': [round(x, 2) for x in t2], 'kernel': 'gemm_split3_kernel '
                            '(tcgen05.mma kind::f16, fp32 TMEM accumulators drained into float64 per 512 features)',
                            'kernel_ms': gms2, 'algorithmic_tflops': flops / (gms2 * 1e-3) / 1e12,
                            'tensor_tflops_executed': 'about 3/2 of algorithmic: three products over the upper-triangular tiles',
                            'peak': 2250.0, 'peak_source': 'nominal dense fp16 (B200_PROFILING.md)',
                            'frac_of_fp16_peak': 1.5 * flops / (gms2 * 1e-3) / 1e12 / 2250.0,
                            'stated_tolerance': '1e-5 relative to |x||y| on Gram entries (tests/test_mapper_gpu.py)',
                            'graph_identical_to_fp64_build': same, 'nodes': int(res2.V), 'edges': int(res2.E)}
    # end to end through the reference-facing call: host table and host lens in, networkx graph,
    # node lists and patches out (H2D of the table, D2H of the CSR / edges, graph object built on the host)
    if with_e2e and world == 1:
        Xh = X.cpu().numpy()
        lens_h = lens.cpu().numpy()
        mapper.mapper_graph(Xh[:2000], lens_h[:2000], r, gain, metric=metric, verbose=False)       # warm-up
        t0 = time.perf_counter()
        Gx, clusters, _ = mapper.mapper_graph(Xh, lens_h, r, gain, metric=metric, verbose=False)
        dt = time.perf_counter() - t0
        out['e2e'] = {'value': N / dt, 'unit': 'cells/s', 'ms': 1e3 * dt, 'h2d_bytes': int(Xh.nbytes + lens_h.nbytes),
                      'd2h_bytes': int(4 * (2 * prof['Mtot'] + 2 * prof['E'] + prof['V'] + r * r + 2)),
                      'api': 'sctda_b200.mapper.mapper_graph(host table, host lens) -> (networkx.Graph, all_clusters, patches)'}
        assert Gx.number_of_nodes() == prof['V'] and Gx.number_of_edges() == prof['E']
        del Xh
    del X
    torch.cuda.empty_cache()
    return out


def run_mapper(a, ctx, dev, steps, warmup, rank=0, world=1):
    """BASELINE.json's Mapper metric: configs[2], 100,000 cells x 2,000 genes (euclidean, r=30, gain 3), at
    every N: one GPU builds it alone, N > 1 shard the per-patch clustering by descending m^2 (labels
    exchanged by one all_gather, the rest replicated).  At N = 1 configs[1] (20k x 5k, correlation) is
    reported as `config2`."""
    peak = fp64_probe_tflops(dev)
    r, gain = a.mapper_resolution, a.mapper_gain
    out = mapper_build_record(a, ctx, dev, 100000, 2000, 'euclidean', r, gain, steps, warmup, rank, world, peak, True)
    if world == 1:
        out['config2'] = mapper_build_record(a, ctx, dev, a.mapper_cells, a.mapper_genes, a.mapper_metric, r, gain,
                                             max(2, steps // 2), 1, rank, world, peak, False)
    return out


# ---------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------
def run_config5(a, ctx, dev, rank, world):
    """BASELINE.json configs[4]: RootedGraph on a 50,000-cell Mapper graph through the reference-facing classes --
    root finding (all-source BFS + Spearman scores on the device), root-distance pseudo-time regression, then the
    connectivity test with the centroid / dispersion columns for every gene (RootedGraph.table, n = 500 shared
    permutations), the genes sharded over the ranks inside UnrootedGraph.connectivity_pvalues.  Graph and table
    are written once by rank 0 in the reference's formats (the table through the binary side-car)."""
    import shutil
    import tempfile
    import torch
    import torch.distributed as dist
    from sctda_b200 import formats, graph, mapper, synth
    N, G, n_perm, r, gain = 50000, 2000, 500, 30, 3.0
    tmp = [None]
    if rank == 0:
        tmp[0] = tempfile.mkdtemp(prefix='sctda_c5_')
    if world > 1:
        dist.broadcast_object_list(tmp, src=0)
    base = os.path.join(tmp[0], 'c5')
    out = {'config': {'workload': 'RootedGraph on a %d-cell Mapper graph (euclidean, PCA lens, resolution %d, gain %g), '
                                  '%d genes + time point column, %d shared permutations, genes sharded over %d B200'
                                  % (N, r, gain, G, n_perm, world)}}
    t, _ = synth._cell_latents(N, synth.SEED)
    if rank == 0:
        Yg = synth.expression_genes_torch(G, N, seed=synth.SEED, device=dev)          # gene-major [G, N]
        X = Yg.t().contiguous()
        lens = mapper.pca_lens_device(X).cpu().numpy()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        Gx, clusters, _ = mapper.mapper_graph(X.cpu().numpy(), lens, r, gain, metric='euclidean', verbose=False)
        torch.cuda.synchronize()
        out['mapper_graph_s'] = time.perf_counter() - t0
        out['nodes'], out['edges'] = Gx.number_of_nodes(), Gx.number_of_edges()
        formats.write_mapper_json(base + '_g.json', clusters)
        formats.write_gexf(base + '_g.gexf', Gx.number_of_nodes(), sorted(Gx.edges()))
        names = ['ID', 'timepoint', 'lib'] + ['G%05d' % i for i in range(G)]
        with open(base + '.all.tsv', 'w') as f:
            f.write('\t'.join(names) + '\n')
        tp = numpy.floor(t * 5.0)
        Y = numpy.empty((len(names), N), dtype=numpy.float64)
        Y[0] = 0.0
        Y[1] = tp
        Y[2] = 0.0
        Y[3:] = Yg.cpu().numpy()
        cdr = (Y[3:] > 0).sum(axis=0) / float(G)
        formats.write_table_sidecar(base + '.all.tsv', names, Y, ['D%d_%d' % (int(tp[i]), i) for i in range(N)],
                                    ['L'] * N, cdr)
        del Y, Yg, X
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    c = graph.RootedGraph(base + '_g', base + '.all.tsv', groups=False, device=dev.index)
    torch.cuda.synchronize()
    init_s = time.perf_counter() - t0
    dg = c._device_graph()
    perms = synth.permutation_block_torch(n_perm, int(dg.S), seed=7, device=dev).cpu().numpy()
    c.table(n=n_perm, perms=perms)                                                    # warm-up (workspaces, stats)
    times = []
    for _ in range(3):
        c._stats = None
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        header, rows = c.table(n=n_perm, perms=perms)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    tt = torch.tensor([float(numpy.median(times)), init_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    table_s, init_max = float(tt[0].item()), float(tt[1].item())
    cen = numpy.array([row[8] for row in rows if row[8] is not None], dtype=numpy.float64)
    out.update({'metric': 'RootedGraph all-gene table: gene*perms/s', 'value': len(rows) * n_perm / table_s,
                'unit': 'gene*perms/s', 'n_gpus': world, 'table_s': table_s, 'table_s_runs': [round(x, 4) for x in times],
                'rooted_init_s': init_max, 'rows': len(rows), 'nodes_lcc': len(c.pl), 'cells_lcc': int(dg.S),
                'memberships': int(dg.Mtot), 'root': str(c.root), 'leaf': str(c.leaf),
                'skeleton': {'levels': len(c.dicdend), 'nodes': c.g3.number_of_nodes(), 'edges': c.g3.number_of_edges(),
                             'path_edges': len(c.select_diff_path())},
                'pseudotime_regression': {'slope': float(c.po[0]), 'intercept': float(c.po[1]), 'r': float(c.po[2])},
                'centroid_range': [float(cen.min()), float(cen.max())] if cen.size else None,
                'significant_q05': int(sum(1 for row in rows if row[7] <= 0.05)),
                'includes': 'node statistics of every table column (get_gene / delta / expr / connectivity / centroid sums), '
                            'the permutation test of the kept genes (genes sharded, counts all-gathered), BH, the rows; '
                            'rooted_init_s (untimed in value): files -> graph, all-source BFS + Spearman root scores, '
                            'skeleton, regression'})
    if world > 1:
        dist.barrier()
    if rank == 0:
        shutil.rmtree(tmp[0], ignore_errors=True)
    return out


def run_own(a):
    import torch
    import torch.distributed as dist
    from sctda_b200 import _lib, graph, synth
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a B200; there is no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    ctx = _lib.context(local)
    ctx.set_kernel_timing(True)

    # ---- inputs, resident in HBM before the timed region ----
    from sctda_b200 import geneorder, parallel
    members, adj, _ = synth.connectivity_case(a.cells, a.nodes, 1, seed=synth.SEED)
    dg = graph.DeviceGraph(members, adj, numpy.arange(a.cells), a.cells, local)
    shard_pos = [parallel.tile_shard_positions(a.genes, r, world) for r in range(world)]   # positions of the planned order
    G = int(shard_pos[rank].size)
    Gmax = max(int(p.size) for p in shard_pos)                    # the largest shard (gather buffers are padded to it)
    g_lo = (a.genes * rank) // world                              # one rank: the whole table
    perms = synth.permutation_block_torch(a.perms, a.cells, seed=synth.SEED + 2, device=dev)
    # The tile layout of the table (a plan, like the CSR of the graph: made once per table): genes with similar
    # supports share a 128-gene tile.  With N ranks the plan is made on the WHOLE table (rank 0, broadcast) and a
    # rank takes every N-th 128-gene tile of the planned order (parallel.tile_shard_positions), so that its tiles are
    # as homogeneous as on one GPU and every rank gets the same mix of sparse and dense tiles (a shard planned on its
    # own 1/N of the genes reaches 0.50 instead of 0.57 of the roofline at N = 8; contiguous ranges of the planned
    # order leave the dense end to one rank).  Every rank builds the synthetic table it would otherwise load from
    # the file.
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if world > 1:
        Yfull = synth.expression_genes_torch(a.genes, a.cells, seed=synth.SEED, device=dev)
        plan = geneorder.gene_order(Yfull) if rank == 0 else torch.empty(a.genes, dtype=torch.int32, device=dev)
        dist.broadcast(plan, src=0)
        mine = plan.long().index_select(0, torch.from_numpy(shard_pos[rank]).to(dev))
        Y = Yfull.index_select(0, mine)
        del Yfull
        torch.cuda.empty_cache()
        order = None                                              # the shard is already in planned order
        unplan = plan.long().index_select(0, torch.from_numpy(numpy.concatenate(shard_pos)).to(dev))   # gathered slot -> gene
    else:
        Y = synth.expression_genes_torch(G, a.cells, seed=synth.SEED, device=dev, gene_offset=g_lo)
        order = geneorder.gene_order(Y)
        unplan = None
    torch.cuda.synchronize()
    plan_ms = 1e3 * (time.perf_counter() - t0)
    observed = graph.node_stats(ctx, dg, Y, log2=True)['conn']
    torch.cuda.synchronize()
    if order is not None:
        sector_density = geneorder.sector_density(Y[:4096], geneorder.gene_order(Y[:4096]))
    else:
        sector_density = geneorder.sector_density(Y[:4096])
    sector_density_natural = geneorder.sector_density(Y[:4096]) if order is not None else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gathered = [None]

    def step():
        """One pass of the test over this rank's gene shard, then the final gather ("... to
        exceed counts for all genes on device 0"): the only collective of the path."""
        ex, ties, _ = graph.conn_perm_test(ctx, dg, Y, perms, observed, log2=True, gene_order=order)
        g = parallel.gather_counts_device(ctx, ex, Gmax)                # sctda_gather_results (NCCL on device buffers)
        if unplan is not None:                                          # counts back at their gene positions
            flat = torch.cat([g[r, :int(shard_pos[r].size)] for r in range(world)])
            g = torch.empty_like(flat).index_copy_(0, unplan, flat)
        gathered[0] = g
        return ex, ties

    for _ in range(a.warmup):
        step()
    barrier()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    with ClockSampler(local) as clocks:
        barrier()
        e0.record()
        for _ in range(a.steps):
            ex, ties = step()
            kernel_ms.append(ctx.last_kernel_ms())                # waits for this step's kernel (events recorded by the library)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    total_units = float(a.genes) * a.perms * a.steps
    value = total_units / (ms_max * 1e-3)

    # ---- e2e: pinned host buffers through the public host API ----
    e2e = None
    if not a.no_e2e:
        Yh = torch.empty(Y.shape, dtype=Y.dtype, pin_memory=True)
        Yh.copy_(Y)
        Ph = torch.empty(perms.shape, dtype=perms.dtype, pin_memory=True)
        Ph.copy_(perms)
        Oh = observed.cpu().pin_memory()
        del Y, perms
        torch.cuda.empty_cache()

        def e2e_step():
            counts = graph.conn_perm_test_host(ctx, dg, Yh, Ph, Oh, log2=True)
            if world > 1:                                          # the per-rank counts meet on every rank, then the host
                counts = parallel.gather_counts_device(ctx, counts.to(dev), Gmax).cpu()
            return counts
        e2e_step()                                                 # warm-up (allocations)
        barrier()
        e2e_steps = min(a.steps, 3)
        e0.record()
        for _ in range(e2e_steps):
            counts = e2e_step()
        e1.record()
        barrier()
        t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        assert int(counts.sum()) >= 0
        e2e = {'value': float(a.genes) * a.perms * e2e_steps / (float(t2.item()) * 1e-3), 'unit': UNIT, 'steps': e2e_steps,
               'h2d_bytes_per_step': int(Yh.numel() * 8 + Ph.numel() * 4 + Oh.numel() * 8),
               'd2h_bytes_per_step': int(G * 4), 'ms_per_step': float(t2.item()) / e2e_steps,
               'includes': 'H2D of the table, the permutation block and the observed scores; the tile layout '
                           '(gene order) of the uploaded table; the kernel; D2H of the counts; for N > 1 the gather on rank 0'}

    # ---- roofline of the dominant kernel ----
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            peak, peak_src = float(json.load(f)['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs (of measured)'
    else:
        peak, peak_src = 6650.0, 'B200_PROFILING.md fallback (of fallback)'
    bytes_per_unit = 4.0 * a.cells + 8.0 * a.cells / a.perms
    k_ms = float(numpy.mean(kernel_ms))                          # mean over all timed steps
    achieved = bytes_per_unit * G * a.perms / (k_ms * 1e-3) / 1e9
    traffic = None                                               # DRAM bytes of this launch, from the committed ncu capture
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        with open(tpath) as f:
            per_unit = json.load(f).get('conn_perm_wide_kernel_dram_bytes_per_gene_perm')
        if per_unit:
            traffic = per_unit * G * a.perms
    # on-chip view: with the permutation block shared, every gene*perm gathers 8*Mtot bytes of table
    # rows from L2 plus the float32 node values of the sparse product (model, matches ncu's
    # l1tex__m_xbar2l1tex_read_bytes to a few per cent); ceiling = measured L2 -> SM read bandwidth
    l2_bytes_per_unit = 8.0 * dg.Mtot * sector_density + 4.0 * dg.V + 4.0 * (dg.nnz + dg.V) + 4.0 * dg.Mtot / 64.0
    l2_peak = ctx.probe_l2_read_gbs()
    l2_ach = l2_bytes_per_unit * G * a.perms / (k_ms * 1e-3) / 1e9
    roof = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
            'traffic': traffic, 'traffic_note': 'bytes per launch = ncu dram__bytes_read+write per gene*perm (profiles/traffic.json) x units of this launch; algorithmic bytes per launch = algorithmic_bytes_per_unit x units',
            'kernel': 'conn_perm_wide_kernel', 'kernel_ms': k_ms, 'kernel_ms_steps': [round(x, 2) for x in kernel_ms], 'peak_source': peak_src,
            'algorithmic_bytes_per_unit': bytes_per_unit,
            'onchip': {'bound': 'l2', 'achieved': l2_ach, 'peak': l2_peak, 'unit': 'GB/s', 'frac': l2_ach / l2_peak,
                       'bytes_per_unit': l2_bytes_per_unit, 'peak_source': 'sctda_probe_l2_read_gbs in this run'},
            'note': 'permutation block shared by all genes: index traffic is amortised over the gene shard; the kernel gathers '
                    'only the 32-byte sectors of the transposed table that hold a non-zero (sector_density x 8*Mtot bytes per '
                    'gene*perm from L2) and is bound by the bytes it can keep in flight (register landing zone x L2 latency) '
                    'and by the LSU data pipe (one wavefront per quad of lanes with an active lane; DESIGN.md)'}

    mapper_line = None
    if not a.no_mapper:
        if not a.no_e2e:
            del Yh, Ph
        else:
            del Y, perms
        torch.cuda.empty_cache()
        mapper_line = run_mapper(a, ctx, dev, 5, 2, rank, world)    # 5 timed builds, median reported
        if rank == 0 and world == 1 and not a.no_cpu:
            mapper_line['cpu_baseline'] = mapper_cpu_sample()
    rooted_line = None
    if not a.no_rooted:
        torch.cuda.empty_cache()
        rooted_line = run_config5(a, ctx, dev, rank, world)
    if rank == 0:
        cpu = None
        if world == 1 and not a.no_cpu:
            arm = CpuArm(a, 1000, 1, 8)
            v, dt = arm.sample(4)
            cpu = {'value': v, 'unit': UNIT, 'cores': 1, 'kind': 'port', 'blas_threads_per_worker': 1,
                   'sample': '4 genes x 1,000 permutations of the same graph, one single-threaded process, %.1f s '
                             '(oracle/conn_oracle.py)' % dt}
        roof['sector_density'] = {'planned_order': sector_density, 'table_order': sector_density_natural,
                                  'note': 'share of 32-byte sectors (4 genes x 1 cell) of the transposed table that hold a '
                                          'non-zero, first 4,096 genes of this rank: the kernel requests only those'}
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps,
                'warmup': a.warmup, 'ms_per_step': ms_max / a.steps, 'higher_is_better': True, 'scaling': 'strong',
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(a, world),
                'clocks': clocks.summary(), 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roof,
                'cpu_baseline': cpu, 'plan_ms_untimed': plan_ms,
                'degenerate_genes_with_exact_ties': int((ties > 0).sum().item()), 'mapper': mapper_line,
                'rooted': rooted_line}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_own(args)
