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

JSON line: see the task contract.  `value` = device-resident throughput; `e2e` = the same
through graph.conn_perm_test_host with pinned HOST buffers (H2D of table + permutations and
D2H of the counts inside the timed region); `roofline` = the dominant kernel
(conn_perm_kernel) against the measured HBM copy bandwidth with SURVEY.md §8d's
algorithmic bytes (4*S + 8*S/n per gene·perm); `cpu_baseline` = the oracle port on a
bounded sample of the same workload.
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


def cpu_sample(a, n_genes, n_perms, procs):
    """Times the oracle on `n_genes` genes x `n_perms` permutations of the bench workload
    with `procs` worker processes.  Returns (gene*perms/s, seconds)."""
    import multiprocessing as mp
    from sctda_b200 import synth
    members, adj, _ = synth.connectivity_case(a.cells, a.nodes, 1, seed=synth.SEED)
    Y = synth.expression_genes_numpy(n_genes, a.cells, seed=synth.SEED)
    rng = numpy.random.RandomState(7)
    perms = numpy.stack([rng.permutation(a.cells) for _ in range(n_perms)]).astype(numpy.int32)
    dense = adj.toarray()
    if procs <= 1:
        _cpu_init(members, dense, Y, perms)
        t0 = time.perf_counter()
        for i in range(n_genes):
            _cpu_gene(i)
        dt = time.perf_counter() - t0
    else:
        ctx = mp.get_context('fork')
        _cpu_init(members, dense, Y, perms)            # inherited by the forked workers
        with ctx.Pool(procs) as pool:
            pool.map(_cpu_gene, range(min(procs, n_genes)))        # spin-up outside the timed region
            t0 = time.perf_counter()
            pool.map(_cpu_gene, range(n_genes), chunksize=1)
            dt = time.perf_counter() - t0
    return n_genes * n_perms / dt, dt


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_perms = 200
    n_genes = 2 * cores
    for _ in range(a.warmup):
        cpu_sample(a, min(cores, 4), 20, min(cores, 4))
    vals, times = [], []
    for _ in range(a.steps):
        v, dt = cpu_sample(a, n_genes, n_perms, cores)
        vals.append(v)
        times.append(dt)
    total = n_genes * n_perms * a.steps / sum(times)
    sample = ('%d genes x %d permutations of the same graph per step, one process per core, '
              'oracle/conn_oracle.py (numpy restatement of ref:966-992)' % (n_genes, n_perms))
    line = {'metric': METRIC, 'value': total, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': 1e3 * sum(times) / a.steps, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'impl': 'reference',
            'config': workload_config(a, a.gpus),
            'cpu_baseline': {'value': total, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': total, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
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


def run_mapper(a, ctx, dev, steps, warmup, rank=0, world=1):
    """world == 1: BASELINE.json configs[1] (20k x 5k, correlation, r=30, gain 3) on one GPU.
    world > 1: configs[2] (100k cells x 2k genes, euclidean), per-patch clustering sharded over the
    ranks by descending m^2, labels exchanged by one all_gather (NCCL), the rest replicated."""
    import torch
    import torch.distributed as dist
    from sctda_b200 import mapper, synth
    if world > 1:
        N, G, metric = 100000, 2000, 'euclidean'
    else:
        N, G, metric = a.mapper_cells, a.mapper_genes, a.mapper_metric
    r, gain = a.mapper_resolution, a.mapper_gain
    X = synth.expression_genes_torch(G, N, seed=synth.SEED + 11, device=dev).T.contiguous()      # cells x genes
    t0 = time.perf_counter()
    lens = mapper.pca_lens_device(X)
    if world > 1:
        dist.broadcast(lens, src=0)                         # every rank must cut the same cover
    torch.cuda.synchronize()
    lens_ms = 1e3 * (time.perf_counter() - t0)

    def build(prof):
        if world > 1:
            return mapper.build_graph_distributed(ctx, X, lens, r, gain, metric=metric, profile=prof)
        return mapper.build_graph_device(ctx, X, lens, r, gain, metric=metric, profile=prof)
    prof = {}
    for _ in range(max(1, warmup)):
        build(prof)
    times, gemm = [], []
    for _ in range(steps):
        prof = {}
        if world > 1:
            dist.barrier()
        res = build(prof)
        t = torch.tensor([prof['total_ms']], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
        gemm.append(prof['gemm_ms'])
    bin_off = res.bin_off.cpu().numpy().astype(numpy.int64)
    m = bin_off[1:] - bin_off[:-1]
    sum_m2 = float((m * m).sum())
    flops = 2.0 * G * sum_m2 / world                          # this rank's share (LPT balances m^2)
    # the build has one host synchronisation (patch sizes come back to plan the batches): on a busy
    # host single builds are delayed by tens of ms, so the figure is the MEDIAN build; every build is listed
    ms = float(numpy.median(times))
    gms = float(numpy.median(gemm))
    peak = fp64_probe_tflops(dev)
    out = {'metric': 'Mapper build cells/s', 'value': N / (ms * 1e-3), 'unit': 'cells/s', 'ms_per_build': ms, 'ms_builds': [round(x, 2) for x in times], 'ms_per_build_mean': float(numpy.mean(times)), 'n_gpus': world,
           'config': {'workload': 'Mapper build: %d cells x %d genes, %s metric, PCA lens, resolution %d, gain %g, '
                                  'max_K 5, %d B200%s' % (N, G, metric, r, gain, world,
                                                          ', patches sharded by descending m^2' if world > 1 else ''),
                      'cover_multiplicity': float(prof['Mtot']) / N, 'nodes': prof['V'], 'edges': prof['E'],
                      'largest_patch': int(m.max()), 'sum_m2': sum_m2},
           'stages_ms': {k[:-3]: round(float(v), 3) for k, v in prof.items()
                         if k.endswith('_ms') and k not in ('total_ms', 'gemm_ms', 'host_ms') and isinstance(v, (int, float))},
           'pca_lens_ms_untimed': lens_ms, 'host_ms': prof.get('host_ms'),
           'roofline': {'bound': 'tensor', 'achieved': flops / (gms * 1e-3) / 1e12, 'peak': peak, 'unit': 'TFLOP/s',
                        'frac': flops / (gms * 1e-3) / 1e12 / peak, 'traffic': None, 'kernel': 'gemm_kernel (FP64 DMMA)',
                        'kernel_ms': gms, 'flops': flops,
                        'executed_flops': 'upper-triangular 128x128 tiles only (the Gram tile is symmetric bit for bit): '
                                          'about half of `flops`, which is SURVEY.md 8d\'s algorithmic 2*G*sum m_b^2 '
                                          'without symmetry credit, so frac can exceed 1',
                        'peak_source': 'torch.matmul float64 6144^3 (cuBLAS DGEMM) probe in this run'}}
    # the opt-in split-precision contraction (sctda_set_precision(ctx, 1): tcgen05 kind::f16 on (hi, lo)
    # half-precision operand pairs, distances within 1e-5) on the same build; the headline above stays
    # the FP64 build whose graph is identical to the oracle's
    with ctx.precision('fp16x3'):
        build({})
        t2, g2 = [], []
        for _ in range(steps):
            prof2 = {}
            if world > 1:
                dist.barrier()
            res2 = build(prof2)
            t = torch.tensor([prof2['total_ms']], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t2.append(float(t.item()))
            g2.append(prof2['gemm_ms'])
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
    if world == 1:
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
    del X
    torch.cuda.empty_cache()
    return out


def mapper_cpu_sample():
    """The Mapper oracle on BASELINE.json configs[0] scale (1,000 cells x 2,000 genes, r=20, gain 2)."""
    from oracle import mapper_oracle as mo
    from sctda_b200 import synth
    X = synth.expression_genes_numpy(2000, 1000, seed=synth.SEED + 11).T.copy()
    Xc = X - X.mean(axis=0)
    u, s_, _ = numpy.linalg.svd(Xc, full_matrices=False)
    lens = u[:, :2] * s_[:2]
    t0 = time.perf_counter()
    edges, clusters, _ = mo.mapper_graph(X, lens, 20, 2.0, metric='correlation')
    dt = time.perf_counter() - t0
    return {'value': 1000 / dt, 'unit': 'cells/s', 'cores': os.cpu_count(), 'kind': 'port',
            'sample': 'oracle/mapper_oracle.py on 1,000 cells x 2,000 genes, r=20, gain 2 (%d nodes, %d edges), %.1f s; '
                      'Gram matrices by an OpenMP fma loop, the rest single-threaded numpy' % (len(clusters), len(edges), dt)}


# ---------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------
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
    members, adj, _ = synth.connectivity_case(a.cells, a.nodes, 1, seed=synth.SEED)
    dg = graph.DeviceGraph(members, adj, numpy.arange(a.cells), a.cells, local)
    g_lo = (a.genes * rank) // world
    g_hi = (a.genes * (rank + 1)) // world
    G = g_hi - g_lo
    Y = synth.expression_genes_torch(G, a.cells, seed=synth.SEED, device=dev, gene_offset=g_lo)
    perms = synth.permutation_block_torch(a.perms, a.cells, seed=synth.SEED + 2, device=dev)
    observed = graph.node_stats(ctx, dg, Y, log2=True)['conn']
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return graph.conn_perm_test(ctx, dg, Y, perms, observed, log2=True)

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
            ex, ties, _ = step()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    # the dominant kernel alone (events recorded by the library around conn_perm_kernel, last timed step)
    kernel_ms.append(ctx.last_kernel_ms())
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
        graph.conn_perm_test_host(ctx, dg, Yh, Ph, Oh, log2=True)       # warm-up (allocations)
        barrier()
        e2e_steps = min(a.steps, 2)
        e0.record()
        for _ in range(e2e_steps):
            counts = graph.conn_perm_test_host(ctx, dg, Yh, Ph, Oh, log2=True)
        e1.record()
        barrier()
        t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        assert int(counts.sum()) >= 0
        e2e = {'value': float(a.genes) * a.perms * e2e_steps / (float(t2.item()) * 1e-3), 'unit': UNIT, 'steps': e2e_steps,
               'h2d_bytes_per_step': int(Yh.numel() * 8 + Ph.numel() * 4 + Oh.numel() * 8),
               'd2h_bytes_per_step': int(counts.numel() * 4), 'ms_per_step': float(t2.item()) / e2e_steps}

    # ---- roofline of the dominant kernel ----
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            peak, peak_src = float(json.load(f)['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs (of measured)'
    else:
        peak, peak_src = 6650.0, 'B200_PROFILING.md fallback (of fallback)'
    bytes_per_unit = 4.0 * a.cells + 8.0 * a.cells / a.perms
    k_ms = float(numpy.mean(kernel_ms))
    achieved = bytes_per_unit * G * a.perms / (k_ms * 1e-3) / 1e9
    traffic = None                                               # DRAM bytes of this launch, from the committed ncu capture
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        with open(tpath) as f:
            per_unit = json.load(f).get('conn_perm_kernel_dram_bytes_per_gene_perm')
        if per_unit:
            traffic = per_unit * G * a.perms
    # on-chip view: with the permutation block shared, every gene*perm gathers 8*Mtot bytes of table
    # rows from L2 plus the float32 node values of the sparse product (model, matches ncu's
    # l1tex__m_xbar2l1tex_read_bytes to a few per cent); ceiling = measured L2 -> SM read bandwidth
    l2_bytes_per_unit = 8.0 * dg.Mtot + 4.0 * dg.V + 4.0 * (dg.nnz + dg.V) + 4.0 * dg.Mtot / 32.0
    l2_peak = ctx.probe_l2_read_gbs()
    l2_ach = l2_bytes_per_unit * G * a.perms / (k_ms * 1e-3) / 1e9
    roof = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
            'traffic': traffic, 'traffic_note': 'bytes per launch = ncu dram__bytes_read+write per gene*perm (profiles/traffic.json) x units of this launch; algorithmic bytes per launch = algorithmic_bytes_per_unit x units',
            'kernel': 'conn_perm_kernel', 'kernel_ms': k_ms, 'peak_source': peak_src,
            'algorithmic_bytes_per_unit': bytes_per_unit,
            'onchip': {'bound': 'l2', 'achieved': l2_ach, 'peak': l2_peak, 'unit': 'GB/s', 'frac': l2_ach / l2_peak,
                       'bytes_per_unit': l2_bytes_per_unit, 'peak_source': 'sctda_probe_l2_read_gbs in this run'},
            'note': 'permutation block shared by all genes: index traffic is amortised over the gene shard, '
                    'the binding resource is the L2->SM gather of 8*Mtot bytes per gene*perm (DESIGN.md)'}

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
    if rank == 0:
        cpu = None
        if world == 1 and not a.no_cpu:
            v, dt = cpu_sample(a, 8, 400, 1)
            cpu = {'value': v, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                   'sample': '8 genes x 400 permutations of the same graph, single process, %.1f s' % dt}
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps,
                'warmup': a.warmup, 'ms_per_step': ms_max / a.steps, 'higher_is_better': True, 'scaling': 'strong',
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(a, world),
                'clocks': clocks.summary(), 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roof,
                'cpu_baseline': cpu, 'tie_genes': int((ties > 0).sum().item()), 'mapper': mapper_line}
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
