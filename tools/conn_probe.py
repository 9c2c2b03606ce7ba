#!/usr/bin/env python
"""tools/conn_probe.py — the permutation kernel alone on a slice of the bench workload
(SURVEY.md 8d C4 graph), for kernel experiments: time per launch, units/s, roofline fraction,
and agreement of the counts / scores with another run (--save / --check).

  python tools/conn_probe.py --genes 2048 --perms 512 [--order pca|density|none] [--save f.npz] [--check f.npz]
  SCTDA_CONN_LEGACY=1 ... runs the 32-gene dense-row kernel, SCTDA_CONN64_WARPS=16|24|32 the CTA size."""
import argparse
import json
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--genes', type=int, default=2048)
    ap.add_argument('--perms', type=int, default=512)
    ap.add_argument('--cells', type=int, default=20000)
    ap.add_argument('--nodes', type=int, default=2000)
    ap.add_argument('--order', default='none')
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--save', default=None)
    ap.add_argument('--check', default=None)
    ap.add_argument('--scores', action='store_true')
    a = ap.parse_args()
    import torch
    from sctda_b200 import _lib, graph, synth, geneorder
    dev = torch.device('cuda', 0)
    ctx = _lib.context(0)
    ctx.set_kernel_timing(True)
    members, adj, _ = synth.connectivity_case(a.cells, a.nodes, 1, seed=synth.SEED)
    dg = graph.DeviceGraph(members, adj, numpy.arange(a.cells), a.cells, 0)
    Y = synth.expression_genes_torch(a.genes, a.cells, seed=synth.SEED, device=dev)
    perms = synth.permutation_block_torch(a.perms, a.cells, seed=synth.SEED + 2, device=dev)
    observed = graph.node_stats(ctx, dg, Y, log2=True)['conn']
    order = None
    t_order = 0.0
    if a.order != 'none':
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        kw = {k: int(os.environ[e]) for k, e in (('sample', 'GO_SAMPLE'), ('dims', 'GO_DIMS'), ('leaf', 'GO_LEAF')) if e in os.environ}
        order = geneorder.gene_order(Y, method=a.order, **kw)
        torch.cuda.synchronize()
        t_order = time.perf_counter() - t0
    ms = []
    for _ in range(a.reps):
        ex, ti, sc = graph.conn_perm_test(ctx, dg, Y, perms, observed, log2=True, gene_order=order,
                                          want_scores=a.scores)
        torch.cuda.synchronize()
        ms.append(ctx.last_kernel_ms())
    units = float(a.genes) * a.perms
    best = min(ms)
    out = {'genes': a.genes, 'perms': a.perms, 'order': a.order, 'legacy': bool(os.environ.get('SCTDA_CONN_LEGACY')),
           'warps': os.environ.get('SCTDA_CONN64_WARPS'), 'kernel_ms': [round(x, 3) for x in ms],
           'units_per_s': units / (best * 1e-3), 'frac_of_8.2e7': units / (best * 1e-3) / (6545.6e9 / 80016.0),
           'order_s': round(t_order, 3), 'exceed_sum': int(ex.sum().item()), 'tie_genes': int((ti > 0).sum().item())}
    if order is not None:
        nz = (Y != 0)
        o = order.long()
        G4 = (a.genes // 4) * 4
        out['sector_density'] = float(nz[o[:G4]].view(G4 // 4, 4, -1).any(dim=1).float().mean().item())
        out['sector_density_natural'] = float(nz[:G4].view(G4 // 4, 4, -1).any(dim=1).float().mean().item())
    if a.save:
        numpy.savez(a.save, ex=ex.cpu().numpy(), sc=sc.cpu().numpy() if sc is not None else numpy.zeros(0))
    if a.check:
        ref = numpy.load(a.check)
        out['counts_equal'] = bool((ref['ex'] == ex.cpu().numpy()).all())
        if sc is not None and ref['sc'].size:
            out['scores_bit_equal'] = bool((ref['sc'].view(numpy.int64) == sc.cpu().numpy().view(numpy.int64)).all())
            out['scores_max_rel'] = float(numpy.nanmax(numpy.abs(ref['sc'] - sc.cpu().numpy()) / numpy.abs(ref['sc'])))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
