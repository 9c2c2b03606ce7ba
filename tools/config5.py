"""BASELINE.json configs[4] at scale on one B200: RootedGraph on a 50k-cell Mapper graph --
root finding (all-source BFS + Spearman on the device), root-distance pseudo-time regression,
connectivity test + centroid / dispersion columns for every gene.

    python tools/config5.py [--cells 50000 --genes 2000 --perms 500 --resolution 30 --gain 3]

The table reaches the classes through the binary side-car of formats.read_table (SURVEY.md §8f row
2); the TSV next to it only carries the header.  The permutation block is shared by the genes
(`perms=`): the reference's per-gene numpy stream would need cells x perms x genes = 5e10 draws.
Prints one JSON line.
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cells', type=int, default=50000)
    ap.add_argument('--genes', type=int, default=2000)
    ap.add_argument('--perms', type=int, default=500)
    ap.add_argument('--resolution', type=int, default=30)
    ap.add_argument('--gain', type=float, default=3.0)
    a = ap.parse_args()
    import torch
    from sctda_b200 import formats, graph, mapper, synth
    dev = torch.device('cuda', 0)
    out = {'config': 'BASELINE.json configs[4]: RootedGraph on a %d-cell Mapper graph, %d genes, %d permutations, 1 B200'
                     % (a.cells, a.genes, a.perms)}
    t0 = time.perf_counter()
    X, t, _ = synth.expression_matrix(a.cells, a.genes, seed=synth.SEED)
    out['synth_s'] = time.perf_counter() - t0
    genes = ['G%05d' % i for i in range(a.genes)]
    tmp = tempfile.mkdtemp(prefix='sctda_c5_')
    base = os.path.join(tmp, 'c5')
    # ---- Mapper graph (path a) through the reference-facing call ----
    t0 = time.perf_counter()
    lens = mapper.pca_lens_device(torch.from_numpy(X).to(dev)).cpu().numpy()
    out['pca_lens_s'] = time.perf_counter() - t0
    t0 = time.perf_counter()
    G, clusters, _ = mapper.mapper_graph(X, lens, a.resolution, a.gain, metric='euclidean', verbose=False)
    torch.cuda.synchronize()
    out['mapper_graph_s'] = time.perf_counter() - t0
    formats.write_mapper_json(base + '_g.json', clusters)
    formats.write_gexf(base + '_g.gexf', G.number_of_nodes(), sorted(G.edges()))
    out['nodes'], out['edges'] = G.number_of_nodes(), G.number_of_edges()
    # ---- table: header-only TSV + binary side-car ----
    names = ['ID', 'timepoint', 'lib'] + genes
    with open(base + '.all.tsv', 'w') as f:
        f.write('\t'.join(names) + '\n')
    tp = numpy.floor(t * 5.0)
    Y = numpy.empty((len(names), a.cells), dtype=numpy.float64)
    Y[0] = 0.0
    Y[1] = tp
    Y[2] = 0.0
    Y[3:] = X.T
    cdr = (X > 0).sum(axis=1) / float(a.cells)
    formats.write_table_sidecar(base + '.all.tsv', names, Y, ['D%d_%d' % (int(tp[i]), i) for i in range(a.cells)],
                                ['L'] * a.cells, cdr)
    del Y
    # ---- RootedGraph (ref:1449-1567) ----
    t0 = time.perf_counter()
    c = graph.RootedGraph(base + '_g', base + '.all.tsv', groups=False)
    torch.cuda.synchronize()
    out['rooted_init_s'] = time.perf_counter() - t0
    out['nodes_lcc'] = len(c.pl)
    out['root'], out['leaf'] = str(c.root), str(c.leaf)
    out['pseudotime_regression'] = {'slope': float(c.po[0]), 'intercept': float(c.po[1]), 'r': float(c.po[2])}
    dg = c._device_graph()
    out['cells_lcc'], out['memberships'] = int(dg.S), int(dg.node_cols.numel())
    perms = synth.permutation_block_torch(a.perms, int(dg.S), seed=7, device=dev).cpu().numpy()
    t0 = time.perf_counter()
    c.save(n=a.perms, perms=perms)
    header, rows = c.last_table
    torch.cuda.synchronize()
    out['save_s'] = time.perf_counter() - t0
    out['rows'] = len(rows)
    out['gene_perms_per_s_e2e'] = len(rows) * a.perms / out['save_s']
    cen = numpy.array([r[8] for r in rows if r[8] is not None], dtype=numpy.float64)
    out['centroid_range'] = [float(cen.min()), float(cen.max())] if cen.size else None
    out['significant_q05'] = int(sum(1 for r in rows if r[7] <= 0.05))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
