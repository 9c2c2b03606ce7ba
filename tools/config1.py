"""BASELINE.json configs[0] end to end — the reference's own CPU-runnable case — through the
reference-facing classes, next to the CPU oracle on the same files:

    synthetic 1,000 cells x 2,000 genes -> .mapper.tsv / .all.tsv
    TopologicalRepresentation(lens='mds', metric='correlation').save(resolution=20, gain=2)
    UnrootedGraph(...).save(n=500)

    python tools/config1.py [--oracle] [--cells 1000 --genes 2000 --perms 500]

With --oracle the Mapper oracle and the connectivity oracle are run on the same lens / the same
numpy seed, the graphs are compared exactly and the gene tables row by row (p-values exact on
tie-free genes).  Prints one JSON line.  The oracle is the checker here, never the product.
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
    ap.add_argument('--cells', type=int, default=1000)
    ap.add_argument('--genes', type=int, default=2000)
    ap.add_argument('--perms', type=int, default=500)
    ap.add_argument('--resolution', type=int, default=20)
    ap.add_argument('--gain', type=float, default=2.0)
    ap.add_argument('--oracle', action='store_true')
    print(json.dumps(run(ap.parse_args())))


def run(a):
    """a: namespace with cells, genes, perms, resolution, gain, oracle.  Returns the record (a dict)."""
    import torch
    from sctda_b200 import formats, graph, mapper, synth
    X, t, _ = synth.expression_matrix(a.cells, a.genes, seed=synth.SEED)
    genes = ['G%05d' % i for i in range(a.genes)]
    tmp = tempfile.mkdtemp(prefix='sctda_c1_')
    base = os.path.join(tmp, 'c1')
    t0 = time.perf_counter()
    with open(base + '.mapper.tsv', 'w') as f:
        f.write('\t'.join(genes) + '\n')
        for row in X:
            f.write('\t'.join(formats.py2_str(v) for v in row) + '\n')
    with open(base + '.all.tsv', 'w') as f:
        f.write('ID\ttimepoint\tlib\t' + '\t'.join(genes) + '\n')
        for i, row in enumerate(X):
            f.write('D%d_L_%d\t%d\tL\t' % (int(t[i] * 4), i, int(t[i] * 4)) + '\t'.join(formats.py2_str(v) for v in row) + '\n')
    out = {'config': 'BASELINE.json configs[0]: %d cells x %d genes, mds lens, correlation, r=%d, gain=%g, n=%d'
                     % (a.cells, a.genes, a.resolution, a.gain, a.perms), 'write_tsv_s': time.perf_counter() - t0}
    # ---- product arm ----
    t0 = time.perf_counter()
    tr = mapper.TopologicalRepresentation(base, lens='mds', metric='correlation', random_state=0, n_init=1, max_iter=100)
    out['lens_s'] = time.perf_counter() - t0
    t0 = time.perf_counter()
    tr.save(base + '_g', a.resolution, a.gain)
    torch.cuda.synchronize()
    out['mapper_save_s'] = time.perf_counter() - t0
    t0 = time.perf_counter()
    c = graph.UnrootedGraph(base + '_g', base + '.all.tsv', groups=False)
    out['load_s'] = time.perf_counter() - t0
    numpy.random.seed(2017)
    t0 = time.perf_counter()
    c.save(n=a.perms)
    header, rows = c.last_table
    torch.cuda.synchronize()
    out['save_s'] = time.perf_counter() - t0
    out['nodes_lcc'] = len(c.pl)
    out['rows'] = len(rows)
    out['gene_perms_per_s_e2e'] = len(rows) * a.perms / out['save_s']
    out['cells_per_s_e2e'] = a.cells / out['mapper_save_s']
    if a.oracle:
        from oracle import conn_oracle as co
        from oracle import mapper_oracle as mo
        import pandas
        df = pandas.read_table(base + '.mapper.tsv')
        lens = tr.lens_data_mds.values
        t0 = time.perf_counter()
        edges, clusters, _ = mo.mapper_graph(df.values, lens, a.resolution, a.gain)
        out['oracle_mapper_s'] = time.perf_counter() - t0
        with open(base + '_g.json') as f:
            dic = json.load(f)
        out['graph_nodes_identical'] = dic == {str(k): [int(x) for x in cl] for k, cl in enumerate(clusters)}
        import networkx
        g = networkx.read_gexf(base + '_g.gexf')
        out['graph_edges_identical'] = sorted((min(int(u), int(v)), max(int(u), int(v))) for u, v in g.edges()) == edges
        st = co.load_state(base + '_g', base + '.all.tsv')
        st.pl = list(c.pl)                                   # same node order as the product (order-invariant up to rounding)
        st.adj = c.adj
        st.samples = numpy.array(c.samples)
        numpy.random.seed(2017)
        t0 = time.perf_counter()
        with numpy.errstate(invalid='ignore', divide='ignore'):
            h2, rows2 = co.save_rows(st, n=a.perms)
        out['oracle_save_s'] = time.perf_counter() - t0
        out['oracle_gene_perms_per_s'] = len(rows2) * a.perms / out['oracle_save_s']
        pv, ties = None, None
        numpy.random.seed(2017)
        keep = [r[0] for r in rows]
        pv, ties = c.connectivity_pvalues(keep, n=a.perms, return_ties=True)
        same_rows = len(rows) == len(rows2) and all(x[0] == y[0] and x[1] == y[1] for x, y in zip(rows, rows2))
        p_exact = sum(1 for x, y, tt in zip(rows, rows2, ties) if tt == 0 and x[6] == y[6])
        p_tied = int((numpy.asarray(ties) > 0).sum())
        p_bad = sum(1 for x, y, tt in zip(rows, rows2, ties) if tt == 0 and x[6] != y[6])
        rel = max(abs(x[5] - y[5]) / abs(y[5]) for x, y in zip(rows, rows2) if y[5] != 0)
        out.update({'table_rows_identical': same_rows, 'p_values_exact': p_exact, 'p_values_tied_genes': p_tied,
                    'p_values_wrong': p_bad, 'max_rel_err_connectivity': rel})
    return out


if __name__ == '__main__':
    main()
