"""Generate tests/golden/* by running the REFERENCE'S OWN SOURCE for hot path (b).

TEST INFRASTRUCTURE ONLY (runs in the build container, where /root/reference is
mounted; never on the GPU box, never from the product path).

The reference (/root/reference/scTDA/main.py) is Python 2 and imports packages
that are not installed here (sakmapper, numexpr, matplotlib_venn, pylab, requests).
This script does NOT copy any of it into the repo.  At run time it reads the source
text from /root/reference, applies the textual Py2->Py3 shim listed below, executes
it as a module and drives UnrootedGraph / RootedGraph on small synthetic inputs that
this script also writes.  Outputs go to tests/golden/ (a few hundred KB).

Shim (everything that differs from running the file under Python 2):
  S1  `print x` statements           -> print(x)                (textual)
  S2  xrange                          -> range                   (module global)
  S3  map                             -> eager list(map(...))    (module global; Py2 map is eager,
                                                                   ref:976 relies on it for the shuffle)
  S4  networkx 1.x API                -> connected_component_subgraphs, to_numpy_matrix provided
                                         on a proxy module (networkx 3 equivalents)
  S5  numexpr.evaluate                -> numpy evaluation of the three expressions used at
                                         ref:982,983,986 (sequential row sums; <= 1 ulp)
  S6  pickle                          -> dump/load are no-ops (layout files, viz only; Py3 pickle
                                         cannot write to the text-mode handles of ref:819-826)
  S7  pylab, matplotlib_venn, requests, sakmapper, mpl_toolkits.mplot3d -> inert stand-ins
  S8  numpy.infty                     -> numpy.inf               (textual; attribute removed in numpy 2)
  S9  dict.values() inside numpy.array -> list(dict.values())    (textual, ref:1118 only; a Py2 dict.values()
                                                                   is a list)
  S10 sch.dendrogram                  -> called with no_plot=True (matplotlib is not installed; the returned
                                         record does not depend on the plot)

Run:  python oracle/gen_golden.py          (rewrites tests/golden/)
"""
import builtins
import io
import json
import os
import re
import sys
import types
from unittest import mock

import networkx
import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = '/root/reference/scTDA/main.py'
GOLD = os.path.join(ROOT, 'tests', 'golden')
sys.path.insert(0, ROOT)


# ----------------------------------------------------------------------------------------------
# the shim
# ----------------------------------------------------------------------------------------------
def _py2_print_to_call(src):
    """S1: rewrite `print <expr>` statements (possibly continued over lines) as calls."""
    lines = src.split('\n')
    out = []
    i = 0
    pat = re.compile(r'^(\s*)print\s+(.*)$')
    while i < len(lines):
        m = pat.match(lines[i])
        if not m:
            out.append(lines[i])
            i += 1
            continue
        indent, expr = m.group(1), m.group(2)
        # gather continuation lines (backslash or open brackets)
        while expr.rstrip().endswith('\\') or (expr.count('(') + expr.count('[') >
                                               expr.count(')') + expr.count(']')):
            if expr.rstrip().endswith('\\'):
                expr = expr.rstrip()[:-1]
            i += 1
            expr += ' ' + lines[i].strip()
        out.append('%sprint(%s)' % (indent, expr))
        i += 1
    return '\n'.join(out)


class _Capture(object):
    """Records the einsum result of ref:989 so the golden file can hold the permuted scores."""
    last_einsum = None


def _numexpr_evaluate(expr):
    """S5: the three expressions the reference evaluates (ref:982, 983, 986)."""
    loc = sys._getframe(1).f_locals

    def seq_sum(a):
        acc = numpy.zeros(a.shape[0], dtype=numpy.float64)
        for c in range(a.shape[1]):
            acc += a[:, c]
        return acc
    if expr == 'sum(2**pk - 1, 1)':
        return seq_sum(2.0 ** loc['pk'] - 1.0)
    if expr == 'log1p(t1)':
        return numpy.log1p(loc['t1'])
    if expr == 'sum(pk, 1)':
        return seq_sum(numpy.asarray(loc['pk'], dtype=numpy.float64))
    raise NotImplementedError(expr)


class _NumpyProxy(object):
    """numpy with the einsum result of ref:989 recorded."""

    def __getattr__(self, name):
        return getattr(numpy, name)

    @staticmethod
    def einsum(*a, **k):
        r = numpy.einsum(*a, **k)
        _Capture.last_einsum = r
        return r


class _NetworkxProxy(object):
    """S4."""
    def __getattr__(self, name):
        return getattr(networkx, name)

    @staticmethod
    def connected_component_subgraphs(g):
        return (g.subgraph(c) for c in networkx.connected_components(g))

    @staticmethod
    def to_numpy_matrix(g, nodelist=None):
        return networkx.to_numpy_array(g, nodelist=list(nodelist) if nodelist is not None else None)


def load_reference():
    """Execute the reference source under the shim; returns the module object."""
    with io.open(REF, 'r', encoding='utf-8') as f:
        src = f.read()
    src = _py2_print_to_call(src)
    src = src.replace('numpy.infty', 'numpy.inf')                                # S8 (textual)
    s9 = 'self.get_gene(genis)[0].values() for genis in lista'
    assert src.count(s9) == 1
    src = src.replace(s9, 'list(self.get_gene(genis)[0].values()) for genis in lista')      # S9 (textual)
    # import the real third-party modules first so that they never see the stand-ins below
    import pandas, scipy.cluster.hierarchy, scipy.interpolate, scipy.optimize, scipy.signal  # noqa: F401
    import scipy.spatial.distance, scipy.stats, sklearn.cluster, sklearn.metrics.pairwise  # noqa: F401
    inert = ['matplotlib_venn', 'pylab', 'requests', 'sakmapper', 'mpl_toolkits',
             'mpl_toolkits.mplot3d']                                             # S7
    saved = {k: sys.modules.get(k) for k in inert + ['numexpr', 'pickle']}
    for k in inert:
        sys.modules[k] = mock.MagicMock(name=k)
    ne = types.ModuleType('numexpr')
    ne.evaluate = _numexpr_evaluate
    sys.modules['numexpr'] = ne
    pk = types.ModuleType('pickle')                                              # S6
    pk.dump = lambda *a, **k: None
    pk.load = lambda *a, **k: {}
    sys.modules['pickle'] = pk
    mod = types.ModuleType('scTDA_reference_under_shim')
    mod.__file__ = REF
    try:
        code = compile(src, REF, 'exec')
        exec(code, mod.__dict__)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    mod.xrange = range                                                           # S2
    mod.map = lambda f, *a: list(builtins.map(f, *a))                            # S3
    mod.numpy = _NumpyProxy()
    mod.networkx = _NetworkxProxy()
    mod.numexpr = ne
    mod.pickle = pk
    import scipy.cluster.hierarchy as real_sch

    class _Sch(object):                                                          # S10
        def __getattr__(self, name):
            return getattr(real_sch, name)

        @staticmethod
        def dendrogram(*a, **k):
            k['no_plot'] = True
            return real_sch.dendrogram(*a, **k)
    mod.sch = _Sch()
    return mod


# ----------------------------------------------------------------------------------------------
# small synthetic inputs in the reference's file formats
# ----------------------------------------------------------------------------------------------
def write_case(prefix, n_cells, n_genes, n_nodes, seed, extra_component=True):
    """Writes <prefix>.all.tsv, <prefix>.json, <prefix>.gexf; returns gene names."""
    from sctda_b200 import synth
    rng = numpy.random.RandomState(seed)
    X, t, _ = synth.expression_matrix(n_cells, n_genes, seed=seed)
    genes = ['G%03d' % i for i in range(n_genes)]
    # a constant gene, an all-zero gene and a single-cell gene: edge cases of the filter / ties
    X[:, 0] = 3.25
    X[:, 1] = 0.0
    X[:, 2] = 0.0
    X[n_cells // 3, 2] = 7.5
    timepoint = numpy.floor(t * 4.0)
    with open(prefix + '.all.tsv', 'w') as f:
        f.write('ID\ttimepoint\tlib\t' + '\t'.join(genes) + '\n')
        for i in range(n_cells):
            f.write('D%d_L%d_%d\t%d\tL%d\t' % (timepoint[i], i % 3, i, timepoint[i], i % 3))
            f.write('\t'.join(repr(float(numpy.float64('%.12g' % v))) for v in X[i]) + '\n')
    # Mapper-like cover of the pseudo-time axis: overlapping windows, 2 "clusters" per window
    order = numpy.argsort(t, kind='stable')
    win = max(4, (3 * n_cells) // n_nodes)
    step = max(1, win // 3)
    dic = {}
    k = 0
    lo = 0
    while lo < n_cells and k < n_nodes:
        cells = order[lo:lo + win]
        half = rng.rand(len(cells)) < 0.5
        for part in (cells[half], cells[~half]):
            if len(part) > 0:
                dic[str(k)] = [int(c) for c in part]
                k += 1
        lo += step
    if extra_component:
        # two small nodes sharing a cell that appears in no other node: a second component
        lone = [int(order[-1]), int(order[-2])]
        for key in list(dic.keys()):
            dic[key] = [c for c in dic[key] if c not in lone] or dic[key][:1]
        dic[str(k)] = [lone[0], lone[1]]
        dic[str(k + 1)] = [lone[1]]
        k += 2
    V = k
    sets = [set(dic[str(i)]) for i in range(V)]
    g = networkx.Graph()
    for i in range(V):
        g.add_node(i)
    for i in range(V):
        for j in range(i):
            if sets[i] & sets[j]:
                g.add_edge(i, j, weight=1.0)
    with open(prefix + '.json', 'w') as f:
        json.dump(dic, f)
    networkx.write_gexf(g, prefix + '.gexf')
    with open(prefix + '.gexf') as f:                                            # keep the file independent of the day it is made
        txt = re.sub(r'lastmodifieddate="[^"]*"', 'lastmodifieddate="2026-10-19"', f.read())
    with open(prefix + '.gexf', 'w') as f:
        f.write(txt)
    return genes


def run_case(ref, prefix, n_perm, seed, rooted):
    numpy.random.seed(seed)
    if rooted:
        c = ref.RootedGraph(prefix, prefix + '.all.tsv', groups=False)
    else:
        c = ref.UnrootedGraph(prefix, prefix + '.all.tsv', groups=False)
    out = {'pl': [str(x) for x in c.pl], 'samples': [int(x) for x in c.samples],
           'n_perm': n_perm, 'seed': seed, 'rooted': bool(rooted)}
    if rooted:
        out['root'] = str(c.root)
        out['leaf'] = str(c.leaf)
        out['po'] = [float(c.po[0]), float(c.po[1])]
        out['dicdis'] = {str(k): int(v) for k, v in c.dicdis.items()}
        out['dendrite'] = {str(k): float(v) for k, v in c.get_dendrite().items()}
    # per-gene direct calls with a private seed, scores captured
    genes = sorted(c.dicgenes.keys())
    per_gene = {}
    for gi in genes[:12]:
        numpy.random.seed(seed + 7)
        p = c.connectivity_pvalue(gi, n=n_perm)
        scores = (float(len(c.pl)) / float(len(c.pl) - 1)) * _Capture.last_einsum
        dic, tol = c.get_gene(gi)
        per_gene[str(gi)] = {
            'p_value': float(p),
            'scores': [float(x) for x in scores],
            'connectivity': float(c.connectivity(gi)),
            'delta': [float(x) for x in c.delta(gi)],
            'expr': int(c.expr(gi)),
            'tol': float(tol),
            'p_nodes': [float(dic[k]) for k in c.pl],
        }
        if rooted:
            cen = c.centroid(gi)
            per_gene[str(gi)]['centroid'] = [None if v is None else float(v) for v in cen]
    out['per_gene'] = per_gene
    # the lib_ / ID_ / _CDR colourings of get_gene and count_gene (ref:902-928, 1347-1383)
    col = {}
    for key in ('lib_L1', 'lib_L2', 'ID_' + str(c.cellID[5]), '_CDR'):
        dic, tol = c.get_gene(key)
        col[key] = {'tol': float(tol), 'p_nodes': [float(dic[k]) for k in c.pl]}
    dic, tol = c.count_gene('timepoint', 2.0)
    col['count timepoint == 2'] = {'tol': float(tol), 'p_nodes': [float(dic[k]) for k in c.pl]}
    out['colourings'] = col
    # gene x gene matrices over the nodes (ref:1111-1190) on a few genes (the constant, the all-zero
    # and the single-cell gene are left out: their profiles divide by zero in the reference)
    lista = genes[3:11]
    with numpy.errstate(invalid='ignore', divide='ignore'):
        out['gene_matrices'] = {
            'genes': [str(g) for g in lista],
            'jsd': [[float(v) for v in row] for row in c.JSD_matrix(lista)],
            'cor': [[float(v) for v in row] for row in c.cor_matrix(lista)],
            'adjacency': [[float(v) for v in row] for row in c.adjacency_matrix(lista)],
            'adjacency_2': [[float(v) for v in row] for row in c.adjacency_matrix(lista, ind=2)],
        }
    if rooted:
        # the skeleton of ref:1499-1556 and the path of ref:1569-1593
        out['skeleton'] = {
            'nodes': [str(x) for x in c.g3.nodes()],
            'edges': [[str(a), str(b)] for a, b in c.g3.edges()],
            'dicdend': {str(n): [sorted(str(x) for x in comp) for comp in comps] for n, comps in c.dicdend.items()},
            'edgesize': [int(x) for x in c.edgesize],
            'nodesize': [int(x) for x in c.nodesize],
            'select_diff_path': [[str(a), str(b)] for a, b in c.select_diff_path()],
        }
    # the table writer, RNG stream consumed gene by gene from `seed`
    numpy.random.seed(seed)
    c.save(n=n_perm, filtercells=0, filterexp=0.0, annotation={'setA': genes[3:9], 'setB': genes[5:6]})
    # ref:1423-1446 on the table just written (every gene expressed in more than 5 cells: the q-value bound is
    # lifted so that the small case has genes to cluster)
    with numpy.errstate(invalid='ignore', divide='ignore'):
        if rooted:                                                               # ref:1950-2040, the 'js' route
            cl = c.cellular_subpopulations(1e9, threshold=1.5, min_cells=5, method='js', clus_thres=0.65)
        else:
            cl = c.cellular_subpopulations(threshold=1.5, min_cells=5, clus_thres=0.65)
        out['cellular_subpopulations'] = {'threshold': 1.5, 'min_cells': 5, 'clus_thres': 0.65, 'clusters': cl}
    os.replace(prefix + '.genes.tsv', prefix + '.ref.genes.tsv')
    with open(prefix + '.ref.json', 'w') as f:
        json.dump(out, f)


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref = load_reference()
    # global helper: BH on a fixed vector with ties
    rs = numpy.random.RandomState(5)
    pv = [float(x) for x in numpy.round(rs.rand(40), 2)] + [0.0, 0.0, 1.0, 0.5, 0.5]
    with open(os.path.join(GOLD, 'bh.ref.json'), 'w') as f:
        json.dump({'p': pv, 'q': [float(x) for x in ref.benjamini_hochberg(pv)]}, f)
    cases = [
        ('unrooted_a', dict(n_cells=240, n_genes=24, n_nodes=60, seed=11), 64, 12345, False),
        ('rooted_b', dict(n_cells=180, n_genes=16, n_nodes=40, seed=23), 48, 777, True),
    ]
    for name, kw, n_perm, seed, rooted in cases:
        prefix = os.path.join(GOLD, name)
        write_case(prefix, **kw)
        run_case(ref, prefix, n_perm, seed, rooted)
        for ext in ('.posg', '.posgl', '.genes.tsv'):                            # by-products of the reference's constructor
            if os.path.exists(prefix + ext):
                os.remove(prefix + ext)
        print('golden case', name, 'written')


if __name__ == '__main__':
    main()
