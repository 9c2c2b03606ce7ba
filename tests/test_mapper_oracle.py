"""The Mapper oracle (oracle/mapper_oracle.py) against what can be checked without sakmapper
(absent third-party dependency: parity with the reference is UNPINNED for this half): scipy's
single linkage, sklearn's Davies-Bouldin index, brute-force cover and nerve, and the call-site
contract of /root/reference/scTDA/main.py:768-777.  CPU only."""
import collections

import numpy
import pytest

from oracle import mapper_oracle as mo
from sctda_b200 import synth


@pytest.fixture(scope='module')
def data():
    X, t, _ = synth.expression_matrix(500, 120, seed=7)
    from sklearn.decomposition import PCA
    lens = PCA(n_components=2).fit_transform(X)
    return X, lens


def _partition(labels):
    d = collections.defaultdict(list)
    for i, l in enumerate(labels):
        d[int(l)].append(i)
    return sorted(d.values())


def test_gram_is_the_fma_chain_and_matches_blas(data):
    X, _ = data
    Xn, sq = mo.row_normalize(X, 'euclidean')
    rows = numpy.array([5, 3, 3, 499, 0])
    G = mo.gram(Xn, rows, rows)
    numpy.testing.assert_allclose(G, X[rows] @ X[rows].T, rtol=1e-13)
    assert (G == G.T).all()                                   # products commute: bitwise symmetric
    Xc, sqc = mo.row_normalize(X, 'correlation')
    numpy.testing.assert_allclose(sqc, 1.0, rtol=1e-14)
    numpy.testing.assert_allclose(1.0 - mo.gram(Xc, rows, rows), 1.0 - numpy.corrcoef(X[rows]), atol=1e-13)
    # constant row -> zeros
    Y = X.copy()
    Y[4] = 2.0
    Yc, s = mo.row_normalize(Y, 'correlation')
    assert (Yc[4] == 0.0).all() and s[4] == 0.0


def test_pairdist_matches_scipy(data):
    import scipy.spatial.distance as ssd
    X, _ = data
    for metric in ('euclidean', 'correlation'):
        Xn, sq = mo.row_normalize(X[:60], metric)
        D = mo.pairdist_rows(Xn, sq, 10, 30, metric)
        ref = ssd.cdist(X[10:30], X[:60], metric)
        numpy.testing.assert_allclose(D, ref, rtol=1e-9, atol=1e-12)
        assert (D[numpy.arange(20), numpy.arange(10, 30)] == 0.0).all()


@pytest.mark.parametrize('equalize', [True, False])
@pytest.mark.parametrize('gain', [0.4, 2.0])
def test_cover_rule(data, equalize, gain):
    _, lens = data
    r = 6
    bins, (lo0, hi0, lo1, hi1) = mo.cover(lens, r, gain, equalize)
    assert len(bins) == r * r
    seen = numpy.zeros(len(lens), dtype=int)
    for b, cells in enumerate(bins):
        i, j = divmod(b, r)
        brute = [c for c in range(len(lens)) if lo0[i] < lens[c, 0] < hi0[i] and lo1[j] < lens[c, 1] < hi1[j]]
        assert list(cells) == brute
        seen[cells] += 1
    assert (seen >= 1).all()                                  # every cell is covered when gain > 0
    if equalize:
        posts = mo.fenceposts(lens[:, 0], r, True)
        inside = [(lens[:, 0] >= posts[k]) & (lens[:, 0] < posts[k + 1]) for k in range(r)]
        sizes = [int(x.sum()) for x in inside]
        assert max(sizes) - min(sizes) <= 2                   # equal-population intervals before widening
    # monotone in gain
    bins2, _ = mo.cover(lens, r, gain * 1.5, equalize)
    for a, b2 in zip(bins, bins2):
        assert set(a) <= set(b2)


def test_single_linkage_matches_scipy_and_db_matches_sklearn(data):
    import scipy.cluster.hierarchy as h
    import sklearn.metrics
    X, lens = data
    Xn, _ = mo.row_normalize(X, 'euclidean')
    bins, _ = mo.cover(lens, 4, 0.5, True)
    cells = max(bins, key=len)
    Gm = mo.gram(Xn, cells, cells)
    parent, w, order = mo.prim_mst(Gm)
    Z = h.linkage(Xn[cells], 'single')
    numpy.testing.assert_allclose(numpy.sort(numpy.sqrt(w[1:])), Z[:, 2], rtol=1e-9)   # MST weights = merge heights
    for K in range(2, 6):
        lab = mo._components(len(cells), parent, set(mo.heaviest_edges(w, order, K - 1)))
        assert _partition(lab) == _partition(h.fcluster(Z, K, 'maxclust'))
        assert [min(p) for p in _partition(lab)] == sorted(min(p) for p in _partition(lab))
        db = mo.db_index_from_gram(Gm, lab, numpy.arange(K))
        assert db == pytest.approx(sklearn.metrics.davies_bouldin_score(Xn[cells], lab), rel=1e-6)
    labels, K, scores = mo.cluster_bin(Gm, max_K=5)
    assert K == 2 + int(numpy.argmin(scores)) and len(scores) == 4
    # through the finest partition == directly
    heavy = mo.heaviest_edges(w, order, 4)
    fine = mo._components(len(cells), parent, set(heavy))
    lab3 = mo._components(len(cells), parent, set(heavy[:2]))
    cof = numpy.zeros(5, dtype=int)
    cof[fine] = lab3
    assert mo.db_index_from_gram(Gm, fine, cof) == pytest.approx(mo.db_index_from_gram(Gm, lab3, numpy.arange(3)), rel=1e-12)


def test_small_patches():
    rng = numpy.random.RandomState(0)
    X = rng.rand(7, 5)
    G1 = mo.gram(X, [0], [0])
    assert mo.cluster_bin(G1)[1] == 1
    G2 = mo.gram(X, [0, 1], [0, 1])
    lab, K, _ = mo.cluster_bin(G2)
    assert K == 2 and list(lab) == [0, 1]                     # a 2-cell patch always splits (K >= 2)
    G5 = mo.gram(X, range(5), range(5))
    assert mo.cluster_bin(G5, max_K=5)[1] == 2                # m <= 5 -> K_max = 2
    G7 = mo.gram(X, range(7), range(7))
    assert 2 <= mo.cluster_bin(G7, max_K=5)[1] <= 3           # K_max = min(7 // 2, 5)
    # duplicates: zero distances, ties broken by index
    Xd = numpy.vstack([X[:3], X[:3]])
    lab, K, _ = mo.cluster_bin(mo.gram(Xd, range(6), range(6)), max_K=5)
    assert lab[0] == lab[3] and lab[1] == lab[4] and lab[2] == lab[5]


def test_mapper_graph_contract(data, tmp_path, capsys):
    X, lens = data
    edges, clusters, patches = mo.mapper_graph(X, lens, 6, 0.4, verbose=True)
    out = capsys.readouterr().out.split('\n')
    assert out[0].startswith('total of ') and out[0].endswith(' patches required clustering')
    assert out[1] == 'this implies %d nodes in the mapper graph' % len(clusters)
    # nodes of a patch partition it; members ascending; every multi-cell patch yields >= 2 nodes
    total = sum(len(c) for c in clusters)
    assert total == sum(len(p) for p in patches.values())
    for c in clusters:
        assert len(c) > 0 and list(c) == sorted(c)
    # nerve == brute force, sorted, no self loops
    sets = [set(c.tolist()) for c in clusters]
    brute = [(i, j) for i in range(len(sets)) for j in range(i + 1, len(sets)) if sets[i] & sets[j]]
    assert edges == brute
    # ref:772-777: json + gexf round trip through the product writers and networkx
    import json
    import networkx
    from sctda_b200 import formats
    p = str(tmp_path / 'm')
    formats.write_mapper_json(p + '.json', clusters)
    formats.write_gexf(p + '.gexf', len(clusters), edges)
    g = networkx.read_gexf(p + '.gexf')
    with open(p + '.json') as f:
        dic = json.load(f)
    assert sorted(g.nodes(), key=int) == [str(k) for k in range(len(clusters))] == sorted(dic.keys(), key=int)
    assert g.number_of_edges() == len(edges)
    assert dic['0'] == [int(x) for x in clusters[0]]


def test_histgap_rule():
    rng = numpy.random.RandomState(3)
    A = rng.normal(0, 0.05, (20, 3))
    B = rng.normal(5, 0.05, (15, 3))
    X = numpy.vstack([A, B])
    lab, K, _ = mo.cluster_bin(mo.gram(X, range(35), range(35)), stat='histgap')
    assert K == 2 and set(lab[:20]) == {0} and set(lab[20:]) == {1}
    lab, K, _ = mo.cluster_bin(mo.gram(A, range(20), range(20)), stat='histgap', hist_bins=3)
    assert K >= 1
