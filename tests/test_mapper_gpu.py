"""Parity of the CUDA Mapper path (through the C-ABI) with the CPU oracle
(oracle/mapper_oracle.py; UNPINNED against the reference: sakmapper is absent).

Bars: row normalisation, Gram matrices and distances BIT-EXACT (float64, fixed operation order:
one fma chain over the features per element); cover memberships, cluster labels, node lists and
edges EXACT (integers).  Needs a B200: `pytest -m gpu`."""
import json
import os

import numpy
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')


def _gpu():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from sctda_b200 import _lib
    return _lib.context(0)


def _data(n, g, seed):
    from sctda_b200 import synth
    X, t, _ = synth.expression_matrix(n, g, seed=seed)
    Xc = X - X.mean(axis=0)
    u, s, vt = numpy.linalg.svd(Xc, full_matrices=False)
    return X, u[:, :2] * s[:2]


def _bits(a):
    return numpy.ascontiguousarray(a, dtype=numpy.float64).view(numpy.int64)


@pytest.mark.parametrize('shape', [(70, 37), (300, 200), (129, 16), (5, 1001), (260, 50), (150, 64)])
@pytest.mark.parametrize('bulk', ['0', '1'])
def test_row_normalize_and_gram_bit_exact(shape, bulk, monkeypatch):
    import subprocess
    import sys
    if bulk == '1':
        # the TMA bulk-copy staging is selected once per process: run this case in a child process
        code = ("import os, sys, importlib.util; os.environ['SCTDA_GEMM_TMA_BULK'] = '1'; sys.path.insert(0, %r); "
                "spec = importlib.util.spec_from_file_location('tm', %r); m = importlib.util.module_from_spec(spec); "
                "spec.loader.exec_module(m); m.test_row_normalize_and_gram_bit_exact(%r, '0', None)"
                % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.abspath(__file__), shape))
        if not torch.cuda.is_available():
            pytest.skip('no CUDA device')
        r = subprocess.run([sys.executable, '-c', code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        assert r.returncode == 0, r.stdout.decode()[-2000:]
        return
    from oracle import mapper_oracle as mo
    from sctda_b200 import mapper
    ctx = _gpu()
    n, g = shape
    X, _ = _data(n, g, seed=n + g)
    X[n // 2] = 1.5                                            # constant row
    Xd = mapper._padded_device_table(X, torch.device('cuda', 0))
    rng = numpy.random.RandomState(1)
    ra = rng.randint(0, n, size=min(n, 150))
    rb = rng.randint(0, n, size=min(n, 131))
    for metric in ('euclidean', 'correlation'):
        Xn_ref, sq_ref = mo.row_normalize(X, metric)
        Xn, sq = mapper.row_normalize(ctx, Xd, metric)
        assert (_bits(Xn.cpu().numpy()) == _bits(Xn_ref)).all()
        assert (_bits(sq.cpu().numpy()) == _bits(sq_ref)).all()
        G = mapper.gram_gather(ctx, Xn, ra, rb).cpu().numpy()
        G_ref = mo.gram(Xn_ref, ra, rb)
        numpy.testing.assert_allclose(G, G_ref, rtol=1e-13, atol=1e-13)
        assert (_bits(G) == _bits(G_ref)).all(), 'DMMA result is not the sequential fma chain'
        if g % 2 == 0:
            # rows without zero padding to a multiple of 16 columns (always the cp.async staging)
            G2 = mapper.gram_gather(ctx, Xn.contiguous(), ra, rb).cpu().numpy()
            assert (_bits(G2) == _bits(G_ref)).all()


@pytest.mark.parametrize('metric', ['euclidean', 'correlation'])
def test_pairdist_rows(metric):
    import scipy.spatial.distance as ssd
    from oracle import mapper_oracle as mo
    from sctda_b200 import mapper
    ctx = _gpu()
    X, _ = _data(333, 150, seed=2)
    Xd = mapper._padded_device_table(X, torch.device('cuda', 0))
    D = mapper.pairwise_distances(ctx, Xd, metric, row_block=100).cpu().numpy()
    Xn, sq = mo.row_normalize(X, metric)
    ref = mo.pairdist_rows(Xn, sq, 0, 333, metric)
    assert (_bits(D) == _bits(ref)).all()
    numpy.testing.assert_allclose(D, ssd.cdist(X, X, metric), rtol=1e-9, atol=1e-7)
    assert (numpy.diag(D) == 0.0).all()
    # one call over all rows: upper triangle of tiles, mirrored -- the same bits
    Ds = mapper.pairwise_distances(ctx, Xd, metric).cpu().numpy()
    assert (_bits(Ds) == _bits(ref)).all()


def _device_result(ctx, X, lens, r, gain, equalize=True, metric='euclidean', max_K=5, stat='db'):
    from sctda_b200 import mapper
    Xd = mapper._padded_device_table(X, torch.device('cuda', 0))
    return mapper.build_graph_device(ctx, Xd, lens, r, gain, equalize=equalize, max_K=max_K, metric=metric, stat=stat)


@pytest.mark.parametrize('case', [(400, 60, 6, 0.4, True, 'euclidean', 5), (400, 60, 5, 2.0, False, 'euclidean', 5),
                                  (900, 101, 9, 1.0, True, 'correlation', 5), (250, 40, 3, 0.3, True, 'euclidean', 8),
                                  (120, 30, 1, 0.5, True, 'euclidean', 5), (500, 50, 4, 0.5, True, 'euclidean', 5, 'histgap'),
                                  (700, 20, 2, 0.3, True, 'correlation', 5, 'histgap')])
def test_graph_identical_to_oracle(case):
    from oracle import mapper_oracle as mo
    ctx = _gpu()
    stat = case[7] if len(case) > 7 else 'db'
    n, g, r, gain, equalize, metric, max_K = case[:7]
    X, lens = _data(n, g, seed=n + r)
    lens[5] = lens[6]                                          # duplicate lens values
    X[11] = X[10]                                              # duplicate cells (zero distances)
    res = _device_result(ctx, X, lens, r, gain, equalize, metric, max_K, stat)
    bin_off = res.bin_off.cpu().numpy()
    members = res.members.cpu().numpy()
    labels = res.labels.cpu().numpy()
    bin_K = res.bin_K.cpu().numpy()
    bin_db = res.bin_db.cpu().numpy()
    # stage by stage against the oracle
    bins, _ = mo.cover(lens, r, gain, equalize)
    Xn, _ = mo.row_normalize(X, metric)
    for b, cells in enumerate(bins):
        assert list(members[bin_off[b]:bin_off[b + 1]]) == list(cells), 'cover patch %d' % b
        if len(cells) == 0:
            assert bin_K[b] == 0
            continue
        lab, K, scores = mo.cluster_bin(mo.gram(Xn, cells, cells), max_K=max_K, stat=stat)
        assert bin_K[b] == K, 'K of patch %d (%d cells): %r vs %r' % (b, len(cells), bin_db[b], scores)
        assert list(labels[bin_off[b]:bin_off[b + 1]]) == list(lab)
        if scores:
            assert (_bits(bin_db[b, :len(scores)]) == _bits(numpy.array(scores))).all()
    edges, clusters, _ = mo.mapper_graph(X, lens, r, gain, equalize=equalize, max_K=max_K, metric=metric, stat=stat)
    node_off = res.node_off.cpu().numpy()
    node_members = res.node_members.cpu().numpy()
    assert res.V == len(clusters)
    for k, c in enumerate(clusters):
        assert list(node_members[node_off[k]:node_off[k + 1]]) == list(c)
    assert [tuple(e) for e in res.edges.cpu().numpy().tolist()] == edges


def test_graph_identical_to_oracle_5k_cells():
    """A build the size of a real patch workload against the oracle: 5,000 cells x 500 genes, r = 8,
    gain 1 (64 patches of ~300 cells; the oracle's C Gram helper keeps it to seconds): node lists
    and edges identical."""
    from oracle import mapper_oracle as mo
    ctx = _gpu()
    X, lens = _data(5000, 500, seed=77)
    for metric in ('euclidean', 'correlation'):
        res = _device_result(ctx, X, lens, 8, 1.0, True, metric, 5, 'db')
        edges, clusters, _ = mo.mapper_graph(X, lens, 8, 1.0, metric=metric)
        node_off = res.node_off.cpu().numpy()
        node_members = res.node_members.cpu().numpy()
        assert res.V == len(clusters)
        for k, c in enumerate(clusters):
            assert list(node_members[node_off[k]:node_off[k + 1]]) == list(c)
        assert [tuple(e) for e in res.edges.cpu().numpy().tolist()] == edges


@pytest.mark.parametrize('linkage', ['average', 'ward', 'complete'])
@pytest.mark.parametrize('case', [(600, 80, 5, 0.6, 'euclidean', 'db'), (900, 64, 4, 1.0, 'correlation', 'db'),
                                  (500, 50, 3, 0.5, 'euclidean', 'histgap')])
def test_other_linkages_match_oracle_and_sklearn(linkage, case):
    """cluster='agglomerative' with sklearn's other linkages (ref:758-771; linkage_kernel: nearest-neighbour chain
    with Lance-Williams updates): node lists, K and edges equal the oracle's (scipy.cluster.hierarchy.linkage on
    the same distances + the same model selection), and on euclidean patches the chosen partitions equal
    sklearn.cluster.AgglomerativeClustering(linkage=...) + davies_bouldin_score on the feature rows."""
    import sklearn.cluster
    import sklearn.metrics
    from oracle import mapper_oracle as mo
    from sctda_b200 import mapper
    ctx = _gpu()
    n, g, r, gain, metric, stat = case
    X, lens = _data(n, g, seed=n + r + len(linkage))
    Xd = mapper._padded_device_table(X, torch.device('cuda', 0))
    res = mapper.build_graph_device(ctx, Xd, lens, r, gain, metric=metric, stat=stat, linkage=linkage)
    assert ctx.lib.sctda_get_linkage(ctx.handle) == 0                      # the setting is restored
    edges, clusters, _ = mo.mapper_graph(X, lens, r, gain, metric=metric, stat=stat, linkage=linkage)
    node_off = res.node_off.cpu().numpy()
    node_members = res.node_members.cpu().numpy()
    assert res.V == len(clusters)
    for k, c in enumerate(clusters):
        assert list(node_members[node_off[k]:node_off[k + 1]]) == list(c)
    assert [tuple(e) for e in res.edges.cpu().numpy().tolist()] == edges
    if metric == 'euclidean' and stat == 'db':
        bin_off = res.bin_off.cpu().numpy()
        members = res.members.cpu().numpy()
        labels = res.labels.cpu().numpy()
        bin_K = res.bin_K.cpu().numpy()
        for b in range(r * r):
            cells = members[bin_off[b]:bin_off[b + 1]]
            m = len(cells)
            if m < 3:
                continue
            best = None
            for K in range(2, (2 if m <= 5 else min(m // 2, 5)) + 1):
                lab = sklearn.cluster.AgglomerativeClustering(n_clusters=K, linkage=linkage).fit_predict(X[cells])
                sc = sklearn.metrics.davies_bouldin_score(X[cells], lab)
                if best is None or sc < best[0]:
                    best = (sc, K, lab)
            mine = labels[bin_off[b]:bin_off[b + 1]]
            assert sorted(sorted(int(c) for c in cells[best[2] == k]) for k in range(best[1])) == \
                sorted(sorted(int(c) for c in cells[mine == k]) for k in range(bin_K[b]))
    # the public seam
    G, all_clusters, _ = mapper.mapper_graph(X, lens, r, gain, metric=metric, stat=stat, linkage=linkage, verbose=False)
    assert G.number_of_nodes() == res.V and G.number_of_edges() == res.E
    with pytest.raises(NotImplementedError):
        mapper.mapper_graph(X, lens, r, gain, clust='kmeans', verbose=False)


def test_graph_against_sklearn_single_linkage_and_davies_bouldin():
    """The whole graph against an INDEPENDENT restatement that shares nothing with the device arithmetic
    (no Gram trick, no fma chain, no lane-ordered sums): BASELINE.json configs[0]'s scale (1,000 cells x
    2,000 genes, r = 20, gain 2).  Per patch: sklearn.cluster.AgglomerativeClustering(linkage='single')
    on the feature rows for K = 2..K_max, K chosen by sklearn.metrics.davies_bouldin_score (smallest
    index, first K on ties), clusters numbered by smallest member; nerve by set intersection.
    The clusterings agree wherever the model selection is not decided by the last bits: node sets
    must match on all patches but a handful, and the graphs must be isomorphic there."""
    import sklearn.cluster
    import sklearn.metrics
    from sctda_b200 import mapper
    ctx = _gpu()
    X, lens = _data(1000, 2000, seed=5)
    r, gain, max_K = 20, 2.0, 5
    res = _device_result(ctx, X, lens, r, gain, True, 'euclidean', max_K, 'db')
    bin_off = res.bin_off.cpu().numpy()
    members = res.members.cpu().numpy()
    labels = res.labels.cpu().numpy()
    bin_K = res.bin_K.cpu().numpy()
    bin_db = res.bin_db.cpu().numpy()
    lo0, hi0, lo1, hi1 = mapper.cover_bounds(lens, r, gain, True)
    n_patch = n_same = n_close = 0
    for b in range(r * r):
        i, j = divmod(b, r)
        cells = numpy.nonzero((lens[:, 0] > lo0[i]) & (lens[:, 0] < hi0[i]) & (lens[:, 1] > lo1[j]) & (lens[:, 1] < hi1[j]))[0]
        assert list(members[bin_off[b]:bin_off[b + 1]]) == list(cells)               # the cover, by brute force
        m = len(cells)
        if m < 2:
            assert bin_K[b] == m
            continue
        if m == 2:                              # sklearn's index needs fewer clusters than cells: two singletons
            assert bin_K[b] == 2
            continue
        n_patch += 1
        kmax = 2 if m <= 5 else min(m // 2, max_K)
        best = None
        scores = []
        for K in range(2, kmax + 1):
            lab = sklearn.cluster.AgglomerativeClustering(n_clusters=K, linkage='single').fit_predict(X[cells])
            sc = sklearn.metrics.davies_bouldin_score(X[cells], lab)
            scores.append(sc)
            if best is None or sc < best[0]:
                best = (sc, K, lab)
        mine = labels[bin_off[b]:bin_off[b + 1]]
        part_ref = sorted(sorted(int(c) for c in cells[best[2] == k]) for k in range(best[1]))
        part_me = sorted(sorted(int(c) for c in cells[mine == k]) for k in range(bin_K[b]))
        numpy.testing.assert_allclose(bin_db[b, :len(scores)], scores, rtol=1e-5)    # the index itself, every K (the
        # device forms centroid distances from the Gram tile: cancellation costs a few digits, 2e-7 at worst here)
        if part_ref == part_me:
            n_same += 1
        else:                                                                         # two K within rounding of each other
            srt = sorted(scores)
            assert srt[1] - srt[0] <= 1e-9 * abs(srt[0]), (b, scores, bin_K[b], best[1])
            n_close += 1
    assert n_patch > 200 and n_same >= n_patch - 3


def test_topological_representation_files(tmp_path):
    """The reference-facing classes end to end: .mapper.tsv -> TopologicalRepresentation.save ->
    name.json / name.gexf -> UnrootedGraph (ref:735-778, 785-806)."""
    _gpu()
    import networkx
    from oracle import mapper_oracle as mo
    from sctda_b200 import formats, graph, mapper
    X, _ = _data(300, 40, seed=4)
    genes = ['G%02d' % i for i in range(40)]
    base = str(tmp_path / 't')
    with open(base + '.mapper.tsv', 'w') as f:
        f.write('\t'.join(genes) + '\n')
        for row in X:
            f.write('\t'.join(formats.py2_str(v) for v in row) + '\n')
    with open(base + '.all.tsv', 'w') as f:
        f.write('ID\ttimepoint\tlib\t' + '\t'.join(genes) + '\n')
        for i, row in enumerate(X):
            f.write('D0_L_%d\t0\tL\t' % i + '\t'.join(formats.py2_str(v) for v in row) + '\n')
    t = mapper.TopologicalRepresentation(base, lens='pca')
    patches = t.save(base + '_g', 5, 0.5)
    assert len(patches) == 25
    import pandas
    df = pandas.read_table(base + '.mapper.tsv')
    lens = t.lens_data_mds.values
    # the PCA lens agrees with sklearn (component signs included)
    from sklearn.decomposition import PCA
    numpy.testing.assert_allclose(lens, PCA(n_components=2).fit_transform(df.values), rtol=1e-7, atol=1e-8)
    edges, clusters, _ = mo.mapper_graph(df.values, lens, 5, 0.5)
    with open(base + '_g.json') as f:
        dic = json.load(f)
    assert dic == {str(k): [int(x) for x in c] for k, c in enumerate(clusters)}
    g = networkx.read_gexf(base + '_g.gexf')
    assert sorted((min(int(a), int(b)), max(int(a), int(b))) for a, b in g.edges()) == edges
    assert all(d['weight'] == 1.0 for _, _, d in g.edges(data=True))
    c = graph.UnrootedGraph(base + '_g', base + '.all.tsv', groups=False)
    assert len(c.pl) >= 2 and c.connectivity('G00') >= 0.0
    # MDS lens: the device dissimilarity feeds sklearn's MDS (ref:746-751)
    t2 = mapper.TopologicalRepresentation(base, lens='mds', metric='correlation', random_state=0, n_init=1, max_iter=30)
    assert t2.lens_data_mds.shape == (300, 2)


def test_properties_at_scale():
    """Invariants at a size the oracle does not reach in seconds."""
    ctx = _gpu()
    n, g, r, gain = 20000, 256, 20, 1.0
    X, lens = _data(n, g, seed=9)
    res = _device_result(ctx, X, lens, r, gain)
    bin_off = res.bin_off.cpu().numpy().astype(numpy.int64)
    members = res.members.cpu().numpy()
    labels = res.labels.cpu().numpy()
    bin_K = res.bin_K.cpu().numpy()
    node_off = res.node_off.cpu().numpy().astype(numpy.int64)
    node_members = res.node_members.cpu().numpy()
    node_bin = res.node_bin.cpu().numpy()
    edges = res.edges.cpu().numpy()
    assert numpy.bincount(members, minlength=n).min() >= 1                     # every cell covered
    sizes = bin_off[1:] - bin_off[:-1]
    for b in range(r * r):
        seg = members[bin_off[b]:bin_off[b + 1]]
        assert (numpy.diff(seg) > 0).all()                                      # ascending, no duplicates
        if sizes[b] >= 2:
            assert 2 <= bin_K[b] <= 5
            assert set(labels[bin_off[b]:bin_off[b + 1]]) == set(range(bin_K[b]))
    assert node_off[-1] == bin_off[-1] == len(node_members)                     # nodes partition the patches
    assert res.V == bin_K.sum()
    sets = [set(node_members[node_off[k]:node_off[k + 1]].tolist()) for k in range(res.V)]
    for b in range(r * r):
        ks = numpy.nonzero(node_bin == b)[0]
        if len(ks):
            assert set().union(*[sets[k] for k in ks]) == set(members[bin_off[b]:bin_off[b + 1]].tolist())
    # nerve: sorted, i < j, exactly the intersecting pairs (checked through a cell -> nodes index)
    assert (edges[:, 0] < edges[:, 1]).all()
    assert (numpy.lexsort((edges[:, 1], edges[:, 0])) == numpy.arange(len(edges))).all()
    brute = set()
    owner = [[] for _ in range(n)]
    for k in range(res.V):
        for c in node_members[node_off[k]:node_off[k + 1]]:
            owner[c].append(k)
    for lst in owner:
        for x in range(len(lst)):
            for y in range(x + 1, len(lst)):
                brute.add((min(lst[x], lst[y]), max(lst[x], lst[y])))
    assert set(map(tuple, edges.tolist())) == brute
    # nodes of one patch never connect
    assert (node_bin[edges[:, 0]] != node_bin[edges[:, 1]]).all()


def test_patches_above_the_shared_memory_caps():
    """One patch above the shared-memory caps of the MST / level-label kernels (10,239 cells) in
    the same batch as small patches: the global-memory fallbacks give the oracle's labels."""
    from oracle import mapper_oracle as mo
    ctx = _gpu()
    n, g, big = 12000, 24, 10800
    X, _ = _data(n, g, seed=31)
    rng = numpy.random.RandomState(2)
    lens = numpy.empty((n, 2))
    lens[:big] = 0.4 * rng.rand(big, 2)
    lens[big:] = 0.6 + 0.4 * rng.rand(n - big, 2)
    lens[big:big + 300, 1] = 0.3 * rng.rand(300)
    res = _device_result(ctx, X, lens, 2, 0.1, equalize=False)
    bin_off = res.bin_off.cpu().numpy()
    members = res.members.cpu().numpy()
    labels = res.labels.cpu().numpy()
    bin_K = res.bin_K.cpu().numpy()
    bin_db = res.bin_db.cpu().numpy()
    bins, _ = mo.cover(lens, 2, 0.1, False)
    sizes = [len(b) for b in bins]
    assert max(sizes) > 10239 and 1 < sorted(sizes)[-2] < 7000
    Xn, _ = mo.row_normalize(X, 'euclidean')
    for b, cells in enumerate(bins):
        assert list(members[bin_off[b]:bin_off[b + 1]]) == list(cells)
        if len(cells) == 0:
            assert bin_K[b] == 0
            continue
        lab, K, scores = mo.cluster_bin(mo.gram(Xn, cells, cells), max_K=5)
        assert bin_K[b] == K
        assert (labels[bin_off[b]:bin_off[b + 1]] == lab).all()
        assert (_bits(bin_db[b, :len(scores)]) == _bits(numpy.array(scores))).all()


def test_mapper_abi_errors_and_empty_inputs():
    ctx = _gpu()
    import ctypes
    from sctda_b200 import _lib, mapper
    dev = torch.device('cuda', 0)
    X = torch.zeros((8, 6), dtype=torch.float64, device=dev)
    lens = torch.zeros((8, 2), dtype=torch.float64, device=dev)
    off = torch.zeros(2, dtype=torch.int32, device=dev)
    b = torch.zeros(1, dtype=torch.float64, device=dev)
    p = _lib.ptr
    L = ctx.lib
    # resolution above the served maximum
    rc = L.sctda_cover_count(ctx.handle, 8, p(lens), 2, 129, p(b), p(b), p(b), p(b), p(off), None)
    assert rc == -4 and b'resolution' in L.sctda_last_error(ctx.handle)
    # NULL lens
    assert L.sctda_cover_count(ctx.handle, 8, None, 2, 1, p(b), p(b), p(b), p(b), p(off), None) == -1
    # max_K out of range, unsupported statistic, odd leading dimension
    lab = torch.zeros(8, dtype=torch.int32, device=dev)
    K = torch.zeros(1, dtype=torch.int32, device=dev)
    mem = torch.arange(8, dtype=torch.int32, device=dev)
    assert L.sctda_bin_cluster(ctx.handle, 6, p(X), 6, 1, p(off), p(mem), 9, 0, p(lab), p(K), None, None) == -4
    assert L.sctda_bin_cluster(ctx.handle, 6, p(X), 6, 1, p(off), p(mem), 5, 7, p(lab), p(K), None, None) == -4
    Xo = torch.zeros((8, 7), dtype=torch.float64, device=dev)
    assert L.sctda_bin_cluster(ctx.handle, 7, p(Xo), 7, 1, p(off), p(mem), 5, 0, p(lab), p(K), None, None) == -4
    # fill before count with a larger problem is refused
    big = torch.zeros(4, dtype=torch.int32, device=dev)
    assert L.sctda_nerve_fill(ctx.handle, 100000, p(big), None) == -1
    # a cover in which every patch is empty (all bounds exclude the cells) and constant data
    res = mapper.build_graph_device(ctx, X, lens.cpu().numpy(), 3, 0.5,
                                    bounds=[numpy.full(3, 5.0), numpy.full(3, 6.0), numpy.full(3, 5.0), numpy.full(3, 6.0)])
    assert res.V == 0 and res.E == 0 and int(res.bin_off[-1]) == 0
    # identical cells: all distances zero, ties broken by index, every patch still splits in two
    res = mapper.build_graph_device(ctx, X, numpy.random.RandomState(0).rand(8, 2), 1, 0.5)
    assert res.V == 2 and res.node_members.cpu().tolist() == list(range(8))
    assert res.node_off.cpu().tolist() == [0, 7, 8] and res.E == 0      # the oracle's tie rule: the last-inserted edge is cut


def test_every_mst_kernel_gives_the_same_graph(monkeypatch):
    """The MST kernels (one CTA of 128 / 256 / 512 threads with register keys, clusters of 8 CTAs with
    4 / 8 keys per thread, the generic global-memory kernel) are selected by patch size; the test hook
    lowers the size classes so that one small input reaches all of them, and the graph must not change."""
    from oracle import mapper_oracle as mo
    ctx = _gpu()
    X, lens = _data(900, 32, seed=77)
    edges, clusters, _ = mo.mapper_graph(X, lens, 3, 0.8)
    sizes = sorted(len(p) for p in mo.cover(lens, 3, 0.8)[0])
    assert sizes[0] < 120 and sizes[-1] > 330
    for thr in (None, '150,250,400', '100,200,300', '60,61,62', '400,500,600,250,150', '330,400,500,200,120'):
        if thr is None:
            monkeypatch.delenv('SCTDA_TEST_PRIM_THRESHOLDS', raising=False)
        else:
            monkeypatch.setenv('SCTDA_TEST_PRIM_THRESHOLDS', thr)
        res = _device_result(ctx, X, lens, 3, 0.8)
        off = res.node_off.cpu().numpy()
        mem = res.node_members.cpu().numpy()
        assert res.V == len(clusters)
        assert all(list(mem[off[k]:off[k + 1]]) == list(c) for k, c in enumerate(clusters)), thr
        assert [tuple(e) for e in res.edges.cpu().numpy().tolist()] == edges


def test_distributed_build_equals_single_device_build():
    """build_graph_distributed with a single rank (no process group) walks the sharded code path
    (LPT assignment, sub-CSR, label exchange) and must give the graph of build_graph_device."""
    ctx = _gpu()
    from sctda_b200 import mapper
    X, lens = _data(1500, 64, seed=12)
    dev = torch.device('cuda', 0)
    Xd = mapper._padded_device_table(X, dev)
    a = mapper.build_graph_device(ctx, Xd, lens, 6, 0.7)
    b = mapper.build_graph_distributed(ctx, Xd, lens, 6, 0.7)
    assert a.V == b.V and a.E == b.E
    for name in ('bin_off', 'members', 'labels', 'bin_K', 'node_off', 'node_members', 'node_bin', 'edges'):
        assert torch.equal(getattr(a, name), getattr(b, name)), name


@pytest.mark.parametrize('shape', [(3000, 400), (400, 1500), (1000, 2000), (5000, 64), (257, 257)])
def test_pca_lens_matches_sklearn(shape):
    """apply_lens(lens='pca') (ref:745, 753): the device lens (sctda_pca_lens_f64: own Gram contraction + subspace
    iteration) against sklearn.decomposition.PCA(n_components=2, svd_solver='full').fit_transform, within 1e-7 of
    the lens scale, for tables with more cells than genes (covariance form) and more genes than cells (dual form)."""
    import sklearn.decomposition
    from sctda_b200 import mapper, synth
    _gpu()
    n, g = shape
    X, _, _ = synth.expression_matrix(n, g, seed=n + g)
    ref = sklearn.decomposition.PCA(n_components=2, svd_solver='full').fit_transform(X)
    lens, info = mapper.pca_lens_device(torch.from_numpy(X).cuda(), return_info=True)
    assert info['status'] == 1 and max(info['residual']) <= 1e-12
    got = lens.cpu().numpy()
    scale = numpy.abs(ref).max()
    numpy.testing.assert_allclose(got, ref, rtol=0, atol=1e-7 * scale)
    # the public seam
    import pandas
    df = pandas.DataFrame(X)
    out = mapper.apply_lens(df, lens='pca')
    numpy.testing.assert_allclose(out.values, ref, rtol=0, atol=1e-7 * scale)
    # reproducible bit for bit
    again = mapper.pca_lens_device(torch.from_numpy(X).cuda()).cpu().numpy()
    assert (_bits(again) == _bits(got)).all()


def test_smacof_matches_sklearn():
    """The device SMACOF (one sctda_smacof_pass per iteration) against sklearn.manifold.smacof from the
    same start: same number of iterations, same stress, same configuration."""
    import warnings
    import sklearn.manifold
    ctx = _gpu()
    from sctda_b200 import mapper
    X, _ = _data(300, 60, seed=8)
    dev = torch.device('cuda', 0)
    for metric in ('correlation', 'euclidean'):
        D = mapper.pairwise_distances(ctx, mapper._padded_device_table(X, dev), metric)
        Dh = D.cpu().numpy()
        assert (Dh == Dh.T).all()
        X0 = numpy.random.RandomState(3).uniform(size=(300, 2))
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            ref_pos, ref_stress, ref_it = sklearn.manifold.smacof(Dh, metric=True, n_components=2, init=X0.copy(), n_init=1,
                                                                  max_iter=300, eps=1e-6, normalized_stress=False,
                                                                  return_n_iter=True)
        pos, stress, it = mapper.smacof_device(ctx, D, init=X0.copy(), max_iter=300, eps=1e-6, return_n_iter=True)
        assert abs(it - ref_it) <= 1
        assert stress == pytest.approx(ref_stress, rel=1e-6)
        if it == ref_it:
            numpy.testing.assert_allclose(pos.cpu().numpy(), ref_pos, rtol=1e-5, atol=1e-7)
    # random restarts draw their starts from the RandomState exactly as sklearn does
    pos1, s1 = mapper.smacof_device(ctx, D, n_init=2, max_iter=40, random_state=5)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ref_pos, ref_s = sklearn.manifold.smacof(Dh, metric=True, n_components=2, n_init=2, max_iter=40, eps=1e-6, random_state=5,
                                                 normalized_stress=False)
    assert s1 == pytest.approx(ref_s, rel=1e-6)
    numpy.testing.assert_allclose(pos1.cpu().numpy(), ref_pos, rtol=1e-5, atol=1e-7)


# ---- opt-in split-precision contraction (tcgen05, sctda_set_precision(ctx, 1)) ---------------------

SPLIT_TOL = 1e-5            # the tolerance include/sctda_b200.h states for precision mode 1


@pytest.mark.parametrize('shape', [(300, 77), (129, 64), (1000, 1000), (2500, 333), (64, 5000)])
@pytest.mark.parametrize('metric', ['correlation', 'euclidean'])
def test_split_precision_distances_within_stated_tolerance(shape, metric):
    """Gram entries within 1e-5 of the FP64 kernel relative to |x||y| (correlation: rows are unit
    vectors, so absolute 1e-5 on 1 - r; euclidean: on the squared distance over |x|^2 + |y|^2)."""
    from sctda_b200 import mapper
    ctx = _gpu()
    n, g = shape
    X, _ = _data(n, g, seed=3 * n + g)
    X[n // 3] = X[n // 3 + 1]                                  # duplicate rows: distance exactly 0 on the diagonal only
    Xd = mapper._padded_device_table(X, torch.device('cuda', 0))
    D0 = mapper.pairwise_distances(ctx, Xd, metric)
    with ctx.precision('fp16x3'):
        assert ctx.get_precision() == 'fp16x3'
        D1 = mapper.pairwise_distances(ctx, Xd, metric)
    assert ctx.get_precision() == 'fp64'
    assert torch.isfinite(D1).all()
    assert (torch.diagonal(D1) == 0).all()
    if metric == 'correlation':
        err = (D1 - D0).abs().max().item()
    else:
        sq = (Xd * Xd).sum(dim=1)
        err = ((D1 * D1 - D0 * D0).abs() / (sq[:, None] + sq[None, :]).clamp_min(1e-300)).max().item()
    assert err <= SPLIT_TOL, err
    assert err > 0.0                                           # it really is the other kernel


@pytest.mark.parametrize('scale', [2.0 ** -40, 1.0, 3.0e7])
def test_split_precision_is_scale_invariant(scale):
    """The power-of-two pre-scaling keeps the FP16 operands in range whatever the units of the table."""
    from sctda_b200 import mapper
    ctx = _gpu()
    X, _ = _data(400, 150, seed=5)
    Xd = mapper._padded_device_table(X * scale, torch.device('cuda', 0))
    D0 = mapper.pairwise_distances(ctx, Xd, 'euclidean')
    with ctx.precision('3xtf32'):                              # the north star's name is accepted as an alias
        D1 = mapper.pairwise_distances(ctx, Xd, 'euclidean')
    sq = (Xd * Xd).sum(dim=1)
    err = ((D1 * D1 - D0 * D0).abs() / (sq[:, None] + sq[None, :]).clamp_min(1e-300)).max().item()
    assert err <= SPLIT_TOL, err
    Z = torch.zeros((130, 40), dtype=torch.float64, device='cuda')
    with ctx.precision('fp16x3'):
        assert (mapper.pairwise_distances(ctx, Z, 'euclidean') == 0).all()


def test_split_precision_graph_and_mode_handling():
    """The graph build in the opt-in mode: same cover (it does not depend on the contraction),
    node count and labels equal to the FP64 build on this data set except where a near-tie in the
    linkage order flips (none here; the mode does not GUARANTEE identity); gram_gather stays FP64."""
    from sctda_b200 import _lib, mapper
    ctx = _gpu()
    X, lens = _data(3000, 400, seed=11)
    Xd = mapper._padded_device_table(X, torch.device('cuda', 0))
    a = mapper.build_graph_device(ctx, Xd, lens, 12, 3.0, metric='correlation')
    with ctx.precision('fp16x3'):
        b = mapper.build_graph_device(ctx, Xd, lens, 12, 3.0, metric='correlation')
        rows = torch.arange(0, 200, dtype=torch.int32, device='cuda')
        Xn, _ = mapper.row_normalize(ctx, Xd, 'correlation')
        g1 = mapper.gram_gather(ctx, Xn, rows, rows)
    g0 = mapper.gram_gather(ctx, Xn, rows, rows)
    assert (g0 == g1).all()                                    # sctda_gram_gather_f64 is always FP64
    assert (a.bin_off == b.bin_off).all() and (a.members == b.members).all()
    same_K = (a.bin_K == b.bin_K).float().mean().item()
    assert same_K >= 0.99, same_K
    assert abs(int(a.V) - int(b.V)) <= max(1, int(a.V) // 100)
    with pytest.raises(ValueError):
        ctx.set_precision('fp8')
    assert ctx.lib.sctda_set_precision(ctx.handle, 7) == -4
    assert ctx.get_precision() == 'fp64'
    # single-feature and tiny problems go through the same kernel
    x1 = mapper._padded_device_table(numpy.arange(1.0, 10.0).reshape(9, 1), torch.device('cuda', 0))
    with ctx.precision('fp16x3'):
        d = mapper.pairwise_distances(ctx, x1, 'euclidean').cpu().numpy()
    ref = numpy.abs(numpy.arange(1.0, 10.0)[:, None] - numpy.arange(1.0, 10.0)[None, :])
    assert numpy.abs(d - ref).max() <= 1e-3                    # sqrt of a 1e-5-relative squared distance near 0
