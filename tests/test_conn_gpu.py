"""Parity of the CUDA connectivity path (through the C-ABI, sctda_b200/_lib.py) with the CPU
oracle (oracle/conn_oracle.py) and with the golden vectors the reference's own source produced
(tests/golden, oracle/gen_golden.py).  Needs a B200: `pytest -m gpu`.

Tolerances (BASELINE.json north_star): permuted scores within 1e-9 relative, exceed counts /
p-values exact.  Observed statistics (pure float64, no float32 step): 1e-12 relative.
A (gene, permutation) pair whose permuted score lies within 1e-11 (relative: the worst-case
rounding of the two evaluations) of the observed score is a TIE: the two scores are the same
number in exact arithmetic, which needs a degenerate gene (constant over the component's cells,
or expressed in one or two of them).  The library counts such pairs in `ties` and does not count
them as exceeding (strict `>`, ref:990); the reference decides them by the summation order of its
BLAS build.  The tests assert: every gene without a tie is exact; a tie only ever occurs on a
degenerate gene; on tied genes the count equals the oracle's count outside the band.
"""
import json
import os
import shutil

import numpy
import pytest
import scipy.sparse

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')


def _gpu():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from sctda_b200 import _lib
    return _lib.context(0)


def _oracle_state(members, adj, Y, log2=True):
    from oracle import conn_oracle as co
    V = len(members)
    return co.GraphState([str(k) for k in range(V)], adj.toarray() if scipy.sparse.issparse(adj) else adj,
                         {str(k): [int(c) for c in members[k]] for k in range(V)},
                         {'g%d' % i: Y[i] for i in range(Y.shape[0])}, log2=log2)


def _random_case(seed, n_cells, n_nodes, n_genes, ragged=False, weights=False):
    """Random connected cover: ring + chords; optionally ragged node sizes (1..), unsorted
    member lists, non-unit edge weights and cells missing from the graph."""
    from sctda_b200 import synth
    rng = numpy.random.RandomState(seed)
    members, adj, Y = synth.connectivity_case(n_cells, n_nodes, n_genes, seed=seed)
    if ragged:
        members = [rng.permutation(m)[:max(1, rng.randint(1, len(m) + 1))] for m in members]
        members[0] = members[0][:1]
    if weights:
        adj = adj.tocoo()
        w = rng.uniform(0.5, 2.0, adj.nnz)
        adj = scipy.sparse.coo_matrix((w, (adj.row, adj.col)), shape=adj.shape)
        adj = ((adj + adj.T) * 0.5).tocsr()
    Y[0] = 0.0                       # all-zero gene: tol = 0
    Y[1] = 2.5                       # constant gene: every permuted score equals the observed one (ties)
    if n_genes > 2:
        Y[2] = 0.0
        Y[2, members[0][0]] = 6.0    # single expressing cell
    return members, adj.tocsr(), Y


@pytest.mark.parametrize('log2', [True, False])
@pytest.mark.parametrize('shape', [(400, 40, 37, 48, False, False), (900, 150, 70, 33, True, True),
                                   (64, 2, 5, 16, False, False)])
def test_perm_test_matches_oracle(log2, shape):
    from oracle import conn_oracle as co
    from sctda_b200 import graph
    ctx = _gpu()
    n_cells, n_nodes, n_genes, n, ragged, weights = shape
    members, adj, Y = _random_case(17 + n_nodes, n_cells, n_nodes, n_genes, ragged, weights)
    if n_nodes == 2:
        adj = scipy.sparse.csr_matrix(numpy.array([[0.0, 1.0], [1.0, 0.0]]))
    st = _oracle_state(members, adj, Y, log2=log2)
    dg = graph.DeviceGraph(members, adj, st.samples, n_cells, 0)
    Yd = torch.from_numpy(Y).cuda()
    stats = graph.node_stats(ctx, dg, Yd, log2=log2, want_p=True)
    rng = numpy.random.RandomState(99)
    shared = co.perm_indices(n, dg.S, rng)
    per_gene = numpy.stack([co.perm_indices(n, dg.S, rng) for _ in range(n_genes)])
    ex_s, ti_s, sc_s = graph.conn_perm_test(ctx, dg, Yd, torch.from_numpy(shared).cuda(), stats['conn'], log2=log2,
                                            want_scores=True)
    ex_g, ti_g, sc_g = graph.conn_perm_test(ctx, dg, Yd, torch.from_numpy(per_gene).cuda(), stats['conn'], log2=log2,
                                            per_gene=True, want_scores=True)
    torch.cuda.synchronize()
    tie_genes = 0
    for i in range(n_genes):
        gi = 'g%d' % i
        # observed statistics
        dic, tol = co.get_gene(st, gi)
        assert stats['tol'][i].item() == pytest.approx(tol, rel=1e-12, abs=0)
        numpy.testing.assert_allclose(stats['p'][i].cpu().numpy(), [dic[k] for k in st.pl], rtol=1e-12, atol=0)
        obs = co.connectivity(st, gi)
        assert stats['conn'][i].item() == pytest.approx(obs, rel=1e-12, abs=0)
        numpy.testing.assert_allclose([stats[k][i].item() for k in ('mean', 'min', 'max')], co.delta(st, gi),
                                      rtol=1e-12, atol=0)
        assert stats['expr'][i].item() == co.expr(st, gi)
        # permuted scores and exceed counts
        for ex, ti, sc, perms in ((ex_s, ti_s, sc_s, shared), (ex_g, ti_g, sc_g, per_gene[i])):
            with numpy.errstate(invalid='ignore', divide='ignore'):
                ref = co.permuted_scores(st, gi, n=n, perms=perms)
            got = sc[i].cpu().numpy()
            if tol > 0.0:
                numpy.testing.assert_allclose(got, ref, rtol=1e-9, atol=0)
            else:
                assert numpy.isnan(got).all() and numpy.isnan(ref).all()
            if ti[i].item() == 0:
                assert ex[i].item() == int(numpy.sum(ref > obs))
            else:
                # A tie needs a DEGENERATE gene: constant over the component's cells (every permutation reproduces
                # the observed configuration) or expressed in at most two of them (a permuted configuration can
                # coincide with the observed one).  Both scores are then the same rational number; the reference's
                # `conn > observed` compares two roundings of it that come from its BLAS build (the summation
                # order and blocking of numpy.dot, ref:989), the library treats them as equal.
                x = Y[i][st.samples]
                # (a two-node graph is degenerate as a whole: swapping the two node values leaves the score unchanged)
                assert n_nodes == 2 or numpy.ptp(x) == 0.0 or numpy.count_nonzero(x) <= 2, 'tie on a non-degenerate gene %d' % i
                tie_genes += 1
                band = numpy.abs(ref - obs) <= 1e-11 * abs(obs)
                assert int(band.sum()) == ti[i].item()
                assert ex[i].item() == int(numpy.sum((ref > obs) & ~band))
    assert n_nodes == 2 or tie_genes <= 2 * 3   # the three planted degenerate genes (constant, all-zero, single-cell), both blocks


@pytest.mark.parametrize('name', ['unrooted_a', 'rooted_b'])
def test_golden_reference_vectors(golden_dir, tmp_path, name):
    """The product classes on the files the reference itself was run on (gen_golden.py)."""
    _gpu()
    from sctda_b200 import graph
    for ext in ('.gexf', '.json', '.all.tsv'):
        shutil.copy(os.path.join(golden_dir, name + ext), str(tmp_path / (name + ext)))
    with open(os.path.join(golden_dir, name + '.ref.json')) as f:
        ref = json.load(f)
    prefix = str(tmp_path / name)
    rooted = ref['rooted']
    if rooted:
        c = graph.RootedGraph(prefix, prefix + '.all.tsv', groups=False)
        assert (str(c.root), str(c.leaf)) == (ref['root'], ref['leaf'])
        assert {k: int(v) for k, v in c.dicdis.items()} == ref['dicdis']
        assert [float(c.po[0]), float(c.po[1])] == pytest.approx(ref['po'], rel=1e-12)
        den = c.get_dendrite()
        for k, v in ref['dendrite'].items():
            assert den[k] == pytest.approx(v, rel=1e-9, abs=1e-12, nan_ok=True)
    else:
        c = graph.UnrootedGraph(prefix, prefix + '.all.tsv', groups=False)
    assert sorted(c.pl) == sorted(ref['pl'])
    assert [int(x) for x in c.samples] == ref['samples']
    n = ref['n_perm']
    order = [ref['pl'].index(k) for k in c.pl]
    for gi, r in ref['per_gene'].items():
        dic, tol = c.get_gene(gi)
        assert tol == pytest.approx(r['tol'], rel=1e-12, abs=0)
        numpy.testing.assert_allclose([dic[k] for k in c.pl], numpy.array(r['p_nodes'])[order], rtol=1e-12, atol=0)
        assert c.connectivity(gi) == pytest.approx(r['connectivity'], rel=1e-12, abs=0)
        numpy.testing.assert_allclose(c.delta(gi), r['delta'], rtol=1e-12, atol=0)
        assert c.expr(gi) == r['expr']
        numpy.random.seed(ref['seed'] + 7)
        pv, ties = c.connectivity_pvalues([gi], n=n, return_ties=True)
        if ties[0] == 0:
            assert pv[0] == r['p_value']
        else:
            assert abs(pv[0] - r['p_value']) <= ties[0] / float(n) + 1e-15
        if rooted:
            cen = c.centroid(gi)
            if r['centroid'][0] is None:
                assert cen == [None, None]
            else:
                numpy.testing.assert_allclose(cen, r['centroid'], rtol=1e-10)
    # the lib_ / ID_ / _CDR colourings and count_gene (ref:902-928, 1347-1383) against the reference's own values
    for key, r in ref['colourings'].items():
        dic, tol = c.count_gene('timepoint', 2.0) if key.startswith('count ') else c.get_gene(key)
        assert tol == pytest.approx(r['tol'], rel=1e-12, abs=0)
        numpy.testing.assert_allclose([dic[k] for k in c.pl], numpy.array(r['p_nodes'])[order], rtol=1e-12, atol=0)
    # the table writer, RNG stream consumed gene by gene from `seed` (ref:1068-1071)
    genes = sorted(c.gene_row.keys())
    numpy.random.seed(ref['seed'])
    assert c.save(n=n, annotation={'setA': genes[3:9], 'setB': genes[5:6]}) is None       # ref:1053-1109 returns nothing
    header, rows = c.last_table
    with open(os.path.join(golden_dir, name + '.ref.genes.tsv')) as f:
        gold = [l[:-1].split('\t') for l in f]
    with open(prefix + '.genes.tsv') as f:
        mine = [l[:-1].split('\t') for l in f]
    assert mine[0] == gold[0] == header
    assert len(mine) == len(gold)
    _, ties = (None, None)
    for a, b in zip(mine[1:], gold[1:]):
        assert a[0] == b[0] and a[1] == b[1]
        for j in range(2, len(gold[0])):
            if gold[0][j] in ('setA', 'setB') or b[j] == 'None':
                assert a[j] == b[j]
            elif gold[0][j] in ('p_value', 'q-value (BH)'):
                continue
            else:
                assert float(a[j]) == pytest.approx(float(b[j]), rel=6e-12, abs=1e-300)
    # p-values: exact wherever no permuted score ties with the observed one
    numpy.random.seed(ref['seed'])
    keep = [r[0] for r in rows]
    pv, ties = c.connectivity_pvalues(keep, n=n, return_ties=True)
    n_tied = 0
    for a, b, p, t in zip(mine[1:], gold[1:], pv, ties):
        assert float(a[6]) == float('%.12g' % p)
        if t == 0:
            assert p == float(b[6])          # exact (the golden file holds Python-3 repr floats)
        else:
            # a permuted configuration reproduces the observed one (sparse / constant gene): the
            # reference decides `conn > observed` between two roundings of the same number, the
            # library counts the pair as equal
            n_tied += 1
            assert abs(float(a[6]) - float(b[6])) <= t / float(n) + 1e-15
    assert n_tied <= 3                      # the planted constant / single-cell genes of the golden cases, nothing else
    for a, b, p, t in zip(mine[1:], gold[1:], pv, ties):
        if t and numpy.ptp(c.Y[c.gene_row[a[0]]][c.samples]) == 0.0:
            assert p == float(b[6]) == 0.0  # a constant gene never exceeds itself: the reference run agrees
    for a, t in zip(mine[1:], ties):
        if t:
            x = c.Y[c.gene_row[a[0]]][c.samples]
            assert numpy.ptp(x) == 0.0 or numpy.count_nonzero(x) <= 2, 'tie on the non-degenerate gene %s' % a[0]
    if n_tied == 0:
        for a, b in zip(mine[1:], gold[1:]):
            assert float(a[7]) == pytest.approx(float(b[7]), rel=6e-12)


def test_invariants_at_scale():
    """Size-independent properties at a size the oracle cannot reach."""
    ctx = _gpu()
    from sctda_b200 import graph, synth
    n_cells, n_nodes, n_genes, n = 20000, 2000, 96, 64
    members, adj, Y = synth.connectivity_case(n_cells, n_nodes, n_genes, seed=5)
    Yd = torch.from_numpy(Y).cuda()
    dg = graph.DeviceGraph(members, adj, numpy.arange(n_cells), n_cells, 0)
    rng = numpy.random.RandomState(1)
    perms = numpy.stack([rng.permutation(n_cells) for _ in range(n)]).astype(numpy.int32)
    perms[0] = numpy.arange(n_cells)             # identity permutation first
    pd = torch.from_numpy(perms).cuda()
    # (1) p * n is an integer in [0, n]; scores finite; identity permutation reproduces the observed
    #     score up to the float32 rounding of ref:972 (linear mode: <= 1e-6 relative)
    st = graph.node_stats(ctx, dg, Yd, log2=False)
    ex, ti, sc = graph.conn_perm_test(ctx, dg, Yd, pd, st['conn'], log2=False, want_scores=True)
    ex = ex.cpu().numpy()
    sc = sc.cpu().numpy()
    obs = st['conn'].cpu().numpy()
    assert ((ex >= 0) & (ex <= n)).all()
    expressed = st['tol'].cpu().numpy() > 0.0        # an all-zero gene has tot = 0: NaN scores, p = 0 (ref:988-990)
    assert expressed.sum() >= n_genes - 8
    assert numpy.isfinite(sc[expressed]).all() and numpy.isnan(sc[~expressed]).all()
    assert (ex[~expressed] == 0).all()
    numpy.testing.assert_allclose(sc[expressed, 0], obs[expressed], rtol=1e-6)
    with numpy.errstate(invalid='ignore'):
        band = numpy.abs(sc - obs[:, None]) <= 1e-11 * numpy.abs(obs[:, None])      # ties do not exceed (header of this file)
        numpy.testing.assert_array_equal(ex, ((sc > obs[:, None]) & ~band).sum(axis=1))
    # (2) invariance under node relabelling (<= 1e-12): reverse the node order
    rev = list(range(n_nodes))[::-1]
    P = scipy.sparse.csr_matrix((numpy.ones(n_nodes), (numpy.arange(n_nodes), rev)), shape=(n_nodes, n_nodes))
    dg2 = graph.DeviceGraph([members[k] for k in rev], (P @ adj @ P.T).tocsr(), numpy.arange(n_cells), n_cells, 0)
    st2 = graph.node_stats(ctx, dg2, Yd, log2=False)
    _, _, sc2 = graph.conn_perm_test(ctx, dg2, Yd, pd, st2['conn'], log2=False, want_scores=True)
    numpy.testing.assert_allclose(st2['conn'].cpu().numpy()[expressed], obs[expressed], rtol=1e-12)
    numpy.testing.assert_allclose(sc2.cpu().numpy()[expressed], sc[expressed], rtol=1e-12)
    # (3) a uniform p has connectivity V/(V-1) * 2E/V^2: a constant gene in linear mode
    Yc = torch.full((1, n_cells), 3.0, dtype=torch.float64, device='cuda')
    stc = graph.node_stats(ctx, dg, Yc, log2=False)
    E2 = adj.nnz                                  # 2E
    assert stc['conn'][0].item() == pytest.approx(n_nodes / (n_nodes - 1.0) * E2 / n_nodes ** 2, rel=1e-12)
    # (4) shared block == the same block replicated per gene: two kernels (128-gene tiles / one gene per lane) with
    #     different but fixed partial-sum groupings of the node loop: identical counts, scores to the last bits
    few = 40
    rep = pd[:8].unsqueeze(0).repeat(few, 1, 1).contiguous()
    stl = graph.node_stats(ctx, dg, Yd[:few], log2=True)
    exa, _, sca = graph.conn_perm_test(ctx, dg, Yd[:few], pd[:8].contiguous(), stl['conn'], want_scores=True)
    exb, _, scb = graph.conn_perm_test(ctx, dg, Yd[:few], rep, stl['conn'], per_gene=True, want_scores=True)
    assert torch.equal(exa, exb)
    numpy.testing.assert_allclose(sca.cpu().numpy(), scb.cpu().numpy(), rtol=1e-13, atol=0, equal_nan=True)


def test_edge_cases_and_errors():
    ctx = _gpu()
    import ctypes
    from sctda_b200 import _lib, graph, synth
    members, adj, Y = synth.connectivity_case(200, 20, 3, seed=9)
    dg = graph.DeviceGraph(members, adj, numpy.arange(200), 200, 0)
    Yd = torch.from_numpy(Y).cuda()
    st = graph.node_stats(ctx, dg, Yd)
    # n = 0 permutations: counts are zero
    ex, ti, _ = graph.conn_perm_test(ctx, dg, Yd, torch.zeros((0, 200), dtype=torch.int32, device='cuda'), st['conn'])
    assert ex.sum().item() == 0
    # G = 0 genes
    ex, ti, _ = graph.conn_perm_test(ctx, dg, Yd[:0], torch.zeros((4, 200), dtype=torch.int32, device='cuda'),
                                     st['conn'][:0])
    assert ex.numel() == 0
    # a graph with a single node is rejected (V/(V-1)), with a message
    g1 = _lib.Graph(V=1, S=200, Ncells=200, Mtot=0, d_node_off=dg.node_off.data_ptr(),
                    d_node_cols=dg.node_cols.data_ptr(), d_adj_off=dg.adj_off.data_ptr(),
                    d_adj_idx=dg.adj_idx.data_ptr(), d_adj_w=dg.adj_w.data_ptr(), d_samples=dg.samples.data_ptr())
    rc = ctx.lib.sctda_node_stats(ctx.handle, ctypes.byref(g1), 3, _lib.ptr(Yd), 200, 1, None, None, None, None, None,
                                  None, None, None, None, None, None)
    assert rc == -1 and b'V >= 2' in ctx.lib.sctda_last_error(ctx.handle)
    # missing required pointer
    rc = ctx.lib.sctda_conn_perm_test(ctx.handle, ctypes.byref(dg.struct), 3, _lib.ptr(Yd), 200, 1, 4, None, 0,
                                      _lib.ptr(st['conn']), 0.0, None, None, None, None)
    assert rc == -1
    # cells outside the largest component: samples is a strict subset of the table rows
    sub = numpy.arange(0, 200, 2)
    members2 = [numpy.array([c for c in m if c % 2 == 0] or [0]) for m in members]
    dg2 = graph.DeviceGraph(members2, adj, sub, 200, 0)
    from oracle import conn_oracle as co
    state = co.GraphState([str(k) for k in range(20)], adj.toarray(),
                          {str(k): [int(c) for c in members2[k]] for k in range(20)}, {'g0': Y[0]})
    assert [int(x) for x in state.samples] == [int(x) for x in sub]
    st2 = graph.node_stats(ctx, dg2, Yd[:1])
    assert st2['conn'][0].item() == pytest.approx(co.connectivity(state, 'g0'), rel=1e-12)
    assert st2['expr'][0].item() == co.expr(state, 'g0')      # expr counts ALL table rows (ref:1050)


def test_reciprocal_division_is_ieee_division():
    """The kernel's div_rn (reciprocal + one fused correction) returns the bits of the IEEE
    division for the divisors of the hot path: node sizes and the constant 0.693147."""
    ctx = _gpu()
    import ctypes
    bad = ctypes.c_int64(-1)
    ctx.check(ctx.lib.sctda_selftest_div(ctx.handle, 1 << 26, ctypes.byref(bad)))
    assert bad.value == 0


def test_own_log1p_is_sub_ulp():
    """log1p_pos (argument reduction + degree-7 minimax series) against CUDA's log1p."""
    ctx = _gpu()
    import ctypes
    d = ctypes.c_int64(-1)
    ctx.check(ctx.lib.sctda_selftest_log1p(ctx.handle, 1 << 26, ctypes.byref(d)))
    assert 0 <= d.value <= 2          # two sub-ulp results differ by at most 2 units in the last place


def test_gene_gene_matrices(golden_dir, tmp_path):
    """JSD_matrix / cor_matrix / adjacency_matrix (ref:1111-1190) of the product class against the
    oracle restatement on a golden graph."""
    _gpu()
    from oracle import conn_oracle as co
    from sctda_b200 import graph
    for ext in ('.gexf', '.json', '.all.tsv'):
        shutil.copy(os.path.join(golden_dir, 'unrooted_a' + ext), str(tmp_path / ('unrooted_a' + ext)))
    prefix = str(tmp_path / 'unrooted_a')
    c = graph.UnrootedGraph(prefix, prefix + '.all.tsv', groups=False)
    st = co.load_state(prefix, prefix + '.all.tsv')
    st.pl = list(c.pl)
    st.adj = c.adj
    genes = ['G%03d' % i for i in (0, 3, 4, 5, 7, 8, 9, 10, 11, 12, 13, 20, 21, 22, 23, 6, 14, 15)]
    jsd, jsdiv = co.jsd_matrix(st, genes)
    got = c.JSD_matrix(genes)
    assert got.shape == (len(genes), len(genes)) and (numpy.diag(got) == 0.0).all()
    numpy.testing.assert_allclose(got ** 2, jsdiv, rtol=1e-9, atol=1e-13)
    numpy.testing.assert_allclose(c.adjacency_matrix(genes), co.adjacency_matrix(st, genes), rtol=1e-11, atol=0)
    numpy.testing.assert_allclose(c.adjacency_matrix(genes[:5], ind=2), co.adjacency_matrix(st, genes[:5], ind=2), rtol=1e-11, atol=0)
    sel = [g for g in genes if numpy.std(c.dicgenes[g]) > 0]         # a constant gene has no correlation distance
    numpy.testing.assert_allclose(c.cor_matrix(sel), co.cor_matrix(st, sel), rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize('name', ['unrooted_a', 'rooted_b'])
def test_gene_gene_matrices_match_reference_golden(golden_dir, tmp_path, name):
    """The product's JSD_matrix / cor_matrix / adjacency_matrix against the matrices the reference's own
    methods (ref:1111-1190) returned when its source was run on these files (oracle/gen_golden.py)."""
    _gpu()
    from sctda_b200 import graph
    for ext in ('.gexf', '.json', '.all.tsv'):
        shutil.copy(os.path.join(golden_dir, name + ext), str(tmp_path / (name + ext)))
    with open(os.path.join(golden_dir, name + '.ref.json')) as f:
        ref = json.load(f)['gene_matrices']
    prefix = str(tmp_path / name)
    c = graph.UnrootedGraph(prefix, prefix + '.all.tsv', groups=False)
    genes = ref['genes']
    numpy.testing.assert_allclose(c.JSD_matrix(genes) ** 2, numpy.array(ref['jsd']) ** 2, rtol=1e-9, atol=1e-13)
    numpy.testing.assert_allclose(c.cor_matrix(genes), numpy.array(ref['cor']), rtol=1e-9, atol=1e-12)
    numpy.testing.assert_allclose(c.adjacency_matrix(genes), numpy.array(ref['adjacency']), rtol=1e-11, atol=0)
    numpy.testing.assert_allclose(c.adjacency_matrix(genes, ind=2), numpy.array(ref['adjacency_2']), rtol=1e-11, atol=0)


@pytest.mark.parametrize('name', ['unrooted_a', 'rooted_b'])
def test_skeleton_and_subpopulations_match_reference_golden(golden_dir, tmp_path, name):
    """cellular_subpopulations (ref:1423-1446, 1950-2040 'js' route) on the gene table the reference wrote, and the
    skeleton of RootedGraph (dendritic_graph, edge / node sizes, select_diff_path: ref:1499-1593), against what the
    reference's own methods returned on these files (oracle/gen_golden.py)."""
    _gpu()
    from sctda_b200 import graph
    for ext in ('.gexf', '.json', '.all.tsv'):
        shutil.copy(os.path.join(golden_dir, name + ext), str(tmp_path / (name + ext)))
    shutil.copy(os.path.join(golden_dir, name + '.ref.genes.tsv'), str(tmp_path / (name + '.genes.tsv')))
    with open(os.path.join(golden_dir, name + '.ref.json')) as f:
        ref = json.load(f)
    prefix = str(tmp_path / name)
    r = ref['cellular_subpopulations']
    if ref['rooted']:
        c = graph.RootedGraph(prefix, prefix + '.all.tsv', groups=False)
        got = c.cellular_subpopulations(1e9, threshold=r['threshold'], min_cells=r['min_cells'], method='js',
                                        clus_thres=r['clus_thres'])
        k = ref['skeleton']
        assert [str(x) for x in c.g3.nodes()] == k['nodes']
        assert [[str(a), str(b)] for a, b in c.g3.edges()] == k['edges']
        assert {str(n): [sorted(comp) for comp in comps] for n, comps in c.dicdend.items()} == k['dicdend']
        assert c.edgesize == k['edgesize'] and c.nodesize == k['nodesize']
        assert [[a, b] for a, b in c.select_diff_path()] == k['select_diff_path']
        # the centroid route: two planted groups of centroids come back as the two clusters
        rows = ['\t'.join(['Gene', 'Cells', 'Mean', 'Min', 'Max', 'Connectivity', 'p_value', 'q-value (BH)', 'Centroid',
                           'Dispersion'])]
        rs = numpy.random.RandomState(4)
        for i in range(40):
            cen = (0.2 if i < 25 else 0.8) + 0.01 * rs.randn()
            rows.append('\t'.join(['g%02d' % i, '50', '1', '0', '2', '0.5', '0.001', '0.01', repr(float(cen)), '0.1']))
        with open(prefix + '.genes.tsv', 'w') as f:
            f.write('\n'.join(rows) + '\n')
        groups = c.cellular_subpopulations(0.5, max_K=5)
        assert sorted(sorted(g) for g in groups) == [['g%02d' % i for i in range(25)], ['g%02d' % i for i in range(25, 40)]]
    else:
        c = graph.UnrootedGraph(prefix, prefix + '.all.tsv', groups=False)
        got = c.cellular_subpopulations(threshold=r['threshold'], min_cells=r['min_cells'], clus_thres=r['clus_thres'])
    assert got == r['clusters']


@pytest.mark.parametrize('n_cells', [30000, 60000])
def test_large_components_use_the_other_staging_paths(n_cells):
    """S = 30,000: the permutation row is staged in shared memory without double buffering;
    S = 60,000: it does not fit and is gathered from global memory (and the lane masks of the 128-gene kernel do
    not fit its shared memory: the 64-gene kernel serves the shared block).  Both must give exactly the counts
    of the per-gene path (which the oracle tests pin) on a replicated block, and its scores to the last bits
    (the kernels group the partial sums of the node loop differently)."""
    ctx = _gpu()
    from sctda_b200 import graph, synth
    n_nodes, n_genes, n = 300, 40, 6
    members, adj, Y = synth.connectivity_case(n_cells, n_nodes, n_genes, seed=13)
    Yd = torch.from_numpy(Y).cuda()
    dg = graph.DeviceGraph(members, adj, numpy.arange(n_cells), n_cells, 0)
    rng = numpy.random.RandomState(2)
    perms = torch.from_numpy(numpy.stack([rng.permutation(n_cells) for _ in range(n)]).astype(numpy.int32)).cuda()
    st = graph.node_stats(ctx, dg, Yd, log2=True)
    exa, tia, sca = graph.conn_perm_test(ctx, dg, Yd, perms, st['conn'], want_scores=True)
    rep = perms.unsqueeze(0).repeat(n_genes, 1, 1).contiguous()
    exb, tib, scb = graph.conn_perm_test(ctx, dg, Yd, rep, st['conn'], per_gene=True, want_scores=True)
    assert torch.equal(exa, exb) and torch.equal(tia, tib)
    numpy.testing.assert_allclose(sca.cpu().numpy(), scb.cpu().numpy(), rtol=1e-13, atol=0, equal_nan=True)
    assert torch.isfinite(sca).any()


def _sharded_pvalues_worker(rank, ws, port, prefix, out_dir):
    import torch.distributed as dist
    from sctda_b200 import graph
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    try:
        c = graph.UnrootedGraph(prefix, prefix + '.all.tsv', groups=False)
        genes = sorted(c.gene_row.keys())
        perms = graph.reference_perm_indices(32, int(c.samples.size), numpy.random.RandomState(3))
        pv, ties = c.connectivity_pvalues(genes, n=32, perms=perms, return_ties=True)
        numpy.save(os.path.join(out_dir, 'pv%d.npy' % rank), numpy.asarray(pv))
        numpy.save(os.path.join(out_dir, 'ti%d.npy' % rank), numpy.asarray(ties))
    finally:
        dist.destroy_process_group()


def test_gene_sharded_pvalues_under_a_process_group_equal_the_single_process_result(tmp_path):
    """UnrootedGraph.connectivity_pvalues(perms=...) under torch.distributed (two ranks sharing cuda:0, gloo): the tile
    plan is broadcast, every rank takes every second 128-gene tile of the planned order, the counts are gathered and
    put back at their gene positions -- p-values and tie counts of every gene equal the single-process call."""
    _gpu()
    import socket
    import torch.multiprocessing as mp
    from oracle import gen_golden
    from sctda_b200 import graph
    prefix = str(tmp_path / 'case')
    gen_golden.write_case(prefix, n_cells=300, n_genes=420, n_nodes=50, seed=5)       # 421 columns: four tiles
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_sharded_pvalues_worker, args=(2, port, prefix, str(tmp_path)), nprocs=2, join=True)
    c = graph.UnrootedGraph(prefix, prefix + '.all.tsv', groups=False)
    genes = sorted(c.gene_row.keys())
    perms = graph.reference_perm_indices(32, int(c.samples.size), numpy.random.RandomState(3))
    pv, ties = c.connectivity_pvalues(genes, n=32, perms=perms, return_ties=True)
    assert len(genes) > 3 * 128
    for r in range(2):
        numpy.testing.assert_array_equal(numpy.load(str(tmp_path / ('pv%d.npy' % r))), numpy.asarray(pv))
        numpy.testing.assert_array_equal(numpy.load(str(tmp_path / ('ti%d.npy' % r))), numpy.asarray(ties))
    assert 0.0 < float(numpy.mean(pv)) < 1.0
