"""The CPU oracle (oracle/conn_oracle.py) against vectors produced by the reference's own
source (oracle/gen_golden.py).  CPU only."""
import json
import os

import numpy
import pytest

from oracle import conn_oracle as co


def _load(golden_dir, name):
    prefix = os.path.join(golden_dir, name)
    with open(prefix + '.ref.json') as f:
        ref = json.load(f)
    st = co.load_state(prefix, prefix + '.all.tsv')
    return prefix, ref, st


@pytest.mark.parametrize('name', ['unrooted_a', 'rooted_b'])
def test_state_matches_reference(golden_dir, name):
    _, ref, st = _load(golden_dir, name)
    assert st.pl == ref['pl']
    assert [int(x) for x in st.samples] == ref['samples']


@pytest.mark.parametrize('name', ['unrooted_a', 'rooted_b'])
def test_per_gene_statistics(golden_dir, name):
    _, ref, st = _load(golden_dir, name)
    for gi, r in ref['per_gene'].items():
        dic, tol = co.get_gene(st, gi)
        assert tol == pytest.approx(r['tol'], rel=1e-14, abs=0)
        numpy.testing.assert_allclose([dic[k] for k in st.pl], r['p_nodes'], rtol=1e-14, atol=0)
        assert co.connectivity(st, gi) == pytest.approx(r['connectivity'], rel=1e-13, abs=0)
        numpy.testing.assert_allclose(co.delta(st, gi), r['delta'], rtol=1e-13, atol=0)
        assert co.expr(st, gi) == r['expr']


@pytest.mark.parametrize('name', ['unrooted_a', 'rooted_b'])
def test_permutation_scores_and_pvalues(golden_dir, name):
    _, ref, st = _load(golden_dir, name)
    n = ref['n_perm']
    for gi, r in ref['per_gene'].items():
        numpy.random.seed(ref['seed'] + 7)
        scores = co.permuted_scores(st, gi, n=n)
        if r['tol'] > 0.0:
            numpy.testing.assert_allclose(scores, r['scores'], rtol=1e-12, atol=0)
        numpy.random.seed(ref['seed'] + 7)
        assert co.connectivity_pvalue(st, gi, n=n) == r['p_value']
        # the index-array form consumes the identical stream
        numpy.random.seed(ref['seed'] + 7)
        perms = co.perm_indices(n, len(st.samples))
        assert co.connectivity_pvalue(st, gi, n=n, perms=perms) == r['p_value']


def _read_tsv(path):
    with open(path) as f:
        lines = [l[:-1].split('\t') for l in f]
    return lines[0], lines[1:]


@pytest.mark.parametrize('name,rooted', [('unrooted_a', False), ('rooted_b', True)])
def test_save_table(golden_dir, tmp_path, name, rooted):
    prefix, ref, st = _load(golden_dir, name)
    genes = sorted(st.dicgenes.keys())
    if rooted:
        co.rooted(st, ref['root'])
        assert {k: int(v) for k, v in st.dicdis.items()} == ref['dicdis']
        assert [float(st.po[0]), float(st.po[1])] == pytest.approx(ref['po'], rel=1e-12)
    numpy.random.seed(ref['seed'])
    header, rows = co.save_rows(st, n=ref['n_perm'], annotation={'setA': genes[3:9], 'setB': genes[5:6]},
                                rooted_cols=rooted)
    out = str(tmp_path / 'o.genes.tsv')
    co.write_genes_tsv(out, header, rows)
    h_ref, r_ref = _read_tsv(prefix + '.ref.genes.tsv')
    h_new, r_new = _read_tsv(out)
    assert h_new == h_ref
    assert len(r_new) == len(r_ref)
    for a, b in zip(r_new, r_ref):
        assert a[0] == b[0] and a[1] == b[1]
        for j in range(2, len(h_ref)):
            if h_ref[j] in ('setA', 'setB') or b[j] == 'None':
                assert a[j] == b[j]
            else:
                # the golden file was written by Python 3 (repr floats); ours prints Python-2
                # str() = '%.12g', i.e. 12 significant digits
                assert float(a[j]) == pytest.approx(float(b[j]), rel=6e-12, abs=1e-300)
    for row, b in zip(rows, r_ref):
        for j in range(2, 8 + (2 if rooted else 0)):
            if b[j] != 'None':
                assert float(row[j]) == pytest.approx(float(b[j]), rel=2e-13, abs=1e-300)
        assert float(row[6]) == float(b[6])        # p-value: exact


def test_rooted_root_finding(golden_dir):
    _, ref, st = _load(golden_dir, 'rooted_b')
    den = co.dendrite(st)
    for k, v in ref['dendrite'].items():
        assert den[k] == pytest.approx(v, rel=1e-12, abs=1e-15, nan_ok=True)
    root, leaf = co.find_root(den)
    assert (str(root), str(leaf)) == (ref['root'], ref['leaf'])


def test_centroid(golden_dir):
    _, ref, st = _load(golden_dir, 'rooted_b')
    co.rooted(st, ref['root'])
    for gi, r in ref['per_gene'].items():
        cen = co.centroid(st, gi)
        if r['centroid'][0] is None:
            assert cen == [None, None]
        else:
            numpy.testing.assert_allclose(cen, r['centroid'], rtol=1e-12)


def test_benjamini_hochberg(golden_dir):
    with open(os.path.join(golden_dir, 'bh.ref.json')) as f:
        d = json.load(f)
    assert co.benjamini_hochberg(d['p']) == d['q']


def test_py2_str():
    assert co.py2_str(0.0) == '0.0'
    assert co.py2_str(0.904) == '0.904'
    assert co.py2_str(0.02489375434321234) == '0.0248937543432'
    assert co.py2_str(710) == '710'
    assert co.py2_str(1e-20) == '1e-20'
    assert co.py2_str(3.0) == '3.0'
    assert co.py2_str(None) == 'None'


@pytest.mark.parametrize('name', ['unrooted_a', 'rooted_b'])
def test_gene_gene_matrices_match_reference(golden_dir, name):
    """The oracle's JSD / correlation / adjacency matrices against the matrices the reference's own
    JSD_matrix, cor_matrix and adjacency_matrix (ref:1111-1190) returned under the shim."""
    with open(os.path.join(golden_dir, name + '.ref.json')) as f:
        ref = json.load(f)['gene_matrices']
    prefix = os.path.join(golden_dir, name)
    st = co.load_state(prefix, prefix + '.all.tsv')
    genes = ref['genes']
    jsd, _ = co.jsd_matrix(st, genes)
    numpy.testing.assert_allclose(jsd, numpy.array(ref['jsd']), rtol=1e-9, atol=1e-7)     # sqrt near 0 amplifies the last bits
    numpy.testing.assert_allclose(co.cor_matrix(st, genes), numpy.array(ref['cor']), rtol=1e-12, atol=1e-14)
    numpy.testing.assert_allclose(co.adjacency_matrix(st, genes), numpy.array(ref['adjacency']), rtol=1e-12, atol=0)
    numpy.testing.assert_allclose(co.adjacency_matrix(st, genes, ind=2), numpy.array(ref['adjacency_2']), rtol=1e-12, atol=0)

