"""CPU-side checks: the C-ABI library builds, loads and exports every symbol the header
declares (no compute calls); host logic (formats, BH, permutation source) against the oracle
and the golden vectors."""
import ctypes
import json
import os
import re

import numpy
import pytest

from oracle import conn_oracle as co

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib_path():
    from sctda_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(lib_path):
    from sctda_b200 import _lib
    with open(os.path.join(ROOT, 'include', 'sctda_b200.h')) as f:
        text = re.sub(r'/\*.*?\*/', '', f.read(), flags=re.S)
    declared = set(re.findall(r'\b(sctda_[a-z0-9_]+)\s*\(', text))
    assert declared, 'no declarations found'
    lib = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(lib, name), name
    # the ctypes binding covers exactly the declared surface
    assert declared == set(_lib.SIGNATURES.keys())
    loaded = _lib.load()
    assert loaded.sctda_abi_version() == 1


def test_no_gpu_means_loud_failure(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from sctda_b200 import _lib
    with pytest.raises(_lib.SctdaError):
        _lib.Context(0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, 'sctda_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), fn


def test_benjamini_hochberg_matches_reference_vectors(golden_dir):
    from sctda_b200 import graph
    with open(os.path.join(golden_dir, 'bh.ref.json')) as f:
        d = json.load(f)
    assert graph.benjamini_hochberg(d['p']) == d['q']
    rs = numpy.random.RandomState(0)
    for n in (1, 2, 7, 100):
        p = list(numpy.round(rs.rand(n), 2))
        assert graph.benjamini_hochberg(p) == co.benjamini_hochberg(p)
    assert graph.benjamini_hochberg([]) == []


def test_permutation_source_is_the_reference_stream():
    from sctda_b200 import graph
    numpy.random.seed(4242)
    a = graph.reference_perm_indices(5, 37)
    numpy.random.seed(4242)
    x = numpy.tile(numpy.arange(37) * 1.5, (5, 1))
    for row in x:                       # ref:976
        numpy.random.shuffle(row)
    numpy.testing.assert_array_equal(x, (numpy.arange(37) * 1.5)[a])
    assert sorted(a[3]) == list(range(37))
    # the C generator (sctda_mt19937_shuffle_rows) is numpy's, draw for draw, and hands the state back
    for seed, n, S in ((1, 7, 50), (99, 33, 213), (5, 3, 1), (9, 20, 4097), (3, 2, 70000), (8, 0, 10)):
        numpy.random.seed(seed)
        a = graph.reference_perm_indices(n, S)
        after_a = numpy.random.randint(0, 1 << 30, 4)
        numpy.random.seed(seed)
        b = co.perm_indices(n, S)
        after_b = numpy.random.randint(0, 1 << 30, 4)
        assert (a == b).all() and (after_a == after_b).all()
    rs, rs2 = numpy.random.RandomState(77), numpy.random.RandomState(77)
    assert (graph.reference_perm_indices(5, 100, rs) == co.perm_indices(5, 100, rs2)).all() and rs.rand() == rs2.rand()


@pytest.mark.parametrize('name', ['unrooted_a', 'rooted_b'])
def test_read_table_matches_oracle_parse(golden_dir, name):
    from sctda_b200 import formats
    path = os.path.join(golden_dir, name + '.all.tsv')
    names, Y, cell_id, libs, cdr = formats.read_table(path)
    dicgenes, cell_id2, libs2, cdr2 = co.parse_table(path)
    assert names == list(dicgenes.keys())
    for i, q in enumerate(names):
        numpy.testing.assert_array_equal(Y[i], dicgenes[q])
    assert list(cell_id) == cell_id2 and list(libs) == libs2
    numpy.testing.assert_array_equal(cdr, cdr2)
    # shift variants (ref:875-878)
    n3, Y3, _, _, _ = formats.read_table(path, shift=3)
    assert n3 == names[3:]
    numpy.testing.assert_array_equal(Y3, Y[3:])
    n4, Y4, _, _, _ = formats.read_table(path, shift=[3, 5])
    assert n4 == [names[3], names[5]]


def test_table_sidecar_round_trip(golden_dir, tmp_path):
    import shutil
    from sctda_b200 import formats
    path = str(tmp_path / 't.all.tsv')
    shutil.copy(os.path.join(golden_dir, 'unrooted_a.all.tsv'), path)
    names, Y, cell_id, libs, cdr = formats.read_table(path)
    formats.write_table_sidecar(path, names, Y, cell_id, libs, cdr)
    n2, Y2, c2, l2, cdr2 = formats.read_table(path)
    assert n2 == names and list(c2) == list(cell_id) and list(l2) == list(libs)
    numpy.testing.assert_array_equal(Y2, Y)
    numpy.testing.assert_array_equal(cdr2, cdr)
    assert formats._read_sidecar(path) is not None
    assert formats._read_sidecar(path, csv=True) is None            # written for the tab-separated reading
    # a table rewritten after its side-car (other size or mtime) is parsed again
    os.utime(path, ns=(1, 1))
    assert formats._read_sidecar(path) is None
    # a side-car that holds pickled objects is never loaded
    formats.write_table_sidecar(path, names, Y, cell_id, libs, cdr)
    numpy.savez(path + '.npz', names=numpy.array(names, dtype=object), Y=Y, table=formats._table_fingerprint(path),
                csv=numpy.array([0]), cell_id=numpy.array(cell_id, dtype=object), libs=numpy.array(libs, dtype=object),
                cdr=cdr)
    assert formats._read_sidecar(path) is None


def test_py2_str_and_writers(tmp_path):
    from sctda_b200 import formats
    for v in (0.0, 0.904, 0.02489375434321234, 710, 1e-20, 3.0, None, 1e16, 123456789012.0, float('nan')):
        assert formats.py2_str(v) == co.py2_str(v)
    import networkx
    p = str(tmp_path / 'g')
    formats.write_gexf(p + '.gexf', 4, [(0, 1), (1, 3)])
    g = networkx.read_gexf(p + '.gexf')
    assert sorted(g.nodes()) == ['0', '1', '2', '3']
    assert sorted(tuple(sorted(e)) for e in g.edges()) == [('0', '1'), ('1', '3')]
    assert all(d['weight'] == 1.0 for _, _, d in g.edges(data=True))
    formats.write_mapper_json(p + '.json', [[3, 1], numpy.array([2])])
    with open(p + '.json') as f:
        assert json.load(f) == {'0': [3, 1], '1': [2]}


def test_cover_bounds_device_form_is_numpy_percentile():
    """The torch form of the cover fenceposts (used on the device) against numpy.percentile / linspace,
    bit for bit (run here on CPU tensors; the GPU parity tests compare the resulting patches)."""
    import torch
    from oracle import mapper_oracle as mo
    from sctda_b200 import mapper
    rng = numpy.random.RandomState(0)
    for n in (7, 100, 1001):
        lens = rng.normal(size=(n, 2))
        lens[3] = lens[4]
        for r in (1, 5, 20, 30):
            for eq in (True, False):
                for gain in (0.4, 3.0):
                    d = mapper.cover_bounds_device(torch.from_numpy(lens), r, gain, eq)
                    h = mapper.cover_bounds(lens, r, gain, eq)
                    for a, b in zip(d, h):
                        assert (a.numpy().view(numpy.int64) == numpy.ascontiguousarray(b).view(numpy.int64)).all()
                    lo0, hi0 = mo.intervals(mo.fenceposts(lens[:, 0], r, eq), gain)
                    assert (d[0].numpy() == lo0).all() and (d[1].numpy() == hi0).all()


def test_package_exports_the_reference_names_on_the_path():
    import sctda_b200 as scTDA
    for name in ('TopologicalRepresentation', 'UnrootedGraph', 'RootedGraph', 'benjamini_hochberg'):
        assert callable(getattr(scTDA, name))
    with pytest.raises(AttributeError):
        scTDA.Preprocess


def test_native_table_parser_is_python_float(lib_path, golden_dir, tmp_path):
    """sctda_parse_table_f64 (all host cores) against the per-field float() loop of the reference
    (ref:859-874): same values bit for bit, same non-numeric fields, on the golden tables and on a
    table of awkward fields."""
    import random
    from sctda_b200 import formats
    for name in ('unrooted_a', 'rooted_b'):
        path = os.path.join(golden_dir, name + '.all.tsv')
        a = formats.read_table(path, native=True)
        b = formats.read_table(path, native=False)
        assert a[0] == b[0] and list(a[2]) == list(b[2]) and list(a[3]) == list(b[3])
        assert (a[1].view(numpy.int64) == b[1].view(numpy.int64)).all()
        numpy.testing.assert_array_equal(a[4], b[4])
        for shift in (3, [3, 5]):
            x, y = formats.read_table(path, shift=shift, native=True), formats.read_table(path, shift=shift, native=False)
            assert x[0] == y[0] and (x[1].view(numpy.int64) == y[1].view(numpy.int64)).all()
    rnd = random.Random(5)
    rs = numpy.random.RandomState(5)
    odd = ['', 'nan', 'inf', '-inf', ' 1.5', '1.5 ', '1_0', '0x10', '1e', 'e5', '.', '-', '+', '1.', '.5', '+.5e-3', '-0', '-0.0',
           '1e400', '1e-400', '4.9e-324', '2.2250738585072011e-308', '17976931348623157' + '0' * 292, '0.1', '0.30000000000000004',
           '123456789012345678901234567890', '0.000000000000000000000000000001234567890123456789', '9007199254740993',
           '9007199254740992.5', '1E5', '1e+5', '00012', 'abc', 'D3_L_7', 'L', '1,5', '1|2', '३']
    vals = []
    for _ in range(3000):
        k = rnd.randrange(6)
        x = float(rs.standard_normal() * 10.0 ** rnd.randrange(-30, 30))
        vals.append([repr(x), '%.12g' % x, '%d' % int(rs.randint(-10 ** 9, 10 ** 9)), '%.3e' % x, '%.25f' % (x % 1000.0),
                     rnd.choice(odd)][k])
    ncol = 40
    rows = [vals[i:i + ncol] for i in range(0, len(vals) - ncol + 1, ncol)]
    header = ['ID', 'timepoint', 'lib'] + ['g%d' % i for i in range(ncol - 3)]
    for sep, csv, trailing in (('\t', False, True), ('\t', False, False), (',', True, True)):
        path = str(tmp_path / ('odd%d%d.tsv' % (csv, trailing)))
        body = [[f.replace(',', ';') for f in r] for r in rows] if csv else rows
        with open(path, 'w', encoding='utf-8') as f:
            f.write(sep.join(header) + '\n' + '\n'.join(sep.join(r) for r in body) + ('\n' if trailing else ''))
        a = formats.read_table(path, csv=csv, native=True)
        b = formats.read_table(path, csv=csv, native=False)
        assert a[0] == b[0] and list(a[2]) == list(b[2]) and list(a[3]) == list(b[3])
        same = (a[1].view(numpy.int64) == b[1].view(numpy.int64)) | (numpy.isnan(a[1]) & numpy.isnan(b[1]))
        assert same.all(), numpy.argwhere(~same)[:5]
        numpy.testing.assert_array_equal(a[4], b[4])
    # a row with a missing / an extra field, an empty line: errors, as in the Python path
    for bad in ('a\tb\n1\t2\n3\n', 'a\tb\n1\t2\t3\n', 'a\tb\n1\t2\n\n3\t4\n'):
        path = str(tmp_path / 'bad.tsv')
        with open(path, 'w') as f:
            f.write(bad)
        with pytest.raises(ValueError):
            formats.read_table(path, native=True)
        with pytest.raises(ValueError):
            formats.read_table(path, native=False)
    # carriage returns: the text-mode path decides (universal newlines)
    path = str(tmp_path / 'crlf.tsv')
    with open(path, 'wb') as f:
        f.write(b'a\tb\r\n1\t2\r\n')
    assert formats._parse_native(path, False) is None
    numpy.testing.assert_array_equal(formats.read_table(path)[1], [[1.0], [2.0]])


def test_native_table_parser_fuzz(lib_path, tmp_path):
    """Random small tables over an alphabet of digits, signs, exponents, separators and junk: the
    native reader and the per-field float() reader agree on values, strings and on which inputs are errors."""
    import random
    from sctda_b200 import formats
    rnd = random.Random(11)
    alphabet = '0123456789.eE+- ,|_xnaif'
    p = str(tmp_path / 'f.tsv')
    agreed = 0
    for _ in range(600):
        ncol, nrow = rnd.randrange(1, 5), rnd.randrange(0, 5)
        lines = ['\t'.join('h%d' % i for i in range(ncol))]
        for _r in range(nrow):
            k = ncol if rnd.random() < 0.9 else rnd.randrange(0, 6)
            lines.append('\t'.join(''.join(rnd.choice(alphabet) for _ in range(rnd.randrange(0, 8))) for _ in range(k)))
        with open(p, 'w') as f:
            f.write('\n'.join(lines) + ('\n' if rnd.random() < 0.7 else ''))
        for csv in (False, True):
            res = []
            for native in (True, False):
                try:
                    res.append(formats.read_table(p, csv=csv, native=native))
                except ValueError:
                    res.append(None)
            a, b = res
            assert (a is None) == (b is None)
            if a is None:
                continue
            assert a[0] == b[0] and list(a[2]) == list(b[2]) and list(a[3]) == list(b[3]) and a[1].shape == b[1].shape
            assert ((a[1].view(numpy.int64) == b[1].view(numpy.int64)) | (numpy.isnan(a[1]) & numpy.isnan(b[1]))).all()
            assert numpy.array_equal(a[4], b[4], equal_nan=True)
            agreed += 1
    assert agreed > 500


def _preprocess_save_reference_rows(genes, counts, long, batch, tot, ws, wg):
    """ref:663-728 restated with Python-3 formatting helpers (the rows of name.all.tsv / name.mapper.tsv)."""
    datat = counts.T
    all_rows, map_rows = ['ID\ttimepoint\tlib\t' + '\t'.join(genes)], ['\t'.join(g for g, m in zip(genes, wg) if m)]
    for n, m in enumerate(ws):
        if m:
            p = 'D' + co.py2_str(long[n]) + '_' + batch[n] + '_' + str(n) + '\t' + co.py2_str(long[n]) + '\t' + batch[n] + '\t'
            all_rows.append(p + '\t'.join(co.py2_str(numpy.log2(1.0 + 1000000.0 * t / float(tot[n]))) for t in datat[n]))
            map_rows.append('\t'.join(co.py2_str(numpy.log2(1.0 + 1000000.0 * t / float(tot[n]))) for t in datat[n, wg]))
    return all_rows, map_rows


def test_preprocess_table_writer_round_trip(tmp_path):
    """formats.write_preprocess_tables (Preprocess.save, ref:663-728): every byte as the reference's loop
    writes it, and the files read back through read_table / pandas give the values to 12 digits."""
    from sctda_b200 import formats
    rng = numpy.random.RandomState(4)
    G, ncell = 57, 230
    genes = ['g%d' % i for i in range(G)]
    counts = rng.poisson(rng.gamma(0.3, 8.0, (G, ncell))).astype(numpy.float64)
    counts[:, 7] = 0.0
    counts[3, 7] = 1.0                                           # a cell with one transcript: log2(1 + 1e6) fields
    long = rng.randint(0, 5, ncell)
    batch = ['L%d' % (i % 3) for i in range(ncell)]
    tot = counts.sum(axis=0)
    ws = rng.rand(ncell) < 0.8
    ws[7] = True
    wg = rng.rand(G) < 0.5
    name = str(tmp_path / 'pp')
    formats.write_preprocess_tables(name, genes, counts, long, batch, totaltran=tot, which_samples=ws, which_genes=wg,
                                    block_rows=64)
    all_rows, map_rows = _preprocess_save_reference_rows(genes, counts, long, batch, tot, ws, wg)
    with open(name + '.all.tsv') as f:
        assert f.read() == '\n'.join(all_rows) + '\n'
    with open(name + '.mapper.tsv') as f:
        assert f.read() == '\n'.join(map_rows) + '\n'
    names, Y, cell_id, libs, _ = formats.read_table(name + '.all.tsv', sidecar=False)
    assert names[3:] == genes and list(libs) == [batch[n] for n in numpy.nonzero(ws)[0]]
    expect = numpy.log2(1.0 + 1000000.0 * counts[:, ws] / tot[ws][None, :])
    numpy.testing.assert_allclose(Y[3:], expect, rtol=1e-11, atol=0)
    # subsampled state: the extra file, the subsampling target as denominator in all three
    sub = numpy.minimum(counts, 3.0)
    formats.write_preprocess_tables(name + 's', genes, counts, long, batch, which_samples=ws, which_genes=wg,
                                    counts_subsampled=sub, target_subsample=500)
    _, Ys, _, _, _ = formats.read_table(name + 's.all.tsv', sidecar=False)
    _, Yn, _, _, _ = formats.read_table(name + 's.no_subsampling.tsv', sidecar=False)
    numpy.testing.assert_allclose(Ys[3:], numpy.log2(1.0 + 1000000.0 * sub[:, ws] / 500.0), rtol=1e-11)
    numpy.testing.assert_allclose(Yn[3:], numpy.log2(1.0 + 1000000.0 * counts[:, ws] / 500.0), rtol=1e-11)


def test_native_formatter_fuzz():
    """sctda_format_table_f64 against formats.py2_str on every class of value."""
    from sctda_b200 import formats
    rng = numpy.random.RandomState(9)
    v = rng.lognormal(0.0, 12.0, (200, 64)) * rng.choice([-1.0, 1.0], (200, 64))
    v[rng.rand(200, 64) < 0.3] = 0.0
    v[0, :8] = [numpy.nan, numpy.inf, -numpy.inf, 1e16, 123456789012.0, -0.0, 0.1 + 0.2, 1e-5]
    v[1] = numpy.round(v[1])
    v[2, :4] = [5e-324, 1.7976931348623157e308, 999999999999.5, 99999999999.95]
    rows = formats.format_rows(v)
    for r in range(v.shape[0]):
        assert rows[r].decode() == '\t'.join(formats.py2_str(x) for x in v[r])


def test_hierarchical_clustering_and_find_clusters():
    """ref:204-266 without the figure: the colour groups partition the leaves; every group but the
    default-colour one is a flat cluster of scipy's fcluster at the same height."""
    import scipy.cluster.hierarchy as sch
    import scipy.spatial.distance
    from sctda_b200 import graph
    rs = numpy.random.RandomState(2)
    pts = numpy.concatenate([rs.randn(6, 3) + 8.0, rs.randn(5, 3) - 8.0, 40.0 * rs.randn(3, 3)])
    mat = scipy.spatial.distance.squareform(scipy.spatial.distance.pdist(pts))
    for cluster_distance in (True, False):
        z = graph.hierarchical_clustering(mat, cluster_distance=cluster_distance, thres=0.3)
        groups = graph.find_clusters(z)
        leaves = sorted(x for g in groups.values() for x in g)
        assert leaves == list(range(len(pts)))
        tri = scipy.spatial.distance.pdist(mat) if cluster_distance else scipy.spatial.distance.squareform(mat)
        Y = sch.linkage(tri, method='average')
        flat = sch.fcluster(Y, 0.3 * Y[:, 2].max(), 'distance')
        big = sorted(sorted(numpy.nonzero(flat == f)[0].tolist()) for f in set(flat) if (flat == f).sum() > 1)
        mine = sorted(sorted(g) for g in groups.values())
        for b in big:
            assert b in mine
