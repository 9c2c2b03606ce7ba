"""File formats on the hot path: the expression table reader, the
Python-2 float formatting of `name.genes.tsv`, the `name.json` / `name.gexf` writers.

Reference: /root/reference/scTDA/main.py ("ref:") — table parse ref:835-885, genes.tsv rows
ref:1063-1086, json/gexf ref:772-777.
"""
import json
import math

import numpy


def py2_str(v):
    """What Python 2's str() printed for the values ref:1078-1079 writes: 12 significant digits
    for floats, with '.0' appended when that looks like an integer."""
    if v is None:
        return 'None'
    if isinstance(v, str):
        return v
    if isinstance(v, (bool, numpy.bool_)):
        return 'True' if v else 'False'
    if isinstance(v, (int, numpy.integer)):
        return '%d' % int(v)
    x = float(v)
    if math.isnan(x):
        return 'nan'
    if math.isinf(x):
        return '-inf' if x < 0 else 'inf'
    s = '%.12g' % x
    if not any(ch in s for ch in '.e'):
        s += '.0'
    return s


def _to_float_column(strings):
    """Column of table fields -> (float64 array, is_number mask), ref:118-126 semantics:
    a field is a number iff float() accepts it; everything else becomes 0.0."""
    try:
        return numpy.asarray(strings, dtype=numpy.float64), None
    except ValueError:
        pass
    out = numpy.zeros(len(strings), dtype=numpy.float64)
    ok = numpy.zeros(len(strings), dtype=bool)
    for i, s in enumerate(strings):
        try:
            out[i] = float(s)
            ok[i] = True
        except ValueError:
            pass
    return out, ok


def write_table_sidecar(path, names, Y, cell_id=(), libs=(), cdr=None):
    """Binary side-car `<table>.npz` of an expression table: the TSV stays the
    interchange format of the reference; when the side-car is present and newer than the TSV,
    read_table loads it instead of parsing 10^8-10^9 text fields in Python."""
    numpy.savez(path + '.npz', names=numpy.array(list(names), dtype=object), Y=numpy.asarray(Y, dtype=numpy.float64),
                cell_id=numpy.array(list(cell_id), dtype=object), libs=numpy.array(list(libs), dtype=object),
                cdr=numpy.zeros(0) if cdr is None else numpy.asarray(cdr, dtype=numpy.float64))


def _read_sidecar(path):
    import os
    side = path + '.npz'
    if not os.path.exists(side) or os.path.getmtime(side) < os.path.getmtime(path):
        return None
    z = numpy.load(side, allow_pickle=True)
    return list(z['names']), numpy.ascontiguousarray(z['Y']), list(z['cell_id']), list(z['libs']), z['cdr']


def read_table(path, shift=None, csv=False, sidecar=True):
    """ref:835-885.  Returns (names, Y, cell_id, libs, cdr):
    names   column names that become keys of dicgenes (header order),
    Y       float64 [len(names), Ncells], gene-major (one row per column of the file),
    cell_id / libs  the non-numeric fields of column 0 / column 'lib' (shift=None only),
    cdr     fraction of columns with value > 0 per cell, excluding ID / lib / timepoint."""
    if sidecar and shift is None:
        got = _read_sidecar(path)
        if got is not None:
            return got
    sep = ',' if csv else '\t'
    with open(path, 'r') as f:
        lines = f.read().split('\n')
    if lines and lines[-1] == '':
        lines.pop()
    rows = []
    for ln in lines:
        sp = ln.split(sep)
        if csv:
            sp = [x.split('|')[0] for x in sp]
        rows.append(sp)
    header, body = rows[0], rows[1:]
    ncol = len(header)
    for i, r in enumerate(body):
        if len(r) != ncol:
            raise ValueError('%s: row %d has %d fields, header has %d' % (path, i + 1, len(r), ncol))
    ncell = len(body)
    cols = list(zip(*body)) if body else [() for _ in header]
    cell_id, libs = [], []
    if shift is None:
        names = list(header)
        coyt = header.index('lib') if 'lib' in header else -1
        ttt = header.index('timepoint') if 'timepoint' in header else -1
        Y = numpy.zeros((ncol, ncell), dtype=numpy.float64)
        cdr = numpy.zeros(ncell, dtype=numpy.float64)
        for c in range(ncol):
            vals, ok = _to_float_column(cols[c])
            Y[c] = vals
            isnum = numpy.ones(ncell, dtype=bool) if ok is None else ok
            if c != 0 and c != coyt and c != ttt:
                cdr += (isnum & (vals > 0.0))
            if ok is not None:
                bad = numpy.nonzero(~ok)[0]
                if c == coyt:
                    libs = [cols[c][i] for i in bad]
                elif c == 0:
                    cell_id = [cols[c][i] for i in bad]
        cdr = cdr / float(max(ncell, 1))
    else:
        if isinstance(shift, int):
            sel = list(range(shift, ncol))
        else:
            sel = list(numpy.arange(ncol)[shift])
        names = [header[c] for c in sel]
        Y = numpy.zeros((len(sel), ncell), dtype=numpy.float64)
        for k, c in enumerate(sel):
            Y[k] = numpy.asarray(cols[c], dtype=numpy.float64)    # the reference float()s every field here
        cdr = numpy.zeros(0)
    return names, Y, cell_id, libs, cdr


def write_genes_tsv(path, header, rows):
    """ref:1063-1067, 1078-1085."""
    with open(path, 'w') as f:
        f.write('\t'.join(header) + '\n')
        for row in rows:
            f.write('\t'.join(py2_str(x) for x in row) + '\n')


def write_mapper_json(path, clusters):
    """ref:772-776: {str(node): [int cell rows]}."""
    dic = {}
    for k, members in enumerate(clusters):
        dic[str(k)] = [int(x) for x in members]
    with open(path, 'w') as f:
        json.dump(dic, f)


def write_gexf(path, n_nodes, edges, weights=None):
    """GEXF 1.2draft as networkx.write_gexf emits for the sakmapper graph (ref:777;
    SURVEY.md appendix B): undirected static graph, node id == label == str(k), one
    <edge> per pair with weight="1.0".  Written directly (V, E can be 1e5+)."""
    with open(path, 'w') as f:
        f.write("<?xml version='1.0' encoding='utf-8'?>\n")
        f.write('<gexf xmlns="http://www.gexf.net/1.2draft" '
                'xmlns:xsi="http://www.w3.org/2001/XMLSchema-instance" '
                'xsi:schemaLocation="http://www.gexf.net/1.2draft http://www.gexf.net/1.2draft/gexf.xsd" '
                'version="1.2">\n')
        f.write('  <meta>\n    <creator>sctda_b200</creator>\n  </meta>\n')
        f.write('  <graph defaultedgetype="undirected" mode="static" name="">\n')
        f.write('    <nodes>\n')
        for k in range(int(n_nodes)):
            f.write('      <node id="%d" label="%d" />\n' % (k, k))
        f.write('    </nodes>\n    <edges>\n')
        for e, (i, j) in enumerate(edges):
            w = 1.0 if weights is None else float(weights[e])
            f.write('      <edge source="%d" target="%d" id="%d" weight="%s" />\n' % (int(i), int(j), e, repr(w)))
        f.write('    </edges>\n  </graph>\n</gexf>\n')
