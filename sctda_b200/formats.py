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


def _table_fingerprint(path):
    import os
    st = os.stat(path)
    return numpy.array([st.st_size, st.st_mtime_ns], dtype=numpy.int64)


def write_table_sidecar(path, names, Y, cell_id=(), libs=(), cdr=None, csv=False):
    """Binary side-car `<table>.npz` of an expression table: the TSV stays the
    interchange format of the reference; read_table loads the side-car instead of parsing 10^8-10^9
    text fields when it was written for exactly this file (size and mtime of the table are recorded)
    and for the same `csv` flag.  Strings are stored as fixed-width unicode arrays: nothing in the
    file is a pickle."""
    numpy.savez(path + '.npz', names=numpy.array(list(names), dtype=numpy.str_),
                Y=numpy.asarray(Y, dtype=numpy.float64), cell_id=numpy.array(list(cell_id), dtype=numpy.str_),
                libs=numpy.array(list(libs), dtype=numpy.str_),
                cdr=numpy.zeros(0) if cdr is None else numpy.asarray(cdr, dtype=numpy.float64),
                table=_table_fingerprint(path), csv=numpy.array([1 if csv else 0], dtype=numpy.int64))


def _read_sidecar(path, csv=False):
    import os
    side = path + '.npz'
    if not os.path.exists(side):
        return None
    try:
        z = numpy.load(side, allow_pickle=False)
        if 'table' not in z.files or 'csv' not in z.files:
            return None
        if not (z['table'] == _table_fingerprint(path)).all() or int(z['csv'][0]) != (1 if csv else 0):
            return None                                   # written for another version of the table
        return ([str(x) for x in z['names']], numpy.ascontiguousarray(z['Y']), [str(x) for x in z['cell_id']],
                [str(x) for x in z['libs']], z['cdr'])
    except (ValueError, OSError, KeyError):
        return None                                       # not one of ours (e.g. an object array): parse the text


def _parse_native(path, csv):
    """Header fields and the table body parsed by sctda_parse_table_f64 (all host cores):
    (header, Y [ncol, nrow] float64, bad) with bad = [(row, col, text)] for every field the native
    parser left to Python's float(), in row-major order.  None when the native path does not apply
    (library not built, carriage returns in the file, empty file)."""
    import ctypes
    try:
        from . import _lib
        lib = _lib.load()
    except (ImportError, OSError):
        return None
    with open(path, 'rb') as f:
        data = f.read()
    if not data or b'\r' in data:
        return None
    nl = data.find(b'\n')
    head = data if nl < 0 else data[:nl]
    body = b'' if nl < 0 else data[nl + 1:]
    sep = ',' if csv else '\t'
    header = head.decode('utf-8').split(sep)
    if csv:
        header = [x.split('|')[0] for x in header]
    ncol = len(header)
    nrow = body.count(b'\n') + (1 if body and not body.endswith(b'\n') else 0)
    Y = numpy.empty((ncol, nrow), dtype=numpy.float64)
    cap = max(1024, 4 * nrow)
    while True:
        pos = numpy.empty(2 * cap, dtype=numpy.int64)
        ln = numpy.empty(cap, dtype=numpy.int32)
        n_bad, bad_row = ctypes.c_int64(0), ctypes.c_int64(-1)
        rc = lib.sctda_parse_table_f64(body, len(body), ord(sep), 1 if csv else 0, nrow, ncol, Y.ctypes.data_as(ctypes.c_void_p),
                                       pos.ctypes.data_as(ctypes.c_void_p), ln.ctypes.data_as(ctypes.c_void_p), cap,
                                       ctypes.byref(n_bad), ctypes.byref(bad_row), 0)
        if rc != 0:
            r = int(bad_row.value)
            raise ValueError('%s: row %d does not have the %d fields of the header' % (path, r + 1 if r >= 0 else -1, ncol))
        if n_bad.value <= cap:
            break
        cap = int(n_bad.value)
    nb = int(n_bad.value)
    flat, off = pos[0:2 * nb:2], pos[1:2 * nb:2]
    order = numpy.argsort(flat, kind='stable')
    bad = []
    for i in order:
        k, o, L = int(flat[i]), int(off[i]), int(ln[i])
        bad.append((k // ncol, k % ncol, body[o:o + L].decode('utf-8')))
    return header, Y, bad


def read_table(path, shift=None, csv=False, sidecar=True, native=True):
    """ref:835-885.  Returns (names, Y, cell_id, libs, cdr):
    names   column names that become keys of dicgenes (header order),
    Y       float64 [len(names), Ncells], gene-major (one row per column of the file),
    cell_id / libs  the non-numeric fields of column 0 / column 'lib' (shift=None only),
    cdr     fraction of columns with value > 0 per cell, excluding ID / lib / timepoint."""
    if sidecar and shift is None:
        got = _read_sidecar(path, csv)
        if got is not None:
            return got
    got = _parse_native(path, csv) if native else None
    if got is not None:
        return _table_from_native(path, got, shift)
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


def _table_from_native(path, got, shift):
    """read_table's result from the natively parsed body: the fields the parser left undecided go
    through Python's float() exactly as ref:118-126 / ref:859-874 treat them."""
    header, Y, bad = got
    ncol, ncell = Y.shape
    strings = {}                                               # col -> [(row, text)] of the non-numeric fields
    for r, c, text in bad:
        try:
            Y[c, r] = float(text)
        except ValueError:
            strings.setdefault(c, []).append((r, text))
    if shift is None:
        coyt = header.index('lib') if 'lib' in header else -1
        ttt = header.index('timepoint') if 'timepoint' in header else -1
        cell_id = [t for _, t in strings.get(0, [])]
        libs = [t for _, t in strings.get(coyt, [])] if coyt >= 0 else []
        if coyt == 0:
            cell_id = []                                       # ref:867-872: the 'lib' test comes first
        sel = [c for c in range(ncol) if c != 0 and c != coyt and c != ttt]
        cdr = numpy.zeros(ncell, dtype=numpy.float64)
        for lo in range(0, len(sel), 256):                     # non-numeric fields are 0.0, hence never > 0
            cdr += (Y[sel[lo:lo + 256]] > 0.0).sum(axis=0)
        cdr = cdr / float(max(ncell, 1))
        return list(header), Y, cell_id, libs, cdr
    sel = list(range(shift, ncol)) if isinstance(shift, int) else list(numpy.arange(ncol)[shift])
    for c in sel:
        if c in strings:                                       # the reference float()s every field here
            raise ValueError('could not convert string to float: %r' % (strings[c][0][1],))
    return [header[c] for c in sel], numpy.ascontiguousarray(Y[sel]), [], [], numpy.zeros(0)


def format_rows(values, denom=None, log2_tpm=False, sep='\t', nthreads=0):
    """The numeric part of table rows as Preprocess.save writes it (ref:686-693): for every row of `values`
    [rows, cols] the fields str(v) (log2_tpm: str(numpy.log2(1.0 + 1000000.0 * v / denom[row]))) under
    Python 2's float printing, joined by `sep`.  Formatted by all host cores (sctda_format_table_f64);
    returns a list of bytes objects, one per row."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    v = numpy.ascontiguousarray(values, dtype=numpy.float64)
    rows, cols = v.shape
    if rows == 0:
        return []
    den = None if denom is None else numpy.ascontiguousarray(denom, dtype=numpy.float64)
    out = numpy.empty(rows * cols * 25, dtype=numpy.uint8)
    end = numpy.empty(rows, dtype=numpy.int64)
    rc = lib.sctda_format_table_f64(v.ctypes.data, cols, rows, cols, None if den is None else den.ctypes.data,
                                    1 if log2_tpm else 0, ord(sep), out.ctypes.data, out.size, end.ctypes.data, int(nthreads))
    if rc != 0:
        raise _lib.SctdaError(rc, 'sctda_format_table_f64')
    slot = cols * 25
    return [out[r * slot:r * slot + end[r]].tobytes() for r in range(rows)]


def write_preprocess_tables(name, genes, counts, long, batch, totaltran=None, which_samples=None, which_genes=None,
                            counts_subsampled=None, target_subsample=None, block_rows=4096):
    """Preprocess.save (ref:663-728): writes name.all.tsv and name.mapper.tsv (and name.no_subsampling.tsv
    when subsampled counts are given) from the state a Preprocess object holds:
      genes [G] names; counts [G, cells] transcript counts (self.data); long [cells] sampling time points;
      batch [cells] library ids; totaltran [cells] total transcripts per cell (default: column sums);
      which_samples [cells] bool mask of the cells to write; which_genes [G] bool mask of the genes of the
      Mapper table; counts_subsampled / target_subsample: self.data_subsampled / self.target_subsample.
    Every value is str(numpy.log2(1.0 + 1000000.0 * t / cv)) as Python 2 printed it, cv = total transcripts
    of the cell, or the subsampling target when subsampled (also for name.no_subsampling.tsv: ref:714-717).
    The numeric fields are formatted natively by all host cores, block by block of rows."""
    genes = [str(g) for g in genes]
    counts = numpy.asarray(counts, dtype=numpy.float64)
    G, ncell = counts.shape
    if len(genes) != G:
        raise ValueError('genes and counts disagree: %d names, %d rows' % (len(genes), G))
    subsampled = counts_subsampled is not None
    if subsampled and target_subsample is None:
        raise ValueError('target_subsample is needed with counts_subsampled')
    ws = numpy.ones(ncell, dtype=bool) if which_samples is None else numpy.asarray(which_samples, dtype=bool)
    wg = numpy.ones(G, dtype=bool) if which_genes is None else numpy.asarray(which_genes, dtype=bool)
    tot = counts.sum(axis=0) if totaltran is None else numpy.asarray(totaltran, dtype=numpy.float64)
    cells = numpy.nonzero(ws)[0]
    prefix = [('D' + py2_str(long[n]) + '_' + str(batch[n]) + '_' + str(n) + '\t' + py2_str(long[n]) + '\t' + str(batch[n])
               + '\t').encode() for n in cells]
    cv = numpy.full(cells.size, float(target_subsample), dtype=numpy.float64) if subsampled else tot[cells]

    def dump(path, header, data, gene_mask, with_prefix):
        cols = numpy.nonzero(gene_mask)[0]
        with open(path, 'wb') as f:
            f.write(header.encode() + b'\n')
            for lo in range(0, cells.size, block_rows):
                sel = cells[lo:lo + block_rows]
                blockv = numpy.ascontiguousarray(data[numpy.ix_(cols, sel)].T)
                rows = format_rows(blockv, denom=cv[lo:lo + block_rows], log2_tpm=True)
                if with_prefix:
                    f.write(b''.join(prefix[lo + i] + r + b'\n' for i, r in enumerate(rows)))
                else:
                    f.write(b''.join(r + b'\n' for r in rows))
    all_genes = numpy.ones(G, dtype=bool)
    head_all = 'ID\ttimepoint\tlib\t' + '\t'.join(genes)
    dump(name + '.all.tsv', head_all, counts_subsampled if subsampled else counts, all_genes, True)
    if subsampled:
        dump(name + '.no_subsampling.tsv', head_all, counts, all_genes, True)
    dump(name + '.mapper.tsv', '\t'.join(g for g, m in zip(genes, wg) if m),
         counts_subsampled if subsampled else counts, wg, False)


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
