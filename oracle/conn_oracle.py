"""CPU ORACLE for hot path (b): the per-gene connectivity permutation test.

TEST INFRASTRUCTURE ONLY.  Nothing under sctda_b200/ imports this module; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may.  It is a Python-3 numpy restatement of the reference's algorithm,
function by function, each citing the reference lines it follows
(/root/reference/scTDA/main.py, abbreviated "ref:").

PARITY STATUS: pinned.  The reference ships no tests or fixtures, and it is
Python 2, so it cannot be imported as is.  oracle/gen_golden.py executes the
reference's own source text for this path (UnrootedGraph / RootedGraph /
benjamini_hochberg) under a small, listed Py2->Py3 shim and stores its outputs
in tests/golden/; tests/test_oracle_golden.py checks this restatement against
those vectors (scores <= 1e-12 relative, p-values and integer columns exact).

Deviations from the reference that cannot be avoided (documented, tested):
  * numexpr is not installed: `numexpr.evaluate('sum(2**pk - 1, 1)')` (ref:982)
    and `log1p` (ref:983) are evaluated with numpy; <= 1 ulp before the float32
    rounding of ref:972,983.
  * dict / set iteration order is CPython-3 order (insertion order for dicts).
"""
import json
import math

import networkx
import numpy


# ----------------------------------------------------------------------------------------------
# state (ref:800-806, 828-829, 839-891)
# ----------------------------------------------------------------------------------------------
class GraphState(object):
    """The attributes of UnrootedGraph that the test reads:
    pl, adj, dic, samples, dicgenes, log2."""

    def __init__(self, pl, adj, dic, dicgenes, log2=True):
        self.pl = list(pl)                      # LCC node ids (strings), ref:805
        self.adj = numpy.asarray(adj, dtype=numpy.float64)   # dense V x V, ref:806
        self.dic = dic                          # str node id -> list of int cell rows, ref:828-829
        self.dicgenes = dicgenes                # column name -> sequence of float, ref:880-885
        self.log2 = log2                        # ref:807
        samples = []
        for i in self.pl:                       # ref:886-889
            samples += list(self.dic[i])
        self.samples = numpy.array(list(set(samples)))
        # filled by rooted()
        self.dicdis = None
        self.po = None


def parse_table(path, csv=False):
    """ref:835-885 with shift=None: every column becomes a key of dicgenes; non-numeric
    fields (cell ID, lib) become 0.0.  Returns (dicgenes: dict name -> float64 array in
    header order, cellID list, libs list, cdr array)."""
    cx = ',' if csv else '\t'
    dicgenes = {}
    cols = []
    cell_id, libs, cdr = [], [], []
    with open(path, 'r') as f:
        for n, line in enumerate(f):
            sp = line[:-1].split(cx)
            if csv:
                sp = [x.split('|')[0] for x in sp]
            if n == 0:
                cols = list(sp)
                coyt = cols.index('lib') if 'lib' in cols else -1      # ref:851-854
                ttt = cols.index('timepoint') if 'timepoint' in cols else -1   # ref:855-858
                for q in cols:
                    dicgenes[q] = []
            else:
                cdr.append(0.0)
                for n2, mji in enumerate(sp):
                    try:                                   # is_number, ref:118-126
                        v = float(mji)
                        ok = True
                    except ValueError:
                        ok = False
                    if ok:
                        if v > 0.0 and n2 != 0 and n2 != coyt and n2 != ttt:   # ref:865-866
                            cdr[-1] += 1.0
                    elif n2 == coyt:
                        libs.append(mji)
                        v = 0.0
                    elif n2 == 0:
                        cell_id.append(mji)
                        v = 0.0
                    else:
                        v = 0.0
                    dicgenes[cols[n2]].append(v)
    for q in dicgenes:
        dicgenes[q] = numpy.array(dicgenes[q], dtype=numpy.float64)
    cdr = numpy.array(cdr) / float(max(len(cdr), 1))        # ref:891
    return dicgenes, cell_id, libs, cdr


def load_state(name, table, log2=True, csv=False):
    """ref:800-806 (+828-829, 839-891) with the networkx-1.x calls replaced by their
    networkx-3 equivalents."""
    g = networkx.read_gexf(name + '.gexf')                                  # ref:801
    comps = [g.subgraph(c) for c in networkx.connected_components(g)]       # ref:802
    sizes = [len(c.nodes()) for c in comps]
    gl = comps[sizes.index(numpy.max(sizes))]                               # ref:803-804
    pl = list(gl.nodes())                                                   # ref:805
    adj = numpy.array(networkx.to_numpy_array(gl, nodelist=pl))             # ref:806
    with open(name + '.json', 'r') as h:
        dic = json.load(h)                                                  # ref:828-829
    dicgenes, _, _, _ = parse_table(table, csv=csv)
    st = GraphState(pl, adj, dic, dicgenes, log2=log2)
    st.gl = gl
    return st


# ----------------------------------------------------------------------------------------------
# node means / observed statistic (ref:893-964, 994-1051)
# ----------------------------------------------------------------------------------------------
def get_gene(st, genin, ignore_log=False):
    """ref:929-964 with con=True: p_k in `dic` key order restricted to LCC nodes,
    normalised by tol = sum of the un-normalised values.  Returns (dict node -> p, tol)."""
    if not isinstance(genin, list):
        genin = [genin]
    plset = set(st.pl)
    genecolor = {}
    lista = []
    for i in st.dic.keys():                               # ref:934-938
        if str(i) in plset:
            genecolor[str(i)] = 0.0
            lista.append(i)
    for mju in genin:
        if mju is None:                                   # ref:943-945
            for i in sorted(lista):
                genecolor[str(i)] = 0.0
        else:
            geys = st.dicgenes[mju]
            for i in sorted(lista):                       # ref:949-959
                pol = 0.0
                if st.log2 and not ignore_log:
                    for j in st.dic[i]:
                        pol += (numpy.power(2, float(geys[j])) - 1.0)
                    pol = numpy.log2(1.0 + (pol / float(len(st.dic[i]))))
                else:
                    for j in st.dic[i]:
                        pol += float(geys[j])
                    pol /= float(len(st.dic[i]))
                genecolor[str(i)] += pol
    tol = sum(genecolor.values())                         # ref:960
    if tol > 0.0:
        for ll in genecolor.keys():                       # ref:961-963
            genecolor[ll] = genecolor[ll] / tol
    return genecolor, tol


def connectivity(st, genis, ind=1):
    """ref:994-1005."""
    dicgen = get_gene(st, genis)[0]
    ex = numpy.array([dicgen[uu] for uu in st.pl])
    cor = float(len(st.pl)) / float(len(st.pl) - 1)
    return cor * numpy.dot(numpy.transpose(ex), numpy.dot(numpy.linalg.matrix_power(st.adj, ind), ex))


def delta(st, genis):
    """ref:1007-1028 with group=None."""
    dicgen, tot = get_gene(st, genis)
    mi = [dicgen[node] for node in st.pl]
    if numpy.mean(mi) > 0.0:
        return numpy.mean(mi) * tot, numpy.min(mi) * tot, numpy.max(mi) * tot
    return 0.0, 0.0, 0.0


def expr(st, genis):
    """ref:1050: cells of the WHOLE table with value > 0."""
    return int(numpy.count_nonzero(numpy.asarray(st.dicgenes[genis]) > 0.0))


# ----------------------------------------------------------------------------------------------
# permutation source (ref:975-976)
# ----------------------------------------------------------------------------------------------
def perm_indices(n, S, rng=None):
    """n x S int32 index arrays drawn exactly as ref:976 draws them: row r of the
    tiled gene matrix is shuffled in place by numpy.random.shuffle, rows in order, from
    the global legacy MT19937 stream.  Shuffling arange(S) consumes the identical draw
    sequence because Fisher-Yates never looks at the values, so
    shuffled_row == row[perm_indices[r]]."""
    if rng is None:
        rng = numpy.random      # the global legacy RandomState, as the reference uses
    out = numpy.empty((n, S), dtype=numpy.int32)
    base = numpy.arange(S, dtype=numpy.int32)
    for r in range(n):
        out[r] = base
        rng.shuffle(out[r])
    return out


# ----------------------------------------------------------------------------------------------
# the permutation test (ref:966-992)
# ----------------------------------------------------------------------------------------------
def permuted_scores(st, genin, n=500, perms=None, rng=None):
    """ref:970-989: the n permuted connectivity scores (float64).  `perms` (n x S int)
    overrides the RNG draw; otherwise rows are shuffled from `rng` / global numpy.random."""
    jk = len(st.pl)
    pm = numpy.zeros((jk, n), dtype='float32')                               # ref:972
    llm = list(numpy.arange(numpy.max(st.samples) + 1)[st.samples])          # ref:973
    koi = {k: u for u, k in enumerate(llm)}                                  # ref:974
    geys = numpy.tile(numpy.asarray(st.dicgenes[genin]), (n, 1))[:, st.samples]   # ref:975
    if perms is None:
        r = numpy.random if rng is None else rng
        for row in geys:                                                     # ref:976 (eager map in Py2)
            r.shuffle(row)
    else:
        geys = numpy.take_along_axis(geys, numpy.asarray(perms, dtype=numpy.int64), axis=1)
    tot = numpy.zeros(n)                                                     # ref:977
    for k, i in enumerate(st.pl):                                            # ref:978-987
        pk = geys[:, numpy.array([koi[x] for x in st.dic[i]], dtype=numpy.int64)]
        q = pk.shape[1]
        if st.log2:
            t1 = _row_sum(2.0 ** pk - 1.0) / q                               # ref:982 (numexpr -> numpy)
            pm[k, :] = numpy.log1p(t1) / 0.693147                            # ref:983, float32 store
            tot += pm[k, :]                                                  # ref:984
        else:
            pm[k, :] = _row_sum(pk) / q                                      # ref:986
            tot += pm[k, :]
    pm = pm / tot                                                            # ref:988 -> float64
    conn = (float(jk) / float(jk - 1)) * numpy.einsum('ij,ij->j', numpy.dot(st.adj, pm), pm)   # ref:989
    return conn


def _row_sum(a):
    """Sequential left-to-right row sums (numexpr's reduction order for `sum(x, 1)`;
    numpy.sum would use pairwise blocks: same value to <= 1 ulp)."""
    acc = numpy.zeros(a.shape[0], dtype=numpy.float64)
    for c in range(a.shape[1]):
        acc += a[:, c]
    return acc


def connectivity_pvalue(st, genin, n=500, perms=None, rng=None):
    """ref:966-992."""
    if genin is None:
        return 0.0                                                           # ref:991-992
    conn = permuted_scores(st, genin, n=n, perms=perms, rng=rng)
    return numpy.mean(conn > connectivity(st, genin))                        # ref:990


# ----------------------------------------------------------------------------------------------
# BH, rooted columns, table writer (ref:92-115, 1724-1743, 1053-1086, 1879-1914)
# ----------------------------------------------------------------------------------------------
def benjamini_hochberg(pvalues):
    """ref:92-115."""
    pvalues = numpy.array(pvalues)
    n = float(pvalues.shape[0])
    new_pvalues = numpy.empty(int(n))
    values = [(pvalue, i) for i, pvalue in enumerate(pvalues)]
    values.sort()
    values.reverse()
    new_values = []
    for i, vals in enumerate(values):
        rank = n - i
        pvalue, index = vals
        new_values.append((n / rank) * pvalue)
    for i in range(0, int(n) - 1):
        if new_values[i] < new_values[i + 1]:
            new_values[i + 1] = new_values[i]
    for i, vals in enumerate(values):
        pvalue, index = vals
        new_pvalues[index] = new_values[i]
    return list(new_pvalues)


def get_distroot(st, root):
    """ref:1454-1461: BFS hop distance from `root` to every LCC node."""
    d = networkx.single_source_shortest_path_length(st.gl, str(root))
    return {str(i): d[i] for i in sorted(st.pl)}


def rooted(st, root, rootlane='timepoint'):
    """ref:1562-1567: dicdis and the pseudo-time regression po."""
    import scipy.stats
    st.dicdis = get_distroot(st, root)
    pel2, tol = get_gene(st, rootlane, ignore_log=True)
    pel = numpy.array([pel2[m] for m in st.pl]) * tol
    dr = numpy.array([st.dicdis[m] for m in st.pl])
    st.po = scipy.stats.linregress(pel, dr)
    return st


def dendrite(st, rootlane='timepoint'):
    """ref:1463-1479: for every LCC node i, -spearman(timepoint(node) - min, dist_i(node))
    over the nodes that are not at the maximal distance from i."""
    import scipy.stats
    out = {}
    daycolor = get_gene(st, rootlane, ignore_log=True)[0]
    dmin = min(daycolor.values())
    for i in st.pl:
        distroot = get_distroot(st, i)
        mx = max(distroot.values())
        x, y = [], []
        for q in distroot.keys():
            if distroot[q] != mx:
                x.append(daycolor[q] - dmin)
                y.append(distroot[q])
        out[str(i)] = -scipy.stats.spearmanr(x, y)[0]
    return out


def find_root(dendritem):
    """ref:1481-1497."""
    q, q2 = 1000.0, -1000.0
    ind, ind2 = 0, 0
    for n in dendritem.keys():
        if -2.0 < dendritem[n] < q:
            q = dendritem[n]
            ind = n
        if dendritem[n] > q2 and dendritem[n] > -2.0:
            q2 = dendritem[n]
            ind2 = n
    return ind, ind2


def centroid(st, genin, ignore_log=False):
    """ref:1724-1743."""
    dicge = get_gene(st, genin, ignore_log=ignore_log)[0]
    pel1 = pel2 = pel3 = 0.0
    for node in st.pl:
        pel1 += st.dicdis[node] * dicge[node]
        pel2 += dicge[node]
    if pel2 > 0.0:
        cen = float(pel1) / float(pel2)
        for node in st.pl:
            pel3 += numpy.power(st.dicdis[node] - cen, 2) * dicge[node]
        return [(cen - st.po[1]) / st.po[0], (numpy.sqrt(pel3 / float(pel2)) - st.po[1]) / st.po[0]]
    return [None, None]


# ----------------------------------------------------------------------------------------------
# gene x gene matrices (ref:1111-1190)
# ----------------------------------------------------------------------------------------------
def jsd_matrix(st, lista):
    """ref:1117-1155 (the un-chunked branch; the chunked one computes the same numbers)."""
    ge = numpy.array([[get_gene(st, g)[0][k] for k in st.pl] for g in lista])
    ge2 = numpy.copy(ge)
    ge2[ge2 == 0] = 1
    plogp = numpy.sum(ge2 * numpy.log2(ge2), axis=1)
    plogptile = numpy.tile(plogp, (len(lista), 1))
    ge_tile = numpy.tile(ge, (len(lista), 1, 1))
    ge_tile_T = numpy.transpose(ge_tile, [1, 0, 2])
    pq = 0.5 * (ge_tile + ge_tile_T)
    pq[pq == 0] = 1
    pqlogpq = numpy.sum(pq * numpy.log2(pq), axis=2)
    jsdiv = 0.5 * (plogptile + plogptile.T - 2 * pqlogpq)
    with numpy.errstate(invalid='ignore'):
        return numpy.sqrt(jsdiv), jsdiv


def cor_matrix(st, lista):
    """ref:1157-1163."""
    import sklearn.metrics.pairwise
    geys = numpy.array([st.dicgenes[m] for m in lista])
    return sklearn.metrics.pairwise.pairwise_distances(geys, metric='correlation')


def adjacency_matrix(st, lista, ind=1):
    """ref:1165-1190."""
    cor = float(len(st.pl)) / float(len(st.pl) - 1)
    pol = {g: get_gene(st, g)[0] for g in lista}
    Ai = numpy.linalg.matrix_power(st.adj, ind)
    mat = []
    for m1 in lista:
        ex1 = numpy.array([pol[m1][u] for u in st.pl])
        mat.append([cor * numpy.dot(numpy.transpose(ex1), numpy.dot(Ai, numpy.array([pol[m2][u] for u in st.pl]))) for m2 in lista])
    return numpy.array(mat)


def py2_str(v):
    """Python-2 `str()` of the values ref:1078-1079 writes: '%.12g' for floats (with a
    trailing '.0' when that prints an integer), plain str for ints / None."""
    if v is None:
        return 'None'
    if isinstance(v, (bool, numpy.bool_)):
        return str(bool(v))
    if isinstance(v, (int, numpy.integer)):
        return str(int(v))
    v = float(v)
    if math.isnan(v):
        return 'nan'
    if math.isinf(v):
        return 'inf' if v > 0 else '-inf'
    s = '%.12g' % v
    if '.' not in s and 'e' not in s:
        s += '.0'
    return s


def save_rows(st, n=500, filtercells=0, filterexp=0.0, annotation=None, rooted_cols=False,
              perms_for=None, rng=None):
    """ref:1062-1086 (rooted_cols: ref:1888-1914).  Returns (header, rows) where each row
    is the list of python values before string formatting.  `perms_for(gene) -> n x S`
    optionally supplies the permutations; otherwise the RNG stream is consumed gene by
    gene in sorted order exactly as the reference does."""
    annotation = annotation or {}
    header = ['Gene', 'Cells', 'Mean', 'Min', 'Max', 'Connectivity', 'p_value', 'q-value (BH)']
    if rooted_cols:
        header += ['Centroid', 'Dispersion']
    header += sorted(annotation.keys())
    pol = []
    lp = sorted(st.dicgenes.keys())                                          # ref:1068
    for gi in lp:                                                            # ref:1069-1071
        if expr(st, gi) > filtercells and delta(st, gi)[2] > filterexp:
            perms = perms_for(gi) if perms_for is not None else None
            pol.append(connectivity_pvalue(st, gi, n=n, perms=perms, rng=rng))
    por = benjamini_hochberg(pol) if pol else []                             # ref:1072
    rows = []
    mj = 0
    for gi in lp:                                                            # ref:1074-1086
        po = expr(st, gi)
        m1, m2, m3 = delta(st, gi)
        if rooted_cols:
            p1, p2 = centroid(st, gi)
        if po > filtercells and m3 > filterexp:
            row = [gi, po, m1, m2, m3, connectivity(st, gi), pol[mj], por[mj]]
            if rooted_cols:
                row += [p1, p2]
            for m in sorted(annotation.keys()):
                row.append('Y' if gi in annotation[m] else 'N')
            rows.append(row)
            mj += 1
    return header, rows


def write_genes_tsv(path, header, rows):
    """ref:1063-1067,1078-1085: tab-separated, '\\n' line ends, Python-2 str() floats."""
    with open(path, 'w') as f:
        f.write('\t'.join(header) + '\n')
        for row in rows:
            f.write('\t'.join(x if isinstance(x, str) else py2_str(x) for x in row) + '\n')
