"""CPU ORACLE for hot path (a): the Mapper graph build behind TopologicalRepresentation.save().

TEST INFRASTRUCTURE ONLY (see oracle/conn_oracle.py for the rule).

PARITY STATUS: UNPINNED against the reference.  The Mapper algorithm is not in the reference
tree: scTDA calls the third-party package `sakmapper` (PyPI `sakmapper`, GitHub szairis/sakmapper,
version un-pinned in /root/reference/setup.py:12), at /root/reference/scTDA/main.py:745-753
(apply_lens) and :768-771 (mapper_graph).  The package is not installed and cannot be fetched
here, and the reference ships no tests or fixtures for it.  This file therefore restates the
algorithm sakmapper publishes (grid cover of a 2-D lens widened by `gain`, per-patch clustering
with model selection up to max_K, nerve) with every choice the call sites leave open fixed
explicitly, and it is anchored on what CAN be checked here:
  * the call-site contract (ref:768-777): return triple, node ids 0..V-1 enumerate all_clusters,
    members castable to int, gexf/json round trip  -> tests/test_mapper_oracle.py;
  * the cover against a brute-force evaluation of the published rule (percentile fenceposts,
    intervals widened by gain/2 of their width on each side, strict inequalities);
  * the single-linkage partitions against scipy.cluster.hierarchy (linkage 'single' + fcluster);
  * the Davies-Bouldin index against sklearn.metrics.davies_bouldin_score;
  * the nerve against brute-force set intersection.

Pinned choices (all are parameters or documented constants):
  cover      r x r grid; fenceposts numpy.percentile (equalize) or linspace(min,max); interval k
             = (post_k - gain*w_k/2, post_{k+1} + gain*w_k/2), w_k = post_{k+1}-post_k; strict <, >.
  distances  squared Euclidean between rows (metric='euclidean') or between centred, unit-norm
             rows (metric='correlation': 1 - r = d^2/2, same ordering), from the Gram matrix:
             d2(p,q) = max(0, (G_pp + G_qq) - 2 G_pq), G_pq = fma chain over features k = 0..G-1.
  clustering single linkage = Prim MST from local vertex 0 (ties: smaller index), cut = remove the
             K-1 heaviest edges (ties: later-inserted first); K_max = 2 if m <= 5 else
             min(m // 2, max_K); K = argmin Davies-Bouldin over 2..K_max (ties: smaller K);
             1 cell -> one singleton node.  statistics='histgap': first-empty-histogram-bin cut.
  DB index   sklearn's definition evaluated from the Gram matrix (no feature-space pass), sums in
             ascending index order through the column sums of the finest partition.
  node order bins row-major (i*r + j), clusters of a bin by smallest member.
  nerve      edge (i<j) iff the member sets intersect; edges sorted lexicographically; weight 1.0.
"""
import ctypes
import os
import subprocess

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
_REFLIB = os.path.join(HERE, '_ref', 'libmapper_oracle.so')
_C_SRC = os.path.join(HERE, 'mapper_oracle.c')
_clib = None


def build_c():
    """Compiles oracle/mapper_oracle.c (the fma-chain Gram matrix) into oracle/_ref/."""
    os.makedirs(os.path.dirname(_REFLIB), exist_ok=True)
    if (not os.path.exists(_REFLIB)) or os.path.getmtime(_REFLIB) < os.path.getmtime(_C_SRC):
        subprocess.check_call(['gcc', '-O2', '-fPIC', '-shared', '-ffp-contract=off', '-fopenmp', '-o', _REFLIB,
                               _C_SRC, '-lm'])
    return _REFLIB


def _lib():
    global _clib
    if _clib is None:
        _clib = ctypes.CDLL(build_c())
        _clib.gram_fma.restype = None
        _clib.gram_fma.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                   ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
    return _clib


# ----------------------------------------------------------------------------------------------
# distances
# ----------------------------------------------------------------------------------------------
def _tree32(parts):
    """The reduction order of the device kernels: 32 lane partials, then the xor butterfly
    16, 8, 4, 2, 1 (lane 0's value).  parts: [..., 32]."""
    a = numpy.array(parts, dtype=numpy.float64, copy=True)
    idx = numpy.arange(32)
    for o in (16, 8, 4, 2, 1):
        a = a + a[..., idx ^ o]
    return a[..., 0]


def _lane_sums(M):
    """Row sums in the device order: lane l adds elements l, l+32, ... sequentially."""
    n, g = M.shape
    acc = numpy.zeros((n, 32), dtype=numpy.float64)
    for k0 in range(0, g, 32):
        blk = M[:, k0:k0 + 32]
        acc[:, :blk.shape[1]] += blk
    return _tree32(acc)


def row_normalize(X, metric):
    """metric='euclidean': rows unchanged; 'correlation': centred rows scaled to unit norm
    (constant rows -> zeros).  Returns (Xn, sq) with sq_i = |Xn_i|^2 in the device order."""
    X = numpy.ascontiguousarray(X, dtype=numpy.float64)
    if metric == 'correlation':
        mean = _lane_sums(X) / float(X.shape[1])
        xc = X - mean[:, None]
        nrm = numpy.sqrt(_lane_sums(xc * xc))
        with numpy.errstate(invalid='ignore', divide='ignore'):
            Xn = numpy.where(nrm[:, None] > 0.0, xc / nrm[:, None], 0.0)
    elif metric == 'euclidean':
        Xn = X
    else:
        raise ValueError('metric must be euclidean or correlation')
    Xn = numpy.ascontiguousarray(Xn)
    return Xn, _lane_sums(Xn * Xn)


def gram(Xn, rows_a, rows_b):
    """G[i, j] = sum_k Xn[rows_a[i], k] * Xn[rows_b[j], k] as ONE fma chain over k ascending
    (what a float64 tensor-core MMA chain with a sequential K loop computes)."""
    Xn = numpy.ascontiguousarray(Xn, dtype=numpy.float64)
    ra = numpy.ascontiguousarray(rows_a, dtype=numpy.int64)
    rb = numpy.ascontiguousarray(rows_b, dtype=numpy.int64)
    out = numpy.empty((ra.size, rb.size), dtype=numpy.float64)
    _lib().gram_fma(Xn.ctypes.data, Xn.shape[0], Xn.shape[1], ra.ctypes.data, ra.size, rb.ctypes.data, rb.size,
                    out.ctypes.data)
    return out


def pairdist_rows(Xn, sq, r0, r1, metric):
    """Rows [r0, r1) of the N x N distance matrix in `metric` (the MDS-lens input)."""
    n = Xn.shape[0]
    G = gram(Xn, numpy.arange(r0, r1), numpy.arange(n))
    if metric == 'correlation':
        D = 1.0 - G
    else:
        D = numpy.sqrt(numpy.maximum((sq[r0:r1, None] + sq[None, :]) - 2.0 * G, 0.0))
    D[numpy.arange(r1 - r0), numpy.arange(r0, r1)] = 0.0
    return D


# ----------------------------------------------------------------------------------------------
# cover
# ----------------------------------------------------------------------------------------------
def fenceposts(values, resolution, equalize=True):
    v = numpy.asarray(values, dtype=numpy.float64)
    if equalize:
        return numpy.array([numpy.percentile(v, k * 100.0 / resolution) for k in range(resolution + 1)])
    return numpy.linspace(v.min(), v.max(), resolution + 1)


def intervals(posts, gain):
    w = posts[1:] - posts[:-1]
    return posts[:-1] - (0.5 * gain) * w, posts[1:] + (0.5 * gain) * w


def cover(lens, resolution, gain, equalize=True):
    """List of r*r ascending cell-index arrays, bin b = i*r + j (axis-0 interval i, axis-1 j)."""
    lens = numpy.asarray(lens, dtype=numpy.float64)
    r = int(resolution)
    lo0, hi0 = intervals(fenceposts(lens[:, 0], r, equalize), gain)
    lo1, hi1 = intervals(fenceposts(lens[:, 1], r, equalize), gain)
    in0 = (lens[:, 0][None, :] > lo0[:, None]) & (lens[:, 0][None, :] < hi0[:, None])
    in1 = (lens[:, 1][None, :] > lo1[:, None]) & (lens[:, 1][None, :] < hi1[:, None])
    bins = []
    for i in range(r):
        for j in range(r):
            bins.append(numpy.nonzero(in0[i] & in1[j])[0])
    return bins, (lo0, hi0, lo1, hi1)


# ----------------------------------------------------------------------------------------------
# per-bin clustering
# ----------------------------------------------------------------------------------------------
def prim_mst(Gm):
    """Prim from vertex 0 on d2(p,q) = max(0, (G_pp+G_qq) - 2 G_pq).  Returns (parent, weight,
    order): vertex u was attached to `parent[u]` by an edge of weight `weight[u]` as the
    `order[u]`-th edge (vertex 0: parent -1)."""
    m = Gm.shape[0]
    gpp = numpy.diag(Gm).copy()
    intree = numpy.zeros(m, dtype=bool)
    parent = numpy.full(m, -1, dtype=numpy.int64)
    weight = numpy.zeros(m)
    order = numpy.full(m, -1, dtype=numpy.int64)
    key = numpy.full(m, numpy.inf)
    intree[0] = True
    u = 0
    for t in range(m - 1):
        d = numpy.maximum((gpp[u] + gpp) - 2.0 * Gm[u], 0.0)
        upd = (~intree) & (d < key)
        key[upd] = d[upd]
        parent[upd] = u
        cand = numpy.where(intree, numpy.inf, key)
        u = int(numpy.argmin(cand))                 # first minimum = smallest index among ties
        intree[u] = True
        weight[u] = key[u]
        order[u] = t
    return parent, weight, order


def _components(m, parent, removed):
    """Labels of the forest obtained by deleting the tree edges of the vertices in `removed`;
    clusters numbered by their smallest vertex."""
    root = numpy.arange(m)

    def find(a):
        while root[a] != a:
            root[a] = root[root[a]]
            a = root[a]
        return a
    for u in range(1, m):
        if u not in removed:
            a, b = find(u), find(int(parent[u]))
            if a != b:
                root[max(a, b)] = min(a, b)
    reps = numpy.array([find(a) for a in range(m)])
    _, labels = numpy.unique(reps, return_inverse=True)     # representatives are the minima: ascending
    return labels


def heaviest_edges(weight, order, count):
    """Vertices whose tree edge is among the `count` heaviest (ties: later-inserted first)."""
    us = [u for u in range(len(weight)) if order[u] >= 0]
    us.sort(key=lambda u: (-weight[u], -order[u]))
    return us[:count]


def db_index_from_gram(Gm, fine_labels, coarse_of_fine):
    """Davies-Bouldin index (sklearn.metrics.davies_bouldin_score's definition) of the partition
    coarse_of_fine[fine_labels[p]], from the Gram matrix only:
      T[p, f]  = sum_{q in fine cluster f} G[q, p]          (q ascending)
      x_p.c_I  = (sum_{f in I} T[p, f]) / n_I               (f ascending)
      c_I.c_J  = (sum_{p in I, ascending} sum_{f in J} T[p, f]) / (n_I n_J)
      |x_p - c_I|^2 = max(0, (G_pp - 2 x_p.c_I) + c_I.c_I),  S_I = mean of the square roots
      M_IJ = sqrt(max(0, (c_I.c_I - 2 c_I.c_J) + c_J.c_J)),  R_IJ = (S_I + S_J) / M_IJ (0 if M_IJ = 0)
      DB = (sum_I max_{J != I} R_IJ) / K;  0 if every S_I is 0."""
    m = Gm.shape[0]
    F = int(fine_labels.max()) + 1
    T = numpy.zeros((m, F))
    for q in range(m):
        T[:, fine_labels[q]] += Gm[q]
    coarse_of_fine = numpy.asarray(coarse_of_fine)
    K = int(coarse_of_fine.max()) + 1
    lab = coarse_of_fine[fine_labels]
    TK = numpy.zeros((m, K))
    for f in range(F):
        TK[:, coarse_of_fine[f]] += T[:, f]
    n = numpy.bincount(lab, minlength=K).astype(numpy.float64)
    cc = numpy.zeros((K, K))
    for p in range(m):
        cc[lab[p]] += TK[p]
    cc = cc / (n[:, None] * n[None, :])
    xc = TK[numpy.arange(m), lab] / n[lab]
    d2 = numpy.maximum((numpy.diag(Gm) - 2.0 * xc) + cc[lab, lab], 0.0)
    dist = numpy.sqrt(d2)
    S = numpy.zeros(K)
    for p in range(m):
        S[lab[p]] += dist[p]
    S = S / n
    if not (S > 0.0).any():
        return 0.0
    tot = 0.0
    for i in range(K):
        best = 0.0
        for j in range(K):
            if j == i:
                continue
            M2 = max((cc[i, i] - 2.0 * cc[i, j]) + cc[j, j], 0.0)
            Mij = numpy.sqrt(M2)
            R = (S[i] + S[j]) / Mij if Mij > 0.0 else 0.0
            if R > best:
                best = R
        tot += best
    return tot / float(K)


def linkage_tree(Gm, linkage):
    """Average / Ward / complete linkage of one bin (scipy.cluster.hierarchy.linkage on the Euclidean distances
    read off the Gram matrix) in the form prim_mst returns for single linkage: the cluster that disappears in a
    merge is a vertex hanging under the surviving one (the cluster containing the smaller smallest member),
    with the squared merge height as edge weight and the merge index as insertion order."""
    import scipy.cluster.hierarchy
    import scipy.spatial.distance
    m = Gm.shape[0]
    g = numpy.diag(Gm)
    d2 = numpy.maximum((g[:, None] + g[None, :]) - 2.0 * Gm, 0.0)
    d = numpy.sqrt(d2)
    numpy.fill_diagonal(d, 0.0)
    Z = scipy.cluster.hierarchy.linkage(scipy.spatial.distance.squareform(0.5 * (d + d.T), checks=False), method=linkage)
    rep = list(range(m)) + [0] * (m - 1)          # smallest member of every (original or merged) cluster
    parent = numpy.full(m, -1, dtype=numpy.int64)
    weight = numpy.zeros(m)
    order = numpy.full(m, -1, dtype=numpy.int64)
    for t in range(m - 1):
        a, b = int(Z[t, 0]), int(Z[t, 1])
        ra, rb = rep[a], rep[b]
        keep, gone = (ra, rb) if ra < rb else (rb, ra)
        parent[gone] = keep
        weight[gone] = Z[t, 2] ** 2
        order[gone] = t
        rep[m + t] = keep
    return parent, weight, order


def cluster_bin(Gm, max_K=5, stat='db', hist_bins=10, linkage='single'):
    """Labels (0..K-1, clusters numbered by smallest member) of one bin from its Gram matrix."""
    m = Gm.shape[0]
    if m == 1:
        return numpy.zeros(1, dtype=numpy.int64), 1, []
    parent, weight, order = prim_mst(Gm) if linkage == 'single' else linkage_tree(Gm, linkage)
    if stat == 'histgap':
        h = numpy.sqrt(weight[1:])
        lo, hi = h.min(), h.max()
        if not hi > lo:
            return numpy.zeros(m, dtype=numpy.int64), 1, []
        width = (hi - lo) / float(hist_bins)
        idx = numpy.minimum(((h - lo) / width).astype(numpy.int64), hist_bins - 1)
        counts = numpy.bincount(idx, minlength=hist_bins)
        empty = numpy.nonzero(counts == 0)[0]
        if empty.size == 0:
            return numpy.zeros(m, dtype=numpy.int64), 1, []
        removed = set(int(u) + 1 for u in numpy.nonzero(idx > empty[0])[0])
        labels = _components(m, parent, removed)
        return labels, int(labels.max()) + 1, []
    if stat != 'db':
        raise ValueError("stat must be 'db' or 'histgap'")
    kmax = 2 if m <= 5 else min(m // 2, int(max_K))
    heavy = heaviest_edges(weight, order, kmax - 1)
    fine = _components(m, parent, set(heavy))
    scores = []
    best = None
    for K in range(2, kmax + 1):
        labK = _components(m, parent, set(heavy[:K - 1]))
        # map fine clusters to the K-partition (each fine cluster lies inside one coarse cluster)
        cof = numpy.zeros(int(fine.max()) + 1, dtype=numpy.int64)
        cof[fine] = labK
        s = db_index_from_gram(Gm, fine, cof)
        scores.append(s)
        if best is None or s < best[0]:
            best = (s, K, labK)
    return best[2], best[1], scores


# ----------------------------------------------------------------------------------------------
# nerve and the whole graph
# ----------------------------------------------------------------------------------------------
def nerve(clusters):
    sets = [set(int(x) for x in c) for c in clusters]
    edges = []
    for i in range(len(sets)):
        for j in range(i + 1, len(sets)):
            if sets[i] & sets[j]:
                edges.append((i, j))
    return edges


def mapper_graph(X, lens, resolution, gain, equalize=True, clust='agglomerative', stat='db', max_K=5,
                 metric='euclidean', verbose=False, linkage='single'):
    """The sakmapper.mapper_graph contract (ref:768-771): returns (edges, all_clusters, patches)
    with all_clusters[k] = ascending cell rows of node k and patches[b] = cell rows of bin b."""
    if clust != 'agglomerative':
        raise ValueError("only clust='agglomerative' (single linkage) is restated")
    Xn, _ = row_normalize(X, metric)
    bins, _ = cover(lens, resolution, gain, equalize)
    all_clusters = []
    node_bin = []
    patches = {}
    n_patch = 0
    for b, cells in enumerate(bins):
        patches[b] = cells
        if cells.size == 0:
            continue
        n_patch += 1
        Gm = gram(Xn, cells, cells)
        labels, K, _ = cluster_bin(Gm, max_K=max_K, stat=stat, linkage=linkage)
        for c in range(K):
            all_clusters.append(cells[labels == c])
            node_bin.append(b)
    if verbose:
        print('total of %d patches required clustering' % n_patch)
        print('this implies %d nodes in the mapper graph' % len(all_clusters))
    return nerve(all_clusters), all_clusters, patches
