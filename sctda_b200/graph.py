"""Host-side mirror of scTDA.UnrootedGraph / scTDA.RootedGraph for the hot path (b):
same constructor arguments, method names, return values and file formats as the reference
(/root/reference/scTDA/main.py, "ref:"), with the per-gene work done in batches by the CUDA
kernels behind include/sctda_b200.h.  No CPU fallback: every statistic comes from the GPU.

Mirrored: __init__ (ref:785-891), get_gene (ref:893-964), connectivity_pvalue (ref:966-992),
connectivity (ref:994-1005), delta (ref:1007-1028), expr (ref:1030-1051), save (ref:1053-1086),
benjamini_hochberg (ref:92-115); RootedGraph.__init__ root/pseudo-time state (ref:1522-1567),
centroid (ref:1724-1743), save (ref:1879-1914).
count_gene (ref:1347-1383) and the lib_ / ID_ / _CDR colourings of get_gene (ref:902-928).
Not mirrored (SURVEY.md §2 "out of scope"): layouts, pickles, drawing, plotting.
"""
import json

import networkx
import numpy
import scipy.sparse
import scipy.sparse.csgraph
import scipy.stats
import torch

from . import _lib
from . import formats
from . import geneorder


def benjamini_hochberg(pvalues):
    """BH step-up q-values with the reference's tie order (ref:92-115): p sorted descending,
    equal p by descending index; q = running minimum of (n/rank) * p."""
    p = numpy.asarray(pvalues, dtype=numpy.float64)
    n = p.shape[0]
    if n == 0:
        return []
    order = numpy.lexsort((numpy.arange(n), p))[::-1]
    rank = float(n) - numpy.arange(n, dtype=numpy.float64)
    vals = (float(n) / rank) * p[order]
    vals = numpy.minimum.accumulate(vals)
    out = numpy.empty(n, dtype=numpy.float64)
    out[order] = vals
    return list(out)


def hierarchical_clustering(mat, method='average', cluster_distance=True, labels=None, thres=0.65):
    """ref:204-255 without the figure: linkage (`method`) of the rows of `mat` under the euclidean
    distance (cluster_distance=True) or of `mat` itself taken as a distance matrix; returns the
    dendrogram record scipy colours at `thres` times the largest merge height (what find_clusters
    reads).  `labels` only annotated the reference's figure.  scipy's linkage / dendrogram are the
    reference's own calls; cellular_subpopulations computes the row distances on the device."""
    return _hierarchical_clustering(mat, method, cluster_distance, thres, None)


def _hierarchical_clustering(mat, method, cluster_distance, thres, ctx):
    import scipy.cluster.hierarchy as sch
    import scipy.spatial.distance
    D = numpy.array(mat, dtype=numpy.float64)
    if not cluster_distance:
        tri = scipy.spatial.distance.squareform(D)                              # ref:215
    elif ctx is not None and D.shape[0] > 1:
        from . import mapper
        dev = torch.device('cuda', ctx.device)
        full = mapper.pairwise_distances(ctx, mapper._padded_device_table(torch.from_numpy(D).to(dev), dev),
                                         'euclidean').cpu().numpy()
        tri = full[numpy.triu_indices(D.shape[0], k=1)]                         # pdist order, ref:217
    else:
        tri = scipy.spatial.distance.pdist(D, metric='euclidean')
    Y = sch.linkage(tri, method=method)                                         # ref:220
    return sch.dendrogram(Y, orientation='right', color_threshold=thres * max(Y[:, 2]), no_plot=True)


def find_clusters(z):
    """ref:258-266: leaves of the dendrogram record `z` grouped by the colour of the link they hang from
    (all leaves that join above the colour threshold share the default colour, hence one group)."""
    clus = {}
    for colour, xs, ys in zip(z['color_list'], z['icoord'], z['dcoord']):
        for x, y in zip(xs, ys):
            if y == 0.0:                                                         # a leaf end of the link
                clus.setdefault(colour, []).append(z['leaves'][int((x - 5.0) / 10.0)])
    return clus


def reference_perm_indices(n, S, rng=None):
    """The permutation source of ref:975-976 as index arrays: row r of an n x S block is
    arange(S) shuffled by numpy.random.shuffle, rows in order, from the GLOBAL legacy
    MT19937 stream (or the RandomState `rng`).  Fisher-Yates never looks at the values, so the
    draws are those the reference consumes and shuffled_row == row[perm[r]].
    The draws are made by sctda_mt19937_shuffle_rows (the same generator and rejection rule in
    C) on the state taken from numpy, which is then put back advanced."""
    import ctypes
    st = (numpy.random if rng is None else rng).get_state()
    if st[0] != 'MT19937':
        raise ValueError('the reference stream is the legacy MT19937 generator')
    key = numpy.ascontiguousarray(st[1], dtype=numpy.uint32).copy()
    pos = ctypes.c_int32(int(st[2]))
    out = numpy.empty((n, S), dtype=numpy.int32)
    lib = _lib.load()
    rc = lib.sctda_mt19937_shuffle_rows(key.ctypes.data, ctypes.byref(pos), int(n), int(S), out.ctypes.data)
    if rc != 0:
        raise _lib.SctdaError(rc, 'sctda_mt19937_shuffle_rows')
    (numpy.random if rng is None else rng).set_state((st[0], key, int(pos.value), st[3], st[4]))   # shuffle leaves a cached normal deviate alone
    return out


class DeviceGraph(object):
    """Device image of (pl, adj, dic, samples): struct sctda_graph (include/sctda_b200.h)."""

    def __init__(self, node_members, adj_csr, samples, n_cells, device):
        """node_members: list of int sequences (table rows), in node order;
        adj_csr: scipy CSR V x V float64 in the same node order; samples: int array."""
        dev = torch.device('cuda', device)
        samples = numpy.asarray(samples, dtype=numpy.int64)
        V = len(node_members)
        if samples.size and (int(samples.min()) < 0 or int(samples.max()) >= int(n_cells)):
            raise ValueError('samples must be table rows in [0, %d): got [%d, %d] (name.json and the table disagree; '
                             'the reference raises IndexError here)' % (n_cells, samples.min(), samples.max()))
        pos = numpy.full(int(samples.max()) + 1 if samples.size else 1, -1, dtype=numpy.int64)
        pos[samples] = numpy.arange(samples.size)                     # koi of ref:973-974
        sizes = numpy.array([len(m) for m in node_members], dtype=numpy.int64)
        off = numpy.zeros(V + 1, dtype=numpy.int64)
        numpy.cumsum(sizes, out=off[1:])
        flat = numpy.concatenate([numpy.asarray(m, dtype=numpy.int64) for m in node_members]) if V else numpy.zeros(0, numpy.int64)
        if flat.size and (int(flat.min()) < 0 or int(flat.max()) >= pos.size):
            raise ValueError('node member lists must be non-empty and contained in samples')
        cols = pos[flat]
        if (cols < 0).any() or (sizes == 0).any():
            raise ValueError('node member lists must be non-empty and contained in samples')
        if off[-1] >= 2 ** 31 or samples.size >= 2 ** 31:
            raise ValueError('graph too large for int32 indices')
        adj_csr = scipy.sparse.csr_matrix(adj_csr, dtype=numpy.float64)
        # the library only evaluates the quadratic form p'Ap (include/sctda_b200.h): hand it the
        # upper-triangular matrix with the same form, half the non-zeros of the symmetric adjacency
        adj_csr = scipy.sparse.csr_matrix(scipy.sparse.triu(adj_csr + adj_csr.T, k=1)
                                          + scipy.sparse.diags(adj_csr.diagonal()))
        adj_csr.sort_indices()
        self.V, self.S, self.Ncells = V, int(samples.size), int(n_cells)
        self.Mtot = int(off[-1])
        self.nnz = int(adj_csr.nnz)
        self.node_off = torch.from_numpy(off.astype(numpy.int32)).to(dev)
        self.node_cols = torch.from_numpy(cols.astype(numpy.int32)).to(dev)
        self.adj_off = torch.from_numpy(adj_csr.indptr.astype(numpy.int32)).to(dev)
        self.adj_idx = torch.from_numpy(adj_csr.indices.astype(numpy.int32)).to(dev)
        self.adj_w = torch.from_numpy(adj_csr.data.astype(numpy.float64)).to(dev)
        self.samples = torch.from_numpy(samples.astype(numpy.int32)).to(dev)
        self.device = dev
        self.struct = _lib.Graph(V=self.V, S=self.S, Ncells=self.Ncells, Mtot=self.Mtot,
                                 d_node_off=self.node_off.data_ptr(), d_node_cols=self.node_cols.data_ptr(),
                                 d_adj_off=self.adj_off.data_ptr(), d_adj_idx=self.adj_idx.data_ptr(),
                                 d_adj_w=self.adj_w.data_ptr(), d_samples=self.samples.data_ptr())


def node_stats(ctx, dg, Y, log2=True, dist=None, want_p=False):
    """sctda_node_stats on a device table Y [G, >=Ncells] (float64, gene-major).
    Returns a dict of device tensors."""
    import ctypes
    G = int(Y.shape[0])
    dev = dg.device
    out = {'expr': torch.empty(G, dtype=torch.int32, device=dev)}
    for k in ('tol', 'mean', 'min', 'max', 'conn'):
        out[k] = torch.empty(G, dtype=torch.float64, device=dev)
    if dist is not None:
        out['cen'] = torch.empty(G, dtype=torch.float64, device=dev)
        out['disp'] = torch.empty(G, dtype=torch.float64, device=dev)
    if want_p:
        out['p'] = torch.empty((G, dg.V), dtype=torch.float64, device=dev)
    p = _lib.ptr
    rc = ctx.lib.sctda_node_stats(ctx.handle, ctypes.byref(dg.struct), G, p(Y), int(Y.stride(0)) if G else dg.Ncells,
                                  1 if log2 else 0, p(dist), p(out['expr']), p(out['tol']), p(out['mean']),
                                  p(out['min']), p(out['max']), p(out['conn']), p(out.get('cen')),
                                  p(out.get('disp')), p(out.get('p')),
                                  ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    ctx.check(rc)
    return out


def conn_perm_test(ctx, dg, Y, perms, observed, log2=True, per_gene=False, want_scores=False, tie_rel=1e-11,
                   gene_order=None):
    """sctda_conn_perm_test.  Y [G, >=Ncells] float64 device; perms int32 device, [n, S]
    (shared) or [G, n, S] (per_gene); observed [G] float64 device; gene_order: int32 device
    permutation of the genes (tile layout of the shared-permutation kernel, results unchanged).
    Returns (exceed int32 [G], ties int32 [G], scores float64 [G, n] or None) on the device."""
    import ctypes
    G = int(Y.shape[0])
    dev = dg.device
    if per_gene:
        assert perms.dim() == 3 and perms.shape[0] == G and perms.shape[2] == dg.S
        n = int(perms.shape[1])
        stride = n * dg.S
    else:
        assert perms.dim() == 2 and perms.shape[1] == dg.S
        n = int(perms.shape[0])
        stride = 0
    assert perms.dtype == torch.int32 and perms.is_contiguous()
    exceed = torch.empty(G, dtype=torch.int32, device=dev)
    ties = torch.empty(G, dtype=torch.int32, device=dev)
    scores = torch.empty((G, n), dtype=torch.float64, device=dev) if want_scores else None
    p = _lib.ptr
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    if gene_order is not None and not per_gene:
        assert gene_order.dtype == torch.int32 and gene_order.numel() == G and gene_order.is_contiguous()
        rc = ctx.lib.sctda_conn_perm_test_ordered(ctx.handle, ctypes.byref(dg.struct), G, p(Y),
                                                  int(Y.stride(0)) if G else dg.Ncells, 1 if log2 else 0, n, p(perms),
                                                  p(gene_order), p(observed), float(tie_rel), p(exceed), p(ties),
                                                  p(scores), st)
    else:
        rc = ctx.lib.sctda_conn_perm_test(ctx.handle, ctypes.byref(dg.struct), G, p(Y),
                                          int(Y.stride(0)) if G else dg.Ncells, 1 if log2 else 0, n, p(perms), stride,
                                          p(observed), float(tie_rel), p(exceed), p(ties), p(scores), st)
    ctx.check(rc)
    return exceed, ties, scores


def check_permutation_block(perms, S):
    """A user-supplied shared block must be [n, S] with every row a permutation of arange(S): the kernels
    gather table rows through these indices without bounds checks."""
    perms = numpy.asarray(perms)
    if perms.ndim != 2 or perms.shape[1] != S:
        raise ValueError('perms must be [n, %d] (one row of cell positions per permutation), got %r' % (S, perms.shape))
    if perms.size:
        if int(perms.min()) < 0 or int(perms.max()) >= S:
            raise ValueError('perms entries must lie in [0, %d)' % S)
        srt = numpy.sort(perms, axis=1)
        if not (srt == numpy.arange(S, dtype=srt.dtype)[None, :]).all():
            raise ValueError('every row of perms must be a permutation of arange(%d)' % S)
    return perms


def conn_perm_test_host(ctx, dg, Y_host, perms_host, observed_host, log2=True):
    """The permutation test on HOST buffers (pinned torch tensors or numpy arrays): copies the
    table, the shared permutation block and the observed scores to the device, runs
    sctda_conn_perm_test and returns the exceed counts as a host int32 tensor.  This is the
    call behind UnrootedGraph.connectivity_pvalues and bench.py's end-to-end number."""
    dev = dg.device
    Yd = torch.as_tensor(Y_host).to(dev, non_blocking=True)
    pd = torch.as_tensor(perms_host).to(dev, non_blocking=True)
    od = torch.as_tensor(observed_host).to(dev, non_blocking=True)
    order = geneorder.gene_order(Yd) if pd.dim() == 2 else None        # tile layout of the shared-block kernel
    ex, _, _ = conn_perm_test(ctx, dg, Yd, pd, od, log2=log2, per_gene=(pd.dim() == 3), gene_order=order)
    out = torch.empty(ex.shape, dtype=ex.dtype, pin_memory=True)
    out.copy_(ex, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    return out


class UnrootedGraph(object):
    """Topological analysis of non-longitudinal single cell RNA-seq data (ref:781-784)."""

    def __init__(self, name, table, shift=None, log2=True, posgl=False, csv=False, groups=True, device=0):
        """Same arguments as ref:785; `posgl` is accepted and ignored (layouts are
        visualisation state); `device` selects the GPU."""
        self.name = name
        self.g = networkx.read_gexf(name + '.gexf')                              # ref:801
        comps = list(networkx.connected_components(self.g))
        sizes = [len(c) for c in comps]
        lcc = comps[sizes.index(max(sizes))]                                     # ref:802-804: first largest
        self.pl = [v for v in self.g.nodes() if v in lcc]                        # ref:805 (file order)
        self.gl = self.g.subgraph(self.pl)
        self.adj_csr = scipy.sparse.csr_matrix(
            networkx.to_scipy_sparse_array(self.g, nodelist=self.pl, dtype=numpy.float64, weight='weight', format='csr'))
        self.log2 = log2
        with open(name + '.json', 'r') as h:
            self.dic = json.load(h)                                              # ref:828-829
        if groups:
            with open(name + '.groups.json', 'r') as h:
                self.dicgroups = json.load(h)
        else:
            self.dicgroups = {}
        names, Y, cell_id, libs, cdr = formats.read_table(table, shift=shift, csv=csv)
        self.gene_names = names
        self.gene_row = {}
        for i, q in enumerate(names):                                            # later duplicates win, ref:876-879
            self.gene_row[q] = i
        self.Y = Y
        self.cellID = numpy.array(cell_id)
        self.libs = libs
        self.cdr = cdr
        samples = []
        for i in self.pl:                                                        # ref:886-889
            samples += self.dic[i]
        self.samples = numpy.array(list(set(samples)))
        self.device = device
        self._ctx = None
        self._dg = {}
        self._Yd = None
        self._stats = None

    # -- state -----------------------------------------------------------------------------------
    @property
    def dicgenes(self):
        """ref:880-885: column name -> values over all table rows."""
        return {q: self.Y[i] for q, i in self.gene_row.items()}

    @property
    def adj(self):
        """Dense float64 adjacency of the largest component (ref:806); built on demand."""
        return self.adj_csr.toarray()

    def _context(self):
        if self._ctx is None:
            self._ctx = _lib.context(self.device)
        return self._ctx

    def _device_graph(self, con=True, ind=1):
        key = (bool(con), int(ind))
        if key not in self._dg:
            if con:
                nodes, A = self.pl, self.adj_csr
                samples = self.samples
            else:
                nodes = [str(i) for i in self.dic.keys()]
                A = scipy.sparse.csr_matrix(networkx.to_scipy_sparse_array(
                    self.g, nodelist=nodes, dtype=numpy.float64, weight='weight', format='csr'))
                s = []
                for i in nodes:
                    s += self.dic[i]
                samples = numpy.array(list(set(s)))
            if ind != 1:
                A = scipy.sparse.csr_matrix(A ** int(ind)) if ind > 0 else scipy.sparse.identity(len(nodes), format='csr')
            self._dg[key] = DeviceGraph([self.dic[i] for i in nodes], A, samples, self.Y.shape[1], self.device)
            self._dg[key].nodes = list(nodes)
        return self._dg[key]

    def _device_table(self):
        if self._Yd is None:
            self._Yd = torch.from_numpy(numpy.ascontiguousarray(self.Y)).to(torch.device('cuda', self.device))
        return self._Yd

    def _rows(self, genes):
        return torch.as_tensor([self.gene_row[g] for g in genes], dtype=torch.int64,
                               device=torch.device('cuda', self.device))

    def _all_stats(self):
        """node statistics of every table column, computed once (log2 mode of the object)."""
        if self._stats is None:
            dist = getattr(self, '_dist_dev', None)
            st = node_stats(self._context(), self._device_graph(), self._device_table(), log2=self.log2, dist=dist)
            self._stats = {k: v.cpu().numpy() for k, v in st.items()}
        return self._stats

    # -- per-gene API (ref:893-1051) ---------------------------------------------------------------
    def get_gene(self, genin, ignore_log=False, con=True):
        """ref:893-964: (dict node -> normalised node value, normalisation factor)."""
        if genin is not None and not isinstance(genin, list):
            if 'lib_' in genin[:4]:                                              # ref:902-903
                return self.count_gene('lib', genin[genin.index('_') + 1:], con=con)
            if 'ID_' in genin[:3]:                                               # ref:904-905
                return self.count_gene('ID', genin[genin.index('_') + 1:], con=con)
            if genin == '_CDR':                                                  # ref:906-928: node means of the cellular detection rate
                return self._node_means(numpy.asarray(self.cdr, dtype=numpy.float64), con)
        genes = genin if isinstance(genin, list) else [genin]
        dg = self._device_graph(con=con)
        real = genes
        if None in genes:                                   # a None resets the accumulation, ref:943-945
            real = genes[len(genes) - genes[::-1].index(None):]
        pol = numpy.zeros(dg.V)
        if real:
            Y = self._device_table().index_select(0, self._rows(real))
            st = node_stats(self._context(), dg, Y, log2=(self.log2 and not ignore_log), want_p=True)
            p = st['p'].cpu().numpy()
            tol = st['tol'].cpu().numpy()
            if len(real) == 1:
                return {k: float(v) for k, v in zip(dg.nodes, p[0])}, float(tol[0])
            for i in range(len(real)):
                pol += p[i] * tol[i] if tol[i] > 0.0 else p[i]
        tot = float(numpy.sum(pol))
        if tot > 0.0:
            pol = pol / tot
        return {k: float(v) for k, v in zip(dg.nodes, pol)}, tot

    def _node_means(self, values, con=True):
        """(dict node -> normalised node mean of a per-cell column, normalisation factor): the tail of
        get_gene / count_gene (ref:921-928, 1373-1383) for a column given as an array over the table rows;
        the node means are taken by sctda_node_stats in linear mode."""
        dg = self._device_graph(con=con)
        Y = torch.from_numpy(numpy.ascontiguousarray(values, dtype=numpy.float64).reshape(1, -1)).to(dg.device)
        st = node_stats(self._context(), dg, Y, log2=False, want_p=True)
        p = st['p'].cpu().numpy()[0]
        return {k: float(v) for k, v in zip(dg.nodes, p)}, float(st['tol'].cpu().numpy()[0])

    def count_gene(self, genin, cond, con=True):
        """ref:1347-1383: for every node the fraction of its cells whose column `genin` ('lib', 'ID' or a
        table column) equals `cond`, normalised to sum 1 over the nodes; returns (dict, normalisation)."""
        if genin is None:
            dg = self._device_graph(con=con)
            return {k: 0.0 for k in dg.nodes}, 0.0
        if genin == 'lib':
            geys = numpy.asarray(self.libs, dtype=object)
        elif genin == 'ID':
            geys = numpy.asarray(self.cellID, dtype=object)
        else:
            geys = self.Y[self.gene_row[genin]]
        return self._node_means((geys == cond).astype(numpy.float64), con)

    def connectivity(self, genis, ind=1):
        """ref:994-1005."""
        if isinstance(genis, list) or ind != 1:
            dicgen = self.get_gene(genis)[0]
            ex = numpy.array([dicgen[u] for u in self.pl])
            A = self.adj_csr if ind == 1 else (self.adj_csr ** int(ind))
            cor = float(len(self.pl)) / float(len(self.pl) - 1)
            return cor * float(ex @ (A @ ex))
        return float(self._all_stats()['conn'][self.gene_row[genis]])

    def delta(self, genis, group=None):
        """ref:1007-1028."""
        if group is None and not isinstance(genis, list):
            s = self._all_stats()
            i = self.gene_row[genis]
            return float(s['mean'][i]), float(s['min'][i]), float(s['max'][i])
        dicgen, tot = self.get_gene(genis)
        if group is not None:
            per = []
            if isinstance(group, list):
                for k in group:
                    per += self.dicgroups[k]
                per = list(set(per))
            else:
                per = self.dicgroups[group]
            mi = [dicgen[str(node)] for node in per]
        else:
            mi = [dicgen[node] for node in self.pl]
        if numpy.mean(mi) > 0.0:
            return numpy.mean(mi) * tot, numpy.min(mi) * tot, numpy.max(mi) * tot
        return 0.0, 0.0, 0.0

    def expr(self, genis, group=None):
        """ref:1030-1051."""
        if group is None:
            return int(self._all_stats()['expr'][self.gene_row[genis]])
        per = []
        if isinstance(group, list):
            for k in group:
                per += self.dicgroups[k]
            per = list(set(per))
        else:
            per = self.dicgroups[group]
        po = []
        for q in per:
            po += list(self.dic[str(q)])
        po = list(set(po))
        return int(numpy.count_nonzero(self.Y[self.gene_row[genis]][po] > 0.0))

    def connectivity_pvalue(self, genin, n=500):
        """ref:966-992: n fresh shuffles from the global numpy.random stream."""
        if genin is None:
            return 0.0
        return self.connectivity_pvalues([genin], n=n)[0]

    def connectivity_pvalues(self, genes, n=500, perms=None, batch_bytes=1 << 30, return_ties=False):
        """p-values of `genes`, in order.  perms=None draws a fresh n x S block per gene from
        the global numpy.random stream, gene after gene, exactly as ref:1069-1071 consumes it;
        an int32 array [n, S] is used as one block shared by all genes."""
        ctx, dg = self._context(), self._device_graph()
        dev = dg.device
        stats = self._all_stats()
        rows = [self.gene_row[g] for g in genes]
        obs_all = torch.from_numpy(stats['conn'][rows]).to(dev)
        Yd = self._device_table()
        pvals = numpy.zeros(len(genes))
        ties = numpy.zeros(len(genes), dtype=numpy.int64)
        if len(genes) == 0 or n == 0:
            return (list(pvals), ties) if return_ties else list(pvals)
        if perms is not None:
            perms = check_permutation_block(perms, dg.S)
            if int(perms.shape[0]) != int(n):
                raise ValueError('perms has %d rows but n = %d: pass n = perms.shape[0] (the p-value is the exceed '
                                 'count over the number of permutations applied)' % (perms.shape[0], n))
            pd = torch.as_tensor(numpy.ascontiguousarray(perms, dtype=numpy.int32)).to(dev)
            from . import parallel
            rank, ws = parallel.world()
            rows_all = self._rows(genes)
            if ws > 1:
                # under torch.distributed the genes are sharded over the ranks (no collective on the data path)
                # and the counts gathered on every rank.  The tile plan is made on ALL the genes
                # (rank 0, broadcast: every rank must cut the same shards) and a rank takes every N-th 128-gene tile
                # of the planned order: tiles as homogeneous as on one GPU, the same mix of sparse and dense tiles
                # on every rank.
                import torch.distributed as dist
                plan = geneorder.gene_order(Yd.index_select(0, rows_all)).long() if rank == 0 else \
                    torch.empty(len(genes), dtype=torch.int64, device=dev)
                if dist.get_backend() == 'nccl':
                    dist.broadcast(plan, src=0)
                else:
                    pc = plan.cpu()
                    dist.broadcast(pc, src=0)
                    plan = pc.to(dev)
                pos = [torch.from_numpy(parallel.tile_shard_positions(len(genes), r, ws)).to(dev) for r in range(ws)]
                mine = plan.index_select(0, pos[rank])
                if mine.numel() > 0:
                    ex, ti, _ = conn_perm_test(ctx, dg, Yd.index_select(0, rows_all.index_select(0, mine)), pd,
                                               obs_all.index_select(0, mine), log2=self.log2)
                else:
                    ex = ti = torch.zeros(0, dtype=torch.int32, device=dev)
                ex = torch.cat(parallel.all_gather_varlen(ex.to(torch.int32).reshape(-1)))
                ti = torch.cat(parallel.all_gather_varlen(ti.to(torch.int32).reshape(-1)))
                back = plan.index_select(0, torch.cat(pos))                 # gathered slot -> position in `genes`
                ex = torch.empty_like(ex).index_copy_(0, back, ex)
                ti = torch.empty_like(ti).index_copy_(0, back, ti)
            else:
                Ysel = Yd.index_select(0, rows_all)
                ex, ti, _ = conn_perm_test(ctx, dg, Ysel, pd, obs_all, log2=self.log2, gene_order=geneorder.gene_order(Ysel))
            pvals = ex.cpu().numpy() / float(n)
            ties = ti.cpu().numpy()
        else:
            per = max(1, int(batch_bytes // max(1, 4 * n * dg.S)))
            for lo in range(0, len(genes), per):
                hi = min(len(genes), lo + per)
                block = numpy.empty((hi - lo, n, dg.S), dtype=numpy.int32)
                for i in range(hi - lo):
                    block[i] = reference_perm_indices(n, dg.S)
                pd = torch.from_numpy(block).to(dev)
                ex, ti, _ = conn_perm_test(ctx, dg, Yd.index_select(0, self._rows(genes[lo:hi])), pd, obs_all[lo:hi],
                                           log2=self.log2, per_gene=True)
                pvals[lo:hi] = ex.cpu().numpy() / float(n)
                ties[lo:hi] = ti.cpu().numpy()
        return (list(pvals), ties) if return_ties else list(pvals)

    # -- gene x gene matrices (ref:1111-1190) ---------------------------------------------------------
    def _profiles(self, lista):
        """Normalised node profiles P [len(lista), V] of the genes (device), nodes in self.pl order."""
        Y = self._device_table().index_select(0, self._rows(lista))
        return node_stats(self._context(), self._device_graph(), Y, log2=self.log2, want_p=True)['p']

    def JSD_matrix(self, lista, maximum_matrix_entries=5 * 10 ** 7, verbose=False):
        """ref:1111-1155: Jensen-Shannon distance matrix of the genes' node profiles
        (`maximum_matrix_entries` only bounds the reference's temporaries; accepted and ignored)."""
        import ctypes
        ctx = self._context()
        P = self._profiles(lista)
        g, V = int(P.shape[0]), int(P.shape[1])
        out = torch.empty((g, g), dtype=torch.float64, device=P.device)
        ctx.check(ctx.lib.sctda_jsd_matrix(ctx.handle, g, V, _lib.ptr(P), int(P.stride(0)), _lib.ptr(out), g,
                                           ctypes.c_void_p(torch.cuda.current_stream(P.device).cuda_stream)))
        return out.cpu().numpy()

    def cor_matrix(self, lista, c=1):
        """ref:1157-1163: correlation distance matrix of the genes over the table rows (cells)."""
        from . import mapper
        Y = self._device_table().index_select(0, self._rows(lista))
        return mapper.pairwise_distances(self._context(), mapper._padded_device_table(Y, Y.device), 'correlation').cpu().numpy()

    def adjacency_matrix(self, lista, ind=1, verbose=False):
        """ref:1165-1190: V/(V-1) * P A^ind P' for the genes' node profiles (one contraction of [P; P A^ind])."""
        from . import mapper
        ctx = self._context()
        P = self._profiles(lista)
        g, V = int(P.shape[0]), int(P.shape[1])
        A = self.adj_csr if ind == 1 else (self.adj_csr ** int(ind))
        PA = torch.from_numpy(numpy.ascontiguousarray((A @ P.cpu().numpy().T).T)).to(P.device)
        stacked = mapper._padded_device_table(torch.cat([P, PA], dim=0), P.device)
        rows = numpy.arange(g)
        M = mapper.gram_gather(ctx, stacked, rows, rows + g)
        cor = float(V) / float(V - 1)
        return (cor * M).cpu().numpy()

    def cellular_subpopulations(self, threshold=0.05, min_cells=5, clus_thres=0.65):
        """ref:1423-1446: genes of name.genes.tsv with q-value below `threshold` expressed in more than
        `min_cells` cells, clustered by average linkage on the rows of their Jensen-Shannon distance
        matrix (device: sctda_jsd_matrix, then the row distances by the FP64 contraction); returns the
        list of gene lists, one per dendrogram colour at `clus_thres`."""
        nam = []
        with open(self.name + '.genes.tsv', 'r') as f:
            for n, line in enumerate(f):
                if n > 0:
                    sp = line[:-1].split('\t')
                    if float(sp[7]) < threshold and float(sp[1]) > min_cells:
                        nam.append(sp[0])
        mat2 = self.JSD_matrix(nam)
        z = _hierarchical_clustering(mat2, 'average', True, clus_thres, self._context())
        return [[nam[x] for x in m] for m in find_clusters(z).values()]

    # -- table (ref:1053-1086) -----------------------------------------------------------------------
    def _table_rows(self, n, filtercells, filterexp, annotation, rooted, perms):
        s = self._all_stats()
        lp = sorted(self.gene_row.keys())                                        # ref:1068
        keep = [gi for gi in lp if s['expr'][self.gene_row[gi]] > filtercells and s['max'][self.gene_row[gi]] > filterexp]
        pol = self.connectivity_pvalues(keep, n=n, perms=perms)                  # ref:1069-1071
        por = benjamini_hochberg(pol)                                            # ref:1072
        header = ['Gene', 'Cells', 'Mean', 'Min', 'Max', 'Connectivity', 'p_value', 'q-value (BH)']
        if rooted:
            header += ['Centroid', 'Dispersion']
        header += sorted(annotation.keys())
        rows = []
        for mj, gi in enumerate(keep):                                           # ref:1074-1086
            i = self.gene_row[gi]
            row = [gi, int(s['expr'][i]), float(s['mean'][i]), float(s['min'][i]), float(s['max'][i]),
                   float(s['conn'][i]), float(pol[mj]), float(por[mj])]
            if rooted:
                row += self._centroid_from_raw(s['cen'][i], s['disp'][i])
            for m in sorted(annotation.keys()):
                row.append('Y' if gi in annotation[m] else 'N')
            rows.append(row)
        return header, rows

    def table(self, n=500, filtercells=0, filterexp=0.0, annotation={}, perms=None):
        """(header, rows) of the gene table save() writes, without writing it."""
        return self._table_rows(n, filtercells, filterexp, annotation, isinstance(self, RootedGraph), perms)

    def save(self, n=500, filtercells=0, filterexp=0.0, annotation={}, *, perms=None):
        """ref:1053-1086: writes name.genes.tsv and returns None, as the reference does (the scatter
        plot of ref:1087-1109 is not produced; the rows stay available as self.last_table).  `perms` (keyword only, an
        extension): see connectivity_pvalues."""
        header, rows = self.last_table = self.table(n, filtercells, filterexp, annotation, perms)
        formats.write_genes_tsv(self.name + '.genes.tsv', header, rows)


class RootedGraph(UnrootedGraph):
    """Longitudinal data: root node, pseudo-time regression, centroid / dispersion columns
    (ref:1449-1567, 1724-1743, 1879-1914).  Root finding (get_dendrite, ref:1463-1479: all-pairs
    BFS plus one Spearman correlation per node) runs on the device (rooted.cu); the one-off
    linear regression of ref:1563-1567 is host scipy."""

    def __init__(self, name, table, rootlane='timepoint', shift=None, log2=True, posgl=False, csv=False,
                 groups=True, device=0):
        UnrootedGraph.__init__(self, name, table, shift, log2, posgl, csv, groups, device)
        self.rootlane = rootlane
        self.root, self.leaf = self.find_root(self.get_dendrite())               # ref:1536
        self.dicdis = self.get_distroot(self.root)                               # ref:1562
        self._skeleton_state()                                                   # ref:1537-1556
        pel2, tol = self.get_gene(self.rootlane, ignore_log=True)                # ref:1563
        pel = numpy.array([pel2[m] for m in self.pl]) * tol
        dr = numpy.array([self.dicdis[m] for m in self.pl])
        self.po = scipy.stats.linregress(pel, dr)                                # ref:1566
        self._dist_dev = torch.from_numpy(dr.astype(numpy.float64)).to(torch.device('cuda', self.device))
        self._stats = None

    def _root_scores(self):
        """sctda_root_scores: (scores [V], hop distances [V, V]) for every source node."""
        if not hasattr(self, '_rs'):
            import ctypes
            ctx = self._context()
            dev = torch.device('cuda', self.device)
            daycolor = self.get_gene(self.rootlane, ignore_log=True)[0]
            day = numpy.array([daycolor[k] for k in self.pl])
            x = day - day.min()                                                   # ref:1472-1475
            order = numpy.argsort(x, kind='stable')
            xs = x[order]
            V = len(self.pl)
            new_group = numpy.concatenate([[True], xs[1:] != xs[:-1]])
            starts = numpy.nonzero(new_group)[0]
            gid = numpy.cumsum(new_group) - 1
            ends = numpy.concatenate([starts[1:], [V]])
            A = scipy.sparse.csr_matrix(self.adj_csr + self.adj_csr.T)
            A.sort_indices()
            t = [torch.from_numpy(numpy.ascontiguousarray(a, dtype=numpy.int32)).to(dev)
                 for a in (A.indptr, A.indices, order, starts[gid], ends[gid])]
            score = torch.empty(V, dtype=torch.float64, device=dev)
            dist = torch.empty((V, V), dtype=torch.int32, device=dev)
            ctx.check(ctx.lib.sctda_root_scores(ctx.handle, V, _lib.ptr(t[0]), _lib.ptr(t[1]), _lib.ptr(t[2]), _lib.ptr(t[3]),
                                                _lib.ptr(t[4]), _lib.ptr(score), _lib.ptr(dist),
                                                ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
            self._rs = (score.cpu().numpy(), dist.cpu().numpy())
        return self._rs

    def get_distroot(self, root):
        """ref:1454-1461: hop distance of every node of the largest component to `root`."""
        d = self._root_scores()[1][self.pl.index(str(root))]
        return {k: int(v) for k, v in zip(self.pl, d)}

    def get_dendrite(self):
        """ref:1463-1479: for every node, minus the Spearman correlation between time point and hop
        distance over the nodes that are not at the maximal distance (all sources in one kernel)."""
        sc = self._root_scores()[0]
        return {str(k): float(v) for k, v in zip(self.pl, sc)}

    @staticmethod
    def find_root(dendritem):
        """ref:1481-1497: (root, leaf) = the nodes with the smallest and the largest score among the
        scores above -2 (NaN scores never qualify), the first such node in dictionary order on ties;
        the reference's sentinels are kept: no score below 1000 leaves root = 0, none above -1000
        leaves leaf = 0."""
        keys = list(dendritem.keys())
        if not keys:
            return 0, 0
        v = numpy.array([dendritem[k] for k in keys], dtype=numpy.float64)
        ok = v > -2.0
        lo = numpy.where(ok & (v < 1000.0), v, numpy.inf)
        hi = numpy.where(ok & (v > -1000.0), v, -numpy.inf)
        root = keys[int(numpy.argmin(lo))] if numpy.isfinite(lo).any() else 0
        leaf = keys[int(numpy.argmax(hi))] if numpy.isfinite(hi).any() else 0
        return root, leaf

    def dendritic_graph(self):
        """ref:1499-1520: skeleton of the graph.  Level n = the nodes of the largest component at hop
        distance n from the root (n < diameter - 1; distances and eccentricities from the device BFS of
        sctda_root_scores); dicdend[n] = connected components of the level (ordered by their first node in
        self.pl order, the order networkx yields them in); skeleton nodes 'n_i', an edge between
        components of consecutive levels whose union is connected."""
        import scipy.sparse.csgraph
        dist_all = self._root_scores()[1]
        diam = int(dist_all.max())
        d = numpy.array([self.dicdis[k] for k in self.pl])
        A = scipy.sparse.csr_matrix(((self.adj_csr + self.adj_csr.T) != 0).astype(numpy.int8))
        g3 = networkx.Graph()
        dicdend = {}
        comp_of = {}
        for n in range(diam - 1):
            idx = numpy.nonzero(d == n)[0]
            comps = []
            if idx.size:
                ncomp, lab = scipy.sparse.csgraph.connected_components(A[idx][:, idx], directed=False)
                order = []
                for c in lab:                                                    # components by first appearance
                    if c not in order:
                        order.append(int(c))
                comps = [set(self.pl[i] for i in idx[lab == c]) for c in order]
                for n2, c in enumerate(order):
                    comp_of[(n, n2)] = idx[lab == c]
            dicdend[n] = comps
            for n2 in range(len(comps)):
                g3.add_node('%d_%d' % (n, n2))
                if n > 0:
                    for n3 in range(len(dicdend[n - 1])):
                        # both parts are connected: the union is connected iff an edge joins them
                        if A[comp_of[(n, n2)]][:, comp_of[(n - 1, n3)]].nnz > 0:
                            g3.add_edge('%d_%d' % (n, n2), '%d_%d' % (n - 1, n3))
        self._comp_idx = comp_of
        self._sym = A
        return g3, dicdend

    def _skeleton_state(self):
        """ref:1537-1556: sizes of the skeleton's edges (graph edges between the two components) and nodes
        (distinct cells of a component)."""
        self.g3, self.dicdend = self.dendritic_graph()
        self.edgesize, self.dicedgesize = [], {}
        self.nodesize, self.dicmelisa = [], {}
        self.edgesizeprun, self.nodesizeprun, self.dicmelisaprun = [], [], {}
        key = lambda name: tuple(int(x) for x in name.split('_'))
        for ee in self.g3.edges():
            a, b = self._comp_idx[key(ee[0])], self._comp_idx[key(ee[1])]
            self.edgesize.append(int(self._sym[a][:, b].nnz))
            self.dicedgesize[ee] = self.edgesize[-1]
        for ee in self.g3.nodes():
            cells = set()
            for uu in self.dicdend[key(ee)[0]][key(ee)[1]]:
                cells.update(self.dic[uu])
            self.nodesize.append(len(cells))
            self.dicmelisa[ee] = cells

    def select_diff_path(self):
        """ref:1569-1593: the path through the skeleton that starts at '0_0' and always follows the
        heaviest unused edge towards a deeper level (the first such edge in dictionary order on ties)."""
        level = lambda name: int(name.split('_')[0])
        lista, last = [], '0_0'
        while True:
            best, siz = None, 0
            for ee, sz in self.dicedgesize.items():
                down = (ee[0] == last and level(ee[1]) > level(ee[0])) or (ee[1] == last and level(ee[0]) > level(ee[1]))
                if down and sz > siz and ee not in lista:
                    best, siz = ee, sz
            if best is None:
                return lista
            lista.append(best)
            last = best[1] if level(best[1]) > level(best[0]) else best[0]

    def cellular_subpopulations(self, min_dispersion, threshold=0.05, min_cells=5, max_K=8, method='centroid',
                                clus_thres=0.65):
        """ref:1950-2040.  method='js': the Jensen-Shannon route of UnrootedGraph.cellular_subpopulations.
        method='centroid': genes of name.genes.tsv with dispersion below `min_dispersion`, q-value below
        `threshold`, expressed in more than `min_cells` cells; their centroids are cut by k-means for
        k = 2 .. max_K - 1 (sklearn KMeans, random_state=170, as the reference) and the k with the smallest mean
        of max_j (var_i - var_j) / |mean_i - mean_j| is kept (the reference's own variant of the Davies-Bouldin
        index, ref:1995-2003); returns the gene lists of the clusters.  The two figures are not drawn."""
        if method == 'js':
            return UnrootedGraph.cellular_subpopulations(self, threshold=threshold, min_cells=min_cells,
                                                         clus_thres=clus_thres)
        if method != 'centroid':
            return None                                                          # the reference falls through
        import sklearn.cluster
        con, nam = [], []
        with open(self.name + '.genes.tsv', 'r') as f:
            for n, line in enumerate(f):
                if n > 0:
                    sp = line[:-1].split('\t')
                    if float(sp[9]) < min_dispersion and float(sp[7]) < threshold and float(sp[1]) > min_cells:
                        con.append(float(sp[8]))
                        nam.append(sp[0])
        pts = numpy.array([[c, 0.0] for c in con])

        def labels(m):
            return sklearn.cluster.KMeans(n_clusters=m, random_state=170).fit_predict(pts)
        y = []
        for m in range(2, max_K):
            pred = labels(m)
            clus = [[q for t, q in zip(pred, con) if t == i] for i in range(m)]
            var = [numpy.var(c) for c in clus]
            mean = [numpy.mean(c) for c in clus]
            ri = [max([0.0 if j == i else (var[i] - var[j]) / abs(mean[i] - mean[j]) for j in range(m)]) for i in range(m)]
            y.append(numpy.mean(ri))
        vals = numpy.array(y, dtype=numpy.float64)
        ok = vals < numpy.inf                                                    # NaN and inf never win the `<` of ref:2016-2021
        rn = int(numpy.argmin(numpy.where(ok, vals, numpy.inf))) if ok.any() else -1   # first minimum
        g = [[] for _ in range(rn + 2)]
        for m, n in zip(list(labels(rn + 2)), nam):
            g[m].append(n)
        return g

    def _centroid_from_raw(self, cen, disp):
        if numpy.isnan(cen):
            return [None, None]                                                  # ref:1742-1743
        return [(float(cen) - self.po[1]) / self.po[0], (float(disp) - self.po[1]) / self.po[0]]

    def centroid(self, genin, ignore_log=False):
        """ref:1724-1743."""
        if ignore_log or isinstance(genin, list):
            dicge = self.get_gene(genin, ignore_log=ignore_log)[0]
            p = numpy.array([dicge[k] for k in self.pl])
            d = numpy.array([self.dicdis[k] for k in self.pl], dtype=numpy.float64)
            if p.sum() > 0.0:
                cen = float((d * p).sum() / p.sum())
                return self._centroid_from_raw(cen, float(numpy.sqrt((((d - cen) ** 2) * p).sum() / p.sum())))
            return [None, None]
        s = self._all_stats()
        i = self.gene_row[genin]
        return self._centroid_from_raw(s['cen'][i], s['disp'][i])

    def save(self, n=500, filtercells=0, filterexp=0.0, annotation={}, *, perms=None):
        """ref:1879-1914: as UnrootedGraph.save with the Centroid and Dispersion columns; returns None."""
        header, rows = self.last_table = self.table(n, filtercells, filterexp, annotation, perms)
        formats.write_genes_tsv(self.name + '.genes.tsv', header, rows)
