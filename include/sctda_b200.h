/*
 * sctda_b200 — C-ABI of the B200-native scTDA hot path.
 *
 * The reference (CamaraLab/scTDA, /root/reference/scTDA/main.py, abbreviated "ref:")
 * is pure Python and has no FFI of its own; its seams for this path are the Python
 * methods listed beside each entry point below.  This header is what a
 * ctypes binding of those methods calls; sctda_b200/_lib.py is that binding and
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer on the context's GPU; the caller
 *     allocates and owns all of them (torch tensors: .data_ptr()); the library only
 *     borrows them for the duration of the stream work it enqueues;
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream);
 *     no call synchronises the host unless its comment says so;
 *   - every call returns an int32 status: 0 = ok, <0 = error class below; the message
 *     is kept in the context (sctda_last_error).  Nothing throws, nothing aborts;
 *   - a context is bound to one GPU and may be used by one host thread at a time;
 *   - all index arrays are int32, all expression / score arrays IEEE float64 unless
 *     stated; matrices are row-major with an explicit leading dimension in ELEMENTS.
 */
#ifndef SCTDA_B200_H
#define SCTDA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCTDA_OK                 0
#define SCTDA_ERR_INVALID       -1   /* bad argument (NULL pointer, negative size, ...) */
#define SCTDA_ERR_CUDA          -2   /* a CUDA runtime call or a kernel launch failed    */
#define SCTDA_ERR_NOMEM         -3   /* workspace allocation failed                       */
#define SCTDA_ERR_UNSUPPORTED   -4   /* valid request the library cannot serve            */

#define SCTDA_ABI_VERSION        1

typedef struct sctda_ctx sctda_ctx;

/* ---- context ------------------------------------------------------------------------------ */

/* ABI version of the loaded library (SCTDA_ABI_VERSION at build time). */
int32_t sctda_abi_version(void);

/* Create a context on CUDA device `device` (sets that device current for the calling
 * thread).  Owns a grow-only device workspace and the last-error string. */
int32_t sctda_ctx_create(int32_t device, sctda_ctx** out_ctx);

/* Frees the workspace.  Synchronises the device. */
int32_t sctda_ctx_destroy(sctda_ctx* ctx);

/* Message of the last failing call on this context ("" if none).  Owned by ctx. */
const char* sctda_last_error(const sctda_ctx* ctx);

/* Bytes of device workspace currently held by the context. */
int64_t sctda_workspace_bytes(const sctda_ctx* ctx);

/* Number of kernel launches enqueued through this context since it was created
 * (bench.py reports the per-step difference as "gpu_launches"). */
int64_t sctda_launch_count(const sctda_ctx* ctx);

/* Precision of the cells x cells contraction (BASELINE.json north_star (1)).
 *   0  FP64 on the DMMA tensor pipe, one fused-multiply-add chain per element: the DEFAULT, the
 *      Mapper graph is identical to the oracle's bit for bit;
 *   1  opt-in split precision on the tcgen05 tensor cores: every value is split into two FP16
 *      numbers (hi, scaled residual: 22 significant bits, the "3xTF32" scheme in half the bytes),
 *      three kind::f16 products with fp32 accumulation in TMEM, drained into float64 sums every 512
 *      features.  Gram entries of unit-normalised rows / squared distances within 1e-5 of the
 *      FP64 result relative to |x||y| (measured <= 4e-6: tests/test_mapper_gpu.py); the graph is
 *      NOT guaranteed identical (it was on every input of the bench and the tests).
 * Applies to sctda_bin_cluster and sctda_pairdist_rows_f64 of this context (sctda_gram_gather_f64
 * always computes in FP64).  Replaces nothing in the reference: sakmapper / sklearn compute
 * these distances in float64 (/root/reference/scTDA/main.py:746-751, 768-771). */
int32_t sctda_set_precision(sctda_ctx* ctx, int32_t mode);
int32_t sctda_get_precision(const sctda_ctx* ctx);

/* Linkage of the per-patch agglomerative clustering (mapper_graph(clust='agglomerative'), ref:758-771;
 * sakmapper hands the patch to sklearn.cluster.AgglomerativeClustering, whose linkage the call sites do not pin):
 *   0  single linkage: the minimum spanning tree of the patch (Prim kernels), the DEFAULT
 *   1  average, 2  Ward, 3  complete: nearest-neighbour chain with Lance-Williams updates on a working copy of the
 *      patch's distance matrix (one CTA per patch; twice the matrix memory per batch).
 * The model selection (Davies-Bouldin over K = 2..K_max, or the histogram-gap cut) is the same for all four. */
int32_t sctda_set_linkage(sctda_ctx* ctx, int32_t mode);
int32_t sctda_get_linkage(const sctda_ctx* ctx);

/* Measurement aid (bench.py "roofline"): when enabled, every hot-path call brackets its
 * DOMINANT kernel (conn_perm_kernel for sctda_conn_perm_test) with CUDA events on the
 * caller's stream.  sctda_last_kernel_ms waits for that kernel to finish (host sync on
 * the event) and returns its duration in milliseconds, summed over the launches of the last
 * call (sctda_bin_cluster launches the contraction once per batch of patches). */
int32_t sctda_set_kernel_timing(sctda_ctx* ctx, int32_t enabled);
int32_t sctda_last_kernel_ms(sctda_ctx* ctx, float* out_ms);

/* Yard-stick for bench.py: measured L2 -> SM read bandwidth (GB/s) of this GPU, all SMs streaming a
 * 32 MB L2-resident buffer with 16-byte loads.  The permutation kernel's gathers are served
 * from L2, so this is the on-chip ceiling its achieved gather rate is reported against.
 * Synchronises the host. */
int32_t sctda_probe_l2_read_gbs(sctda_ctx* ctx, double* out_gbs);

/* Side of the square output tiles of the FP64 contraction kernel (symmetric problems run the upper triangle of
 * tiles and mirror it): what a caller needs to count the flops the kernel executes, next to the algorithmic
 * 2*G*m^2 of SURVEY.md 8d (bench.py: roofline.frac_executed). */
int32_t sctda_gemm_tile_fp64(void);

/* ---- hot path (b): connectivity permutation test ------------------------------------------- */

/*
 * Graph and table state shared by the two calls below.  It is the device image of the
 * attributes UnrootedGraph.__init__ builds (ref:800-806, 828-829, 886-889):
 *
 *   V            number of nodes of the largest connected component (len(self.pl))
 *   d_node_off   [V+1]   CSR offsets of the node member lists, nodes in self.pl order
 *   d_node_cols  [Mtot]  members as COLUMN POSITIONS into `samples` (0..S-1), in the
 *                        member-list order of self.dic[node] (the summation order)
 *   d_adj_off    [V+1]   CSR of self.adj (ref:806), rows in self.pl order
 *   d_adj_idx    [nnz]   column (node) indices
 *   d_adj_w      [nnz]   float64 edge weights
 *                        The adjacency enters only through the quadratic form p'Ap
 *                        (ref:989, 1003): any A' with A' + A'^T == A + A^T gives the same
 *                        statistic, so callers may pass the upper-triangular form
 *                        triu(A + A^T, 1) + diag(A) and halve the sparse work.
 *   S            len(self.samples); d_samples [S] = self.samples (table row of each column)
 *   Ncells       rows of the expression table
 */
typedef struct sctda_graph {
    int32_t        V;
    int32_t        S;
    int32_t        Ncells;
    int32_t        Mtot;          /* total number of node members (= d_node_off[V]); 0 = not given: the library reads it
                                     from the device with one synchronising 4-byte copy when it needs it */
    const int32_t* d_node_off;
    const int32_t* d_node_cols;
    const int32_t* d_adj_off;
    const int32_t* d_adj_idx;
    const double*  d_adj_w;
    const int32_t* d_samples;
} sctda_graph;

/*
 * Batched UnrootedGraph.expr / get_gene / delta / connectivity (ref:1030-1051, 893-964,
 * 1007-1028, 994-1005) and, when d_dist != NULL, the raw RootedGraph.centroid sums
 * (ref:1724-1743) for G genes at once.
 *
 *   d_Y        [G, ldy]  expression table, gene-major: d_Y[g*ldy + cell], all Ncells rows
 *                        (the device image of self.dicgenes for the G genes)
 *   log2_mode  self.log2 (ref:807): 1 = values are log2(1+TPM), node value is
 *              log2(1 + mean(2^x - 1)); 0 = plain mean
 *   d_dist     [V] float64 hop distance of each node to the root (self.dicdis) or NULL
 * outputs, each [G] (any may be NULL):
 *   d_expr     int32   number of table rows with value > 0            (expr)
 *   d_tol      float64 sum over nodes of the un-normalised node values (get_gene()[1])
 *   d_mean/d_min/d_max float64 delta(): (mean,min,max of p_k) * tol, zeros if mean <= 0
 *   d_conn     float64 connectivity(): V/(V-1) * p' A p
 *   d_cen/d_disp float64 sum(d_k p_k)/sum(p_k) and sqrt(sum((d_k-cen)^2 p_k)/sum(p_k)),
 *              NaN when sum(p_k) == 0 (the reference returns [None, None])
 *   d_p        [G, V] float64 normalised node values p_k in self.pl order, or NULL
 */
int32_t sctda_node_stats(sctda_ctx* ctx, const sctda_graph* graph,
                         int32_t G, const double* d_Y, int64_t ldy, int32_t log2_mode,
                         const double* d_dist,
                         int32_t* d_expr, double* d_tol,
                         double* d_mean, double* d_min, double* d_max,
                         double* d_conn, double* d_cen, double* d_disp,
                         double* d_p, void* stream);

/*
 * Batched UnrootedGraph.connectivity_pvalue (ref:966-992) for G genes x n permutations.
 *
 *   d_perm            int32 permutation index arrays: shuffled_row[j] = row[perm[j]]
 *                     (ref:975-976; numpy.random.shuffle applied to arange(S) consumes
 *                     the same MT19937 draws as shuffling the values)
 *   perm_gene_stride  0      : one block [n, S] shared by all G genes
 *                     n*S    : a fresh block per gene, [G, n, S] (the reference's stream)
 *   d_observed        [G] float64 observed connectivity (sctda_node_stats d_conn)
 *   tie_rel           relative half-width of the band around `observed` inside which a
 *                     permuted score is also counted in d_ties (<= 0: 1e-11).  The default
 *                     covers the worst-case rounding of both evaluations of the score (a sum
 *                     of at most nnz/2 + V float64 products, 2^-53 each: 1.3e-12 at
 *                     nnz = 24,000) and nothing else: a count in d_ties means the permuted
 *                     and the observed score are the same number up to rounding, which
 *                     needs a degenerate gene (constant over the component, or expressed in
 *                     one or two cells); the reference's own `>` is then decided by the
 *                     summation order of its BLAS build.
 * outputs:
 *   d_exceed  [G] int32  #{r : conn_r > observed, not tied}   (p-value = exceed / n), exact integers
 *   d_ties    [G] int32  #{r : |conn_r - observed| <= tie_rel*|observed|} or NULL.  A tied pair is
 *                        the same number in exact arithmetic and is NOT counted as exceeding (strict `>`,
 *                        ref:990), whatever the roundings of the two evaluations say: the result does not
 *                        depend on the kernel or its summation order.
 *   d_scores  [G, n] float64 the permuted scores conn_r (ref:989) or NULL
 *
 * Arithmetic follows ref:978-989 step by step: float64 sequential member sums of
 * 2^x - 1, /q, log1p, / 0.693147, ONE rounding to float32 (ref:972), float64 from there on.
 */
int32_t sctda_conn_perm_test(sctda_ctx* ctx, const sctda_graph* graph,
                             int32_t G, const double* d_Y, int64_t ldy, int32_t log2_mode,
                             int32_t n, const int32_t* d_perm, int64_t perm_gene_stride,
                             const double* d_observed, double tie_rel,
                             int32_t* d_exceed, int32_t* d_ties, double* d_scores,
                             void* stream);

/*
 * The same test for ONE shared permutation block [n, S] with the genes laid out in tiles in the
 * order d_gene_order [G] (a permutation of 0..G-1; NULL = table order).  Results are written at the
 * ORIGINAL gene positions, so the order never changes what is computed, only which genes share
 * a 32-byte sector of the transposed table: the kernel never requests sectors that hold only
 * zeros, and genes with similar supports empty more sectors and whole rows (the host layer's
 * sctda_b200/geneorder.py proposes an order).
 * sctda_conn_perm_test with perm_gene_stride == 0 is this call with d_gene_order == NULL.
 * Needs S <= 65,535 (16-bit staged permutation rows); larger graphs and per-gene permutation
 * blocks run on the 32-gene kernel behind sctda_conn_perm_test.
 */
int32_t sctda_conn_perm_test_ordered(sctda_ctx* ctx, const sctda_graph* graph,
                                     int32_t G, const double* d_Y, int64_t ldy, int32_t log2_mode,
                                     int32_t n, const int32_t* d_perm, const int32_t* d_gene_order,
                                     const double* d_observed, double tie_rel,
                                     int32_t* d_exceed, int32_t* d_ties, double* d_scores,
                                     void* stream);

/* Self-test of the node epilogue: the table-driven float32(log1p(sum/q)/0.693147) with its guard
 * against the full-accuracy path, bit for bit, on n_trials pseudo-random (sum, q) pairs.
 * *out_slow = trials that took the full-accuracy path (may be NULL).  Synchronises the host. */
int32_t sctda_selftest_pm_fast(sctda_ctx* ctx, int64_t n_trials, int64_t* out_mismatches, int64_t* out_slow);

/* ---- hot path (a): Mapper graph build ------------------------------------------------------- */
/*
 * The reference reaches this path through the third-party package sakmapper:
 *   sakmapper.apply_lens(df, lens='mds', metric=...)           /root/reference/scTDA/main.py:746-751
 *   sakmapper.mapper_graph(df, lens_data, resolution, gain, equalize, clust, stat, max_K)   :768-771
 * The entry points below are the stages of mapper_graph (cover -> per-patch clustering -> nodes
 * -> nerve) and the distance contraction behind apply_lens(lens='mds').  The algorithm and every
 * choice the call sites leave open are stated in oracle/mapper_oracle.py.
 * Variable-size outputs use a count call (sizes to the device, read by the caller) then a fill
 * call; the context keeps the intermediate state between the two.
 *
 * Row padding contract of every entry point that takes (G, d_Xn, ldxn): ldxn must be even.  Rows
 * are staged in 16-byte chunks with the tail zero-filled by the library.  With the environment
 * switch SCTDA_GEMM_TMA_BULK=1 and ldxn >= 16*ceil(G/16), rows are staged with 128-byte TMA bulk
 * copies instead and the columns G .. 16*ceil(G/16)-1 of every row MUST BE ZERO (this is how
 * sctda_row_normalize_f64 writes them); measured slower, kept for experiments.
 */

/* Xn = X (mode 0, metric 'euclidean') or the centred, unit-norm rows of X (mode 1, 'correlation';
 * constant rows become zeros); d_sq[i] = |Xn_i|^2 (may be NULL).  Columns G..ldxn-1 of Xn are
 * zeroed.  Sums run in a fixed order (lane-strided partials, xor butterfly). */
int32_t sctda_row_normalize_f64(sctda_ctx* ctx, int32_t N, int32_t G, const double* d_X, int64_t ldx,
                                int32_t mode, double* d_Xn, int64_t ldxn, double* d_sq, void* stream);

/* out[i][j] = sum_k Xn[rows_a[i]][k] * Xn[rows_b[j]][k]  (FP64 tensor-pipe MMA, one fused
 * multiply-add chain over k ascending per element).  rows_* may be NULL (identity).
 * Rows must be 16-byte aligned: ldxn even. */
int32_t sctda_gram_gather_f64(sctda_ctx* ctx, int32_t G, const double* d_Xn, int64_t ldxn,
                              int32_t na, const int32_t* d_rows_a, int32_t nb, const int32_t* d_rows_b,
                              double* d_out, int64_t ldo, void* stream);

/* Rows [r0, r1) of the N x N dissimilarity that apply_lens(lens='mds', metric=...) hands to MDS:
 * metric 0 = euclidean  sqrt(max(0, (sq_i + sq_j) - 2 G_ij)),  metric 1 = correlation 1 - G_ij on
 * the mode-1 rows; the diagonal is exactly 0.  d_D is [r1-r0, ldd]. */
int32_t sctda_pairdist_rows_f64(sctda_ctx* ctx, int32_t N, int32_t G, const double* d_Xn, int64_t ldxn,
                                const double* d_sq, int32_t r0, int32_t r1, int32_t metric,
                                double* d_D, int64_t ldd, void* stream);

/* One SMACOF pass of the MDS lens (apply_lens(lens='mds'), the reference's default lens; sklearn.manifold
 * MDS, metric variant): for the configuration d_Z [N, 2] and the dissimilarities d_D [N, ldd]
 * (sctda_pairdist_rows_f64), writes the Guttman transform d_Znew [N, 2] and d_stats[0] = stress(Z) =
 * sum_ij (|z_i - z_j| - delta_ij)^2 / 2, d_stats[1] = sum_ij |z_i - z_j|^2.  The iteration, the
 * convergence test and the restarts are driven by the host (sctda_b200.mapper.smacof_device). */
int32_t sctda_smacof_pass(sctda_ctx* ctx, int32_t N, const double* d_D, int64_t ldd, const double* d_Z,
                          double* d_Znew, double* d_stats, void* stream);

/* The PCA lens, apply_lens(df, lens='pca') (ref:745, 753; sklearn.decomposition.PCA(n_components=2)
 * .fit_transform): d_lens [N, ldl >= 2] receives the two leading principal component scores of the rows of
 * d_X [N, ldx >= G], with sklearn's sign rule (the loading of largest magnitude of each component is
 * positive).  Column means, the smaller Gram matrix of the centred table (G x G or N x N, by the FP64
 * contraction of this library) and its two leading eigenpairs by Chebyshev-filtered block subspace iteration
 * with Rayleigh-Ritz, stopped when |A q - theta q| <= tol * theta for both (tol <= 0: 1e-12; max_iter <= 0:
 * 200 outer steps).  d_info [6] (may be NULL): theta_1, theta_2 (squared singular values), the two residuals,
 * outer steps, status (1 converged, 0 iteration limit).  Synchronises the host every 4 outer steps (reads the
 * convergence flag). */
int32_t sctda_pca_lens_f64(sctda_ctx* ctx, int32_t N, int32_t G, const double* d_X, int64_t ldx, double* d_lens,
                           int64_t ldl, double* d_info, int32_t max_iter, double tol, void* stream);

/* Cover (mapper_graph stage 1).  d_lens [N, ldl>=2]: the 2-D lens.  d_lo0, d_hi0, d_lo1, d_hi1 [r]: open interval
 * bounds per axis (fenceposts widened by gain; the host computes them from percentiles).
 * Patch b = i*r + j holds the cells with lo0[i] < lens0 < hi0[i] and lo1[j] < lens1 < hi1[j].
 * count: d_bin_off [r*r + 1] exclusive offsets (last = total memberships).  fill: d_members
 * [total], ascending cell index inside each patch.  r <= 128. */
int32_t sctda_cover_count(sctda_ctx* ctx, int32_t N, const double* d_lens, int64_t ldl, int32_t r,
                          const double* d_lo0, const double* d_hi0, const double* d_lo1,
                          const double* d_hi1, int32_t* d_bin_off, void* stream);
int32_t sctda_cover_fill(sctda_ctx* ctx, int32_t N, int32_t r, const int32_t* d_bin_off,
                         int32_t* d_members, void* stream);

/* Per-patch clustering (mapper_graph stage 2): single linkage (MST) on squared Euclidean
 * distances between the rows of Xn, K in 2..K_max chosen by the Davies-Bouldin index
 * (stat 0 = 'db'), K_max = 2 if m <= 5 else min(m/2, max_K); a one-cell patch is one node.
 * stat 1 = 'histgap': the classic Mapper cut instead — tree edge lengths in 10 equal bins, every
 * edge above the first empty bin is removed (any number of clusters; d_bin_db stays NaN).
 * d_labels [total] cluster of each membership (numbered by smallest member), d_bin_K [nbins],
 * d_bin_db [nbins, 8] the index of K = 2, 3, ... (NaN where not evaluated; may be NULL).
 * Synchronises the host once (patch sizes: batching of the Gram tiles to the free memory,
 * largest patches first).  max_K <= 8. */
int32_t sctda_bin_cluster(sctda_ctx* ctx, int32_t G, const double* d_Xn, int64_t ldxn, int32_t nbins,
                          const int32_t* d_bin_off, const int32_t* d_members, int32_t max_K, int32_t stat,
                          int32_t* d_labels, int32_t* d_bin_K, double* d_bin_db, void* stream);

/* Node lists (all_clusters of ref:772-774): nodes ordered by (patch, cluster).  count:
 * d_node_base [nbins+1] (last = V).  fill: d_node_off [V+1], d_node_members [total] ascending,
 * d_node_bin [V]. */
int32_t sctda_nodes_count(sctda_ctx* ctx, int32_t nbins, const int32_t* d_bin_K, int32_t* d_node_base,
                          void* stream);
int32_t sctda_nodes_fill(sctda_ctx* ctx, int32_t nbins, const int32_t* d_bin_off, const int32_t* d_members,
                         const int32_t* d_labels, const int32_t* d_node_base, int32_t V,
                         int32_t* d_node_off, int32_t* d_node_members, int32_t* d_node_bin, void* stream);

/* Nerve (mapper_graph stage 3): edge (i < j) iff nodes i and j share a cell.  count: *d_edge_count
 * (int64 on the device).  fill: d_edges [E, 2] int32, sorted lexicographically.  V <= 131,072. */
int32_t sctda_nerve_count(sctda_ctx* ctx, int32_t N, int32_t V, const int32_t* d_node_off,
                          const int32_t* d_node_members, int64_t Mtot, int64_t* d_edge_count, void* stream);
int32_t sctda_nerve_fill(sctda_ctx* ctx, int32_t V, int32_t* d_edges, void* stream);

/* ---- multi-GPU glue: one process per GPU, communicator created by the host -------- */
/*
 * nccl_comm is an ncclComm_t (passed as void*) the HOST created with ncclCommInitRank, one rank per GPU; this
 * library never initialises NCCL and resolves ncclAllGather at run time from the libnccl.so.2 loaded in the
 * process.  All buffers are device buffers; nothing synchronises the host.
 *
 * sctda_csr_select: the sub-CSR of the patches d_sel [nsel] (ascending patch ids) of the cover CSR:
 *   d_sub_members[d_sub_off[i] + t] = d_members[d_bin_off[d_sel[i]] + t]  (d_sub_off [nsel+1] given by the caller).
 * sctda_allgather_csr: the ONE exchange of the patch-sharded Mapper build.  This rank clustered its patches
 *   (sctda_bin_cluster on its sub-CSR): d_frag_labels [frag_len], d_frag_K [frag_nb].  Fragments are padded to
 *   max_frag / max_nb (the maxima over the ranks), all-gathered, and every patch b of the full cover is placed:
 *   d_labels[d_bin_off[b] + t] = fragment of rank d_owner[b] at d_frag_off[b] + t, d_bin_K[b] = its K at d_frag_idx[b].
 *   With nranks == 1 the communicator may be NULL (the gather is a local copy).
 * sctda_gather_results: the final gather of the permutation test (no other collective on that path):
 *   d_all [nranks * max_n] receives every rank's d_mine [n_mine], padded to max_n, in rank order.
 */
int32_t sctda_csr_select(sctda_ctx* ctx, int32_t nsel, const int32_t* d_sel, const int32_t* d_bin_off,
                         const int32_t* d_members, const int32_t* d_sub_off, int32_t* d_sub_members, void* stream);
int32_t sctda_allgather_csr(sctda_ctx* ctx, void* nccl_comm, int32_t nranks, const int32_t* d_frag_labels,
                            int64_t frag_len, int64_t max_frag, const int32_t* d_frag_K, int32_t frag_nb,
                            int32_t max_nb, int32_t nbins, const int32_t* d_bin_off, const int32_t* d_owner,
                            const int32_t* d_frag_off, const int32_t* d_frag_idx, int32_t* d_labels,
                            int32_t* d_bin_K, void* stream);
int32_t sctda_gather_results(sctda_ctx* ctx, void* nccl_comm, int32_t nranks, const int32_t* d_mine,
                             int64_t n_mine, int64_t max_n, int32_t* d_all, void* stream);

/* ---- gene x gene matrices over the nodes ---------------------------------- */
/*
 * UnrootedGraph.JSD_matrix (ref:1111-1155): Jensen-Shannon distance between the normalised node
 * profiles d_P [g, ldp>=V] (sctda_node_stats d_p) of g genes; d_out [g, ldo>=g].
 * jsd = sqrt(0.5*(h_i + h_j - 2*sum_k f((p_ik+p_jk)/2))), f(x) = x*log2(x), f(0) = 0; no clamping, as
 * the reference.  (cor_matrix, ref:1157-1163, is sctda_pairdist_rows_f64 on the genes' table rows;
 * adjacency_matrix, ref:1165-1190, is sctda_gram_gather_f64 on [P; P*A].)
 */
int32_t sctda_jsd_matrix(sctda_ctx* ctx, int32_t g, int32_t V, const double* d_P, int64_t ldp, double* d_out,
                         int64_t ldo, void* stream);

/*
 * HOST function (no GPU work): the permutation source of ref:975-976.  Fills out[rows, S] with
 * arange(S) shuffled row after row exactly as numpy.random.shuffle does on the legacy MT19937
 * state (key624[624], *pos as returned by numpy.random.get_state()); the state is advanced in
 * place so that numpy.random.set_state continues the stream where the reference would be.
 */
int32_t sctda_mt19937_shuffle_rows(uint32_t* key624, int32_t* pos, int32_t rows, int32_t S, int32_t* out);

/*
 * HOST function (no GPU work): the numeric fields of a table body, i.e. the per-field float() loop
 * of UnrootedGraph.__init__ (ref:839-885) on all host cores.  body[nbytes]: the file after its header
 * line, rows ending in '\n' (or "\r\n"; the last row may lack it), nrows rows of ncols fields
 * separated by the byte `sep`; strip_bar != 0 keeps only the part of a field before the first '|'
 * (the csv variant, ref:851-853).  out[ncols][nrows] (gene-major) receives the value of every field
 * of the plain decimal form, 0.0 for the others; those others are reported, in no particular order,
 * as (flat field index r*ncols+c, byte offset) pairs in bad_pos[2*i], bad_pos[2*i+1] and lengths
 * bad_len[i] for i < min(*n_bad, bad_cap) (call again with a larger capacity when *n_bad > bad_cap),
 * so that the caller applies float() to them exactly as the reference does.  A row with a different
 * number of fields returns SCTDA_ERR_INVALID with *bad_row = the first such row.  nthreads <= 0: all cores.
 */
int32_t sctda_parse_table_f64(const char* body, int64_t nbytes, int32_t sep, int32_t strip_bar, int64_t nrows,
                              int64_t ncols, double* out, int64_t* bad_pos, int32_t* bad_len, int64_t bad_cap,
                              int64_t* n_bad, int64_t* bad_row, int32_t nthreads);

/*
 * HOST function (no GPU work): the numeric fields Preprocess.save writes (ref:663-728), formatted by all
 * host cores.  counts[nrows][ld] row-major (one row per CELL, ncols genes used); denom[nrows] (the cell's
 * total transcripts or the subsampling target) or NULL; log2_tpm != 0 writes
 * str(log2(1.0 + 1000000.0 * count / denom)) as Python 2 prints it ('%.12g', '.0' appended to integral
 * values), log2_tpm == 0 formats the values as they are.  Row r is written into its own slot
 * out[r*ncols*25 ...] with `sep` between fields and no line end; row_end[r] = bytes used.  cap must be at
 * least nrows*ncols*25.  nthreads <= 0: all cores.
 */
int32_t sctda_format_table_f64(const double* counts, int64_t ld, int64_t nrows, int64_t ncols,
                               const double* denom, int32_t log2_tpm, int32_t sep, char* out, int64_t cap,
                               int64_t* row_end, int32_t nthreads);

/* ---- RootedGraph root finding (first "next" row) ------------------------------- */
/*
 * RootedGraph.get_dendrite (ref:1463-1479) for all V source nodes at once: hop distances from
 * every node (BFS over the FULL symmetric adjacency CSR d_adj_off / d_adj_idx, both directions
 * stored), nodes at the maximal distance dropped, score = -spearman(x, distance) with average
 * ranks for ties (scipy.stats.spearmanr).  x is given through its sorted order, computed once by
 * the caller: d_xorder [V] nodes by ascending x, d_gstart / d_gend [V] the tie group
 * [gstart, gend) of each sorted position.  d_score [V] (NaN where undefined), d_dist [V, V]
 * int32 hop distances (-1: unreachable) or NULL.  V <= 17,000.
 */
int32_t sctda_root_scores(sctda_ctx* ctx, int32_t V, const int32_t* d_adj_off, const int32_t* d_adj_idx,
                          const int32_t* d_xorder, const int32_t* d_gstart, const int32_t* d_gend,
                          double* d_score, int32_t* d_dist, void* stream);

/*
 * Diagnostic: the permutation kernel divides with a reciprocal and one fused correction step
 * instead of the generic IEEE division; this runs n_trials pseudo-random comparisons of the
 * two on the device (numerators 0 and 2^-24..2^40, divisors 1..8192 and 0.693147) and returns
 * the number of results that differ in any bit (expected: 0).  Synchronises the host.
 */
int32_t sctda_selftest_div(sctda_ctx* ctx, int64_t n_trials, int64_t* out_mismatches);

/* Diagnostic: the permutation kernel's own log1p (positive finite arguments) against the generic
 * CUDA log1p over n_trials pseudo-random arguments in 2^-60 .. 2^40: returns the largest
 * difference in units of the last place (both are sub-ulp accurate: expected <= 2).
 * Synchronises the host. */
int32_t sctda_selftest_log1p(sctda_ctx* ctx, int64_t n_trials, int64_t* out_max_ulp);

#ifdef __cplusplus
}
#endif
#endif /* SCTDA_B200_H */
