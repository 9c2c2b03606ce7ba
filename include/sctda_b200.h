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

/* Measurement aid (bench.py "roofline"): when enabled, every hot-path call brackets its
 * DOMINANT kernel (conn_perm_kernel for sctda_conn_perm_test) with CUDA events on the
 * caller's stream.  sctda_last_kernel_ms waits for that kernel to finish (host sync on
 * the event) and returns its duration in milliseconds. */
int32_t sctda_set_kernel_timing(sctda_ctx* ctx, int32_t enabled);
int32_t sctda_last_kernel_ms(sctda_ctx* ctx, float* out_ms);

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
    int32_t        reserved;
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
 *                     permuted score is also counted in d_ties (<= 0: 1e-9)
 * outputs:
 *   d_exceed  [G] int32  #{r : conn_r > observed}   (p-value = exceed / n), exact integers
 *   d_ties    [G] int32  #{r : |conn_r - observed| <= tie_rel*|observed|} or NULL
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
 * Diagnostic: the permutation kernel divides with a reciprocal and one fused correction step
 * instead of the generic IEEE division; this runs n_trials pseudo-random comparisons of the
 * two on the device (numerators 0 and 2^-24..2^40, divisors 1..8192 and 0.693147) and returns
 * the number of results that differ in any bit (expected: 0).  Synchronises the host.
 */
int32_t sctda_selftest_div(sctda_ctx* ctx, int64_t n_trials, int64_t* out_mismatches);

#ifdef __cplusplus
}
#endif
#endif /* SCTDA_B200_H */
