// Hot path (a), stages 2-4: cover assignment, per-patch single-linkage clustering with
// Davies-Bouldin model selection, node lists and the nerve.
//
// Reference seam: sakmapper.mapper_graph(df, lens_data, resolution, gain, equalize, clust, stat,
// max_K) -> (G, all_clusters, patches), called at /root/reference/scTDA/main.py:768-771
// (SURVEY.md §8a rows a4-a6).  sakmapper is an un-vendored third-party package; the algorithm
// is the one restated in oracle/mapper_oracle.py (SURVEY.md appendix C.3 pins the open choices).
//
// Everything here is integer work or float64 work in a FIXED operation order (explicit
// __dadd_rn / __fma_rn, sequential sums in ascending index order), so that patch memberships,
// cluster labels and edges are reproducible bit for bit by the scalar oracle:
//   * cover: per-cell interval bit masks, then one warp per patch scans the cells in order with
//     ballot compaction (members come out ascending, no atomics, no sort);
//   * clustering: one CTA per patch; Prim's MST over d2(p,q) = max(0,(G_pp+G_qq)-2G_pq) read
//     row by row from the patch Gram tile (dist.cu), keys in shared memory, argmin by warp
//     shuffles with (key, index) lexicographic ties; cut = the K-1 heaviest edges; K chosen by
//     the Davies-Bouldin index evaluated from the Gram tile through the column sums of the
//     finest partition;
//   * nerve: cell -> node incidence, pairs OR-ed into a V x V adjacency bitset, rows compacted
//     with popc prefix sums into the lexicographically sorted edge list.
#include <math.h>

#include "common.cuh"

int sctda_internal_gram_bins(sctda_ctx* ctx, int G, const double* d_Xn, int64_t ldxn, int nbins,
                             const int32_t* d_bin_off, const int32_t* d_members, const int64_t* d_tile_prefix,
                             const int64_t* d_gram_off, double* d_gram, cudaStream_t st);

namespace {

constexpr int kMaxK = 8;            // largest max_K served
constexpr int kMaxRes = 128;        // largest resolution served (two 64-bit mask words per axis)

// ---------------------------------------------------------------------------------------------
// exclusive scan, one CTA: out[0] = 0, out[i+1] = out[i] + f(in[i]); optional max of in[]
// ---------------------------------------------------------------------------------------------
enum { MAP_ID = 0, MAP_TILES = 1, MAP_SQUARE = 2 };

template <typename TOut>
__global__ void __launch_bounds__(1024) scan_kernel(const int32_t* __restrict__ in, int64_t n, int map,
                                                    TOut* __restrict__ out, int32_t* __restrict__ out_max) {
    __shared__ long long s_sum[1024];
    __shared__ int s_max[1024];
    const int t = threadIdx.x;
    const int64_t per = (n + 1023) / 1024;
    const int64_t lo = (int64_t)t * per, hi = lo + per < n ? lo + per : n;
    auto f = [map](int32_t v) -> long long {
        if (map == MAP_TILES) { long long q = (v + 127) / 128; return q * (q + 1) / 2; }   // upper-triangular tiles
        if (map == MAP_SQUARE) return (long long)v * v;
        return v;
    };
    long long s = 0;
    int mx = 0;
    for (int64_t i = lo; i < hi; ++i) { s += f(in[i]); mx = max(mx, in[i]); }
    s_sum[t] = s;
    s_max[t] = mx;
    __syncthreads();
    if (t == 0) {
        long long run = 0;
        int m2 = 0;
        for (int i = 0; i < 1024; ++i) { long long v = s_sum[i]; s_sum[i] = run; run += v; m2 = max(m2, s_max[i]); }
        out[n] = (TOut)run;
        if (out_max) *out_max = m2;
    }
    __syncthreads();
    long long run = s_sum[t];
    for (int64_t i = lo; i < hi; ++i) { out[i] = (TOut)run; run += f(in[i]); }
}

__global__ void bin_size_kernel(int n, const int32_t* __restrict__ off, int32_t* __restrict__ c) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n) c[b] = off[b + 1] - off[b];
}

// ---------------------------------------------------------------------------------------------
// cover
// ---------------------------------------------------------------------------------------------
__global__ void cover_mask_kernel(const double* __restrict__ lens, int64_t ldl, int N, int r,
                                  const double* __restrict__ lo0, const double* __restrict__ hi0,
                                  const double* __restrict__ lo1, const double* __restrict__ hi1,
                                  unsigned long long* __restrict__ masks) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    const double x = lens[(int64_t)c * ldl], y = lens[(int64_t)c * ldl + 1];
    unsigned long long m[4] = {0, 0, 0, 0};
    for (int k = 0; k < r; ++k) {
        if (x > lo0[k] && x < hi0[k]) m[k >> 6] |= 1ull << (k & 63);          // strict, as the oracle
        if (y > lo1[k] && y < hi1[k]) m[2 + (k >> 6)] |= 1ull << (k & 63);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) masks[(int64_t)c * 4 + i] = m[i];
}

// one warp per patch b = i*r + j: FILL = false counts, FILL = true writes the ascending members
template <bool FILL>
__global__ void __launch_bounds__(256) cover_scan_kernel(const unsigned long long* __restrict__ masks, int N, int r,
                                                         int32_t* __restrict__ counts,
                                                         const int32_t* __restrict__ bin_off,
                                                         int32_t* __restrict__ members) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= r * r) return;
    const int i = b / r, j = b % r;
    const int wi = i >> 6, wj = 2 + (j >> 6);
    const unsigned long long bi = 1ull << (i & 63), bj = 1ull << (j & 63);
    int total = 0;
    const int base = FILL ? bin_off[b] : 0;
    for (int c0 = 0; c0 < N; c0 += 32) {
        const int c = c0 + lane;
        bool in = false;
        if (c < N) in = (masks[(int64_t)c * 4 + wi] & bi) && (masks[(int64_t)c * 4 + wj] & bj);
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        if (FILL && in) members[base + total + __popc(bal & ((1u << lane) - 1))] = c;
        total += __popc(bal);
    }
    if (!FILL && lane == 0) counts[b] = total;
}

// ---------------------------------------------------------------------------------------------
// Prim's MST per patch
// ---------------------------------------------------------------------------------------------
struct ClusterArgs {
    int nbins, max_K;
    const int32_t* bin_off;
    const int64_t* gram_off;
    const double* gram;
    int32_t* mst_parent;      // [Mtot] parent of local vertex u in the tree (-1 for the root)
    double* mst_w;            // [Mtot] weight of that edge
    int32_t* mst_order;       // [Mtot] insertion step of vertex u (-1 for the root)
    int32_t* mst_seq;         // [Mtot] vertex inserted at step t (seq[0] = 0)
    double* T;                // [Mtot][kMaxK] column sums of the finest partition
    double* TK;               // [Mtot][kMaxK] column sums of the level under evaluation
    double* dist;             // [Mtot] scratch
    int32_t* labels;          // out [Mtot]
    int32_t* bin_K;           // out [nbins]
    double* bin_db;           // out [nbins][kMaxK] Davies-Bouldin score of K = 2.. (NaN where not evaluated)
};

__device__ __forceinline__ double dist2(double gpp_u, double gpp_q, double g) {
    return fmax(fma(-2.0, g, __dadd_rn(gpp_u, gpp_q)), 0.0);
}

__global__ void __launch_bounds__(256) prim_kernel(ClusterArgs a) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ double s_rk[8];
    __shared__ int s_ri[8];
    __shared__ int s_u;
    const int b = blockIdx.x;
    const int off = a.bin_off[b], m = a.bin_off[b + 1] - off;
    if (m == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* key = (double*)s_raw;                 // >= 0: best known distance to the tree; -1: in the tree
    double* gpp = key + m;
    int* par = (int*)(gpp + m);
    const double* Gm = a.gram + a.gram_off[b];
    for (int q = tid; q < m; q += 256) {
        key[q] = INFINITY;
        gpp[q] = Gm[(int64_t)q * m + q];
        par[q] = -1;
    }
    if (tid == 0) {
        a.mst_parent[off] = -1; a.mst_w[off] = 0.0; a.mst_order[off] = -1; a.mst_seq[off] = 0;
    }
    __syncthreads();
    if (tid == 0) key[0] = -1.0;
    int u = 0;
    for (int t = 0; t < m - 1; ++t) {
        __syncthreads();
        const double gu = gpp[u];
        const double* row = Gm + (int64_t)u * m;
        double bk = INFINITY;
        int bi = 0x7fffffff;
        for (int q = tid; q < m; q += 256) {
            double k = key[q];
            if (k >= 0.0) {
                double d = dist2(gu, gpp[q], row[q]);
                if (d < k) { k = d; key[q] = d; par[q] = u; }
                if (k < bk || (k == bk && q < bi)) { bk = k; bi = q; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ok = __shfl_xor_sync(0xffffffffu, bk, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ok < bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; }
        }
        if (lane == 0) { s_rk[warp] = bk; s_ri[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < 8; ++w)
                if (s_rk[w] < bk || (s_rk[w] == bk && s_ri[w] < bi)) { bk = s_rk[w]; bi = s_ri[w]; }
            s_u = bi;
            a.mst_parent[off + bi] = par[bi];
            a.mst_w[off + bi] = bk;
            a.mst_order[off + bi] = t;
            a.mst_seq[off + t + 1] = bi;
            key[bi] = -1.0;
        }
        __syncthreads();
        u = s_u;
    }
}

// ---------------------------------------------------------------------------------------------
// cut + Davies-Bouldin model selection per patch
// ---------------------------------------------------------------------------------------------
// Sum over p ascending of (lab[p] == I ? v[p*stride] : 0.0): the masked terms add +0.0, an exact
// no-op, so this is the sequential sum over the members of cluster I; the loads do not depend on
// the running sum and are issued eight at a time.
__device__ __forceinline__ double masked_seq_sum(const double* __restrict__ v, int stride, const int8_t* lab, int I, int m) {
    double acc = 0.0;
    int p = 0;
    for (; p + 8 <= m; p += 8) {
        double x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = (lab[p + u] == I) ? v[(int64_t)(p + u) * stride] : 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) acc = __dadd_rn(acc, x[u]);
    }
    for (; p < m; ++p) acc = __dadd_rn(acc, (lab[p] == I) ? v[(int64_t)p * stride] : 0.0);
    return acc;
}

__global__ void __launch_bounds__(256) select_kernel(ClusterArgs a) {
    extern __shared__ __align__(16) unsigned char s_dyn[];      // labels of every level: int8 [nlev][m]
    __shared__ double s_rk[8];
    __shared__ int s_ri[8], s_ro[8];
    __shared__ int s_heavy[kMaxK];               // vertices whose edge is the h-th heaviest
    __shared__ int s_cof[kMaxK];                 // fine cluster -> coarse cluster of the level under evaluation
    __shared__ int s_fine_rep[kMaxK];
    __shared__ double s_n[kMaxK];
    __shared__ double s_cc[kMaxK][kMaxK];
    __shared__ double s_S[kMaxK];
    __shared__ double s_db[kMaxK];
    const int b = blockIdx.x;
    const int off = a.bin_off[b], m = a.bin_off[b + 1] - off;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < kMaxK) a.bin_db[(int64_t)b * kMaxK + tid] = NAN;
    if (m == 0) { if (tid == 0) a.bin_K[b] = 0; return; }
    if (m == 1) { if (tid == 0) { a.bin_K[b] = 1; a.labels[off] = 0; } return; }
    const double* Gm = a.gram + a.gram_off[b];
    const int kmax = m <= 5 ? 2 : min(m / 2, a.max_K);
    const int nlev = kmax - 1;                   // level L (0-based) = L+1 edges removed = L+2 clusters
    const double* w = a.mst_w + off;
    const int32_t* ord = a.mst_order + off;
    const int32_t* par = a.mst_parent + off;
    const int32_t* seq = a.mst_seq + off;
    int8_t* lev = (int8_t*)s_dyn;
    // (c) the nlev heaviest tree edges: (weight desc, insertion order desc)
    for (int h = 0; h < nlev; ++h) {
        double bk = -1.0;
        int bo = -1, bi = -1;
        for (int q = tid; q < m; q += 256) {
            int o = ord[q];
            if (o < 0) continue;
            bool taken = false;
            for (int g = 0; g < h; ++g) taken |= (s_heavy[g] == q);
            if (taken) continue;
            double k = w[q];
            if (k > bk || (k == bk && o > bo)) { bk = k; bo = o; bi = q; }
        }
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) {
            double ok = __shfl_xor_sync(0xffffffffu, bk, o2);
            int oo = __shfl_xor_sync(0xffffffffu, bo, o2);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o2);
            if (ok > bk || (ok == bk && oo > bo)) { bk = ok; bo = oo; bi = oi; }
        }
        if (lane == 0) { s_rk[warp] = bk; s_ro[warp] = bo; s_ri[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int x = 1; x < 8; ++x)
                if (s_rk[x] > bk || (s_rk[x] == bk && s_ro[x] > bo)) { bk = s_rk[x]; bo = s_ro[x]; bi = s_ri[x]; }
            s_heavy[h] = bi;
        }
        __syncthreads();
    }
    // (d) labels of every level: thread L walks the vertices in insertion order; a vertex whose edge
    //     is among the L+1 heaviest opens a component, the others inherit their parent's (inserted
    //     earlier).  Components are then renumbered by their smallest vertex.
    if (tid < nlev) {
        const int L = tid;
        int8_t* lab = lev + (size_t)L * m;
        int ncomp = 1;
        int cmin[kMaxK];
        cmin[0] = 0;
        lab[0] = 0;
        for (int t = 1; t < m; ++t) {
            const int v = seq[t];
            bool cut = false;
            for (int g = 0; g <= L; ++g) cut |= (s_heavy[g] == v);
            int c;
            if (cut) { c = ncomp; cmin[ncomp++] = v; }
            else { c = lab[par[v]]; if (v < cmin[c]) cmin[c] = v; }
            lab[v] = (int8_t)c;
        }
        int rank[kMaxK];
        for (int c = 0; c < ncomp; ++c) {
            int rk = 0;
            for (int d = 0; d < ncomp; ++d) rk += (cmin[d] < cmin[c]);
            rank[c] = rk;
        }
        for (int v = 0; v < m; ++v) lab[v] = (int8_t)rank[lab[v]];
        if (L == nlev - 1)
            for (int c = 0; c < ncomp; ++c) s_fine_rep[rank[c]] = cmin[c];
    }
    __syncthreads();
    const int F = kmax;                                        // clusters of the finest level
    const int8_t* fine = lev + (size_t)(nlev - 1) * m;
    double* T = a.T + (int64_t)off * kMaxK;                    // [m][kMaxK]
    double* TK = a.TK + (int64_t)off * kMaxK;                  // [m][kMaxK], per level
    double* dist = a.dist + off;
    // (e) T[p][f] = sum over q ascending of G[q][p] for q in fine cluster f
    for (int p = tid; p < m; p += 256) {
        double acc[kMaxK];
#pragma unroll
        for (int f = 0; f < kMaxK; ++f) acc[f] = 0.0;
        int q = 0;
        for (; q + 4 <= m; q += 4) {
            double g[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) g[u] = Gm[(int64_t)(q + u) * m + p];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int fq = fine[q + u];                    // uniform across the CTA
#pragma unroll
                for (int f = 0; f < kMaxK; ++f)
                    if (fq == f) acc[f] = __dadd_rn(acc[f], g[u]);
            }
        }
        for (; q < m; ++q) {
            const double g = Gm[(int64_t)q * m + p];
            const int fq = fine[q];
#pragma unroll
            for (int f = 0; f < kMaxK; ++f)
                if (fq == f) acc[f] = __dadd_rn(acc[f], g);
        }
#pragma unroll
        for (int f = 0; f < kMaxK; ++f) T[(int64_t)p * kMaxK + f] = acc[f];
    }
    __syncthreads();
    // (f) Davies-Bouldin index of every level
    for (int L = 0; L < nlev; ++L) {
        const int K = L + 2;
        const int8_t* lab = lev + (size_t)L * m;
        if (tid < F) s_cof[tid] = lab[s_fine_rep[tid]];
        __syncthreads();
        for (int p = tid; p < m; p += 256) {                    // TK[p][J] = sum_{f in J, ascending} T[p][f]
            double tk[kMaxK];
#pragma unroll
            for (int J = 0; J < kMaxK; ++J) tk[J] = 0.0;
            for (int f = 0; f < F; ++f) {
                const double t = T[(int64_t)p * kMaxK + f];
                const int J = s_cof[f];
#pragma unroll
                for (int j2 = 0; j2 < kMaxK; ++j2)
                    if (J == j2) tk[j2] = __dadd_rn(tk[j2], t);
            }
#pragma unroll
            for (int J = 0; J < kMaxK; ++J) TK[(int64_t)p * kMaxK + J] = tk[J];
        }
        if (tid < K) {                                          // cluster sizes
            int n = 0;
            for (int p = 0; p < m; ++p) n += (lab[p] == tid);
            s_n[tid] = (double)n;
        }
        __syncthreads();
        if (tid < K * K) {                                      // c_I . c_J, one thread per pair
            const int I = tid / K, J = tid % K;
            const double acc = masked_seq_sum(TK + J, kMaxK, lab, I, m);
            s_cc[I][J] = __ddiv_rn(acc, __dmul_rn(s_n[I], s_n[J]));
        }
        __syncthreads();
        for (int p = tid; p < m; p += 256) {                    // |x_p - c_I|
            const int I = lab[p];
            const double xc = __ddiv_rn(TK[(int64_t)p * kMaxK + I], s_n[I]);
            const double d2 = fmax(__dadd_rn(fma(-2.0, xc, Gm[(int64_t)p * m + p]), s_cc[I][I]), 0.0);
            dist[p] = sqrt(d2);
        }
        __syncthreads();
        if (tid < K) s_S[tid] = __ddiv_rn(masked_seq_sum(dist, 1, lab, tid, m), s_n[tid]);
        __syncthreads();
        if (tid == 0) {
            bool any = false;
            for (int i = 0; i < K; ++i) any |= (s_S[i] > 0.0);
            double db = 0.0;
            if (any) {
                double tot = 0.0;
                for (int i = 0; i < K; ++i) {
                    double best = 0.0;
                    for (int j = 0; j < K; ++j) {
                        if (j == i) continue;
                        const double M2 = fmax(__dadd_rn(fma(-2.0, s_cc[i][j], s_cc[i][i]), s_cc[j][j]), 0.0);
                        const double Mij = sqrt(M2);
                        const double R = Mij > 0.0 ? __ddiv_rn(__dadd_rn(s_S[i], s_S[j]), Mij) : 0.0;
                        if (R > best) best = R;
                    }
                    tot = __dadd_rn(tot, best);
                }
                db = __ddiv_rn(tot, (double)K);
            }
            s_db[L] = db;
            a.bin_db[(int64_t)b * kMaxK + L] = db;
        }
        __syncthreads();
    }
    // (g) first minimum
    int bestL = 0;
    for (int L = 1; L < nlev; ++L)
        if (s_db[L] < s_db[bestL]) bestL = L;
    const int8_t* lab = lev + (size_t)bestL * m;
    for (int p = tid; p < m; p += 256) a.labels[off + p] = lab[p];
    if (tid == 0) a.bin_K[b] = bestL + 2;
}

// ---------------------------------------------------------------------------------------------
// nodes
// ---------------------------------------------------------------------------------------------
__global__ void node_bin_kernel(int nbins, const int32_t* __restrict__ node_base, int32_t* __restrict__ node_bin,
                                int32_t* __restrict__ node_label) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbins) return;
    for (int v = node_base[b]; v < node_base[b + 1]; ++v) { node_bin[v] = b; node_label[v] = v - node_base[b]; }
}

template <bool FILL>
__global__ void __launch_bounds__(256) node_scan_kernel(int V, const int32_t* __restrict__ node_bin,
                                                        const int32_t* __restrict__ node_label,
                                                        const int32_t* __restrict__ bin_off,
                                                        const int32_t* __restrict__ members,
                                                        const int32_t* __restrict__ labels, int32_t* __restrict__ sizes,
                                                        const int32_t* __restrict__ node_off,
                                                        int32_t* __restrict__ node_members) {
    const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (v >= V) return;
    const int b = node_bin[v], c = node_label[v];
    const int lo = bin_off[b], hi = bin_off[b + 1];
    const int base = FILL ? node_off[v] : 0;
    int total = 0;
    for (int p0 = lo; p0 < hi; p0 += 32) {
        const int p = p0 + lane;
        const bool in = p < hi && labels[p] == c;
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        if (FILL && in) node_members[base + total + __popc(bal & ((1u << lane) - 1))] = members[p];
        total += __popc(bal);
    }
    if (!FILL && lane == 0) sizes[v] = total;
}

// ---------------------------------------------------------------------------------------------
// nerve
// ---------------------------------------------------------------------------------------------
__global__ void cell_count_kernel(int64_t Mtot, const int32_t* __restrict__ node_members, int32_t* __restrict__ cnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Mtot) atomicAdd(cnt + node_members[i], 1);
}

__global__ void cell_fill_kernel(int V, const int32_t* __restrict__ node_off, const int32_t* __restrict__ node_members,
                                 const int32_t* __restrict__ cell_off, int32_t* __restrict__ cursor,
                                 int32_t* __restrict__ cell_nodes) {
    const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (v >= V) return;
    for (int i = node_off[v] + lane; i < node_off[v + 1]; i += 32) {
        const int c = node_members[i];
        cell_nodes[cell_off[c] + atomicAdd(cursor + c, 1)] = v;
    }
}

__global__ void pair_kernel(int N, int W, const int32_t* __restrict__ cell_off, const int32_t* __restrict__ cell_nodes,
                            unsigned* __restrict__ bits) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    const int lo = cell_off[c], hi = cell_off[c + 1];
    for (int x = lo; x < hi; ++x)
        for (int y = x + 1; y < hi; ++y) {
            int i = cell_nodes[x], j = cell_nodes[y];
            if (i == j) continue;
            if (i > j) { int t = i; i = j; j = t; }
            atomicOr(bits + (int64_t)i * W + (j >> 5), 1u << (j & 31));
        }
}

template <bool FILL>
__global__ void __launch_bounds__(256) edge_scan_kernel(int V, int W, const unsigned* __restrict__ bits,
                                                        int32_t* __restrict__ row_cnt,
                                                        const int64_t* __restrict__ row_off,
                                                        int32_t* __restrict__ edges) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= V) return;
    const int64_t base = FILL ? row_off[i] : 0;
    int total = 0;
    for (int w0 = 0; w0 < W; w0 += 32) {
        const int wd = w0 + lane;
        const unsigned word = wd < W ? bits[(int64_t)i * W + wd] : 0u;
        const int pc = __popc(word);
        int incl = pc;                                          // inclusive prefix over the lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (FILL) {
            int64_t pos = base + total + (incl - pc);
            unsigned x = word;
            while (x) {
                const int bit = __ffs(x) - 1;
                x &= x - 1;
                edges[2 * pos] = i;
                edges[2 * pos + 1] = wd * 32 + bit;
                ++pos;
            }
        }
        total += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (!FILL && lane == 0) row_cnt[i] = total;
}

}  // namespace

// =============================================================================================
// C-ABI
// =============================================================================================
extern "C" int32_t sctda_cover_count(sctda_ctx* ctx, int32_t N, const double* d_lens, int64_t ldl, int32_t r,
                                     const double* d_lo0, const double* d_hi0, const double* d_lo1,
                                     const double* d_hi1, int32_t* d_bin_off, void* stream) {
    if (!ctx) return SCTDA_ERR_INVALID;
    if (N < 0 || r < 1 || !d_lens || ldl < 2 || !d_lo0 || !d_hi0 || !d_lo1 || !d_hi1 || !d_bin_off)
        return sctda_fail(ctx, SCTDA_ERR_INVALID, "cover_count: bad arguments%s");
    if (r > kMaxRes) return sctda_fail(ctx, SCTDA_ERR_UNSUPPORTED, "cover: resolution above 128 is not served%s");
    cudaStream_t st = (cudaStream_t)stream;
    SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
    const int B = r * r;
    unsigned long long* masks = (unsigned long long*)sctda_ws(ctx, WS_COVER, (size_t)(N > 0 ? N : 1) * 32 + (size_t)B * 4);
    if (!masks) return SCTDA_ERR_NOMEM;
    int32_t* counts = (int32_t*)(masks + (size_t)(N > 0 ? N : 1) * 4);
    if (N > 0) {
        cover_mask_kernel<<<(N + 255) / 256, 256, 0, st>>>(d_lens, ldl, N, r, d_lo0, d_hi0, d_lo1, d_hi1, masks);
        SCTDA_LAUNCH_CHECK(ctx, "cover_mask_kernel");
    }
    cover_scan_kernel<false><<<(unsigned)cdiv64((int64_t)B * 32, 256), 256, 0, st>>>(masks, N, r, counts, nullptr, nullptr);
    SCTDA_LAUNCH_CHECK(ctx, "cover_scan_kernel<count>");
    scan_kernel<int32_t><<<1, 1024, 0, st>>>(counts, B, MAP_ID, d_bin_off, nullptr);
    SCTDA_LAUNCH_CHECK(ctx, "scan_kernel");
    return SCTDA_OK;
}

extern "C" int32_t sctda_cover_fill(sctda_ctx* ctx, int32_t N, int32_t r, const int32_t* d_bin_off,
                                    int32_t* d_members, void* stream) {
    if (!ctx) return SCTDA_ERR_INVALID;
    if (N < 0 || r < 1 || r > kMaxRes || !d_bin_off || !d_members)
        return sctda_fail(ctx, SCTDA_ERR_INVALID, "cover_fill: bad arguments%s");
    if (ctx->slot_bytes[WS_COVER] < (size_t)(N > 0 ? N : 1) * 32)
        return sctda_fail(ctx, SCTDA_ERR_INVALID, "cover_fill: call sctda_cover_count with the same N first%s");
    SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
    const unsigned long long* masks = (const unsigned long long*)ctx->slot_ptr[WS_COVER];
    cover_scan_kernel<true><<<(unsigned)cdiv64((int64_t)r * r * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        masks, N, r, nullptr, d_bin_off, d_members);
    SCTDA_LAUNCH_CHECK(ctx, "cover_scan_kernel<fill>");
    return SCTDA_OK;
}

extern "C" int32_t sctda_bin_cluster(sctda_ctx* ctx, int32_t G, const double* d_Xn, int64_t ldxn, int32_t nbins,
                                     const int32_t* d_bin_off, const int32_t* d_members, int32_t max_K, int32_t stat,
                                     int32_t* d_labels, int32_t* d_bin_K, double* d_bin_db, void* stream) {
    if (!ctx) return SCTDA_ERR_INVALID;
    if (G < 1 || nbins < 0 || !d_Xn || ldxn < G || !d_bin_off || !d_members || !d_labels || !d_bin_K)
        return sctda_fail(ctx, SCTDA_ERR_INVALID, "bin_cluster: bad arguments%s");
    if (max_K < 2 || max_K > kMaxK) return sctda_fail(ctx, SCTDA_ERR_UNSUPPORTED, "bin_cluster: max_K must be in 2..8%s");
    if (stat != 0) return sctda_fail(ctx, SCTDA_ERR_UNSUPPORTED, "bin_cluster: only statistics='db' (0) is served%s");
    if ((ldxn & 1) || ((uintptr_t)d_Xn & 15))
        return sctda_fail(ctx, SCTDA_ERR_UNSUPPORTED, "bin_cluster: rows must be 16-byte aligned (even ldxn)%s");
    if (nbins == 0) return SCTDA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
    // sizes: tile prefix, Gram offsets, largest patch  (one small host sync to size the workspace)
    int64_t* pre = (int64_t*)sctda_ws(ctx, WS_CLUSTER, (size_t)(nbins + 1) * 16 + 64);
    if (!pre) return SCTDA_ERR_NOMEM;
    int64_t* tile_prefix = pre;
    int64_t* gram_off = pre + (nbins + 1);
    int32_t* d_max = (int32_t*)(gram_off + (nbins + 1));
    // bin sizes as an array: reuse d_bin_K as scratch for the counts
    bin_size_kernel<<<(nbins + 255) / 256, 256, 0, st>>>(nbins, d_bin_off, d_bin_K);
    SCTDA_LAUNCH_CHECK(ctx, "bin_size_kernel");
    scan_kernel<int64_t><<<1, 1024, 0, st>>>(d_bin_K, nbins, MAP_TILES, tile_prefix, d_max);
    SCTDA_LAUNCH_CHECK(ctx, "scan_kernel<tiles>");
    scan_kernel<int64_t><<<1, 1024, 0, st>>>(d_bin_K, nbins, MAP_SQUARE, gram_off, nullptr);
    SCTDA_LAUNCH_CHECK(ctx, "scan_kernel<square>");
    int64_t h_gram = 0;
    int32_t h_max = 0, h_mtot = 0;
    SCTDA_CUDA(ctx, cudaMemcpyAsync(&h_gram, gram_off + nbins, 8, cudaMemcpyDeviceToHost, st));
    SCTDA_CUDA(ctx, cudaMemcpyAsync(&h_max, d_max, 4, cudaMemcpyDeviceToHost, st));
    SCTDA_CUDA(ctx, cudaMemcpyAsync(&h_mtot, d_bin_off + nbins, 4, cudaMemcpyDeviceToHost, st));
    SCTDA_CUDA(ctx, cudaStreamSynchronize(st));
    if (h_mtot == 0) {
        SCTDA_CUDA(ctx, cudaMemsetAsync(d_bin_K, 0, (size_t)nbins * 4, st));
        return SCTDA_OK;
    }
    const size_t prim_smem = (size_t)h_max * 20 + 16;
    if (prim_smem > 200 * 1024)
        return sctda_fail(ctx, SCTDA_ERR_UNSUPPORTED, "bin_cluster: a patch with more than 10,000 cells is not served%s");
    double* gram = (double*)sctda_ws(ctx, WS_CLUSTER2, (size_t)h_gram * 8);
    if (!gram) return SCTDA_ERR_NOMEM;
    const size_t M = (size_t)h_mtot;
    const size_t aux_bytes = M * (4 + 8 + 4 + 4 + 16 * kMaxK + 8) + 256;
    unsigned char* aux = (unsigned char*)sctda_ws(ctx, WS_CLUSTER3, aux_bytes);
    if (!aux) return SCTDA_ERR_NOMEM;
    ClusterArgs a;
    a.nbins = nbins; a.max_K = max_K; a.bin_off = d_bin_off; a.gram_off = gram_off; a.gram = gram;
    a.T = (double*)aux;
    a.TK = a.T + M * kMaxK;
    a.mst_w = a.TK + M * kMaxK;
    a.dist = a.mst_w + M;
    a.mst_parent = (int32_t*)(a.dist + M);
    a.mst_order = a.mst_parent + M;
    a.mst_seq = a.mst_order + M;
    a.labels = d_labels; a.bin_K = d_bin_K;
    double* db = d_bin_db;
    if (!db) {
        db = (double*)sctda_ws(ctx, WS_MISC, (size_t)nbins * kMaxK * 8);
        if (!db) return SCTDA_ERR_NOMEM;
    }
    a.bin_db = db;
    int rc = sctda_internal_gram_bins(ctx, G, d_Xn, ldxn, nbins, d_bin_off, d_members, tile_prefix, gram_off, gram, st);
    if (rc) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        SCTDA_CUDA(ctx, cudaFuncSetAttribute(prim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        SCTDA_CUDA(ctx, cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
        attr_set = true;
    }
    prim_kernel<<<nbins, 256, prim_smem, st>>>(a);
    SCTDA_LAUNCH_CHECK(ctx, "prim_kernel");
    select_kernel<<<nbins, 256, (size_t)h_max * (kMaxK - 1) + 16, st>>>(a);
    SCTDA_LAUNCH_CHECK(ctx, "select_kernel");
    return SCTDA_OK;
}

extern "C" int32_t sctda_nodes_count(sctda_ctx* ctx, int32_t nbins, const int32_t* d_bin_K, int32_t* d_node_base,
                                     void* stream) {
    if (!ctx) return SCTDA_ERR_INVALID;
    if (nbins < 0 || !d_bin_K || !d_node_base) return sctda_fail(ctx, SCTDA_ERR_INVALID, "nodes_count: bad arguments%s");
    SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
    scan_kernel<int32_t><<<1, 1024, 0, (cudaStream_t)stream>>>(d_bin_K, nbins, MAP_ID, d_node_base, nullptr);
    SCTDA_LAUNCH_CHECK(ctx, "scan_kernel<nodes>");
    return SCTDA_OK;
}

extern "C" int32_t sctda_nodes_fill(sctda_ctx* ctx, int32_t nbins, const int32_t* d_bin_off, const int32_t* d_members,
                                    const int32_t* d_labels, const int32_t* d_node_base, int32_t V,
                                    int32_t* d_node_off, int32_t* d_node_members, int32_t* d_node_bin, void* stream) {
    if (!ctx) return SCTDA_ERR_INVALID;
    if (nbins < 0 || V < 0 || !d_bin_off || !d_members || !d_labels || !d_node_base || !d_node_off || !d_node_members ||
        !d_node_bin)
        return sctda_fail(ctx, SCTDA_ERR_INVALID, "nodes_fill: bad arguments%s");
    cudaStream_t st = (cudaStream_t)stream;
    SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
    if (V == 0) { SCTDA_CUDA(ctx, cudaMemsetAsync(d_node_off, 0, 4, st)); return SCTDA_OK; }
    int32_t* tmp = (int32_t*)sctda_ws(ctx, WS_NODES, (size_t)V * 8);
    if (!tmp) return SCTDA_ERR_NOMEM;
    int32_t* node_label = tmp;
    int32_t* sizes = tmp + V;
    node_bin_kernel<<<(nbins + 255) / 256, 256, 0, st>>>(nbins, d_node_base, d_node_bin, node_label);
    SCTDA_LAUNCH_CHECK(ctx, "node_bin_kernel");
    const unsigned blocks = (unsigned)cdiv64((int64_t)V * 32, 256);
    node_scan_kernel<false><<<blocks, 256, 0, st>>>(V, d_node_bin, node_label, d_bin_off, d_members, d_labels, sizes,
                                                    nullptr, nullptr);
    SCTDA_LAUNCH_CHECK(ctx, "node_scan_kernel<count>");
    scan_kernel<int32_t><<<1, 1024, 0, st>>>(sizes, V, MAP_ID, d_node_off, nullptr);
    SCTDA_LAUNCH_CHECK(ctx, "scan_kernel<node sizes>");
    node_scan_kernel<true><<<blocks, 256, 0, st>>>(V, d_node_bin, node_label, d_bin_off, d_members, d_labels, nullptr,
                                                   d_node_off, d_node_members);
    SCTDA_LAUNCH_CHECK(ctx, "node_scan_kernel<fill>");
    return SCTDA_OK;
}

extern "C" int32_t sctda_nerve_count(sctda_ctx* ctx, int32_t N, int32_t V, const int32_t* d_node_off,
                                     const int32_t* d_node_members, int64_t Mtot, int64_t* d_edge_count, void* stream) {
    if (!ctx) return SCTDA_ERR_INVALID;
    if (N < 1 || V < 0 || Mtot < 0 || !d_node_off || !d_node_members || !d_edge_count)
        return sctda_fail(ctx, SCTDA_ERR_INVALID, "nerve_count: bad arguments%s");
    if (V > 131072) return sctda_fail(ctx, SCTDA_ERR_UNSUPPORTED, "nerve: more than 131,072 nodes is not served%s");
    cudaStream_t st = (cudaStream_t)stream;
    SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
    if (V == 0) { SCTDA_CUDA(ctx, cudaMemsetAsync(d_edge_count, 0, 8, st)); return SCTDA_OK; }
    const int W = (V + 31) / 32;
    // workspace layout: bits [V*W] u32 | row_off [V+1] i64 | row_cnt [V] i32 | cell_off [N+1] | cursor [N] | cell_nodes [Mtot]
    const size_t bits_b = (size_t)V * W * 4;
    const size_t total = bits_b + (size_t)(V + 1) * 8 + (size_t)V * 4 + (size_t)(N + 1) * 4 + (size_t)N * 4 + (size_t)Mtot * 4 + 64;
    unsigned char* ws = (unsigned char*)sctda_ws(ctx, WS_NERVE, total);
    if (!ws) return SCTDA_ERR_NOMEM;
    unsigned* bits = (unsigned*)ws;
    int64_t* row_off = (int64_t*)(ws + ((bits_b + 7) & ~(size_t)7));
    int32_t* row_cnt = (int32_t*)(row_off + V + 1);
    int32_t* cell_off = row_cnt + V;
    int32_t* cursor = cell_off + N + 1;
    int32_t* cell_nodes = cursor + N;
    SCTDA_CUDA(ctx, cudaMemsetAsync(bits, 0, bits_b, st));
    SCTDA_CUDA(ctx, cudaMemsetAsync(cursor, 0, (size_t)N * 4, st));
    if (Mtot > 0) {
        cell_count_kernel<<<(unsigned)cdiv64(Mtot, 256), 256, 0, st>>>(Mtot, d_node_members, cursor);
        SCTDA_LAUNCH_CHECK(ctx, "cell_count_kernel");
    }
    scan_kernel<int32_t><<<1, 1024, 0, st>>>(cursor, N, MAP_ID, cell_off, nullptr);
    SCTDA_LAUNCH_CHECK(ctx, "scan_kernel<cells>");
    SCTDA_CUDA(ctx, cudaMemsetAsync(cursor, 0, (size_t)N * 4, st));
    const unsigned vblocks = (unsigned)cdiv64((int64_t)V * 32, 256);
    cell_fill_kernel<<<vblocks, 256, 0, st>>>(V, d_node_off, d_node_members, cell_off, cursor, cell_nodes);
    SCTDA_LAUNCH_CHECK(ctx, "cell_fill_kernel");
    pair_kernel<<<(N + 127) / 128, 128, 0, st>>>(N, W, cell_off, cell_nodes, bits);
    SCTDA_LAUNCH_CHECK(ctx, "pair_kernel");
    edge_scan_kernel<false><<<vblocks, 256, 0, st>>>(V, W, bits, row_cnt, nullptr, nullptr);
    SCTDA_LAUNCH_CHECK(ctx, "edge_scan_kernel<count>");
    scan_kernel<int64_t><<<1, 1024, 0, st>>>(row_cnt, V, MAP_ID, row_off, nullptr);
    SCTDA_LAUNCH_CHECK(ctx, "scan_kernel<rows>");
    SCTDA_CUDA(ctx, cudaMemcpyAsync(d_edge_count, row_off + V, 8, cudaMemcpyDeviceToDevice, st));
    return SCTDA_OK;
}

extern "C" int32_t sctda_nerve_fill(sctda_ctx* ctx, int32_t V, int32_t* d_edges, void* stream) {
    if (!ctx) return SCTDA_ERR_INVALID;
    if (V < 0 || !d_edges) return sctda_fail(ctx, SCTDA_ERR_INVALID, "nerve_fill: bad arguments%s");
    if (V == 0) return SCTDA_OK;
    const int W = (V + 31) / 32;
    const size_t bits_b = (size_t)V * W * 4;
    if (ctx->slot_bytes[WS_NERVE] < bits_b + (size_t)(V + 1) * 8)
        return sctda_fail(ctx, SCTDA_ERR_INVALID, "nerve_fill: call sctda_nerve_count with the same V first%s");
    SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned char* ws = (unsigned char*)ctx->slot_ptr[WS_NERVE];
    const unsigned* bits = (const unsigned*)ws;
    const int64_t* row_off = (const int64_t*)(ws + ((bits_b + 7) & ~(size_t)7));
    edge_scan_kernel<true><<<(unsigned)cdiv64((int64_t)V * 32, 256), 256, 0, (cudaStream_t)stream>>>(V, W, bits, nullptr,
                                                                                                  row_off, d_edges);
    SCTDA_LAUNCH_CHECK(ctx, "edge_scan_kernel<fill>");
    return SCTDA_OK;
}
