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

#include <cooperative_groups.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "gemm.cuh"

int sctda_internal_gram_bins(sctda_ctx* ctx, int G, const double* d_Xn, int64_t ldxn, int nprob,
                             const int32_t* d_prob_bin, const int32_t* d_bin_off, const int32_t* d_members,
                             const int64_t* d_tile_prefix, const int64_t* d_gram_off, double* d_gram, cudaStream_t st,
                             int64_t nrows, bool reuse_split, int prob_begin, int prob_end, unsigned long long* d_tile_counter,
                             int grid_ctas, bool untimed, int tiles_per_cta);

namespace {

__global__ void max_member_kernel(int64_t M, const int32_t* __restrict__ members, int32_t* __restrict__ out) {
    int v = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) v = max(v, members[i]);
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, v);
}

constexpr int kMaxK = 8;            // largest max_K served
constexpr int kMaxRes = 128;        // largest resolution served (two 64-bit mask words per axis)

// ---------------------------------------------------------------------------------------------
// exclusive scan, one CTA: out[0] = 0, out[i+1] = out[i] + in[i]
// ---------------------------------------------------------------------------------------------
template <typename TOut>
__global__ void __launch_bounds__(1024) scan_kernel(const int32_t* __restrict__ in, int64_t n, TOut* __restrict__ out) {
    __shared__ long long s_sum[1024];
    const int t = threadIdx.x;
    const int64_t per = (n + 1023) / 1024;
    const int64_t lo = (int64_t)t * per, hi = lo + per < n ? lo + per : n;
    long long s = 0;
    for (int64_t i = lo; i < hi; ++i) s += in[i];
    s_sum[t] = s;
    __syncthreads();
    if (t == 0) {
        long long run = 0;
        for (int i = 0; i < 1024; ++i) { long long v = s_sum[i]; s_sum[i] = run; run += v; }
        out[n] = (TOut)run;
    }
    __syncthreads();
    long long run = s_sum[t];
    for (int64_t i = lo; i < hi; ++i) { out[i] = (TOut)run; run += in[i]; }
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
    int stat;                 // 0: Davies-Bouldin selection over 2..K_max, 1: first-empty-histogram-bin cut
    const int32_t* bin_off;
    const int32_t* prob_bin;  // [nprob] patch index of problem p of the current batch (largest first)
    int prob0;                // first problem served by this launch
    const int64_t* gram_off;  // [nprob+1] element offset of problem p's Gram tile in `gram`
    const double* gram;
    int prim_cap, lev_cap;    // largest m whose Prim state / level labels fit in shared memory
    double* pkey;             // [Mtot] global fallbacks for patches above the caps
    double* pgpp;
    int32_t* ppar;
    int8_t* lev_g;            // [Mtot][kMaxK-1]
    int32_t* fine_rep;        // [nprob][kMaxK] smallest vertex of every cluster of the finest level
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
    double* dmat;             // linkage != single: working distance matrices, same offsets as `gram`
    int linkage;              // 0 single (MST kernels), 1 average, 2 ward, 3 complete (linkage_kernel)
};

__device__ __forceinline__ double dist2(double gpp_u, double gpp_q, double g) {
    return fmax(fma(-2.0, g, __dadd_rn(gpp_u, gpp_q)), 0.0);
}

__global__ void __launch_bounds__(256) prim_kernel(ClusterArgs a) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ double s_rk[8];
    __shared__ int s_ri[8];
    __shared__ int s_u;
    const int prob = a.prob0 + blockIdx.x;
    const int b = a.prob_bin[prob];
    const int off = a.bin_off[b], m = a.bin_off[b + 1] - off;
    if (m == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool big = m > a.prim_cap;
    double* key = big ? a.pkey + off : (double*)s_raw;   // >= 0: best known distance to the tree; -1: in the tree
    double* gpp = big ? a.pgpp + off : key + m;
    int* par = big ? a.ppar + off : (int*)(gpp + m);
    const double* Gm = a.gram + a.gram_off[prob];
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

// Prim with the keys in REGISTERS: 512 threads, thread t owns vertices t, t+512, ... (KPT of them);
// |row|^2 and the parents live in shared memory.  One barrier per step (double-buffered warp
// results); the only serial latency left is the fetch of the next row of the Gram tile.
template <int KPT, int NT>
__global__ void __launch_bounds__(NT) prim_reg_kernel(ClusterArgs a) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ double s_rk[2][NT / 32];
    __shared__ int s_ri[2][NT / 32];
    const int prob = a.prob0 + blockIdx.x;
    const int b = a.prob_bin[prob];
    const int off = a.bin_off[b], m = a.bin_off[b + 1] - off;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* gpp = (double*)s_raw;
    int* par = (int*)(gpp + m);
    const double* Gm = a.gram + a.gram_off[prob];
    double key[KPT];
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
        const int q = tid + NT * i;
        key[i] = (q < m && q != 0) ? INFINITY : -1.0;          // -1: in the tree (or not a vertex)
        if (q < m) { gpp[q] = Gm[(int64_t)q * m + q]; par[q] = -1; }
    }
    if (tid == 0) { a.mst_parent[off] = -1; a.mst_w[off] = 0.0; a.mst_order[off] = -1; a.mst_seq[off] = 0; }
    __syncthreads();
    int u = 0;
    for (int t = 0; t < m - 1; ++t) {
        const double gu = gpp[u];
        const double* row = Gm + (int64_t)u * m;
        double g[KPT];
#pragma unroll
        for (int i = 0; i < KPT; ++i) g[i] = key[i] >= 0.0 ? row[tid + NT * i] : 0.0;
        double bk = INFINITY;
        int bi = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < KPT; ++i) {
            const int q = tid + NT * i;
            double k = key[i];
            if (k >= 0.0) {
                const double d = dist2(gu, gpp[q], g[i]);
                if (d < k) { k = d; key[i] = d; par[q] = u; }
                if (k < bk || (k == bk && q < bi)) { bk = k; bi = q; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ok = __shfl_xor_sync(0xffffffffu, bk, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ok < bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; }
        }
        const int pp = t & 1;
        if (lane == 0) { s_rk[pp][warp] = bk; s_ri[pp][warp] = bi; }
        __syncthreads();
        // every warp reduces the NT/32 warp candidates with shuffles (log steps instead of a sequential
        // scan by every thread; the parity double buffer makes a second barrier unnecessary)
        bk = lane < NT / 32 ? s_rk[pp][lane] : INFINITY;
        bi = lane < NT / 32 ? s_ri[pp][lane] : 0x7fffffff;
#pragma unroll
        for (int o = NT / 64; o > 0; o >>= 1) {
            const double ok = __shfl_xor_sync(0xffffffffu, bk, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ok < bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; }
        }
        bk = __shfl_sync(0xffffffffu, bk, 0);
        bi = __shfl_sync(0xffffffffu, bi, 0);
        u = bi;
        const int sel = (u % NT) == tid ? (u / NT) : -1;
#pragma unroll
        for (int i = 0; i < KPT; ++i) key[i] = (sel == i) ? -1.0 : key[i];     // stays in registers
        if (sel >= 0) {
            a.mst_parent[off + u] = par[u];
            a.mst_w[off + u] = bk;
            a.mst_order[off + u] = t;
            a.mst_seq[off + t + 1] = u;
        }
    }
}

// Prim for LARGE patches: one thread-block CLUSTER of 8 CTAs (4,096 threads) per patch.  Vertex q is
// owned by cluster thread q mod 4096 (keys and parents in registers); every CTA keeps its own copy
// of |row|^2 in shared memory.  Per step each CTA reduces its candidates, publishes (key, index) in
// its shared memory, one cluster barrier, and every CTA reads the eight candidates through
// distributed shared memory: the per-step latency is one row fetch of m/8 elements per CTA plus
// one cluster barrier instead of a whole-row fetch by a single SM.
constexpr int kClusterCtas = 8;

template <int KPT>
__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(512) prim_cluster_kernel(ClusterArgs a) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ double s_rk[16];
    __shared__ int s_ri[16];
    __shared__ __align__(16) unsigned long long s_slot[2][kClusterCtas][2];   // candidates (key bits, index) of all CTAs, by step parity
    __shared__ __align__(8) unsigned long long s_xbar[2];                      // their arrival barriers
    const int rank = (int)cluster.block_rank();
    const int prob = a.prob0 + blockIdx.x / kClusterCtas;
    const int b = a.prob_bin[prob];
    const int off = a.bin_off[b], m = a.bin_off[b + 1] - off;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gt = rank * 512 + tid;             // cluster-wide thread index
    constexpr int STRIDE = kClusterCtas * 512;
    double* gpp = (double*)s_raw;
    const double* Gm = a.gram + a.gram_off[prob];
    for (int q = tid; q < m; q += 512) gpp[q] = Gm[(int64_t)q * m + q];
    double key[KPT];
    int par[KPT];
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
        const int q = gt + STRIDE * i;
        key[i] = (q < m && q != 0) ? INFINITY : -1.0;
        par[i] = -1;
    }
    if (gt == 0) { a.mst_parent[off] = -1; a.mst_w[off] = 0.0; a.mst_order[off] = -1; a.mst_seq[off] = 0; }
    if (tid == 0) {
        for (int i = 0; i < 2; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_xbar[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster.sync();                               // every CTA's barriers exist before anybody sends
    unsigned slot_remote[kClusterCtas], xbar_remote[kClusterCtas];       // shared::cluster addresses (used by thread 0)
#pragma unroll
    for (int r = 0; r < kClusterCtas; ++r) {
        asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(slot_remote[r]) : "r"((unsigned)__cvta_generic_to_shared(&s_slot[0][rank][0])), "r"(r));
        asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(xbar_remote[r]) : "r"((unsigned)__cvta_generic_to_shared(&s_xbar[0])), "r"(r));
    }
    int u = 0;
    for (int t = 0; t < m - 1; ++t) {
        const double gu = gpp[u];
        const double* row = Gm + (int64_t)u * m;
        double g[KPT];
#pragma unroll
        for (int i = 0; i < KPT; ++i) g[i] = key[i] >= 0.0 ? row[gt + STRIDE * i] : 0.0;
        double bk = INFINITY;
        int bi = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < KPT; ++i) {
            const int q = gt + STRIDE * i;
            double k = key[i];
            if (k >= 0.0) {
                const double d = dist2(gu, gpp[q], g[i]);
                if (d < k) { k = d; key[i] = d; par[i] = u; }
                if (k < bk || (k == bk && q < bi)) { bk = k; bi = q; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ok = __shfl_xor_sync(0xffffffffu, bk, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ok < bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; }
        }
        if (lane == 0) { s_rk[warp] = bk; s_ri[warp] = bi; }
        __syncthreads();
        const int pp = t & 1;
        if (warp == 0) {
            bk = lane < 16 ? s_rk[lane] : INFINITY;
            bi = lane < 16 ? s_ri[lane] : 0x7fffffff;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                const double ok = __shfl_xor_sync(0xffffffffu, bk, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ok < bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; }
            }
            if (lane == 0) {
                // Candidate exchange without a cluster barrier: one 16-byte st.async into slot [pp][rank] of
                // every CTA (this one included), each completing 16 bytes on the receiver's mbarrier; every
                // CTA then waits on its OWN barrier for the 128 bytes of the step.  (barrier.cluster with
                // release semantics costs a MEMBAR.ALL.GPU per step: 30 % of this kernel's stall samples.)
                const unsigned xb = (unsigned)__cvta_generic_to_shared(&s_xbar[pp]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xb), "r"(16u * kClusterCtas) : "memory");
                const unsigned long long kb = (unsigned long long)__double_as_longlong(bk), ib = (unsigned long long)(unsigned)bi;
#pragma unroll
                for (int r = 0; r < kClusterCtas; ++r)
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];"
                                 ::"r"(slot_remote[r] + (unsigned)pp * (kClusterCtas * 16)), "l"(kb), "l"(ib),
                                   "r"(xbar_remote[r] + (unsigned)pp * 8) : "memory");
            }
        }
        {
            const unsigned xb = (unsigned)__cvta_generic_to_shared(&s_xbar[pp]);
            const unsigned parity = (unsigned)(t >> 1) & 1u;
            const long long t0 = clock64();
            for (;;) {
                unsigned ok;
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(xb), "r"(parity) : "memory");
                if (ok) break;
                if (clock64() - t0 > 8000000000LL) __trap();
            }
        }
        // the eight candidates, one per lane, reduced by shuffles in every warp
        bk = lane < kClusterCtas ? __longlong_as_double((long long)s_slot[pp][lane][0]) : INFINITY;
        bi = lane < kClusterCtas ? (int)s_slot[pp][lane][1] : 0x7fffffff;
#pragma unroll
        for (int o = kClusterCtas / 2; o > 0; o >>= 1) {
            const double ok = __shfl_xor_sync(0xffffffffu, bk, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ok < bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; }
        }
        bk = __shfl_sync(0xffffffffu, bk, 0);
        bi = __shfl_sync(0xffffffffu, bi, 0);
        u = bi;
        const int sel = (u % STRIDE) == gt ? (u / STRIDE) : -1;
        int pu = -1;
#pragma unroll
        for (int i = 0; i < KPT; ++i) {
            if (sel == i) { pu = par[i]; key[i] = -1.0; }
        }
        if (sel >= 0) {
            a.mst_parent[off + u] = pu;
            a.mst_w[off + u] = bk;
            a.mst_order[off + u] = t;
            a.mst_seq[off + t + 1] = u;
        }
    }
    cluster.sync();                               // no CTA exits while its shared memory may still be read
}

// ---------------------------------------------------------------------------------------------
// cut + Davies-Bouldin model selection per patch
// ---------------------------------------------------------------------------------------------
// Sum over p ascending of (lab[p] == I ? v[p*stride] : 0.0): the masked terms add +0.0, an exact
// no-op, so this is the sequential sum over the members of cluster I (oracle order).  Done by one
// WARP: lane l loads element p0 + l of each 32-element chunk (coalesced, two chunks ahead of the
// adds), the chunk is then added in element order through shuffles; every lane ends with the same
// value.  The m dependent adds thus never wait for global memory (a single thread issuing eight
// loads at a time spent m/8 round trips on it: 5.4 -> 2.2 ms on a rank of the 8-GPU build).
__device__ __forceinline__ double warp_masked_seq_sum(const double* __restrict__ v, int stride, const int8_t* lab, int I, int m,
                                                      int lane) {
    auto fetch = [&](int p0) -> double {
        const int p = p0 + lane;
        return (p < m && lab[p] == I) ? v[(int64_t)p * stride] : 0.0;
    };
    double acc = 0.0;
    double x0 = fetch(0), x1 = fetch(32);
    for (int p0 = 0; p0 < m; p0 += 32) {
        const double x2 = fetch(p0 + 64);
#pragma unroll
        for (int u = 0; u < 32; ++u) acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, x0, u));
        x0 = x1;
        x1 = x2;
    }
    return acc;
}

// ------------------------------------------------------------------------------------------------
// Average / Ward / complete linkage of one patch (cluster='agglomerative' with sklearn's other linkages,
// /root/reference/scTDA/main.py:758-771): nearest-neighbour chain with Lance-Williams updates on a working copy
// of the patch's distance matrix.  One CTA per patch; every step is one block-wide argmin over a matrix row.
//   init     D[i][j] = squared Euclidean distance from the Gram tile (Ward), or its square root (average, complete)
//   chain    follow nearest neighbours (ties: the previous chain element first, then the smallest index) until two
//            clusters are mutually nearest; merge them into the slot of smaller index
//   update   average  d(k, ab) = (n_a d(k,a) + n_b d(k,b)) / (n_a + n_b)
//            Ward     d2(k, ab) = ((n_a+n_k) d2(k,a) + (n_b+n_k) d2(k,b) - n_k d2(a,b)) / (n_a+n_b+n_k)
//            complete d(k, ab) = max(d(k,a), d(k,b))
// All three are reducible: the chain's merges, ordered by height, are the dendrogram.  The merge tree is written
// in the form the single-linkage MST has (parent / weight / insertion order / insertion sequence): the slot that
// disappears in a merge is a vertex whose parent is the surviving slot and whose edge weight is the squared merge
// height; the sequence lists the last survivor first and then the slots in reverse order of disappearance.
// Removing the K-1 heaviest edges of that tree leaves the K clusters of the dendrogram, so labels_kernel, the
// Davies-Bouldin selection and the histogram-gap cut run unchanged on it.
// ------------------------------------------------------------------------------------------------
constexpr int kLinkThreads = 1024;

__global__ void __launch_bounds__(kLinkThreads) linkage_kernel(ClusterArgs a) {
    __shared__ double s_v[32];
    __shared__ int s_i[32];
    __shared__ int s_pick;
    __shared__ double s_pickv;
    const int prob = a.prob0 + blockIdx.x;
    const int b = a.prob_bin[prob];
    const int off = a.bin_off[b], m = a.bin_off[b + 1] - off;
    if (m < 2) {
        if (m == 1 && threadIdx.x == 0) { a.mst_parent[off] = -1; a.mst_order[off] = -1; a.mst_seq[off] = 0; a.mst_w[off] = 0.0; }
        return;
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* Gm = a.gram + a.gram_off[prob];
    double* D = a.dmat + a.gram_off[prob];
    double* size = a.dist + off;                 // cluster size of an active slot, 0 for a slot that is gone
    int32_t* chain = a.ppar + off;
    int32_t* par = a.mst_parent + off;
    int32_t* ord = a.mst_order + off;
    int32_t* seq = a.mst_seq + off;
    double* w = a.mst_w + off;
    const bool ward = a.linkage == 2;
    for (int64_t e = tid; e < (int64_t)m * m; e += kLinkThreads) {
        const int i = (int)(e / m), j = (int)(e % m);
        const double d2 = dist2(Gm[(int64_t)i * m + i], Gm[(int64_t)j * m + j], Gm[e]);
        D[e] = ward ? d2 : sqrt(d2);
    }
    for (int v = tid; v < m; v += kLinkThreads) { size[v] = 1.0; par[v] = -1; ord[v] = -1; w[v] = 0.0; }
    __syncthreads();
    int chain_len = 0, cursor = 0, merges = 0;   // (uniform across the CTA)
    while (merges < m - 1) {
        if (chain_len == 0) {
            while (size[cursor] == 0.0) ++cursor;         // the lowest active slot starts a chain
            if (tid == 0) chain[0] = cursor;
            chain_len = 1;
            __syncthreads();
        }
        const int x = chain[chain_len - 1];
        const int prev = chain_len > 1 ? chain[chain_len - 2] : -1;
        // nearest active neighbour of x: smallest distance, the previous chain element wins ties, then the smallest index
        const double* row = D + (int64_t)x * m;
        double bv = INFINITY;
        int bi = 0x7fffffff;
        for (int j = tid; j < m; j += kLinkThreads) {
            if (j == x || size[j] == 0.0) continue;
            const double d = row[j];
            if (d < bv || (d == bv && (j == prev || (bi != prev && j < bi)))) { bv = d; bi = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov < bv || (ov == bv && oi != bi && (oi == prev || (bi != prev && oi < bi)))) { bv = ov; bi = oi; }
        }
        if (lane == 0) { s_v[warp] = bv; s_i[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int k = 1; k < kLinkThreads / 32; ++k) {
                const double ov = s_v[k];
                const int oi = s_i[k];
                if (ov < bv || (ov == bv && oi != bi && (oi == prev || (bi != prev && oi < bi)))) { bv = ov; bi = oi; }
            }
            s_pick = bi;
            s_pickv = bv;
        }
        __syncthreads();
        const int y = s_pick;
        const double dxy = s_pickv;
        if (y != prev) {                                   // grow the chain
            if (tid == 0) chain[chain_len] = y;
            ++chain_len;
            __syncthreads();
            continue;
        }
        // x and y are mutually nearest: merge into the slot of smaller index
        const int keep = min(x, y), gone = max(x, y);
        const double na = size[keep], nb = size[gone];
        const double* rk = D + (int64_t)keep * m;
        const double* rg = D + (int64_t)gone * m;
        __syncthreads();                                   // every thread has read size[] of this step
        for (int k = tid; k < m; k += kLinkThreads) {
            const double nk = size[k];
            if (k == keep || k == gone || nk == 0.0) continue;
            const double dk = rk[k], dg = rg[k];
            double d;
            if (a.linkage == 1) d = (na * dk + nb * dg) / (na + nb);
            else if (a.linkage == 2) d = ((na + nk) * dk + (nb + nk) * dg - nk * dxy) / (na + nb + nk);
            else d = fmax(dk, dg);
            D[(int64_t)keep * m + k] = d;
            D[(int64_t)k * m + keep] = d;
        }
        if (tid == 0) {
            size[keep] = na + nb;
            size[gone] = 0.0;
            par[gone] = keep;
            ord[gone] = merges;                            // chronological index of the merge
            w[gone] = ward ? dxy : dxy * dxy;              // squared height, like the MST's squared edge lengths
            seq[m - 1 - merges] = gone;                    // reverse order of disappearance; seq[0] is filled at the end
        }
        ++merges;
        chain_len -= 2;
        __syncthreads();
    }
    if (tid == 0) {
        int root = 0;
        while (size[root] == 0.0) ++root;
        seq[0] = root;
    }
}

constexpr int kSelThreads = 1024;

// Kernel A of the selection: the (K_max - 1) heaviest tree edges and the labels of every level
// (or, for the histogram-gap rule, the final labels).  One CTA per patch.
__global__ void __launch_bounds__(kSelThreads) labels_kernel(ClusterArgs a) {
    extern __shared__ __align__(16) unsigned char s_dyn[];      // labels of every level: int8 [nlev][m]
    __shared__ double s_rk[32];
    __shared__ int s_ri[32], s_ro[32];
    __shared__ int s_heavy[kMaxK];               // vertices whose edge is the h-th heaviest
    __shared__ int s_fine_rep[kMaxK];
    __shared__ double s_cc[kMaxK][kMaxK];        // scratch of the histogram-gap reduction
    const int prob = a.prob0 + blockIdx.x;
    const int b = a.prob_bin[prob];
    const int off = a.bin_off[b], m = a.bin_off[b + 1] - off;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (m < 2) return;                           // handled by trivial_bins_kernel
    const int kmax = m <= 5 ? 2 : min(m / 2, a.max_K);
    const int nlev = kmax - 1;                   // level L (0-based) = L+1 edges removed = L+2 clusters
    const double* w = a.mst_w + off;
    const int32_t* ord = a.mst_order + off;
    const int32_t* par = a.mst_parent + off;
    const int32_t* seq = a.mst_seq + off;
    const bool lev_in_smem = (int64_t)m * (a.max_K - 1) <= a.lev_cap;
    int8_t* lev_out = a.lev_g + (size_t)off * (kMaxK - 1);     // [nlev][m] in global memory for kernels B and C
    int8_t* lev = lev_in_smem ? (int8_t*)s_dyn : lev_out;
    if (a.stat == 1) {
        // Histogram-gap cut (the rule BASELINE.json's north_star names; oracle cluster_bin(stat='histgap')):
        // edge lengths h = sqrt(d2) of the m-1 tree edges, 10 equal bins between min and max, every edge
        // in a bin above the first empty one is removed; no empty bin (or hi == lo) -> one cluster.
        constexpr int NB = 10;
        __shared__ int s_cnt[NB];
        __shared__ double s_lo, s_hi;
        if (tid < NB) s_cnt[tid] = 0;
        double lo = INFINITY, hi = -INFINITY;
        for (int q = tid; q < m; q += kSelThreads)
            if (ord[q] >= 0) { const double h = sqrt(w[q]); lo = fmin(lo, h); hi = fmax(hi, h); }
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) {
            lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o2));
            hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o2));
        }
        if (lane == 0) { s_rk[warp] = lo; s_cc[warp / kMaxK][warp % kMaxK] = hi; }
        __syncthreads();
        if (tid == 0) {
            for (int x = 1; x < kSelThreads / 32; ++x) { lo = fmin(lo, s_rk[x]); hi = fmax(hi, s_cc[x / kMaxK][x % kMaxK]); }
            s_lo = lo; s_hi = hi;
        }
        __syncthreads();
        lo = s_lo; hi = s_hi;
        const bool flat = !(hi > lo);
        const double width = __ddiv_rn(__dsub_rn(hi, lo), (double)NB);
        if (!flat)
            for (int q = tid; q < m; q += kSelThreads)
                if (ord[q] >= 0) atomicAdd(&s_cnt[min((int)__ddiv_rn(__dsub_rn(sqrt(w[q]), lo), width), NB - 1)], 1);
        __syncthreads();
        __shared__ int s_empty;
        __shared__ int s_scan[kSelThreads];
        if (tid == 0) {
            int empty = -1;
            if (!flat)
                for (int i = 0; i < NB && empty < 0; ++i)
                    if (s_cnt[i] == 0) empty = i;
            s_empty = empty;
        }
        __syncthreads();
        const int empty = s_empty;
        // Components of the tree without the cut edges, by the whole CTA: every vertex points at its parent
        // unless its own edge is cut (the root of its component points at itself); pointer jumping doubles the
        // distance covered per round; the component's smallest member names it; components are numbered by
        // smallest member through a prefix count.
        int32_t* comp = a.labels + off;
        int32_t* scratch = (int32_t*)(a.TK + (int64_t)off * kMaxK);          // up [m], cmin [m], first [m]
        int32_t* up = scratch, *cmin = scratch + m, *first = scratch + 2 * (int64_t)m;
        for (int v = tid; v < m; v += kSelThreads) {
            bool cut = v == 0 || ord[v] < 0;                                 // vertex 0 is the root of the tree
            if (!cut && empty >= 0) cut = min((int)__ddiv_rn(__dsub_rn(sqrt(w[v]), lo), width), NB - 1) > empty;
            up[v] = cut ? v : par[v];
            cmin[v] = 0x7fffffff;
            first[v] = 0;
        }
        __syncthreads();
        int32_t* src = up;
        int32_t* dst = first;                                                // free until the prefix count below
        for (int span = 1; span < m; span <<= 1) {                           // ceil(log2 m) rounds, ping-pong buffers
            for (int v = tid; v < m; v += kSelThreads) dst[v] = src[src[v]];
            __syncthreads();
            int32_t* t = src; src = dst; dst = t;
        }
        if (src != up)
            for (int v = tid; v < m; v += kSelThreads) up[v] = src[v];
        __syncthreads();
        for (int v = tid; v < m; v += kSelThreads) first[v] = 0;
        __syncthreads();
        for (int v = tid; v < m; v += kSelThreads) atomicMin(&cmin[up[v]], v);
        __syncthreads();
        for (int v = tid; v < m; v += kSelThreads)
            if (up[v] == v) first[cmin[v]] = 1;                              // v is the smallest member of a component
        __syncthreads();
        // exclusive prefix count of `first` over the vertices: contiguous chunks per thread, chunk totals scanned by thread 0
        const int chunk = (m + kSelThreads - 1) / kSelThreads;
        const int c0 = min(m, tid * chunk), c1 = min(m, c0 + chunk);
        int local = 0;
        for (int v = c0; v < c1; ++v) local += first[v];
        s_scan[tid] = local;
        __syncthreads();
        if (tid == 0) {
            int run = 0;
            for (int i = 0; i < kSelThreads; ++i) { const int x = s_scan[i]; s_scan[i] = run; run += x; }
            a.bin_K[b] = run;
        }
        __syncthreads();
        int run = s_scan[tid];
        for (int v = c0; v < c1; ++v) {
            const int f = first[v];
            first[v] = run;                                                   // number of the component whose smallest member is v
            run += f;
        }
        __syncthreads();
        for (int v = tid; v < m; v += kSelThreads) comp[v] = first[cmin[up[v]]];
        return;
    }
    // (c) the nlev heaviest tree edges: (weight desc, insertion order desc)
    for (int h = 0; h < nlev; ++h) {
        double bk = -1.0;
        int bo = -1, bi = -1;
        for (int q = tid; q < m; q += kSelThreads) {
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
            for (int x = 1; x < kSelThreads / 32; ++x)
                if (s_rk[x] > bk || (s_rk[x] == bk && s_ro[x] > bo)) { bk = s_rk[x]; bo = s_ro[x]; bi = s_ri[x]; }
            s_heavy[h] = bi;
        }
        __syncthreads();
    }
    // (d) labels of every level: thread L walks the vertices in insertion order; a vertex whose edge
    //     is among the L+1 heaviest opens a component, the others inherit their parent's (inserted
    //     earlier).  Components are then renumbered by their smallest vertex.
    __shared__ int8_t s_rank[kMaxK - 1][kMaxK];
    if (warp == 0) {
        // warp 0: lane L < nlev owns level L; all 32 lanes fetch the insertion sequence and the parents
        // (two dependent gathers) 32 at a time, two chunks ahead of the walk
        const bool own = lane < nlev;
        const int L = own ? lane : 0;
        int8_t* lab = lev + (size_t)L * m;
        // branch-free walk: a vertex whose edge is the g-th heaviest (g <= L) opens component g + 1, every
        // other vertex inherits its parent's component; the smallest vertex of every component is tracked
        // in registers.  (Component numbers are temporary: they are re-ranked by smallest vertex below.)
        int hv[kMaxK - 1], cm[kMaxK];
#pragma unroll
        for (int g = 0; g < kMaxK - 1; ++g) hv[g] = (g <= L && g < nlev) ? s_heavy[g] : -1;
#pragma unroll
        for (int d = 0; d < kMaxK; ++d) cm[d] = 0x7fffffff;
        cm[0] = 0;
        if (own) lab[0] = 0;
        int v0, p0, v1, p1;
        { const int t = 1 + lane; v0 = t < m ? seq[t] : 0; p0 = t < m ? par[v0] : 0; }
        { const int t = 33 + lane; v1 = t < m ? seq[t] : 0; p1 = t < m ? par[v1] : 0; }
        for (int t0 = 1; t0 < m; t0 += 32) {
            const int t2 = t0 + 64 + lane;
            const int v2 = t2 < m ? seq[t2] : 0;
            const int p2 = t2 < m ? par[v2] : 0;
            const int cnt = min(32, m - t0);                  // warp-uniform
            for (int u = 0; u < cnt; ++u) {
                const int v = __shfl_sync(0xffffffffu, v0, u), pu = __shfl_sync(0xffffffffu, p0, u);
                int c = own ? (int)lab[pu] : 0;
#pragma unroll
                for (int g = 0; g < kMaxK - 1; ++g) c = (hv[g] == v) ? g + 1 : c;
#pragma unroll
                for (int d = 0; d < kMaxK; ++d) cm[d] = (d == c) ? min(cm[d], v) : cm[d];
                if (own) lab[v] = (int8_t)c;
            }
            v0 = v1; p0 = p1; v1 = v2; p1 = p2;
        }
        if (own) {
            const int ncomp = L + 2;
#pragma unroll
            for (int c = 0; c < kMaxK; ++c) {                  // components renumbered by their smallest vertex
                if (c < ncomp) {
                    int rk = 0;
#pragma unroll
                    for (int d = 0; d < kMaxK; ++d) rk += (d < ncomp && cm[d] < cm[c]);
                    s_rank[L][c] = (int8_t)rk;
                    if (L == nlev - 1) s_fine_rep[rk] = cm[c];
                }
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < nlev * m; i += kSelThreads) {         // relabel (and copy out when the labels live in shared memory)
        const int8_t r = s_rank[i / m][lev[i]];
        lev[i] = r;
        if (lev_in_smem) lev_out[i] = r;
    }
    if (tid < kmax) a.fine_rep[(int64_t)prob * kMaxK + tid] = s_fine_rep[tid];
}

// Kernel B: T[p][f] = sum over q ascending of G[q][p] for q in fine cluster f.  The columns of a patch
// are independent: one CTA per 256 columns, so that a large patch streams its Gram tile through
// many SMs instead of one (a single CTA sustains < 100 GB/s).
constexpr int kColChunk = 256;

__global__ void __launch_bounds__(kColChunk) tcol_kernel(ClusterArgs a, int nprob, const int64_t* __restrict__ chunk_prefix) {
    const int64_t t = blockIdx.x;
    int lo = 0, hi = nprob;                                   // last problem with chunk_prefix[p] <= t
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (chunk_prefix[mid] <= t) lo = mid; else hi = mid;
    }
    const int prob = lo;
    const int b = a.prob_bin[prob];
    const int off = a.bin_off[b], m = a.bin_off[b + 1] - off;
    if (m < 2 || a.stat == 1) return;
    const int p = (int)(t - chunk_prefix[prob]) * kColChunk + threadIdx.x;
    const double* Gm = a.gram + a.gram_off[prob];
    const int kmax = m <= 5 ? 2 : min(m / 2, a.max_K);
    const int nlev = kmax - 1;
    const int8_t* fine = a.lev_g + (size_t)off * (kMaxK - 1) + (size_t)(nlev - 1) * m;
    double* T = a.T + (int64_t)off * kMaxK;
    if (p >= m) return;
    {
        double acc[kMaxK];
#pragma unroll
        for (int f = 0; f < kMaxK; ++f) acc[f] = 0.0;
        int q = 0;
        for (; q + 8 <= m; q += 8) {
            double g[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) g[u] = Gm[(int64_t)(q + u) * m + p];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
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
}

// Kernel C: Davies-Bouldin index of every level from T, choice of K, final labels.  One CTA per patch.
__global__ void __launch_bounds__(kSelThreads) db_kernel(ClusterArgs a) {
    __shared__ int s_cof[kMaxK];                 // fine cluster -> coarse cluster of the level under evaluation
    __shared__ int s_fine_rep[kMaxK];
    __shared__ double s_n[kMaxK];
    __shared__ double s_cc[kMaxK][kMaxK];
    __shared__ double s_S[kMaxK];
    __shared__ double s_db[kMaxK];
    const int prob = a.prob0 + blockIdx.x;
    const int b = a.prob_bin[prob];
    const int off = a.bin_off[b], m = a.bin_off[b + 1] - off;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (m < 2 || a.stat == 1) return;
    const double* Gm = a.gram + a.gram_off[prob];
    const int kmax = m <= 5 ? 2 : min(m / 2, a.max_K);
    const int nlev = kmax - 1;
    const int8_t* lev = a.lev_g + (size_t)off * (kMaxK - 1);
    if (tid < kmax) s_fine_rep[tid] = a.fine_rep[(int64_t)prob * kMaxK + tid];
    __syncthreads();
    const int F = kmax;
    double* T = a.T + (int64_t)off * kMaxK;                    // [m][kMaxK]
    double* TK = a.TK + (int64_t)off * kMaxK;                  // [m][kMaxK], per level
    double* dist = a.dist + off;
    // (f) Davies-Bouldin index of every level
    for (int L = 0; L < nlev; ++L) {
        const int K = L + 2;
        const int8_t* lab = lev + (size_t)L * m;
        if (tid < F) s_cof[tid] = lab[s_fine_rep[tid]];
        __syncthreads();
        for (int p = tid; p < m; p += kSelThreads) {                    // TK[p][J] = sum_{f in J, ascending} T[p][f]
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
        if (warp < K) {                                         // cluster sizes, one warp per cluster
            int n = 0;
            for (int p = lane; p < m; p += 32) n += (lab[p] == warp);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
            if (lane == 0) s_n[warp] = (double)n;
        }
        __syncthreads();
        for (int pr = warp; pr < K * K; pr += kSelThreads / 32) {     // c_I . c_J, one warp per pair
            const int I = pr / K, J = pr % K;
            const double acc = warp_masked_seq_sum(TK + J, kMaxK, lab, I, m, lane);
            if (lane == 0) s_cc[I][J] = __ddiv_rn(acc, __dmul_rn(s_n[I], s_n[J]));
        }
        __syncthreads();
        for (int p = tid; p < m; p += kSelThreads) {                    // |x_p - c_I|
            const int I = lab[p];
            const double xc = __ddiv_rn(TK[(int64_t)p * kMaxK + I], s_n[I]);
            const double d2 = fmax(__dadd_rn(fma(-2.0, xc, Gm[(int64_t)p * m + p]), s_cc[I][I]), 0.0);
            dist[p] = sqrt(d2);
        }
        __syncthreads();
        if (warp < K) {
            const double acc = warp_masked_seq_sum(dist, 1, lab, warp, m, lane);
            if (lane == 0) s_S[warp] = __ddiv_rn(acc, s_n[warp]);
        }
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
    for (int p = tid; p < m; p += kSelThreads) a.labels[off + p] = lab[p];
    if (tid == 0) a.bin_K[b] = bestL + 2;
}

// patches with 0 or 1 cells, and the NaN fill of the score table
__global__ void trivial_bins_kernel(int nbins, const int32_t* __restrict__ bin_off, int32_t* __restrict__ labels,
                                    int32_t* __restrict__ bin_K, double* __restrict__ bin_db) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbins) return;
    for (int i = 0; i < kMaxK; ++i) bin_db[(int64_t)b * kMaxK + i] = NAN;
    const int m = bin_off[b + 1] - bin_off[b];
    if (m == 0) bin_K[b] = 0;
    if (m == 1) { bin_K[b] = 1; labels[bin_off[b]] = 0; }
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
    scan_kernel<int32_t><<<1, 1024, 0, st>>>(counts, B, d_bin_off);
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
    SCTDA_TIME_RESET(ctx);
    if (G < 1 || nbins < 0 || !d_Xn || ldxn < G || !d_bin_off || !d_members || !d_labels || !d_bin_K)
        return sctda_fail(ctx, SCTDA_ERR_INVALID, "bin_cluster: bad arguments%s");
    if (max_K < 2 || max_K > kMaxK) return sctda_fail(ctx, SCTDA_ERR_UNSUPPORTED, "bin_cluster: max_K must be in 2..8%s");
    if (stat != 0 && stat != 1)
        return sctda_fail(ctx, SCTDA_ERR_UNSUPPORTED, "bin_cluster: statistics must be 0 ('db') or 1 ('histgap')%s");
    if ((ldxn & 1) || ((uintptr_t)d_Xn & 15))
        return sctda_fail(ctx, SCTDA_ERR_UNSUPPORTED, "bin_cluster: rows must be 16-byte aligned (even ldxn)%s");
    if (nbins == 0) return SCTDA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
    // patch sizes on the host (one small synchronising copy): batches, launch order, workspace sizes
    std::vector<int32_t> h_off((size_t)nbins + 1);
    SCTDA_CUDA(ctx, cudaMemcpyAsync(h_off.data(), d_bin_off, ((size_t)nbins + 1) * 4, cudaMemcpyDeviceToHost, st));
    SCTDA_CUDA(ctx, cudaStreamSynchronize(st));
    const size_t M = (size_t)h_off[nbins];
    double* db = d_bin_db;
    if (!db) {
        db = (double*)sctda_ws(ctx, WS_MISC, (size_t)nbins * kMaxK * 8);
        if (!db) return SCTDA_ERR_NOMEM;
    }
    trivial_bins_kernel<<<(nbins + 255) / 256, 256, 0, st>>>(nbins, d_bin_off, d_labels, d_bin_K, db);
    SCTDA_LAUNCH_CHECK(ctx, "trivial_bins_kernel");
    std::vector<int32_t> order;
    for (int b = 0; b < nbins; ++b)
        if (h_off[b + 1] - h_off[b] >= 2) order.push_back(b);
    if (order.empty()) return SCTDA_OK;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
        return (h_off[x + 1] - h_off[x]) > (h_off[y + 1] - h_off[y]);          // largest patches first
    });
    // Gram workspace: patches are processed in batches whose tiles fit in one of TWO buffers, so
    // that the contraction of batch i+1 (caller's stream) overlaps the MST / selection kernels of
    // batch i (the context's auxiliary stream): the few largest patches bound the MST time of a
    // batch and would otherwise leave most SMs idle.
    // opt-in split-precision contraction (gemm_split.cu): it splits rows 0..max(member) of Xn into (hi, lo)
    // operands once per call; reserve that workspace before the Gram buffers are sized
    int64_t x_rows = 0;
    if (ctx->precision == 1) {
        int32_t* d_max = (int32_t*)sctda_ws(ctx, WS_MISC2, 256);
        if (!d_max) return SCTDA_ERR_NOMEM;
        SCTDA_CUDA(ctx, cudaMemsetAsync(d_max, 0, 4, st));
        max_member_kernel<<<(unsigned)std::min<size_t>((M + 255) / 256, 1024), 256, 0, st>>>((int64_t)M, d_members, d_max);
        SCTDA_LAUNCH_CHECK(ctx, "max_member_kernel");
        int32_t h_max = 0;
        SCTDA_CUDA(ctx, cudaMemcpyAsync(&h_max, d_max, 4, cudaMemcpyDeviceToHost, st));
        SCTDA_CUDA(ctx, cudaStreamSynchronize(st));
        x_rows = (int64_t)h_max + 1;
        if (!sctda_ws(ctx, WS_SPLIT, (size_t)x_rows * ((G + 63) / 64) * 256 + 256)) return SCTDA_ERR_NOMEM;
    }
    size_t all_bytes = 0;
    for (int b : order) { const size_t mb = (size_t)(h_off[b + 1] - h_off[b]); all_bytes += mb * mb * 8; }
    size_t budget = (size_t)16 << 30;
    if (all_bytes > ctx->slot_bytes[WS_CLUSTER2] || all_bytes > budget) {
        // (steady state skips this driver query: everything fits the buffer of the previous call)
        size_t free_b = 0, total_b = 0;
        SCTDA_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
        free_b += ctx->slot_bytes[WS_CLUSTER2] + ctx->slot_bytes[WS_CLUSTER4];
        budget = free_b / 4;
        if (budget > ((size_t)16 << 30)) budget = (size_t)16 << 30;
        if (budget < ((size_t)256 << 20)) budget = (size_t)256 << 20;
    }
    const int linkage = ctx->linkage;             // 0 single (MST), 1 average, 2 ward, 3 complete (sctda_set_linkage)
    if (linkage != 0) budget /= 2;                // the chain works on a copy of every patch's distance matrix
    // (Cutting a build that fits one buffer into three batches so that the pipeline below can hide the MST chains of
    // its large patches was measured on a rank's share of the 8-GPU 100k x 2k build: 134 -> 158 ms.  Every batch of
    // such a shard still holds patches of 8,000+ cells, so a 20 ms chain stays exposed at the end and the contraction
    // pays for the ones it hides.)
    const char* pipe_env = getenv("SCTDA_PIPELINE");
    const size_t aux_bytes = M * (8 * (2 * kMaxK + 4) + 4 * 4 + (kMaxK - 1)) + 256;
    unsigned char* aux = (unsigned char*)sctda_ws(ctx, WS_CLUSTER3, aux_bytes);
    if (!aux) return SCTDA_ERR_NOMEM;
    ClusterArgs a;
    a.nbins = nbins; a.max_K = max_K; a.stat = stat; a.bin_off = d_bin_off;
    a.T = (double*)aux;
    a.TK = a.T + M * kMaxK;
    a.mst_w = a.TK + M * kMaxK;
    a.dist = a.mst_w + M;
    a.pkey = a.dist + M;
    a.pgpp = a.pkey + M;
    a.mst_parent = (int32_t*)(a.pgpp + M);
    a.mst_order = a.mst_parent + M;
    a.mst_seq = a.mst_order + M;
    a.ppar = a.mst_seq + M;
    a.lev_g = (int8_t*)(a.ppar + M);
    a.labels = d_labels; a.bin_K = d_bin_K; a.bin_db = db;
    a.linkage = linkage; a.dmat = nullptr;
    a.prim_cap = 0;                               // the generic kernel keeps its state in global memory
    a.lev_cap = 160 * 1024 - 16;                  // bytes of shared memory for the level labels, (max_K - 1) * m per patch
    if (!(ctx->attr_done & 1u)) {
        SCTDA_CUDA(ctx, cudaFuncSetAttribute(prim_cluster_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024 - 1024));
        SCTDA_CUDA(ctx, cudaFuncSetAttribute(prim_cluster_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 136 * 1024));
        SCTDA_CUDA(ctx, cudaFuncSetAttribute(prim_reg_kernel<8, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        SCTDA_CUDA(ctx, cudaFuncSetAttribute(labels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        ctx->attr_done |= 1u;
    }
    // size classes of the MST kernels (a test hook lowers them so that small inputs reach every kernel)
    int thr_mid = 4096, thr_big = 16384, thr_huge = 28000, thr_1k = 1024, thr_512 = 512;
    bool split_small = false;                     // separate launches with small CTAs for small patches (test hook only)
    if (const char* e = getenv("SCTDA_TEST_PRIM_THRESHOLDS")) {
        int x = 0, y = 0, z = 0, p = 0, q = 0;
        const int got = sscanf(e, "%d,%d,%d,%d,%d", &x, &y, &z, &p, &q);
        if (got >= 3 && 0 < x && x <= y && y <= z && x <= 4096 && y <= 16384 && z <= 28000) {
            thr_mid = x; thr_big = y; thr_huge = z;
            thr_1k = thr_1k < x ? thr_1k : x;
            thr_512 = thr_512 < x ? thr_512 : x;
            if (got == 5 && 0 < q && q <= p && p <= x && p <= 1024 && q <= 512) { thr_1k = p; thr_512 = q; split_small = true; }
        }
    }
    if (!ctx->fork_stream) {
        int prio_lo = 0, prio_hi = 0;                 // the highest priority: when the Gram tiles of the large patches are
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);      // done, their MST kernels must win the freed SMs over the next contraction
        SCTDA_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->fork_stream, cudaStreamNonBlocking, prio_hi));
        for (int i = 0; i < 2; ++i) SCTDA_CUDA(ctx, cudaEventCreateWithFlags(&ctx->fork_ev[i], cudaEventDisableTiming));
    }
    if (!ctx->aux_stream) {
        SCTDA_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 4; ++i) SCTDA_CUDA(ctx, cudaEventCreateWithFlags(&ctx->pipe_ev[i], cudaEventDisableTiming));
    }

    // batches (host only)
    struct Batch { size_t pos, end, elems; int max_lev; std::vector<int64_t> tiles, goff, chunks; };
    std::vector<Batch> batches;
    for (size_t pos = 0; pos < order.size();) {
        Batch bt;
        bt.pos = pos; bt.elems = 0; bt.max_lev = 0;
        bt.tiles.assign(1, 0);
        bt.goff.assign(1, 0);
        bt.chunks.assign(1, 0);
        size_t end = pos;
        while (end < order.size()) {
            const int64_t m = h_off[order[end] + 1] - h_off[order[end]];
            if (end > pos && (bt.elems + (size_t)(m * m)) * 8 > budget) break;
            bt.elems += (size_t)(m * m);
            const int64_t tile = sctda_gemm_tile(ctx->precision);
            const int64_t q = (m + tile - 1) / tile;
            bt.tiles.push_back(bt.tiles.back() + q * (q + 1) / 2);            // upper-triangular tiles of the contraction kernel
            bt.goff.push_back((int64_t)bt.elems);
            bt.chunks.push_back(bt.chunks.back() + (m + kColChunk - 1) / kColChunk);     // column chunks of the T kernel
            if (m * (max_K - 1) <= a.lev_cap && m > bt.max_lev) bt.max_lev = (int)m;
            ++end;
        }
        bt.end = end;
        batches.push_back(std::move(bt));
        pos = end;
    }
    // Two-stream pipeline (contraction of batch i+1 over the MST / selection of batch i).  Measured on a B200 with the
    // three-CTA-per-SM contraction: the 7-batch 100k x 2k build on one GPU goes from 1,000 to 970 ms (the contraction
    // slows from 824 to 861 ms next to the MST kernels, the MST chains disappear behind it); with one or two batches
    // the cross-stream hand-over only adds jitter (up to +30 % on single-batch builds), so the pipeline runs from
    // three batches on (SCTDA_PIPELINE=0 / 1 forces it off / on from two batches) and the default otherwise runs
    // every kernel on the caller's stream.
    const size_t pipe_min = pipe_env ? (atoi(pipe_env) != 0 ? 2 : (size_t)-1) : 3;
    const bool pipe = linkage == 0 && batches.size() >= pipe_min;
    cudaStream_t aux_st = pipe ? ctx->aux_stream : st;
    // allocate both buffers up front (no reallocation while the pipeline runs)
    size_t need0 = 0, need1 = 0, meta0 = 0, meta1 = 0;
    for (size_t i = 0; i < batches.size(); ++i) {
        const size_t nbb = batches[i].end - batches[i].pos;
        const size_t mb = (nbb + 1) * 24 + nbb * 4 + nbb * kMaxK * 4 + 64;
        if (i % 2 == 0) { need0 = std::max(need0, batches[i].elems * 8); meta0 = std::max(meta0, mb); }
        else { need1 = std::max(need1, batches[i].elems * 8); meta1 = std::max(meta1, mb); }
    }
    double* gram_buf[2] = {(double*)sctda_ws(ctx, WS_CLUSTER2, need0), need1 ? (double*)sctda_ws(ctx, WS_CLUSTER4, need1) : nullptr};
    unsigned char* meta_buf[2] = {(unsigned char*)sctda_ws(ctx, WS_CLUSTER, meta0),
                                  meta1 ? (unsigned char*)sctda_ws(ctx, WS_CLUSTER5, meta1) : nullptr};
    if (!gram_buf[0] || !meta_buf[0] || (need1 && (!gram_buf[1] || !meta_buf[1]))) return SCTDA_ERR_NOMEM;
    double* dmat_buf = nullptr;
    if (linkage != 0) {                            // one working matrix buffer, as large as the largest batch
        dmat_buf = (double*)sctda_ws(ctx, WS_LINK, std::max(need0, need1));
        if (!dmat_buf) return SCTDA_ERR_NOMEM;
    }
    for (size_t i = 0; i < batches.size(); ++i) {
        const Batch& bt = batches[i];
        const int k = (int)(i & 1);
        const int nb = (int)(bt.end - bt.pos);
        if (pipe && i >= 2) SCTDA_CUDA(ctx, cudaStreamWaitEvent(st, ctx->pipe_ev[2 + k], 0));   // buffer k is free again
        int64_t* d_tiles = (int64_t*)meta_buf[k];
        int64_t* d_goff = d_tiles + (nb + 1);
        int64_t* d_chunks = d_goff + (nb + 1);
        int32_t* d_prob = (int32_t*)(d_chunks + (nb + 1));
        a.fine_rep = d_prob + nb;
        SCTDA_CUDA(ctx, cudaMemcpyAsync(d_tiles, bt.tiles.data(), (size_t)(nb + 1) * 8, cudaMemcpyHostToDevice, st));
        SCTDA_CUDA(ctx, cudaMemcpyAsync(d_goff, bt.goff.data(), (size_t)(nb + 1) * 8, cudaMemcpyHostToDevice, st));
        SCTDA_CUDA(ctx, cudaMemcpyAsync(d_chunks, bt.chunks.data(), (size_t)(nb + 1) * 8, cudaMemcpyHostToDevice, st));
        SCTDA_CUDA(ctx, cudaMemcpyAsync(d_prob, order.data() + bt.pos, (size_t)nb * 4, cudaMemcpyHostToDevice, st));
        a.prob_bin = d_prob; a.gram_off = d_goff; a.gram = gram_buf[k];
        // MST size classes: patches are sorted by size; [0,n_huge) above 28,000 cells (generic kernel, state in
        // global memory), [n_huge,n_big) above 16,384 and [n_big,n_mid) above 4,096 (one cluster of 8
        // CTAs per patch, 8 or 4 register keys per thread), the rest one CTA with 8 register keys
        int n_huge = 0, n_big = 0, n_mid = 0, n_1k = 0, n_512 = 0;
        for (size_t j = bt.pos; j < bt.end; ++j) {
            const int m = h_off[order[j] + 1] - h_off[order[j]];
            if (m > thr_huge) ++n_huge;
            if (m > thr_big) ++n_big;
            if (m > thr_mid) ++n_mid;
            if (m > thr_1k) ++n_1k;
            if (m > thr_512) ++n_512;
        }
        // The dependent chain of a large patch's MST (tens of ms on 8 SMs) is hidden behind the contraction of
        // the other patches: the Gram tiles of the large patches are computed first, their cluster MST kernels
        // start on a second stream as soon as those tiles exist, and the contraction of the remaining patches
        // runs beside them on the other SMs (tiles claimed from a counter, so that CTAs which only find a free
        // SM when an MST kernel ends do not hold work back).
        // MEASURED on a rank's share of the 8-GPU 100k x 2k build (tools/mapper_rank.py, B200): 148.3 -> 148.0 ms and
        // 139.5 -> 137.4 ms.  The MST chain does disappear behind the contraction, but the contraction pays it back:
        // 111 -> 132 ms as short-lived CTAs sharing the SMs and the L2 with the MST kernels (a persistent launch on
        // the SMs left over, plus a helper launch afterwards, did no better: 152 ms).  Opt-in (SCTDA_OVERLAP=1).
        static const bool want_overlap = getenv("SCTDA_OVERLAP") && atoi(getenv("SCTDA_OVERLAP")) != 0;
        const bool overlap = want_overlap && !pipe && ctx->precision == 0 && n_mid > 0 && nb > n_mid;
        int rc;
        if (overlap) {
            unsigned long long* ctr = (unsigned long long*)sctda_ws(ctx, WS_MISC2, 256);
            if (!ctr) return SCTDA_ERR_NOMEM;
            SCTDA_CUDA(ctx, cudaMemsetAsync(ctr, 0, 2 * sizeof(unsigned long long), st));
            rc = sctda_internal_gram_bins(ctx, G, d_Xn, ldxn, nb, d_prob, d_bin_off, d_members, d_tiles, d_goff, gram_buf[k], st,
                                          x_rows, i > 0, 0, n_mid, ctr, 0, false, 0);
            if (rc) return rc;
            SCTDA_CUDA(ctx, cudaEventRecord(ctx->fork_ev[0], st));
            SCTDA_CUDA(ctx, cudaStreamWaitEvent(ctx->fork_stream, ctx->fork_ev[0], 0));
            // The contraction of the remaining patches runs as short-lived CTAs (eight tiles each, claimed from the
            // counter): SMs turn over every ~150 us, so the MST clusters of the higher-priority stream get theirs as
            // soon as the large patches' tiles exist, and the contraction takes them back when an MST kernel ends.
            const int64_t tiles2 = bt.tiles[nb] - bt.tiles[n_mid];
            rc = sctda_internal_gram_bins(ctx, G, d_Xn, ldxn, nb, d_prob, d_bin_off, d_members, d_tiles, d_goff, gram_buf[k], st,
                                          x_rows, true, n_mid, nb, ctr + 1, (int)std::max<int64_t>(1, (tiles2 + 7) / 8), false, 8);
            if (rc) return rc;
        } else {
            rc = sctda_internal_gram_bins(ctx, G, d_Xn, ldxn, nb, d_prob, d_bin_off, d_members, d_tiles, d_goff, gram_buf[k], st,
                                          x_rows, i > 0, 0, 0, nullptr, 0, false, 0);
            if (rc) return rc;
        }
        if (pipe) {
            SCTDA_CUDA(ctx, cudaEventRecord(ctx->pipe_ev[k], st));
            SCTDA_CUDA(ctx, cudaStreamWaitEvent(aux_st, ctx->pipe_ev[k], 0));
        }
        // Without the overlap: the cluster launches (few SMs, long dependent chains) and the one-CTA launch (all
        // other patches) still run side by side: the latter is forked onto the second stream, joined before the
        // selection kernels.  With it the roles are swapped: the cluster launches are the ones on the second stream.
        static const bool no_fork = getenv("SCTDA_NO_FORK") && atoi(getenv("SCTDA_NO_FORK")) != 0;
        const bool fork = overlap || (!no_fork && n_mid > 0 && nb > n_mid);
        cudaStream_t reg_st = aux_st, clu_st = aux_st;
        if (overlap) {
            clu_st = ctx->fork_stream;
        } else if (fork) {
            SCTDA_CUDA(ctx, cudaEventRecord(ctx->fork_ev[0], aux_st));
            SCTDA_CUDA(ctx, cudaStreamWaitEvent(ctx->fork_stream, ctx->fork_ev[0], 0));
            reg_st = ctx->fork_stream;
        }
        if (linkage != 0) {                           // average / Ward / complete: one nearest-neighbour chain per patch
            a.prob0 = 0;
            a.dmat = dmat_buf;
            linkage_kernel<<<nb, kLinkThreads, 0, aux_st>>>(a);
            SCTDA_LAUNCH_CHECK(ctx, "linkage_kernel");
            n_huge = n_big = n_mid = n_1k = n_512 = nb;         // no MST launches below
        }
        if (n_huge && linkage == 0) {
            a.prob0 = 0;
            prim_kernel<<<n_huge, 256, 16, clu_st>>>(a);
            SCTDA_LAUNCH_CHECK(ctx, "prim_kernel");
        }
        if (n_big > n_huge) {
            a.prob0 = n_huge;
            const int mm = h_off[order[bt.pos + n_huge] + 1] - h_off[order[bt.pos + n_huge]];
            prim_cluster_kernel<8><<<(n_big - n_huge) * kClusterCtas, 512, (size_t)mm * 8 + 16, clu_st>>>(a);
            SCTDA_LAUNCH_CHECK(ctx, "prim_cluster_kernel<8>");
        }
        if (n_mid > n_big) {
            a.prob0 = n_big;
            const int mm = h_off[order[bt.pos + n_big] + 1] - h_off[order[bt.pos + n_big]];
            prim_cluster_kernel<4><<<(n_mid - n_big) * kClusterCtas, 512, (size_t)mm * 8 + 16, clu_st>>>(a);
            SCTDA_LAUNCH_CHECK(ctx, "prim_cluster_kernel<4>");
        }
        // One CTA of 512 threads per patch, keys in registers, ONE launch for every patch up to 4,096 cells:
        // the launch lasts as long as the dependent chain of its largest patch, and the small patches
        // fill the other SMs meanwhile.  (Measured: separate launches with 128 / 256-thread CTAs for the
        // small patches run those 3x faster but serialise behind the large-patch launch: 7.4 -> 9.5 ms at
        // config 2; the smaller instantiations stay available through the test hook.)
        if (n_1k > n_mid || (nb > n_mid && !split_small)) {
            a.prob0 = n_mid;
            const int last = split_small ? n_1k : nb;
            const int mm = h_off[order[bt.pos + n_mid] + 1] - h_off[order[bt.pos + n_mid]];
            prim_reg_kernel<8, 512><<<last - n_mid, 512, (size_t)mm * 12 + 16, reg_st>>>(a);
            SCTDA_LAUNCH_CHECK(ctx, "prim_reg_kernel<8,512>");
        }
        if (split_small && n_512 > n_1k) {
            a.prob0 = n_1k;
            const int mm = h_off[order[bt.pos + n_1k] + 1] - h_off[order[bt.pos + n_1k]];
            prim_reg_kernel<4, 256><<<n_512 - n_1k, 256, (size_t)mm * 12 + 16, reg_st>>>(a);
            SCTDA_LAUNCH_CHECK(ctx, "prim_reg_kernel<4,256>");
        }
        if (split_small && nb > n_512) {
            a.prob0 = n_512;
            const int mm = h_off[order[bt.pos + n_512] + 1] - h_off[order[bt.pos + n_512]];
            prim_reg_kernel<4, 128><<<nb - n_512, 128, (size_t)mm * 12 + 16, reg_st>>>(a);
            SCTDA_LAUNCH_CHECK(ctx, "prim_reg_kernel<4,128>");
        }
        if (fork) {                                   // the selection kernels need every MST
            SCTDA_CUDA(ctx, cudaEventRecord(ctx->fork_ev[1], ctx->fork_stream));
            SCTDA_CUDA(ctx, cudaStreamWaitEvent(aux_st, ctx->fork_ev[1], 0));
        }
        a.prob0 = 0;
        labels_kernel<<<nb, kSelThreads, (size_t)bt.max_lev * (max_K - 1) + 16, aux_st>>>(a);
        SCTDA_LAUNCH_CHECK(ctx, "labels_kernel");
        if (stat == 0) {
            tcol_kernel<<<(unsigned)bt.chunks.back(), kColChunk, 0, aux_st>>>(a, nb, d_chunks);
            SCTDA_LAUNCH_CHECK(ctx, "tcol_kernel");
            db_kernel<<<nb, kSelThreads, 0, aux_st>>>(a);
            SCTDA_LAUNCH_CHECK(ctx, "db_kernel");
        }
        if (pipe) SCTDA_CUDA(ctx, cudaEventRecord(ctx->pipe_ev[2 + k], aux_st));
    }
    // the caller's stream continues only after every batch is done
    if (pipe) {
        SCTDA_CUDA(ctx, cudaStreamWaitEvent(st, ctx->pipe_ev[2], 0));
        SCTDA_CUDA(ctx, cudaStreamWaitEvent(st, ctx->pipe_ev[3], 0));
    }
    return SCTDA_OK;
}

extern "C" int32_t sctda_nodes_count(sctda_ctx* ctx, int32_t nbins, const int32_t* d_bin_K, int32_t* d_node_base,
                                     void* stream) {
    if (!ctx) return SCTDA_ERR_INVALID;
    if (nbins < 0 || !d_bin_K || !d_node_base) return sctda_fail(ctx, SCTDA_ERR_INVALID, "nodes_count: bad arguments%s");
    SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
    scan_kernel<int32_t><<<1, 1024, 0, (cudaStream_t)stream>>>(d_bin_K, nbins, d_node_base);
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
    scan_kernel<int32_t><<<1, 1024, 0, st>>>(sizes, V, d_node_off);
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
    scan_kernel<int32_t><<<1, 1024, 0, st>>>(cursor, N, cell_off);
    SCTDA_LAUNCH_CHECK(ctx, "scan_kernel<cells>");
    SCTDA_CUDA(ctx, cudaMemsetAsync(cursor, 0, (size_t)N * 4, st));
    const unsigned vblocks = (unsigned)cdiv64((int64_t)V * 32, 256);
    cell_fill_kernel<<<vblocks, 256, 0, st>>>(V, d_node_off, d_node_members, cell_off, cursor, cell_nodes);
    SCTDA_LAUNCH_CHECK(ctx, "cell_fill_kernel");
    pair_kernel<<<(N + 127) / 128, 128, 0, st>>>(N, W, cell_off, cell_nodes, bits);
    SCTDA_LAUNCH_CHECK(ctx, "pair_kernel");
    edge_scan_kernel<false><<<vblocks, 256, 0, st>>>(V, W, bits, row_cnt, nullptr, nullptr);
    SCTDA_LAUNCH_CHECK(ctx, "edge_scan_kernel<count>");
    scan_kernel<int64_t><<<1, 1024, 0, st>>>(row_cnt, V, row_off);
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
