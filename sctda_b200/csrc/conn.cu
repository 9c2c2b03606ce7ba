// Hot path (b): the per-gene connectivity permutation test and the per-gene node statistics.
//
// Reference behaviour (CamaraLab/scTDA, scTDA/main.py, "ref:"):
//   connectivity_pvalue  ref:966-992     get_gene   ref:893-964     connectivity ref:994-1005
//   delta                ref:1007-1028   expr       ref:1030-1051   centroid     ref:1724-1743
//
// Design (B200, sm_100a).  The work is an index-driven gather followed by a small sparse
// product, all in float64 with one float32 rounding; there is no dense contraction, so the
// kernels are built for the memory system, not the tensor cores:
//
//   * layout: lanes of a warp <-> 32 consecutive GENES.  The expression table is transformed
//     once per call (2^x - 1, ref:982) and transposed into gene tiles ET[tile][cell][32], so
//     that every gathered cell is one fully used, aligned 256-byte row shared by the 32 lanes;
//     a gene tile (S*256 B) stays resident in the 126 MB L2 while all n permutations of it
//     are processed;
//   * a permutation enters only through the composite index perm[node_cols[m]]; it is loaded
//     coalesced (32 members at a time) and broadcast with warp shuffles, so the inner loop is
//     SHFL + LDG.64 + DADD with 8 independent loads in flight per lane;
//   * member sums are sequential in member-list order (the reference's order), one node per
//     warp at a time; the float32 node values (ref:972) go to a per-CTA scratch that stays
//     in L2, followed by the CSR product p'Ap and the comparison with the observed score;
//   * a persistent grid (a multiple of the SM count) walks (gene tile, permutation chunk)
//     items tile-major, so that concurrently running CTAs share gene tiles in L2.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "conn_math.cuh"

namespace {

constexpr int kLanes = 32;          // genes per tile
constexpr int kPermChunk = 8;       // permutations per work item

// ------------------------------------------------------------------------------------------
// ET[tile][s][lane] = f(Y[gene = tile*32+lane][samples[s]]),  f(x) = 2^x - 1 (log2 mode) or x.
// Genes beyond G are padded with zeros.  32x32 transpose through shared memory.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) et_prep_kernel(const double* __restrict__ Y, int64_t ldy, int G, int S,
                                                      const int32_t* __restrict__ samples, int log2_mode,
                                                      double* __restrict__ ET) {
    __shared__ double tile[32][33];
    const int tileIdx = blockIdx.y;
    const int s0 = blockIdx.x * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
    const int s = s0 + tx;
    int col = 0;
    if (s < S) col = samples ? samples[s] : s;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int gl = ty + 8 * r;
        int g = tileIdx * 32 + gl;
        double v = 0.0;
        if (g < G && s < S) {
            double x = Y[(int64_t)g * ldy + col];
            v = log2_mode ? (exp2(x) - 1.0) : x;
        }
        tile[gl][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int sl = ty + 8 * r;
        if (s0 + sl < S) ET[((int64_t)tileIdx * S + s0 + sl) * kLanes + tx] = tile[tx][sl];
    }
}

// expr (ref:1050): rows of the WHOLE table with value > 0.  One warp per gene.
__global__ void __launch_bounds__(256) expr_count_kernel(const double* __restrict__ Y, int64_t ldy, int G,
                                                         int Ncells, int32_t* __restrict__ out) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= G) return;
    const double* row = Y + (int64_t)warp * ldy;
    int c = 0;
    for (int i = lane; i < Ncells; i += 32) c += (row[i] > 0.0) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) out[warp] = c;
}

// Where the permutation of the current (gene tile, permutation) item is read from.
enum PermKind {
    PERM_NONE = 0,        // identity (observed statistics)
    PERM_GLOBAL = 1,      // one row shared by the 32 lanes, gathered from global memory
    PERM_SMEM = 2,        // one row shared by the 32 lanes, staged in shared memory by cp.async
    PERM_PER_LANE = 3,    // every lane (gene) has its own row in global memory (reference stream)
};

// Sequential member sum of one node for the 32 genes of a tile.
// Shared permutation: idx(m) = perm[cols[m]] is the same for all lanes: 32 composite indices are
// loaded at a time (coalesced cols, then the gather through perm) and broadcast by shuffles with
// immediate lane numbers.  `base` = tile pointer + lane, one opaque 64-bit value, so that every
// gather address is a single IMAD.WIDE: the inner loop is SHFL + IMAD.WIDE + LDG.64 + DADD.
template <int KIND>
__device__ __forceinline__ double node_sum(const double* base, const int32_t* __restrict__ cols,
                                           const int32_t* perm, int beg, int end, int lane) {
    double acc = 0.0;
    if (KIND != PERM_PER_LANE) {
        for (int b = beg; b < end; b += 32) {
            int m = b + lane;
            int my = 0;
            if (m < end) {
                my = __ldg(cols + m);
                if (KIND == PERM_GLOBAL) my = __ldg(perm + my);
                if (KIND == PERM_SMEM) my = perm[my];
            }
            const int cnt = end - b;                 // > 0
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
                if (cnt >= g8 * 8 + 8) {             // warp-uniform
                    double v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        int i = __shfl_sync(0xffffffffu, my, g8 * 8 + u);
                        v[u] = __ldg(base + (size_t)(unsigned)i * kLanes);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) acc += v[u];
                }
            }
            if (cnt < 32 && (cnt & 7)) {
                for (int t = cnt & ~7; t < cnt; ++t) {
                    int i = __shfl_sync(0xffffffffu, my, t);
                    acc += __ldg(base + (size_t)(unsigned)i * kLanes);
                }
            }
        }
    } else {
        int m = beg;
        for (; m + 4 <= end; m += 4) {
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int i = __ldg(perm + __ldg(cols + m + u));
                v[u] = __ldg(base + (size_t)(unsigned)i * kLanes);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) acc += v[u];
        }
        for (; m < end; ++m) {
            int i = __ldg(perm + __ldg(cols + m));
            acc += __ldg(base + (size_t)(unsigned)i * kLanes);
        }
    }
    return acc;
}

// ---- software-pipelined variant for the shared-permutation kernel ---------------------------
// composite index of member m of a node ending at `end` (0 beyond the end)
template <int KIND>
__device__ __forceinline__ int composite_index(const int32_t* __restrict__ cols, const int32_t* perm, int m, int end) {
    int my = 0;
    if (m < end) {
        my = __ldg(cols + m);
        if (KIND == PERM_GLOBAL) my = __ldg(perm + my);
        if (KIND == PERM_SMEM) my = perm[my];
    }
    return my;
}

#ifndef SCTDA_CONN_PIPELINE
#define SCTDA_CONN_PIPELINE 0
#endif
#define SCTDA_LOAD8(v, my, g8)                                                  \
    _Pragma("unroll") for (int u = 0; u < 8; ++u) {                             \
        const int i_ = __shfl_sync(0xffffffffu, my, (g8) * 8 + u);              \
        v[u] = __ldg(base + (size_t)(unsigned)i_ * kLanes);                     \
    }
#define SCTDA_ADD8(acc, v) _Pragma("unroll") for (int u = 0; u < 8; ++u) acc += v[u];

// Adds the `cnt` (1..32) members whose composite indices sit in the lanes of `my`, in member
// order, with the loads of the next group of eight issued before the adds of the current one
// (up to 16 gathers in flight per lane).
__device__ __forceinline__ double chunk_sum(const double* base, int my, int cnt, double acc) {
#if SCTDA_CONN_PIPELINE
    double va[8], vb[8];
    if (cnt >= 8) { SCTDA_LOAD8(va, my, 0) }
    if (cnt >= 16) { SCTDA_LOAD8(vb, my, 1) }
    if (cnt >= 8) { SCTDA_ADD8(acc, va) }
    if (cnt >= 24) { SCTDA_LOAD8(va, my, 2) }
    if (cnt >= 16) { SCTDA_ADD8(acc, vb) }
    if (cnt >= 32) { SCTDA_LOAD8(vb, my, 3) }
    if (cnt >= 24) { SCTDA_ADD8(acc, va) }
    if (cnt >= 32) { SCTDA_ADD8(acc, vb) }
#else
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) {
        if (cnt >= g8 * 8 + 8) {                      // warp-uniform
            double v[8];
            SCTDA_LOAD8(v, my, g8)
            SCTDA_ADD8(acc, v)
        }
    }
#endif
    if (cnt < 32 && (cnt & 7)) {
        for (int t = cnt & ~7; t < cnt; ++t) {
            const int i = __shfl_sync(0xffffffffu, my, t);
            acc += __ldg(base + (size_t)(unsigned)i * kLanes);
        }
    }
    return acc;
}

__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gmem_src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// 1/q per node (ref:982-986 divide the member sum by the node size q).
__global__ void inv_size_kernel(const int32_t* __restrict__ off, int V, double* __restrict__ inv_q) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < V) inv_q[k] = 1.0 / (double)(off[k + 1] - off[k]);
}

struct PermArgs {
    int V, S, G, n, log2_mode;
    int nchunks;                  // ceil(n / kPermChunk)
    int64_t items;                // tiles * nchunks
    const int32_t* node_off;
    const int32_t* node_cols;
    const int32_t* adj_off;
    const int32_t* adj_idx;
    const double* adj_w;
    const double* ET;
    const double* inv_q;
    const int32_t* perm;
    int64_t perm_gene_stride;
    const double* observed;
    double tie_rel;
    int32_t* exceed;
    int32_t* ties;
    double* scores;
    float* pm;                    // [gridDim.x][V][32]
    unsigned long long* counter;  // work queue head (zeroed before the launch)
};

// STAGE: 0 = permutation rows read from global memory (KIND decides how), 1 = the current row is
// copied to shared memory at the start of each permutation, 2 = double buffered: the next row is
// prefetched by cp.async while the current one is processed (S*8 bytes of dynamic shared memory).
template <int NW, int KIND, int STAGE>
__global__ void __launch_bounds__(NW * 32) conn_perm_kernel(PermArgs a) {
    __shared__ double s_tot[NW][kLanes];
    __shared__ double s_part[NW][kLanes];
    extern __shared__ int32_t s_perm[];              // [STAGE][S]
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    float* pm = a.pm + (size_t)blockIdx.x * a.V * kLanes;
    const double cor = (double)a.V / (double)(a.V - 1);          // ref:989
    int buf = 0;
    // Work items are claimed from a global queue, one ahead: all CTAs stay within ~2*gridDim
    // consecutive items, i.e. on one or two gene tiles, whatever their relative speed (a static
    // stride lets CTAs drift apart over many tiles on long runs and the tiles fall out of L2).
    __shared__ long long s_claim[2];
    if (threadIdx.x == 0) {
        s_claim[0] = (long long)atomicAdd(a.counter, 1ull);
        s_claim[1] = (long long)atomicAdd(a.counter, 1ull);
    }
    __syncthreads();
    int64_t item = s_claim[0], next_item = s_claim[1];
    if (STAGE == 2 && item < a.items) {                           // prologue: first row of this CTA
        const int r0 = (int)(item % a.nchunks) * kPermChunk;
        const int32_t* src = a.perm + (size_t)r0 * a.S;
        for (int i = threadIdx.x; i < a.S; i += NW * 32) cp_async_4(s_perm + i, src + i);
    }

    for (; item < a.items; item = next_item, next_item = s_claim[0]) {
        __syncthreads();                                          // everybody has read s_claim[0]
        if (threadIdx.x == 0) s_claim[0] = (long long)atomicAdd(a.counter, 1ull);   // visible after the next barrier
        const int tileIdx = (int)(item / a.nchunks);
        const int chunk = (int)(item % a.nchunks);
        const int g = tileIdx * kLanes + lane;
        const bool live = g < a.G;
        const double* base = a.ET + (size_t)tileIdx * a.S * kLanes + lane;
        asm volatile("" : "+l"(base));               // opaque: gather address = base + idx*256 in one IMAD.WIDE
        const double obs = live ? a.observed[g] : 0.0;
        int cnt_exceed = 0, cnt_tie = 0;
        const int r_end = min(a.n, (chunk + 1) * kPermChunk);
        for (int r = chunk * kPermChunk; r < r_end; ++r) {
            const int32_t* perm_r;
            if (KIND == PERM_PER_LANE) perm_r = a.perm + (size_t)(live ? g : 0) * a.perm_gene_stride + (size_t)r * a.S;
            else perm_r = a.perm + (size_t)r * a.S;
            if (STAGE == 1) {
                for (int i = threadIdx.x; i < a.S; i += NW * 32) cp_async_4(s_perm + i, perm_r + i);
                cp_async_commit_wait_all();
                __syncthreads();
                perm_r = s_perm;
            }
            if (STAGE == 2) {
                cp_async_commit_wait_all();          // this thread's part of the current row has landed
                __syncthreads();                     // ... and everybody else's
                int rn = r + 1;                      // the row this CTA processes next, if any
                if (rn >= r_end) rn = (next_item < a.items) ? (int)(next_item % a.nchunks) * kPermChunk : -1;
                if (rn >= 0) {
                    const int32_t* src = a.perm + (size_t)rn * a.S;
                    int32_t* dst = s_perm + (size_t)(buf ^ 1) * a.S;
                    for (int i = threadIdx.x; i < a.S; i += NW * 32) cp_async_4(dst + i, src + i);
                }
                perm_r = s_perm + (size_t)buf * a.S;
                buf ^= 1;
            }
            // ---- node values (ref:978-987) ----
            double tot_w = 0.0;
            if (KIND == PERM_PER_LANE) {
                for (int k = w; k < a.V; k += NW) {
                    int beg = __ldg(a.node_off + k), end = __ldg(a.node_off + k + 1);
                    double acc = node_sum<KIND>(base, a.node_cols, perm_r, beg, end, lane);
                    double t1 = div_rn(acc, (double)(end - beg), __ldg(a.inv_q + k));   // ref:982/986: sum / q
                    float pmv = a.log2_mode ? pm_log2_value(t1) : (float)t1;            // ref:983/986, float32 store
                    __stcg(pm + (size_t)k * kLanes + lane, pmv);
                    tot_w += (double)pmv;                                               // ref:984
                }
            } else {
                // the composite indices of the next node's first chunk are fetched while the current
                // node is summed, the second chunk's at the start of the node
                int k = w, beg = 0, end = 0, pref = 0;
                if (k < a.V) {
                    beg = __ldg(a.node_off + k);
                    end = __ldg(a.node_off + k + 1);
                    pref = composite_index<KIND>(a.node_cols, perm_r, beg + lane, end);
                }
                while (k < a.V) {
                    const int my0 = pref;
                    const int len = end - beg;
                    int my1 = 0;
                    if (len > 32) my1 = composite_index<KIND>(a.node_cols, perm_r, beg + 32 + lane, end);
                    const int kn = k + NW;
                    int begn = 0, endn = 0;
                    if (kn < a.V) {
                        begn = __ldg(a.node_off + kn);
                        endn = __ldg(a.node_off + kn + 1);
                        pref = composite_index<KIND>(a.node_cols, perm_r, begn + lane, endn);
                    }
                    double acc = chunk_sum(base, my0, min(32, len), 0.0);
                    if (len > 32) {
                        acc = chunk_sum(base, my1, min(32, len - 32), acc);
                        for (int b = beg + 64; b < end; b += 32)
                            acc = chunk_sum(base, composite_index<KIND>(a.node_cols, perm_r, b + lane, end), min(32, end - b), acc);
                    }
                    double t1 = div_rn(acc, (double)len, __ldg(a.inv_q + k));           // ref:982/986: sum / q
                    float pmv = a.log2_mode ? pm_log2_value(t1) : (float)t1;            // ref:983/986, float32 store
                    __stcg(pm + (size_t)k * kLanes + lane, pmv);
                    tot_w += (double)pmv;                                               // ref:984
                    k = kn; beg = begn; end = endn;
                }
            }
            s_tot[w][lane] = tot_w;
            __syncthreads();
            // ---- p'Ap on the un-normalised values; the 1/tot^2 is applied once (ref:988-989) ----
            double part_w = 0.0;
            for (int k = w; k < a.V; k += NW) {
                int eb = __ldg(a.adj_off + k), ee = __ldg(a.adj_off + k + 1);
                double acc = 0.0;
                for (int e = eb; e < ee; ++e) {
                    int l = __ldg(a.adj_idx + e);
                    double wt = __ldg(a.adj_w + e);
                    acc += wt * (double)__ldcg(pm + (size_t)l * kLanes + lane);
                }
                part_w += (double)__ldcg(pm + (size_t)k * kLanes + lane) * acc;
            }
            s_part[w][lane] = part_w;
            __syncthreads();
            if (w == 0) {
                double tot = 0.0, part = 0.0;
#pragma unroll
                for (int i = 0; i < NW; ++i) {
                    tot += s_tot[i][lane];
                    part += s_part[i][lane];
                }
                double conn = cor * (part / (tot * tot));
                if (live) {
                    const bool tied = fabs(conn - obs) <= a.tie_rel * fabs(obs);       // the same number up to rounding
                    if (conn > obs && !tied) ++cnt_exceed;                             // ref:990, strict: a tie does not exceed
                    if (tied) ++cnt_tie;
                    if (a.scores) a.scores[(size_t)g * a.n + r] = conn;
                }
            }
            __syncthreads();
        }
        if (w == 0 && live) {
            if (cnt_exceed) atomicAdd(a.exceed + g, cnt_exceed);
            if (a.ties && cnt_tie) atomicAdd(a.ties + g, cnt_tie);
        }
    }
}

// ------------------------------------------------------------------------------------------
// node statistics: get_gene / delta / connectivity / centroid for one gene tile per item.
// scratch P[cta][V][32] float64 holds pol_k, then p_k = pol_k / tol.
// ------------------------------------------------------------------------------------------
struct StatArgs {
    int V, S, G, log2_mode, tiles;
    const int32_t* node_off;
    const int32_t* node_cols;
    const int32_t* adj_off;
    const int32_t* adj_idx;
    const double* adj_w;
    const double* ET;
    const double* dist;
    double *tol, *mean, *mn, *mx, *conn, *cen, *disp, *p_out;
    double* P;
};

template <int NW>
__device__ __forceinline__ double block_sum(double v, double (*s)[kLanes], int w, int lane) {
    __syncthreads();
    s[w][lane] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < NW; ++i) t += s[i][lane];
    return t;
}

template <int NW>
__global__ void __launch_bounds__(NW * 32) node_stats_kernel(StatArgs a) {
    __shared__ double s_a[NW][kLanes];
    __shared__ double s_b[NW][kLanes];
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    double* P = a.P + (size_t)blockIdx.x * a.V * kLanes;
    const double cor = (double)a.V / (double)(a.V - 1);
    for (int tileIdx = blockIdx.x; tileIdx < a.tiles; tileIdx += gridDim.x) {
        const int g = tileIdx * kLanes + lane;
        const bool live = g < a.G;
        const double* base = a.ET + (size_t)tileIdx * a.S * kLanes + lane;
        // pol_k (ref:949-959)
        double tol_w = 0.0;
        for (int k = w; k < a.V; k += NW) {
            int beg = __ldg(a.node_off + k), end = __ldg(a.node_off + k + 1);
            double acc = node_sum<PERM_NONE>(base, a.node_cols, nullptr, beg, end, lane);
            double pol = acc / (double)(end - beg);
            if (a.log2_mode) pol = log2(1.0 + pol);
            __stcg(P + (size_t)k * kLanes + lane, pol);
            tol_w += pol;
        }
        const double tol = block_sum<NW>(tol_w, s_a, w, lane);                 // ref:960
        // p_k = pol_k / tol (ref:961-963), running sums for delta (ref:1013-1028)
        double sum_w = 0.0, mn_w = INFINITY, mx_w = -INFINITY, d1_w = 0.0;
        for (int k = w; k < a.V; k += NW) {
            double p = __ldcg(P + (size_t)k * kLanes + lane);
            if (tol > 0.0) p = p / tol;
            __stcg(P + (size_t)k * kLanes + lane, p);
            sum_w += p;
            mn_w = fmin(mn_w, p);
            mx_w = fmax(mx_w, p);
            if (a.dist) d1_w += __ldg(a.dist + k) * p;
        }
        const double psum = block_sum<NW>(sum_w, s_a, w, lane);
        const double pel1 = block_sum<NW>(d1_w, s_b, w, lane);
        __syncthreads();
        s_a[w][lane] = mn_w;
        s_b[w][lane] = mx_w;
        __syncthreads();
        double pmin = INFINITY, pmax = -INFINITY;
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            pmin = fmin(pmin, s_a[i][lane]);
            pmax = fmax(pmax, s_b[i][lane]);
        }
        // connectivity (ref:998-1005) and the dispersion sum (ref:1737-1741)
        const double cen = (psum > 0.0) ? pel1 / psum : NAN;
        double part_w = 0.0, d3_w = 0.0;
        for (int k = w; k < a.V; k += NW) {
            int eb = __ldg(a.adj_off + k), ee = __ldg(a.adj_off + k + 1);
            double acc = 0.0;
            for (int e = eb; e < ee; ++e)
                acc += __ldg(a.adj_w + e) * __ldcg(P + (size_t)__ldg(a.adj_idx + e) * kLanes + lane);
            double p = __ldcg(P + (size_t)k * kLanes + lane);
            part_w += p * acc;
            if (a.dist) {
                double d = __ldg(a.dist + k) - cen;
                d3_w += d * d * p;
            }
            if (a.p_out && live) a.p_out[(size_t)g * a.V + k] = p;
        }
        const double part = block_sum<NW>(part_w, s_a, w, lane);
        const double pel3 = block_sum<NW>(d3_w, s_b, w, lane);
        if (w == 0 && live) {
            const double pmean = psum / (double)a.V;
            const bool pos = pmean > 0.0;
            if (a.tol) a.tol[g] = tol;
            if (a.mean) a.mean[g] = pos ? pmean * tol : 0.0;
            if (a.mn) a.mn[g] = pos ? pmin * tol : 0.0;
            if (a.mx) a.mx[g] = pos ? pmax * tol : 0.0;
            if (a.conn) a.conn[g] = cor * part;
            if (a.cen) a.cen[g] = cen;
            if (a.disp) a.disp[g] = (psum > 0.0) ? sqrt(pel3 / psum) : NAN;
        }
        __syncthreads();
    }
}

// div_rn against the IEEE division for pseudo-random numerators and the divisors the hot path
// uses (node sizes 1..8192 and the constant 0.693147).
__global__ void selftest_div_kernel(int64_t n, uint64_t seed, unsigned long long* mismatches) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t x = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1);
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
    uint64_t mant = x & 0xFFFFFFFFFFFFFull;
    int ex = (int)((x >> 52) & 63) - 24;                       // 2^-24 .. 2^39
    double a = __longlong_as_double((long long)(((uint64_t)(1023 + ex) << 52) | mant));
    if ((x >> 60) == 0) a = 0.0;
    int q = (int)((x >> 40) % 8192) + 1;
    double b = (double)q;
    bool bad = __double_as_longlong(div_rn(a, b, 1.0 / b)) != __double_as_longlong(a / b);
    const double c = 0.693147;
    bad |= __double_as_longlong(div_rn(a, c, 1.0 / 0.693147)) != __double_as_longlong(a / c);
    if (bad) atomicAdd(mismatches, 1ull);
}

// log1p_pos against the generic log1p: largest difference in units of the last place
__global__ void selftest_log1p_kernel(int64_t n, uint64_t seed, unsigned long long* max_ulp) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t x = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1);
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
    const uint64_t mant = x & 0xFFFFFFFFFFFFFull;
    const int ex = (int)((x >> 52) % 100) - 60;                  // 2^-60 .. 2^39
    const double v = __longlong_as_double((long long)(((uint64_t)(1023 + ex) << 52) | mant));
    const long long a = __double_as_longlong(log1p_pos(v)), b = __double_as_longlong(log1p(v));
    const unsigned long long d = (unsigned long long)(a > b ? a - b : b - a);
    atomicMax(max_ulp, d);
}

int check_graph(sctda_ctx* ctx, const sctda_graph* gr) {
    if (!gr) return sctda_fail(ctx, SCTDA_ERR_INVALID, "graph is NULL%s");
    if (gr->V < 2) return sctda_fail(ctx, SCTDA_ERR_INVALID, "graph needs V >= 2 nodes (V/(V-1) factor, ref:989)%s");
    if (gr->S < 1 || gr->Ncells < 1) return sctda_fail(ctx, SCTDA_ERR_INVALID, "graph needs S >= 1 and Ncells >= 1%s");
    if (!gr->d_node_off || !gr->d_node_cols || !gr->d_adj_off || !gr->d_adj_idx || !gr->d_adj_w)
        return sctda_fail(ctx, SCTDA_ERR_INVALID, "graph CSR pointer is NULL%s");
    return SCTDA_OK;
}

int launch_et_prep(sctda_ctx* ctx, const sctda_graph* gr, int G, const double* dY, int64_t ldy, int log2_mode,
                   double** out_ET, cudaStream_t st) {
    const int tiles = (G + kLanes - 1) / kLanes;
    double* ET = (double*)sctda_ws(ctx, WS_ET, (size_t)tiles * gr->S * kLanes * sizeof(double));
    if (!ET) return SCTDA_ERR_NOMEM;
    dim3 grid((gr->S + 31) / 32, tiles), block(32, 8);
    et_prep_kernel<<<grid, block, 0, st>>>(dY, ldy, G, gr->S, gr->d_samples, log2_mode, ET);
    SCTDA_LAUNCH_CHECK(ctx, "et_prep_kernel");
    *out_ET = ET;
    return SCTDA_OK;
}

}  // namespace

extern "C" int32_t sctda_node_stats(sctda_ctx* ctx, const sctda_graph* gr, int32_t G, const double* d_Y,
                                    int64_t ldy, int32_t log2_mode, const double* d_dist, int32_t* d_expr,
                                    double* d_tol, double* d_mean, double* d_min, double* d_max, double* d_conn,
                                    double* d_cen, double* d_disp, double* d_p, void* stream) {
    if (!ctx) return SCTDA_ERR_INVALID;
    int rc = check_graph(ctx, gr);
    if (rc) return rc;
    if (G < 0 || (G > 0 && !d_Y) || ldy < gr->Ncells)
        return sctda_fail(ctx, SCTDA_ERR_INVALID, "node_stats: bad table (G < 0, NULL, or ldy < Ncells)%s");
    if (G == 0) return SCTDA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
    if (d_expr) {
        int blocks = (int)cdiv64((int64_t)G * 32, 256);
        expr_count_kernel<<<blocks, 256, 0, st>>>(d_Y, ldy, G, gr->Ncells, d_expr);
        SCTDA_LAUNCH_CHECK(ctx, "expr_count_kernel");
    }
    double* ET = nullptr;
    rc = launch_et_prep(ctx, gr, G, d_Y, ldy, log2_mode, &ET, st);
    if (rc) return rc;
    constexpr int NW = 8;
    StatArgs a;
    a.V = gr->V; a.S = gr->S; a.G = G; a.log2_mode = log2_mode;
    a.tiles = (G + kLanes - 1) / kLanes;
    a.node_off = gr->d_node_off; a.node_cols = gr->d_node_cols;
    a.adj_off = gr->d_adj_off; a.adj_idx = gr->d_adj_idx; a.adj_w = gr->d_adj_w;
    a.ET = ET; a.dist = d_dist;
    a.tol = d_tol; a.mean = d_mean; a.mn = d_min; a.mx = d_max; a.conn = d_conn;
    a.cen = d_dist ? d_cen : nullptr; a.disp = d_dist ? d_disp : nullptr; a.p_out = d_p;
    int grid = a.tiles < ctx->sm_count * 2 ? a.tiles : ctx->sm_count * 2;
    a.P = (double*)sctda_ws(ctx, WS_PM, (size_t)grid * gr->V * kLanes * sizeof(double));
    if (!a.P) return SCTDA_ERR_NOMEM;
    node_stats_kernel<NW><<<grid, NW * 32, 0, st>>>(a);
    SCTDA_LAUNCH_CHECK(ctx, "node_stats_kernel");
    return SCTDA_OK;
}

static int32_t conn_perm_test_impl(sctda_ctx* ctx, const sctda_graph* gr, int32_t G, const double* d_Y,
                                   int64_t ldy, int32_t log2_mode, int32_t n, const int32_t* d_perm,
                                   int64_t perm_gene_stride, const int32_t* d_order, const double* d_observed,
                                   double tie_rel, int32_t* d_exceed, int32_t* d_ties, double* d_scores,
                                   void* stream) {
    if (!ctx) return SCTDA_ERR_INVALID;
    SCTDA_TIME_RESET(ctx);
    int rc = check_graph(ctx, gr);
    if (rc) return rc;
    if (G < 0 || n < 0 || (G > 0 && !d_Y) || ldy < gr->Ncells)
        return sctda_fail(ctx, SCTDA_ERR_INVALID, "conn_perm_test: bad table (G/n < 0, NULL, or ldy < Ncells)%s");
    if (G == 0) return SCTDA_OK;
    if (!d_exceed || !d_observed || (n > 0 && !d_perm))
        return sctda_fail(ctx, SCTDA_ERR_INVALID, "conn_perm_test: d_exceed, d_observed and d_perm are required%s");
    if (perm_gene_stride != 0 && perm_gene_stride < (int64_t)n * gr->S)
        return sctda_fail(ctx, SCTDA_ERR_INVALID, "conn_perm_test: perm_gene_stride must be 0 or >= n*S%s");
    cudaStream_t st = (cudaStream_t)stream;
    SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
    SCTDA_CUDA(ctx, cudaMemsetAsync(d_exceed, 0, (size_t)G * sizeof(int32_t), st));
    if (d_ties) SCTDA_CUDA(ctx, cudaMemsetAsync(d_ties, 0, (size_t)G * sizeof(int32_t), st));
    if (n == 0) return SCTDA_OK;
    // one permutation block shared by all genes: the 64-gene sector-mask kernel (conn_tile64.cu)
    static const bool legacy = getenv("SCTDA_CONN_LEGACY") != nullptr;
    static const bool tile64 = getenv("SCTDA_CONN_TILE64") != nullptr;
    if (perm_gene_stride == 0 && !legacy && !tile64 && sctda_wide_eligible(gr))
        return sctda_conn_perm_wide(ctx, gr, G, d_Y, ldy, log2_mode, n, d_perm, d_order, d_observed, tie_rel,
                                    d_exceed, d_ties, d_scores, st);
    if (perm_gene_stride == 0 && !legacy && sctda_tile64_eligible(gr))
        return sctda_conn_perm_tile64(ctx, gr, G, d_Y, ldy, log2_mode, n, d_perm, d_order, d_observed, tie_rel,
                                      d_exceed, d_ties, d_scores, st);
    if (d_order) return sctda_fail(ctx, SCTDA_ERR_UNSUPPORTED, "conn_perm_test_ordered: a gene order needs the shared-permutation kernel (S <= 65535)%s");
    double* ET = nullptr;
    rc = launch_et_prep(ctx, gr, G, d_Y, ldy, log2_mode, &ET, st);
    if (rc) return rc;
    double* inv_q = (double*)sctda_ws(ctx, WS_MISC, (size_t)gr->V * sizeof(double) + 64);
    if (!inv_q) return SCTDA_ERR_NOMEM;
    unsigned long long* counter = (unsigned long long*)(inv_q + gr->V);
    SCTDA_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
    inv_size_kernel<<<(gr->V + 255) / 256, 256, 0, st>>>(gr->d_node_off, gr->V, inv_q);
    SCTDA_LAUNCH_CHECK(ctx, "inv_size_kernel");
#ifndef SCTDA_CONN_WARPS
#define SCTDA_CONN_WARPS 32
#endif
    constexpr int NW = SCTDA_CONN_WARPS;
    PermArgs a;
    a.V = gr->V; a.S = gr->S; a.G = G; a.n = n; a.log2_mode = log2_mode;
    a.nchunks = (n + kPermChunk - 1) / kPermChunk;
    const int tiles = (G + kLanes - 1) / kLanes;
    a.items = (int64_t)tiles * a.nchunks;
    a.node_off = gr->d_node_off; a.node_cols = gr->d_node_cols;
    a.adj_off = gr->d_adj_off; a.adj_idx = gr->d_adj_idx; a.adj_w = gr->d_adj_w;
    a.ET = ET; a.inv_q = inv_q; a.perm = d_perm; a.perm_gene_stride = perm_gene_stride;
    a.observed = d_observed; a.tie_rel = tie_rel > 0.0 ? tie_rel : 1e-11;
    a.exceed = d_exceed; a.ties = d_ties; a.scores = d_scores; a.counter = counter;
    int64_t grid64 = (int64_t)ctx->sm_count;          // one 1024-thread CTA per SM: the node-value scratch stays in L2
    if (grid64 > a.items) grid64 = a.items;
    const int grid = (int)grid64;
    a.pm = (float*)sctda_ws(ctx, WS_PM, (size_t)grid * gr->V * kLanes * sizeof(float));
    if (!a.pm) return SCTDA_ERR_NOMEM;
    // shared permutation rows are staged in shared memory when they fit (227 KB per CTA on sm_100a)
    const size_t row_bytes = (size_t)gr->S * sizeof(int32_t);
    const size_t smem_budget = 227 * 1024 - 2 * NW * kLanes * sizeof(double) - 1024;
    SCTDA_TIME_BEGIN(ctx, st);
    if (perm_gene_stride != 0) {
        conn_perm_kernel<NW, PERM_PER_LANE, 0><<<grid, NW * 32, 0, st>>>(a);
    } else if (2 * row_bytes <= smem_budget) {
        auto kern = conn_perm_kernel<NW, PERM_SMEM, 2>;
        SCTDA_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * row_bytes)));
        kern<<<grid, NW * 32, 2 * row_bytes, st>>>(a);
    } else if (row_bytes <= smem_budget) {
        auto kern = conn_perm_kernel<NW, PERM_SMEM, 1>;
        SCTDA_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_bytes));
        kern<<<grid, NW * 32, row_bytes, st>>>(a);
    } else {
        conn_perm_kernel<NW, PERM_GLOBAL, 0><<<grid, NW * 32, 0, st>>>(a);
    }
    SCTDA_TIME_END(ctx, st);
    SCTDA_LAUNCH_CHECK(ctx, "conn_perm_kernel");
    return SCTDA_OK;
}

extern "C" int32_t sctda_conn_perm_test(sctda_ctx* ctx, const sctda_graph* gr, int32_t G, const double* d_Y,
                                        int64_t ldy, int32_t log2_mode, int32_t n, const int32_t* d_perm,
                                        int64_t perm_gene_stride, const double* d_observed, double tie_rel,
                                        int32_t* d_exceed, int32_t* d_ties, double* d_scores, void* stream) {
    return conn_perm_test_impl(ctx, gr, G, d_Y, ldy, log2_mode, n, d_perm, perm_gene_stride, nullptr, d_observed, tie_rel,
                               d_exceed, d_ties, d_scores, stream);
}

extern "C" int32_t sctda_conn_perm_test_ordered(sctda_ctx* ctx, const sctda_graph* gr, int32_t G, const double* d_Y,
                                                int64_t ldy, int32_t log2_mode, int32_t n, const int32_t* d_perm,
                                                const int32_t* d_gene_order, const double* d_observed, double tie_rel,
                                                int32_t* d_exceed, int32_t* d_ties, double* d_scores, void* stream) {
    return conn_perm_test_impl(ctx, gr, G, d_Y, ldy, log2_mode, n, d_perm, 0, d_gene_order, d_observed, tie_rel,
                               d_exceed, d_ties, d_scores, stream);
}

extern "C" int32_t sctda_selftest_div(sctda_ctx* ctx, int64_t n_trials, int64_t* out_mismatches) {
    if (!ctx || !out_mismatches || n_trials < 0) return SCTDA_ERR_INVALID;
    SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned long long* d = (unsigned long long*)sctda_ws(ctx, WS_MISC, 256);
    if (!d) return SCTDA_ERR_NOMEM;
    SCTDA_CUDA(ctx, cudaMemset(d, 0, sizeof(unsigned long long)));
    if (n_trials > 0) {
        selftest_div_kernel<<<(unsigned)cdiv64(n_trials, 256), 256>>>(n_trials, 0x5c7da5eedull, d);
        SCTDA_LAUNCH_CHECK(ctx, "selftest_div_kernel");
    }
    unsigned long long h = 0;
    SCTDA_CUDA(ctx, cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost));
    *out_mismatches = (int64_t)h;
    return SCTDA_OK;
}

extern "C" int32_t sctda_selftest_log1p(sctda_ctx* ctx, int64_t n_trials, int64_t* out_max_ulp) {
    if (!ctx || !out_max_ulp || n_trials < 0) return SCTDA_ERR_INVALID;
    SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned long long* d = (unsigned long long*)sctda_ws(ctx, WS_MISC, 256);
    if (!d) return SCTDA_ERR_NOMEM;
    SCTDA_CUDA(ctx, cudaMemset(d, 0, sizeof(unsigned long long)));
    if (n_trials > 0) {
        selftest_log1p_kernel<<<(unsigned)cdiv64(n_trials, 256), 256>>>(n_trials, 0x10961ull, d);
        SCTDA_LAUNCH_CHECK(ctx, "selftest_log1p_kernel");
    }
    unsigned long long h = 0;
    SCTDA_CUDA(ctx, cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost));
    *out_max_ulp = (int64_t)h;
    return SCTDA_OK;
}
