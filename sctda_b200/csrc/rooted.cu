// RootedGraph root finding (SURVEY.md §8f "next" row 1): RootedGraph.get_dendrite,
// /root/reference/scTDA/main.py:1463-1479, with get_distroot :1454-1461.
//
// For EVERY node i of the largest component the reference runs V single-pair shortest-path
// queries (networkx) to get the hop distances from i, drops the nodes at the maximal distance and
// takes minus the Spearman correlation between the node's mean time point and its distance.
// Here: one CTA per source node; level-synchronous BFS with the distance array in shared memory;
// average ranks of the time points through a prefix sum over their (precomputed) sorted order,
// ranks of the distances through a histogram; Pearson correlation of the ranks in float64.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int RT = 256;

__device__ __forceinline__ double block_sum_d(double v, double* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < RT / 32; ++i) t += s_red[i];
    return t;
}

// s_dist [V] int, s_pre [V+1] int (prefix over the sorted order of x), s_cnt [V+1] int (distance histogram)
__global__ void __launch_bounds__(RT) root_score_kernel(int V, const int32_t* __restrict__ adj_off,
                                                        const int32_t* __restrict__ adj_idx,
                                                        const int32_t* __restrict__ xorder,   // nodes by ascending x
                                                        const int32_t* __restrict__ gstart,   // tie group [gstart, gend) of each sorted position
                                                        const int32_t* __restrict__ gend,
                                                        double* __restrict__ score, int32_t* __restrict__ dist_out) {
    extern __shared__ int s_mem[];
    __shared__ double s_red[RT / 32];
    __shared__ int s_flag, s_max, s_n;
    int* s_dist = s_mem;
    int* s_pre = s_mem + V;
    int* s_cnt = s_pre + V + 1;
    const int src = blockIdx.x, tid = threadIdx.x;
    for (int q = tid; q < V; q += RT) s_dist[q] = (q == src) ? 0 : -1;
    __syncthreads();
    // ---- BFS, one level per sweep (ref:1454-1461) ----
    for (int level = 0;; ++level) {
        if (tid == 0) s_flag = 0;
        __syncthreads();
        for (int q = tid; q < V; q += RT) {
            if (s_dist[q] != level) continue;
            for (int e = adj_off[q]; e < adj_off[q + 1]; ++e) {
                const int nb = adj_idx[e];
                if (s_dist[nb] < 0) { s_dist[nb] = level + 1; s_flag = 1; }     // benign race: same value
            }
        }
        __syncthreads();
        if (!s_flag) { if (tid == 0) s_max = level; break; }
        __syncthreads();
    }
    __syncthreads();
    const int mx = s_max;
    if (dist_out)
        for (int q = tid; q < V; q += RT) dist_out[(int64_t)src * V + q] = s_dist[q];
    // ---- subset = reachable nodes with dist != mx (ref:1472-1476) ----
    // prefix of the inclusion flags over the sorted order of x (thread 0: V is a few thousand)
    for (int d = tid; d <= V; d += RT) s_cnt[d] = 0;
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int p = 0; p < V; ++p) {
            s_pre[p] = run;
            const int q = xorder[p];
            const int d = s_dist[q];
            if (d >= 0 && d != mx) { ++run; s_cnt[d + 1] += 1; }
        }
        s_pre[V] = run;
        s_n = run;
    }
    __syncthreads();
    const int n = s_n;
    if (n < 2) { if (tid == 0) score[src] = NAN; return; }
    // distance ranks: before[d] = number of subset nodes with a smaller distance
    // (s_cnt[d+1] holds the count of distance d; convert to exclusive prefix in place, thread 0)
    if (tid == 0) {
        int run = 0;
        for (int d = 0; d < mx; ++d) { const int c = s_cnt[d + 1]; s_cnt[d + 1] = run; run += c; }
    }
    __syncthreads();
    const double mean = 0.5 * ((double)n + 1.0);
    double sxy = 0.0, sxx = 0.0, syy = 0.0;
    for (int p = tid; p < V; p += RT) {
        const int q = xorder[p];
        const int d = s_dist[q];
        if (d < 0 || d == mx) continue;
        const int gs = gstart[p], ge = gend[p];
        const int before = s_pre[gs];
        const int cnt = s_pre[ge] - before;
        const double rx = (double)before + 0.5 * ((double)cnt + 1.0);
        // count of subset nodes at distance d = before[d+1] - before[d] (before[mx] would be n)
        const int bd = s_cnt[d + 1];
        const int nd = (d + 1 <= mx - 1 ? s_cnt[d + 2] : n) - bd;
        const double ry = (double)bd + 0.5 * ((double)nd + 1.0);
        const double ax = rx - mean, ay = ry - mean;
        sxy += ax * ay;
        sxx += ax * ax;
        syy += ay * ay;
    }
    sxy = block_sum_d(sxy, s_red);
    sxx = block_sum_d(sxx, s_red);
    syy = block_sum_d(syy, s_red);
    if (tid == 0) score[src] = (sxx > 0.0 && syy > 0.0) ? -(sxy / sqrt(sxx * syy)) : NAN;
}

}  // namespace

extern "C" int32_t sctda_root_scores(sctda_ctx* ctx, int32_t V, const int32_t* d_adj_off, const int32_t* d_adj_idx,
                                     const int32_t* d_xorder, const int32_t* d_gstart, const int32_t* d_gend,
                                     double* d_score, int32_t* d_dist, void* stream) {
    if (!ctx) return SCTDA_ERR_INVALID;
    if (V < 1 || !d_adj_off || !d_adj_idx || !d_xorder || !d_gstart || !d_gend || !d_score)
        return sctda_fail(ctx, SCTDA_ERR_INVALID, "root_scores: bad arguments%s");
    const size_t smem = ((size_t)3 * V + 4) * sizeof(int);
    if (smem > 200 * 1024) return sctda_fail(ctx, SCTDA_ERR_UNSUPPORTED, "root_scores: more than 17,000 nodes is not served%s");
    SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!(ctx->attr_done & 4u)) {
        SCTDA_CUDA(ctx, cudaFuncSetAttribute(root_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ctx->attr_done |= 4u;
    }
    root_score_kernel<<<V, RT, smem, (cudaStream_t)stream>>>(V, d_adj_off, d_adj_idx, d_xorder, d_gstart, d_gend, d_score, d_dist);
    SCTDA_LAUNCH_CHECK(ctx, "root_score_kernel");
    return SCTDA_OK;
}
