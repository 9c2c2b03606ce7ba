// Gene x gene Jensen-Shannon distances over the nodes:
// UnrootedGraph.JSD_matrix, /root/reference/scTDA/main.py:1111-1155.
//
//   jsd[i][j] = sqrt( 0.5 * (h_i + h_j - 2 * sum_k f(0.5 * (p_ik + p_jk))) ),  f(x) = x log2 x (f(0) = 0),
//   h_i = sum_k f(p_ik)
//
// where p_i is the normalised node profile of gene i (sctda_node_stats d_p).  The reference tiles
// a [genes, genes, nodes] array through numpy (chunked by maximum_matrix_entries); here one CTA
// computes a 16 x 16 block of gene pairs, staging 64 nodes of both gene blocks in shared memory;
// every pair is one sequential sum over the nodes (fixed order), FP64 log2.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int JT = 16, JK = 64;

__device__ __forceinline__ double xlog2x(double x) { return x > 0.0 ? x * log2(x) : 0.0; }

__global__ void __launch_bounds__(JT * JT) jsd_kernel(int g, int V, const double* __restrict__ P, int64_t ldp,
                                                      double* __restrict__ out, int64_t ldo) {
    __shared__ double sa[JT][JK + 1], sb[JT][JK + 1];
    const int tx = threadIdx.x & (JT - 1), ty = threadIdx.x / JT;
    const int i = blockIdx.y * JT + ty, j = blockIdx.x * JT + tx;
    double hi = 0.0, hj = 0.0, hm = 0.0;
    for (int k0 = 0; k0 < V; k0 += JK) {
        for (int t = threadIdx.x; t < JT * JK; t += JT * JT) {
            const int r = t / JK, c = t % JK;
            const int gi = blockIdx.y * JT + r, gj = blockIdx.x * JT + r;
            sa[r][c] = (gi < g && k0 + c < V) ? P[(int64_t)gi * ldp + k0 + c] : 0.0;
            sb[r][c] = (gj < g && k0 + c < V) ? P[(int64_t)gj * ldp + k0 + c] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int c = 0; c < JK; ++c) {
            const double a = sa[ty][c], b = sb[tx][c];
            hi += xlog2x(a);
            hj += xlog2x(b);
            hm += xlog2x(0.5 * (a + b));
        }
        __syncthreads();
    }
    if (i < g && j < g) out[(int64_t)i * ldo + j] = sqrt(0.5 * (hi + hj - 2.0 * hm));
}

}  // namespace

extern "C" int32_t sctda_jsd_matrix(sctda_ctx* ctx, int32_t g, int32_t V, const double* d_P, int64_t ldp, double* d_out,
                                    int64_t ldo, void* stream) {
    if (!ctx) return SCTDA_ERR_INVALID;
    if (g < 0 || V < 1 || !d_P || !d_out || ldp < V || ldo < g) return sctda_fail(ctx, SCTDA_ERR_INVALID, "jsd_matrix: bad arguments%s");
    if (g == 0) return SCTDA_OK;
    SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
    dim3 grid((g + JT - 1) / JT, (g + JT - 1) / JT);
    jsd_kernel<<<grid, JT * JT, 0, (cudaStream_t)stream>>>(g, V, d_P, ldp, d_out, ldo);
    SCTDA_LAUNCH_CHECK(ctx, "jsd_kernel");
    return SCTDA_OK;
}
