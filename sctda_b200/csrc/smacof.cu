// MDS lens: the SMACOF iteration behind
// sakmapper.apply_lens(df, lens='mds', metric=...) -> sklearn.manifold.MDS, the reference's DEFAULT
// lens (/root/reference/scTDA/main.py:735, 746-751).  The N x N dissimilarity matrix comes from
// sctda_pairdist_rows_f64 and stays on the device; one pass over it per iteration:
//
//   for the current configuration Z [N, 2]:
//     stress(Z)  = sum_{i,j} (|z_i - z_j| - delta_ij)^2 / 2,   ssd(Z) = sum_{i,j} |z_i - z_j|^2
//     Z'_i       = (1/N) * ( (sum_j r_ij) z_i - sum_j r_ij z_j ),  r_ij = delta_ij / max(|z_i - z_j|, "0 -> 1e-5")
//
// i.e. the Guttman transform and the stress exactly as sklearn.manifold._mds._smacof_single
// evaluates them (metric MDS).  Bandwidth-bound: 8*N^2 bytes of dissimilarities per pass
// (3.2 GB at 20,000 cells), Z (16*N bytes) is served from L2.  One CTA per row block,
// fixed-order reductions (results do not depend on scheduling).
#include <math.h>

#include "common.cuh"

namespace {

constexpr int ST = 256;          // threads per CTA, one CTA per row

__device__ __forceinline__ double block_reduce(double v, double* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < ST / 32; ++i) t += s_red[i];
    return t;
}

__global__ void __launch_bounds__(ST) smacof_row_kernel(int N, const double* __restrict__ D, int64_t ldd,
                                                       const double2* __restrict__ Z, double2* __restrict__ Znew,
                                                       double* __restrict__ row_stress, double* __restrict__ row_ssd) {
    __shared__ double s_red[ST / 32];
    const int i = blockIdx.x;
    const double2 zi = Z[i];
    const double* drow = D + (int64_t)i * ldd;
    double sr = 0.0, sx = 0.0, sy = 0.0, st = 0.0, sd = 0.0;
#pragma unroll 4
    for (int j = threadIdx.x; j < N; j += ST) {
        const double2 zj = __ldg(Z + j);
        const double dx = zi.x - zj.x, dy = zi.y - zj.y;
        const double d2 = dx * dx + dy * dy;
        const double delta = __ldg(drow + j);
        // 1/sqrt(d2) from the single-precision estimate and two Newton steps (full double accuracy):
        // the generic sqrt + divide sequences made this bandwidth-bound pass FP64-issue-bound
        double dist, r;
        if (j != i && d2 > 1e-30 && d2 < 1e30) {
            double y = (double)rsqrtf((float)d2);
            y = y * fma(-0.5 * d2, y * y, 1.5);
            y = y * fma(-0.5 * d2, y * y, 1.5);
            dist = d2 * y;
            r = delta * y;
        } else {
            dist = (j == i) ? 0.0 : sqrt(d2);
            r = delta / (dist == 0.0 ? 1e-5 : dist);          // sklearn: distances[distances == 0] = 1e-5
        }
        const double e = dist - delta;
        st += e * e;
        sd += (j == i) ? 0.0 : d2;
        sr += r;
        sx += r * zj.x;
        sy += r * zj.y;
    }
    sr = block_reduce(sr, s_red);
    sx = block_reduce(sx, s_red);
    sy = block_reduce(sy, s_red);
    st = block_reduce(st, s_red);
    sd = block_reduce(sd, s_red);
    if (threadIdx.x == 0) {
        const double inv_n = 1.0 / (double)N;
        Znew[i] = make_double2(inv_n * (sr * zi.x - sx), inv_n * (sr * zi.y - sy));
        row_stress[i] = st;
        row_ssd[i] = sd;
    }
}

// stats[0] = sum(row_stress) / 2, stats[1] = sum(row_ssd); one CTA, fixed order
__global__ void __launch_bounds__(1024) smacof_stats_kernel(int N, const double* __restrict__ row_stress,
                                                           const double* __restrict__ row_ssd, double* __restrict__ stats) {
    __shared__ double s_a[1024], s_b[1024];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < N; i += 1024) { a += row_stress[i]; b += row_ssd[i]; }
    s_a[threadIdx.x] = a;
    s_b[threadIdx.x] = b;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s_a[threadIdx.x] += s_a[threadIdx.x + o]; s_b[threadIdx.x] += s_b[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { stats[0] = 0.5 * s_a[0]; stats[1] = s_b[0]; }
}

}  // namespace

extern "C" int32_t sctda_smacof_pass(sctda_ctx* ctx, int32_t N, const double* d_D, int64_t ldd, const double* d_Z,
                                     double* d_Znew, double* d_stats, void* stream) {
    if (!ctx) return SCTDA_ERR_INVALID;
    if (N < 1 || !d_D || ldd < N || !d_Z || !d_Znew || !d_stats || ((uintptr_t)d_Z & 15) || ((uintptr_t)d_Znew & 15))
        return sctda_fail(ctx, SCTDA_ERR_INVALID, "smacof_pass: bad arguments (Z must be 16-byte aligned)%s");
    cudaStream_t st = (cudaStream_t)stream;
    SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
    double* rows = (double*)sctda_ws(ctx, WS_MISC, (size_t)N * 16);
    if (!rows) return SCTDA_ERR_NOMEM;
    SCTDA_TIME_RESET(ctx);
    SCTDA_TIME_BEGIN(ctx, st);
    smacof_row_kernel<<<N, ST, 0, st>>>(N, d_D, ldd, (const double2*)d_Z, (double2*)d_Znew, rows, rows + N);
    SCTDA_TIME_END(ctx, st);
    SCTDA_LAUNCH_CHECK(ctx, "smacof_row_kernel");
    smacof_stats_kernel<<<1, 1024, 0, st>>>(N, rows, rows + N, d_stats);
    SCTDA_LAUNCH_CHECK(ctx, "smacof_stats_kernel");
    return SCTDA_OK;
}
