// PCA lens: sakmapper.apply_lens(df, lens='pca') as scTDA calls it (/root/reference/scTDA/main.py:745, 753;
// sklearn.decomposition.PCA(n_components=2).fit_transform behind it): the two leading principal component
// scores of the cells x genes table, on the device, without a library eigen-solver.
//
//   1. column means in a fixed order, centred table written transposed (genes x cells) when G <= N, or in
//      place of a copy (cells x genes) when N < G;
//   2. the smaller of the two Gram matrices (G x G covariance * (N-1), or N x N) by the library's own FP64
//      tensor-pipe contraction (dist.cu: upper triangle of tiles, mirrored);
//   3. its two leading eigenpairs by Chebyshev-filtered block subspace iteration (8 vectors).  One outer step:
//      Z = A Q (one bandwidth-bound pass over A), Rayleigh-Ritz on the 8 x 8 projected matrix (cyclic Jacobi),
//      exact residuals |A q - theta q| / theta of the two leading Ritz pairs (they stop the iteration through a
//      device flag), then the basis is filtered by a degree-16 Chebyshev polynomial in A that damps the interval
//      [0, theta_8] below the block (expression tables have a wide noise bulk right under the second component:
//      the plain iteration converges like (lambda_9 / lambda_2)^k, i.e. not at all) and re-orthogonalised
//      (Cholesky QR);
//   4. sklearn's sign rule (svd_flip on the components: the entry of largest magnitude is positive) and the
//      projection of the centred rows.
// Every reduction runs in a fixed order: the lens is reproducible bit for bit from run to run.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int kB = 8;                 // block size of the subspace iteration
constexpr int WS_PCA_T = 3;           // centred (transposed) table
constexpr int WS_PCA_A = 14;          // Gram matrix
constexpr int WS_PCA_V = 15;          // vectors, small matrices, flags

// ---- column means: partial sums over row blocks (fixed order), then the blocks in order ----------------
__global__ void col_partial_kernel(const double* __restrict__ X, int64_t ldx, int N, int G, int rows_per_block,
                                   double* __restrict__ part) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(N, r0 + rows_per_block);
    double s = 0.0;
    for (int r = r0; r < r1; ++r) s += X[(int64_t)r * ldx + g];
    part[(int64_t)blockIdx.y * G + g] = s;
}

__global__ void col_mean_kernel(const double* __restrict__ part, int nblk, int N, int G, double* __restrict__ mean) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += part[(int64_t)b * G + g];
    mean[g] = s / (double)N;
}

// T[g][n] = X[n][g] - mean[g] (TRANSPOSE = true, ldt >= N) or T[n][g] = X[n][g] - mean[g] (ldt >= G); padding zeroed.
template <bool TRANSPOSE>
__global__ void __launch_bounds__(256) centre_kernel(const double* __restrict__ X, int64_t ldx, int N, int G,
                                                     const double* __restrict__ mean, double* __restrict__ T, int64_t ldt) {
    __shared__ double tile[32][33];
    const int n0 = blockIdx.y * 32, g0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int n = n0 + ty + 8 * r, g = g0 + tx;
        double v = 0.0;
        if (n < N && g < G) v = X[(int64_t)n * ldx + g] - mean[g];
        if (TRANSPOSE) tile[ty + 8 * r][tx] = v;
        else if (n < N && g < ldt) T[(int64_t)n * ldt + g] = v;
    }
    if (!TRANSPOSE) return;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int g = g0 + ty + 8 * r, n = n0 + tx;
        if (g < G && n < ldt) T[(int64_t)g * ldt + n] = (n < N) ? tile[tx][ty + 8 * r] : 0.0;
    }
}

// ---- subspace iteration ---------------------------------------------------------------------------------
struct IterState {
    int done;                 // set when the two leading Ritz pairs have converged
    int iters;
    double theta[kB];
    double resid[2];
};

__global__ void init_q_kernel(double* __restrict__ Q, int M) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * kB) return;
    uint64_t x = 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1);
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
    Q[i] = (double)(x >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}

constexpr int kCheb = 16;            // degree of the Chebyshev filter

// Coefficients of filter step i (1..kCheb) for the damped interval [0, b], scaled to 1 at a0 (Zhou & Saad's
// recurrence): out = alpha * (A Y - c Y) - beta * Yprev.  step 0: the plain product out = A Y.
__device__ __forceinline__ void cheb_coeff(int step, double b, double a0, double& alpha, double& beta, double& c) {
    if (step == 0) { alpha = 1.0; beta = 0.0; c = 0.0; return; }
    const double e = 0.5 * b;
    c = 0.5 * b;
    const double sigma1 = e / (a0 - c);
    double sigma = sigma1;
    alpha = sigma1 / e;
    beta = 0.0;
    for (int i = 2; i <= step; ++i) {
        const double s2 = 1.0 / (2.0 / sigma1 - sigma);
        alpha = 2.0 * s2 / e;
        beta = sigma * s2;
        sigma = s2;
    }
}

// Z[i][0..7] = alpha * (sum_j A[i][j] Y[j][0..7] - c Y[i][0..7]) - beta * P[i][0..7]; one warp per row, lane-strided
// partial sums, xor butterfly.  step 0 is the plain product; steps >= 1 are skipped (Z = Y copied) when the
// Ritz values do not give a usable interval yet.
__global__ void __launch_bounds__(256) matmul8_kernel(const double* __restrict__ A, int64_t lda, int M,
                                                      const double* Y, const double* P, double* Z,
                                                      int step, const IterState* st) {
    if (st->done) return;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    double alpha, beta, c;
    const double b = st->theta[kB - 1], a0 = st->theta[0];
    bool usable = step == 0;
    if (step > 0 && st->iters >= 4 && b > 0.0 && a0 > b) {
        // The filter multiplies the leading direction by T_m(x1) and the second one by T_m(x2): beyond a ratio of
        // 1e6 the second direction drowns in the rounding of the first (the block would lose it), so the degree is
        // limited to m with (g(x1)/g(x2))^m <= 1e6, g(x) = x + sqrt(x^2 - 1) the growth factor of T_m per degree.
        const double x1 = (2.0 * a0 - b) / b, x2 = (2.0 * st->theta[1] - b) / b;
        const double g1 = x1 + sqrt(fmax(x1 * x1 - 1.0, 0.0)), g2 = x2 + sqrt(fmax(x2 * x2 - 1.0, 0.0));
        const double ratio = g1 / g2;
        int m = ratio <= 1.0000001 ? kCheb : (int)(13.815510557964274 / log(ratio));
        m = max(2, min(kCheb, m));
        usable = step <= m;
    }
    if (!usable) {                                                // no filter (yet, or degree reached): pass the basis through
        if (lane < kB) Z[(int64_t)row * kB + lane] = Y[(int64_t)row * kB + lane];
        return;
    }
    cheb_coeff(step, b, a0, alpha, beta, c);
    const double* a = A + (int64_t)row * lda;
    double acc[kB];
#pragma unroll
    for (int k = 0; k < kB; ++k) acc[k] = 0.0;
    for (int j = lane; j < M; j += 32) {
        const double av = a[j];
        const double4 q0 = *(const double4*)(Y + (int64_t)j * kB), q1 = *(const double4*)(Y + (int64_t)j * kB + 4);
        acc[0] = fma(av, q0.x, acc[0]); acc[1] = fma(av, q0.y, acc[1]); acc[2] = fma(av, q0.z, acc[2]); acc[3] = fma(av, q0.w, acc[3]);
        acc[4] = fma(av, q1.x, acc[4]); acc[5] = fma(av, q1.y, acc[5]); acc[6] = fma(av, q1.z, acc[6]); acc[7] = fma(av, q1.w, acc[7]);
    }
#pragma unroll
    for (int k = 0; k < kB; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if (lane == 0) {
        double out[kB];
#pragma unroll
        for (int k = 0; k < kB; ++k) {
            double v = acc[k];
            if (step > 0) {
                v = alpha * (v - c * Y[(int64_t)row * kB + k]);
                if (beta != 0.0) v -= beta * P[(int64_t)row * kB + k];
            }
            out[k] = v;
        }
        *(double4*)(Z + (int64_t)row * kB) = make_double4(out[0], out[1], out[2], out[3]);
        *(double4*)(Z + (int64_t)row * kB + 4) = make_double4(out[4], out[5], out[6], out[7]);
    }
}

// sum over the block of `n` values per thread, fixed order (threads in order inside a warp tree, warps in order)
template <int NV>
__device__ void block_sums(double* v, double* s_red /* [32][NV] */, double* out /* [NV] shared */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) s_red[w * NV + i] = x;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double t = 0.0;
        for (int k = 0; k < nw; ++k) t += s_red[k * NV + threadIdx.x];
        out[threadIdx.x] = t;
    }
    __syncthreads();
}

// cyclic Jacobi on the symmetric 8 x 8 matrix H: eigenvalues (descending) in th, eigenvectors in the columns of W
__device__ void jacobi8(double H[kB][kB], double W[kB][kB], double th[kB]) {
    for (int i = 0; i < kB; ++i)
        for (int j = 0; j < kB; ++j) W[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0, dia = 0.0;
        for (int i = 0; i < kB; ++i) {
            dia += H[i][i] * H[i][i];
            for (int j = i + 1; j < kB; ++j) off += H[i][j] * H[i][j];
        }
        if (off <= 1e-34 * dia) break;
        for (int p = 0; p < kB - 1; ++p)
            for (int q = p + 1; q < kB; ++q) {
                if (H[p][q] == 0.0) continue;
                const double tau = (H[q][q] - H[p][p]) / (2.0 * H[p][q]);
                const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                const double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
                for (int k = 0; k < kB; ++k) {
                    const double hkp = H[k][p], hkq = H[k][q];
                    H[k][p] = c * hkp - s * hkq;
                    H[k][q] = s * hkp + c * hkq;
                }
                for (int k = 0; k < kB; ++k) {
                    const double hpk = H[p][k], hqk = H[q][k];
                    H[p][k] = c * hpk - s * hqk;
                    H[q][k] = s * hpk + c * hqk;
                }
                for (int k = 0; k < kB; ++k) {
                    const double wkp = W[k][p], wkq = W[k][q];
                    W[k][p] = c * wkp - s * wkq;
                    W[k][q] = s * wkp + c * wkq;
                }
            }
    }
    int idx[kB];
    for (int i = 0; i < kB; ++i) idx[i] = i;
    for (int i = 0; i < kB; ++i)                                   // selection sort, descending
        for (int j = i + 1; j < kB; ++j)
            if (H[idx[j]][idx[j]] > H[idx[i]][idx[i]]) { int t = idx[i]; idx[i] = idx[j]; idx[j] = t; }
    double Ws[kB][kB];
    for (int c = 0; c < kB; ++c) {
        th[c] = H[idx[c]][idx[c]];
        for (int k = 0; k < kB; ++k) Ws[k][c] = W[k][idx[c]];
    }
    for (int i = 0; i < kB; ++i)
        for (int j = 0; j < kB; ++j) W[i][j] = Ws[i][j];
}

// One CTA.  Given orthonormal Q and Z = A Q: Rayleigh-Ritz (H = Q'Z, H = W diag(theta) W'), Ritz vectors
// Qr = Q W and their images Zr = Z W, exact residuals of the two leading pairs, and the next basis
// Q <- Zr R^-1 with R'R = Zr'Zr (Cholesky).  When the residuals are below tol the Ritz vectors are kept
// in Q and `done` is set.
__global__ void __launch_bounds__(256) ritz_kernel(double* Q, const double* Z, int M, double tol, IterState* st) {
    if (st->done) return;
    const bool ortho_only = tol < 0.0;                             // Q <- orthonormalised Z, nothing else
    __shared__ double s_red[32 * 72];
    __shared__ double s_out[72];
    __shared__ double sW[kB][kB], sT[kB][kB], sTh[kB];
    __shared__ int s_fail;
    // pass 1: H = Q'Z (upper 36) and S0 = Z'Z (upper 36)
    double v[72];
#pragma unroll
    for (int i = 0; i < 72; ++i) v[i] = 0.0;
    for (int r = threadIdx.x; r < M; r += blockDim.x) {
        double q[kB], z[kB];
#pragma unroll
        for (int c = 0; c < kB; ++c) { z[c] = Z[(int64_t)r * kB + c]; q[c] = ortho_only ? z[c] : Q[(int64_t)r * kB + c]; }
        int t = 0;
#pragma unroll
        for (int a = 0; a < kB; ++a)
#pragma unroll
            for (int b = a; b < kB; ++b) {
                v[t] = fma(0.5, fma(q[a], z[b], q[b] * z[a]), v[t]);       // symmetrised projection
                v[36 + t] = fma(z[a], z[b], v[36 + t]);
                ++t;
            }
    }
    block_sums<72>(v, s_red, s_out);
    // the 8 x 8 algebra of thread 0 works on shared arrays: with the matrices on the local stack of a kernel that
    // already spills its 72 accumulators, ptxas 12.9 produced wrong Ritz values (the same source is correct on the host)
    __shared__ double H[kB][kB], S0[kB][kB], W[kB][kB], th[kB], S[kB][kB], tmp[kB][kB], R[kB][kB], Ri[kB][kB];
    if (threadIdx.x == 0) {
        int t = 0;
        for (int a = 0; a < kB; ++a)
            for (int b = a; b < kB; ++b) {
                H[a][b] = H[b][a] = s_out[t];
                S0[a][b] = S0[b][a] = s_out[36 + t];
                ++t;
            }
        jacobi8(H, W, th);
        // S = W' S0 W = Zr'Zr, Cholesky S = R'R (R upper), T = W R^-1: Q_next = Z T
        for (int i = 0; i < kB; ++i)
            for (int j = 0; j < kB; ++j) {
                double s = 0.0;
                for (int k = 0; k < kB; ++k) s += S0[i][k] * W[k][j];
                tmp[i][j] = s;
            }
        for (int i = 0; i < kB; ++i)
            for (int j = 0; j < kB; ++j) {
                double s = 0.0;
                for (int k = 0; k < kB; ++k) s += W[k][i] * tmp[k][j];
                S[i][j] = s;
            }
        int fail = 0;
        double smax = 0.0;
        for (int i = 0; i < kB; ++i) smax = fmax(smax, S[i][i]);
        for (int i = 0; i < kB; ++i)
            for (int j = 0; j < kB; ++j) R[i][j] = 0.0;
        for (int j = 0; j < kB; ++j) {
            double d = S[j][j];
            for (int k = 0; k < j; ++k) d -= R[k][j] * R[k][j];
            if (!(d > 1e-28 * smax)) d = smax > 0.0 ? smax : 1.0;  // a direction the block has lost (rank < 8): that column dies out
            R[j][j] = sqrt(d);
            for (int i = j + 1; i < kB; ++i) {
                double s = S[j][i];
                for (int k = 0; k < j; ++k) s -= R[k][j] * R[k][i];
                R[j][i] = s / R[j][j];
            }
        }
        for (int i = 0; i < kB; ++i)                                // Ri = R^-1 (upper triangular)
            for (int j = 0; j < kB; ++j) Ri[i][j] = 0.0;
        for (int j = 0; j < kB; ++j) {
            Ri[j][j] = 1.0 / R[j][j];
            for (int i = j - 1; i >= 0; --i) {
                double s = 0.0;
                for (int k = i + 1; k <= j; ++k) s += R[i][k] * Ri[k][j];
                Ri[i][j] = -s / R[i][i];
            }
        }
        for (int i = 0; i < kB; ++i)
            for (int j = 0; j < kB; ++j) {
                double s = 0.0;
                for (int k = 0; k < kB; ++k) s += W[i][k] * Ri[k][j];
                sT[i][j] = s;
                sW[i][j] = W[i][j];
            }
        for (int c = 0; c < kB; ++c) sTh[c] = th[c];
        s_fail = fail;
    }
    __syncthreads();
    // pass 2: residuals of the two leading Ritz pairs, |Z w_c - theta_c Q w_c|^2
    double rs[2] = {0.0, 0.0};
    for (int r = threadIdx.x; r < M; r += blockDim.x) {
        double q[kB], z[kB];
#pragma unroll
        for (int c = 0; c < kB; ++c) { q[c] = Q[(int64_t)r * kB + c]; z[c] = Z[(int64_t)r * kB + c]; }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            double zr = 0.0, qr = 0.0;
#pragma unroll
            for (int k = 0; k < kB; ++k) { zr = fma(z[k], sW[k][c], zr); qr = fma(q[k], sW[k][c], qr); }
            const double d = zr - sTh[c] * qr;
            rs[c] = fma(d, d, rs[c]);
        }
    }
    block_sums<2>(rs, s_red, s_out);
    const double r0 = sqrt(s_out[0]) / fabs(sTh[0]), r1 = sqrt(s_out[1]) / fabs(sTh[1]);
    const bool conv = tol >= 0.0 && ((r0 <= tol && r1 <= tol) || s_fail);
    __syncthreads();
    // pass 3: Q <- Ritz vectors (converged) or the re-orthogonalised images
    for (int r = threadIdx.x; r < M; r += blockDim.x) {
        double src[kB], o[kB];
#pragma unroll
        for (int c = 0; c < kB; ++c) src[c] = conv ? Q[(int64_t)r * kB + c] : Z[(int64_t)r * kB + c];
#pragma unroll
        for (int c = 0; c < kB; ++c) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < kB; ++k) s = fma(src[k], conv ? sW[k][c] : sT[k][c], s);
            o[c] = s;
        }
#pragma unroll
        for (int c = 0; c < kB; ++c) Q[(int64_t)r * kB + c] = o[c];
    }
    if (threadIdx.x == 0 && tol >= 0.0) {                          // (tol < 0: the call only orthonormalises Z into Q)
        st->iters += 1;
        for (int c = 0; c < kB; ++c) st->theta[c] = sTh[c];
        st->resid[0] = r0;
        st->resid[1] = r1;
        if (conv) st->done = s_fail ? 2 : 1;
    }
}

// Not converged within max_iter: the last basis is rotated to its Ritz vectors by one more ritz step; here only
// the two leading columns are extracted and normalised.  comp[c][i] = column c of Q (M entries).
__global__ void extract_kernel(const double* __restrict__ Q, int M, double* __restrict__ vec) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    vec[i] = Q[(int64_t)i * kB];
    vec[M + i] = Q[(int64_t)i * kB + 1];
}

// comp[c][g] = sum_n Xc[n][g] u_c[n] (dual form: components from the left vectors, un-normalised)
__global__ void dual_components_kernel(const double* __restrict__ Xc, int64_t ld, int N, int G, const double* __restrict__ u,
                                       double* __restrict__ comp) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    double s0 = 0.0, s1 = 0.0;
    for (int n = 0; n < N; ++n) {
        const double x = Xc[(int64_t)n * ld + g];
        s0 = fma(x, u[n], s0);
        s1 = fma(x, u[N + n], s1);
    }
    comp[g] = s0;
    comp[G + g] = s1;
}

// sign[c] = sign of the entry of largest magnitude of comp[c] (first such entry), one CTA
__global__ void __launch_bounds__(1024) sign_kernel(const double* __restrict__ comp, int G, double* __restrict__ sign) {
    __shared__ double s_val[1024];
    __shared__ int s_idx[1024];
    for (int c = 0; c < 2; ++c) {
        double best = -1.0;
        int bi = 0x7fffffff;
        for (int g = threadIdx.x; g < G; g += blockDim.x) {
            const double a = fabs(comp[(int64_t)c * G + g]);
            if (a > best) { best = a; bi = g; }
        }
        s_val[threadIdx.x] = best;
        s_idx[threadIdx.x] = bi;
        __syncthreads();
        for (int o = blockDim.x / 2; o > 0; o >>= 1) {
            if (threadIdx.x < o) {
                const double a = s_val[threadIdx.x + o];
                const int i = s_idx[threadIdx.x + o];
                if (a > s_val[threadIdx.x] || (a == s_val[threadIdx.x] && i < s_idx[threadIdx.x])) {
                    s_val[threadIdx.x] = a;
                    s_idx[threadIdx.x] = i;
                }
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) sign[c] = (s_idx[0] < G && comp[(int64_t)c * G + s_idx[0]] < 0.0) ? -1.0 : 1.0;
        __syncthreads();
    }
}

// lens[n][c] = sign[c] * sum_g (X[n][g] - mean[g]) v_c[g]; one warp per cell, lane-strided sums, xor butterfly
__global__ void __launch_bounds__(256) project_kernel(const double* __restrict__ X, int64_t ldx, int N, int G,
                                                      const double* __restrict__ mean, const double* __restrict__ v,
                                                      const double* __restrict__ sign, double* __restrict__ lens, int64_t ldl) {
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    double s0 = 0.0, s1 = 0.0;
    for (int g = lane; g < G; g += 32) {
        const double x = X[(int64_t)n * ldx + g] - mean[g];
        s0 = fma(x, v[g], s0);
        s1 = fma(x, v[G + g], s1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (lane == 0) {
        lens[(int64_t)n * ldl] = sign[0] * s0;
        lens[(int64_t)n * ldl + 1] = sign[1] * s1;
    }
}

// dual form: lens[n][c] = sign[c] * sqrt(theta_c) * u_c[n]
__global__ void dual_lens_kernel(const double* __restrict__ u, int N, const IterState* st, const double* __restrict__ sign,
                                 double* __restrict__ lens, int64_t ldl) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    lens[(int64_t)n * ldl] = sign[0] * sqrt(fmax(st->theta[0], 0.0)) * u[n];
    lens[(int64_t)n * ldl + 1] = sign[1] * sqrt(fmax(st->theta[1], 0.0)) * u[N + n];
}

__global__ void info_kernel(const IterState* st, double* info) {
    info[0] = st->theta[0];
    info[1] = st->theta[1];
    info[2] = st->resid[0];
    info[3] = st->resid[1];
    info[4] = (double)st->iters;
    info[5] = (double)st->done;
}

}  // namespace

extern "C" int32_t sctda_pca_lens_f64(sctda_ctx* ctx, int32_t N, int32_t G, const double* d_X, int64_t ldx, double* d_lens,
                                      int64_t ldl, double* d_info, int32_t max_iter, double tol, void* stream) {
    if (!ctx) return SCTDA_ERR_INVALID;
    if (N < 2 || G < 2 || !d_X || !d_lens || ldx < G || ldl < 2)
        return sctda_fail(ctx, SCTDA_ERR_INVALID, "pca_lens: needs N >= 2 cells, G >= 2 genes, ldx >= G, ldl >= 2%s");
    if (max_iter <= 0) max_iter = 200;
    if (!(tol > 0.0)) tol = 1e-12;
    cudaStream_t st = (cudaStream_t)stream;
    SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool primal = G <= N;                                   // eigen-problem of the smaller Gram matrix
    const int M = primal ? G : N;                                 // its order
    const int K = primal ? N : G;                                 // contraction length
    const int64_t ldt = ((int64_t)K + 15) / 16 * 16;
    double* T = (double*)sctda_ws(ctx, WS_PCA_T, (size_t)M * ldt * sizeof(double));
    double* A = (double*)sctda_ws(ctx, WS_PCA_A, (size_t)M * M * sizeof(double));
    const int rows_per_block = 2048;
    const int nblk = (N + rows_per_block - 1) / rows_per_block;
    const size_t maxd = (size_t)(G > N ? G : N);
    const size_t small = ((size_t)nblk * G + (size_t)G + 3 * (size_t)M * kB + 2 * maxd + 2 * (size_t)G + 64) * sizeof(double)
                         + sizeof(IterState) + 1024;
    unsigned char* V = (unsigned char*)sctda_ws(ctx, WS_PCA_V, small);
    if (!T || !A || !V) return SCTDA_ERR_NOMEM;
    auto take = [&](size_t doubles) {                              // 64-byte aligned pieces of the small workspace
        double* p = (double*)V;
        V += (doubles * sizeof(double) + 63) & ~(size_t)63;
        return p;
    };
    double* part = take((size_t)nblk * G);
    double* mean = take((size_t)G);
    double* Q = take((size_t)M * kB);
    double* Z = take((size_t)M * kB);
    double* Y2 = take((size_t)M * kB);
    double* vec = take(2 * maxd);                                  // [2][M]
    double* comp = take(2 * (size_t)G);                            // [2][G] (dual form)
    double* sign = take(8);
    IterState* state = (IterState*)take((sizeof(IterState) + 7) / 8);
    dim3 gm((G + 255) / 256, nblk);
    col_partial_kernel<<<gm, 256, 0, st>>>(d_X, ldx, N, G, rows_per_block, part);
    SCTDA_LAUNCH_CHECK(ctx, "col_partial_kernel");
    col_mean_kernel<<<(G + 255) / 256, 256, 0, st>>>(part, nblk, N, G, mean);
    SCTDA_LAUNCH_CHECK(ctx, "col_mean_kernel");
    if (primal) {
        dim3 gc((G + 31) / 32, (int)((ldt + 31) / 32));
        centre_kernel<true><<<gc, 256, 0, st>>>(d_X, ldx, N, G, mean, T, ldt);
    } else {
        dim3 gc((int)((ldt + 31) / 32), (N + 31) / 32);
        centre_kernel<false><<<gc, 256, 0, st>>>(d_X, ldx, N, G, mean, T, ldt);
    }
    SCTDA_LAUNCH_CHECK(ctx, "centre_kernel");
    const int saved_precision = ctx->precision;
    ctx->precision = 0;                                           // the lens is always computed in FP64
    int rc = sctda_gram_gather_f64(ctx, K, T, ldt, M, nullptr, M, nullptr, A, M, stream);
    ctx->precision = saved_precision;
    if (rc) return rc;
    SCTDA_CUDA(ctx, cudaMemsetAsync(state, 0, sizeof(IterState), st));
    init_q_kernel<<<(M * kB + 255) / 256, 256, 0, st>>>(Q, M);
    SCTDA_LAUNCH_CHECK(ctx, "init_q_kernel");
    // Orthonormalisation of a basis B into Q: the ritz kernel on the pair (Q = B, Z = B) computes S0 = B'B,
    // rotates by its eigenvectors and applies the Cholesky factor (Q <- B W R^-1); with tol < 0 it touches
    // nothing else.
    ritz_kernel<<<1, 256, 0, st>>>(Q, Q, M, -1.0, state);
    SCTDA_LAUNCH_CHECK(ctx, "ritz_kernel");
    const int mm_blocks = (int)cdiv64((int64_t)M * 32, 256);
    for (int it = 0; it < max_iter; ++it) {
        matmul8_kernel<<<mm_blocks, 256, 0, st>>>(A, M, M, Q, Q, Z, 0, state);          // Z = A Q
        ritz_kernel<<<1, 256, 0, st>>>(Q, Z, M, tol, state);                            // Ritz values, residuals, Q <- orth(Z W)
        // Chebyshev filter of the new basis: three buffers rotate (prev, cur, next)
        // Y_1 -> Z, Y_2 -> Y2, then every step overwrites the older of the two in place (a row of the older
        // basis is read and written by the same lane only); Q stays intact until the last step
        double* prev = Q;
        double* cur = Q;
        for (int i = 1; i <= kCheb; ++i) {
            double* nxt = i == 1 ? Z : (i == 2 ? Y2 : prev);
            matmul8_kernel<<<mm_blocks, 256, 0, st>>>(A, M, M, cur, prev, nxt, i, state);
            prev = cur;
            cur = nxt;
        }
        ritz_kernel<<<1, 256, 0, st>>>(Q, cur, M, -1.0, state);                         // Q <- orth(filtered basis); no-op once converged
        ritz_kernel<<<1, 256, 0, st>>>(Q, Q, M, -1.0, state);                           // twice: the filtered columns differ by many orders of magnitude
        ctx->launches += 4 + kCheb;
        if ((it & 3) == 3) {                                       // every 4 outer steps: has the device flag been raised?
            int done = 0;
            SCTDA_CUDA(ctx, cudaMemcpyAsync(&done, &state->done, sizeof(int), cudaMemcpyDeviceToHost, st));
            SCTDA_CUDA(ctx, cudaStreamSynchronize(st));
            if (done) break;
        }
    }
    SCTDA_CUDA(ctx, cudaGetLastError());
    extract_kernel<<<(M + 255) / 256, 256, 0, st>>>(Q, M, vec);
    SCTDA_LAUNCH_CHECK(ctx, "extract_kernel");
    if (primal) {
        sign_kernel<<<1, 1024, 0, st>>>(vec, G, sign);
        SCTDA_LAUNCH_CHECK(ctx, "sign_kernel");
        project_kernel<<<(int)cdiv64((int64_t)N * 32, 256), 256, 0, st>>>(d_X, ldx, N, G, mean, vec, sign, d_lens, ldl);
        SCTDA_LAUNCH_CHECK(ctx, "project_kernel");
    } else {
        dual_components_kernel<<<(G + 255) / 256, 256, 0, st>>>(T, ldt, N, G, vec, comp);
        SCTDA_LAUNCH_CHECK(ctx, "dual_components_kernel");
        sign_kernel<<<1, 1024, 0, st>>>(comp, G, sign);
        SCTDA_LAUNCH_CHECK(ctx, "sign_kernel");
        dual_lens_kernel<<<(N + 255) / 256, 256, 0, st>>>(vec, N, state, sign, d_lens, ldl);
        SCTDA_LAUNCH_CHECK(ctx, "dual_lens_kernel");
    }
    if (d_info) {
        info_kernel<<<1, 1, 0, st>>>(state, d_info);
        SCTDA_LAUNCH_CHECK(ctx, "info_kernel");
    }
    return SCTDA_OK;
}
