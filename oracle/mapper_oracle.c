/* CPU ORACLE helper for hot path (a) — TEST INFRASTRUCTURE ONLY (see oracle/mapper_oracle.py).
 *
 * gram_fma: G[i][j] = sum_k X[ra[i]][k] * X[rb[j]][k] evaluated as ONE chain of fused
 * multiply-adds over k = 0..g-1 starting from 0.0 — the arithmetic a float64 tensor-core MMA
 * pipeline with a sequential K loop performs (each mma.sync m8n8k4 step is four chained FMAs
 * per output element).  Compiled with -ffp-contract=off: only the explicit fma() fuses.
 */
#include <math.h>
#include <stdint.h>

void gram_fma(const double* X, int64_t n, int64_t g, const int64_t* ra, int64_t na, const int64_t* rb, int64_t nb,
              double* out) {
    (void)n;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t i = 0; i < na; ++i) {
        const double* a = X + ra[i] * g;
        for (int64_t j = 0; j < nb; ++j) {
            const double* b = X + rb[j] * g;
            double acc = 0.0;
            for (int64_t k = 0; k < g; ++k) acc = fma(a[k], b[k], acc);
            out[i * nb + j] = acc;
        }
    }
}
