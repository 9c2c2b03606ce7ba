// Problem description shared by the two contraction kernels: the FP64 DMMA kernel (dist.cu, the
// default, bit-exact against the oracle) and the opt-in 3xTF32 tcgen05 kernel (gemm_split.cu).
#pragma once
#include "common.cuh"

enum { EPI_GRAM = 0, EPI_EUCLID = 1, EPI_CORR = 2 };

// Side of the square output tile of the FP64 kernel (dist.cu): 64 (4 warps, two CTAs per SM; measured 5-13 % faster
// than one 8-warp CTA on 128 x 128 tiles: independent CTAs fill each other's per-stage barriers and tile
// prologues / epilogues); the opt-in tcgen05 kernel always works on 128 x 128.
// tile_prefix of a batched launch counts the upper-triangular tiles of THIS size (sctda_gemm_tile(precision)).
#ifndef SCTDA_GEMM_TILE
#define SCTDA_GEMM_TILE 64
#endif
constexpr int kGemmTileFp64 = SCTDA_GEMM_TILE;
static inline int sctda_gemm_tile(int precision) { return precision == 1 ? 128 : kGemmTileFp64; }

// One batched problem: out[i][j] for i < na, j < nb.
struct GemmProblem {
    const int32_t* rows_a;   // row indices into X (NULL: identity + row0_a)
    const int32_t* rows_b;
    int row0_a, row0_b;
    int na, nb;
    double* out;
    int64_t ldo;
};

struct GemmArgs {
    const double* X;         // [*, ldx]
    int64_t nrows;           // rows of X (needed by the 3xTF32 operand split; 0 = unknown)
    int64_t ldx;
    int G;                   // features (K of the GEMM)
    const double* sq;        // |row|^2 (EUCLID epilogue)
    int epilogue;
    // single problem (nprob == 0) or batched patches
    GemmProblem single;
    int symmetric;               // single problem with rows_a == rows_b, na == nb: upper-triangular tiles, mirrored
    int nprob;
    const int32_t* prob_bin;     // [nprob] patch index of problem p
    const int32_t* bin_off;      // member offsets of all patches
    const int32_t* members;      // [Mtot]
    const int64_t* tile_prefix;  // [nprob+1] number of upper-triangular 128x128 tiles before problem p
    const int64_t* gram_off;     // [nprob+1] element offset of problem p's m x m output
    double* gram;
    // FP64 kernel only: the launch covers the tiles of problems [prob_begin, prob_end) (0, 0 = all), and with a
    // counter (zeroed by the host) the CTAs claim tiles from it instead of striding: a launch that shares the
    // GPU with long-running MST kernels must not leave tiles to CTAs that start late
    int prob_begin, prob_end;
    unsigned long long* tile_counter;
    int grid_ctas;               // host: CTAs of the launch (0 = one per SM)
    int tiles_per_cta;           // with a tile counter: a CTA retires after this many tiles (0 = runs until the counter is
                                 // exhausted), so that SMs turn over and kernels of a higher-priority stream get in
    int untimed;                 // host: a helper launch on another stream, not bracketed by the call's timing events
};


// gemm_split.cu: same problems, 3xTF32 on the tcgen05 tensor cores (sctda_set_precision(ctx, 1)).
// reuse_split: the (hi, lo) operands of this X are already in the workspace (later batches of one call).
int sctda_launch_gemm_split(sctda_ctx* ctx, const GemmArgs& a, cudaStream_t st, bool reuse_split);
