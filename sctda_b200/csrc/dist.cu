// Hot path (a), stage 1: row normalisation and the cells x cells Gram / distance contraction.
//
// Reference seam: sakmapper.apply_lens(..., metric=...) builds the N x N dissimilarity for the MDS
// lens (/root/reference/scTDA/main.py:746-751) and sakmapper.mapper_graph clusters every cover
// patch on distances between its rows (:768-771).  Both reduce to
//     G[i][j] = sum_k Xn[ra[i]][k] * Xn[rb[j]][k]
// over row index lists (a row block against all rows, or a patch against itself).
//
// Design (sm_100a).  tcgen05 has no FP64 kind, so the one dense contraction of the path runs
// on the FP64 tensor pipe through warp-level mma.sync.m8n8k4 (DMMA): 128x128 CTA tile, 8 warps
// of 32x64, K staged 16 columns at a time through a 3-stage ring.  Rows are GATHERED (index
// lists), which a tiled TMA box cannot express.  Two stagings exist: 16-byte cp.async with
// zero-fill on the K tail (default), and one-dimensional TMA bulk copies (cp.async.bulk, one
// 128-byte row segment each, SASS UBLKCP, completing on the stage's mbarrier; needs rows
// zero-padded to a multiple of 16 columns).  The bulk path is bit-identical but measured 22 %
// slower on these short segments and is therefore opt-in (see launch_gemm).
// Shared-memory rows are padded to 20 doubles so that the 8x4 fragment loads of each half warp
// hit 32 distinct banks.  K is walked sequentially with no split-K:
// every output element is ONE chain of fused multiply-adds over k ascending, which makes the
// result reproducible bit for bit by a scalar fma loop (oracle/mapper_oracle.c).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "gemm.cuh"

namespace {

#ifndef SCTDA_GEMM_STAGES
#define SCTDA_GEMM_STAGES (SCTDA_GEMM_TILE == 64 ? 4 : 3)
#endif
// CTA tile BM x BN = kGemmTileFp64 squared (gemm.cuh): 128 -> 8 warps (4 x 2, warp tile 32 x 64), one CTA per SM;
// 64 -> 4 warps (2 x 2, warp tile 32 x 32), TWO CTAs per SM with a four-stage ring (164 KB of shared memory, 90 KB
// left to L1, which the .ca copies use).  Measured on the 100k x 2k patch tiles (B200, gemm_ms of tools/mapper_run.py):
// 128-tile / 1 CTA / 3 stages 809 ms; 64-tile: 3 CTAs x 3 stages 785, 2 x 3 775, 2 x 4 767, 2 x 5 782, 4 x 2 832,
// 1 x 8 1,024 (and 888 / 826 for the first two with .cg copies).
constexpr int BM = kGemmTileFp64, BN = kGemmTileFp64, BK = 16, LDS = 20, STAGES = SCTDA_GEMM_STAGES;
constexpr int WARPS_M = BM / 32, WN = BN / 2, NJ = WN / 8;   // warps along M; columns and 8-column blocks of a warp tile
constexpr int GEMM_THREADS = WARPS_M * 2 * 32;
#ifndef SCTDA_GEMM_CTAS
#define SCTDA_GEMM_CTAS 2
#endif
constexpr int GEMM_CTAS_PER_SM = BM == 128 ? 1 : SCTDA_GEMM_CTAS;
static_assert(BM == 128 || BM == 64, "kGemmTileFp64 must be 128 or 64");
static_assert(GEMM_THREADS == 2 * BM, "the staging assignments assume two threads per tile row");
constexpr int STAGE_DOUBLES = (BM + BN) * LDS;
constexpr size_t GEMM_SMEM = (size_t)STAGES * STAGE_DOUBLES * sizeof(double);

__device__ __forceinline__ void cp_async_16_zfill(void* smem_dst, const void* gmem_src, int src_bytes) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    // .ca: through L1.  With .cg every 16-byte copy is its own L2 request and comes back as a whole 32-byte sector:
    // twice the operand bytes on the L2 -> SM fabric (ncu: 553 GB against 276 GB of operands per launch), which the
    // three-CTA 64-tile kernel felt: 826 -> 785 ms on the 100k x 2k patch tiles.
#ifdef SCTDA_GEMM_CPASYNC_CG
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
#else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier -------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    const long long t0 = clock64();
    for (;;) {
        unsigned ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > 8000000000LL) __trap();      // ~4 s: a lost bulk copy must not hang the device
    }
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

__device__ __forceinline__ void dmma_884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// BULK: every K stage is a full 16 columns (the caller guarantees zero padding up to a multiple
// of 16) and is staged by warp 0 with one 128-byte TMA bulk copy per row, completing on the
// stage's mbarrier; `it` counts the K stages this CTA has consumed (stage = it % 3, parity from it / 3).
template <bool BULK>
__device__ __forceinline__ void gemm_tile(const GemmArgs& a, const GemmProblem& pr, int ti, int tj, double* smem,
                                          bool mirror, uint64_t* bars, unsigned& it) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WARPS_M, wn = warp / WARPS_M;      // WARPS_M x 2 warps, warp tile 32 x WN
    const int m0 = ti * BM, n0 = tj * BN;
    // this thread's cp.async assignments: chunk c = tid + 256*i  ->  row c/8, 16-byte column c%8
    const double* srcA[4];
    const double* srcB[4];
    int dst_off[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int c = tid + GEMM_THREADS * i;
        int r = c >> 3, q = c & 7;
        int ra = min(m0 + r, pr.na - 1), rb = min(n0 + r, pr.nb - 1);
        int64_t ga = pr.rows_a ? (int64_t)pr.rows_a[ra] : (int64_t)(pr.row0_a + ra);
        int64_t gb = pr.rows_b ? (int64_t)pr.rows_b[rb] : (int64_t)(pr.row0_b + rb);
        srcA[i] = a.X + ga * a.ldx + q * 2;
        srcB[i] = a.X + gb * a.ldx + q * 2;
        dst_off[i] = r * LDS + q * 2;
    }
    const int ksteps = (a.G + BK - 1) / BK;
    // bulk path: thread t copies row (t & 127) of A (t < 128) or of B: one 128-byte TMA bulk copy per
    // thread and stage (the copy instruction is warp-uniform in hardware, so 32 lanes are 32 issues)
    const double* bulk_row = nullptr;
    const int bulk_dst = ((tid >= BM) ? BM * LDS : 0) + (tid % BM) * LDS;
    if (BULK) {
        const int r = tid % BM;
        if (tid < BM) {
            const int ra = min(m0 + r, pr.na - 1);
            bulk_row = a.X + (pr.rows_a ? (int64_t)pr.rows_a[ra] : (int64_t)(pr.row0_a + ra)) * a.ldx;
        } else {
            const int rb = min(n0 + r, pr.nb - 1);
            bulk_row = a.X + (pr.rows_b ? (int64_t)pr.rows_b[rb] : (int64_t)(pr.row0_b + rb)) * a.ldx;
        }
    }
    auto issue_bulk = [&](int ks, unsigned slot) {            // all threads; expect_tx was posted by thread 0
        bulk_copy_g2s(smem + (size_t)slot * STAGE_DOUBLES + bulk_dst, bulk_row + ks * BK, BK * 8, bars + slot);
    };
    auto issue = [&](int ks) {
        double* sA = smem + (size_t)(ks % STAGES) * STAGE_DOUBLES;
        double* sB = sA + BM * LDS;
        const int k0 = ks * BK;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int q2 = (dst_off[i] % LDS);                      // column (doubles) of this chunk
            int remain = a.G - (k0 + q2);                     // doubles left in the row from here
            int bytes = remain >= 2 ? 16 : (remain == 1 ? 8 : 0);
            cp_async_16_zfill(sA + dst_off[i], bytes ? (const void*)(srcA[i] + k0) : (const void*)a.X, bytes);
            cp_async_16_zfill(sB + dst_off[i], bytes ? (const void*)(srcB[i] + k0) : (const void*)a.X, bytes);
        }
    };
    double acc[4][NJ][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    if (BULK) {
        if (tid == 0)
            for (int s = 0; s < STAGES - 1 && s < ksteps; ++s) mbar_expect_tx(bars + (it + s) % STAGES, (BM + BN) * BK * 8);
        __syncthreads();
        for (int s = 0; s < STAGES - 1 && s < ksteps; ++s) issue_bulk(s, (it + s) % STAGES);
    } else {
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) {
            if (s < ksteps) issue(s);
            cp_async_commit();
        }
    }
    const int fr = lane >> 2, fc = lane & 3;
    for (int ks = 0; ks < ksteps; ++ks) {
        const unsigned slot = BULK ? (it + ks) % STAGES : ks % STAGES;
        if (BULK) {
            const bool more = ks + STAGES - 1 < ksteps;
            const unsigned nslot = (it + ks + STAGES - 1) % STAGES;
            if (tid == 0 && more) mbar_expect_tx(bars + nslot, (BM + BN) * BK * 8);   // its previous phase was consumed by all
            mbar_wait(bars + slot, ((it + ks) / STAGES) & 1);
            __syncthreads();                                  // everybody is done with the slot refilled below
            if (more) issue_bulk(ks + STAGES - 1, nslot);
        } else {
            cp_async_wait<STAGES - 2>();
            __syncthreads();
            if (ks + STAGES - 1 < ksteps) issue(ks + STAGES - 1);
            cp_async_commit();
        }
        const double* sA = smem + (size_t)slot * STAGE_DOUBLES + (wm * 32 + fr) * LDS + fc;
        const double* sB = smem + (size_t)slot * STAGE_DOUBLES + BM * LDS + (wn * WN + fr) * LDS + fc;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double fa[4], fb[NJ];
#pragma unroll
            for (int i = 0; i < 4; ++i) fa[i] = sA[i * 8 * LDS + kk * 4];
#pragma unroll
            for (int j = 0; j < NJ; ++j) fb[j] = sB[j * 8 * LDS + kk * 4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma_884(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
        }
    }
    if (!BULK) cp_async_wait<0>();
    it += (unsigned)ksteps;
    __syncthreads();                                          // smem is reused by the next tile
    // epilogue: thread holds rows fr (+8*i), columns 2*fc, 2*fc+1 (+8*j) of its warp tile
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = m0 + wm * 32 + i * 8 + fr;
        if (r >= pr.na) continue;
        const int64_t gr = pr.rows_a ? (int64_t)pr.rows_a[r] : (int64_t)(pr.row0_a + r);
        const double sqr = (a.epilogue == EPI_EUCLID) ? a.sq[gr] : 0.0;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = n0 + wn * WN + j * 8 + fc * 2 + e;
                if (c >= pr.nb) continue;
                double v = acc[i][j][e];
                if (a.epilogue != EPI_GRAM) {
                    const int64_t gc = pr.rows_b ? (int64_t)pr.rows_b[c] : (int64_t)(pr.row0_b + c);
                    if (a.epilogue == EPI_EUCLID) v = sqrt(fmax(fma(-2.0, v, __dadd_rn(sqr, a.sq[gc])), 0.0));
                    else v = __dsub_rn(1.0, v);
                    if (gr == gc) v = 0.0;
                }
                pr.out[(int64_t)r * pr.ldo + c] = v;
                if (mirror) pr.out[(int64_t)c * pr.ldo + r] = v;     // G is symmetric bit for bit (products commute)
            }
        }
    }
}

// Tile t (global index over the upper-triangular tiles of all patches of the batch) of a batched launch.
template <bool BULK>
__device__ __forceinline__ void batched_tile(const GemmArgs& a, int64_t t, double* smem, uint64_t* bars, unsigned& it) {
    int lo = 0, hi = a.nprob;                                 // last p with tile_prefix[p] <= t
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (a.tile_prefix[mid] <= t) lo = mid; else hi = mid;
    }
    const int bin = a.prob_bin[lo];
    const int m = a.bin_off[bin + 1] - a.bin_off[bin];
    const int tn = (m + BN - 1) / BN;
    int64_t lt = t - a.tile_prefix[lo];                      // index into the upper triangle of tiles, row-major
    int ti = 0;
    while (lt >= tn - ti) { lt -= tn - ti; ++ti; }
    const int tj = ti + (int)lt;
    GemmProblem pr;
    pr.rows_a = pr.rows_b = a.members + a.bin_off[bin];
    pr.row0_a = pr.row0_b = 0;
    pr.na = pr.nb = m;
    pr.out = a.gram + a.gram_off[lo];
    pr.ldo = m;
    gemm_tile<BULK>(a, pr, ti, tj, smem, ti != tj, bars, it);
}

template <bool BULK>
__global__ void __launch_bounds__(GEMM_THREADS, GEMM_CTAS_PER_SM) gemm_kernel(GemmArgs a) {
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t bars[STAGES];
    unsigned it = 0;
    if (BULK) {
        if (threadIdx.x == 0) {
            for (int s = 0; s < STAGES; ++s) mbar_init(bars + s, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    if (a.nprob == 0) {
        const int tn = (a.single.nb + BN - 1) / BN;
        if (a.symmetric) {                                    // square and symmetric: upper triangle of tiles, mirrored
            const int64_t total = (int64_t)tn * (tn + 1) / 2;
            for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
                int64_t lt = t;
                int ti = 0;
                while (lt >= tn - ti) { lt -= tn - ti; ++ti; }
                gemm_tile<BULK>(a, a.single, ti, ti + (int)lt, smem, lt != 0, bars, it);
            }
            return;
        }
        const int64_t total = (int64_t)((a.single.na + BM - 1) / BM) * tn;
        for (int64_t t = blockIdx.x; t < total; t += gridDim.x)
            gemm_tile<BULK>(a, a.single, (int)(t / tn), (int)(t % tn), smem, false, bars, it);
        return;
    }
    const int pb = a.prob_end > 0 ? a.prob_begin : 0, pe = a.prob_end > 0 ? a.prob_end : a.nprob;
    const int64_t first = a.tile_prefix[pb], total = a.tile_prefix[pe];
    if (!a.tile_counter) {                                    // static striding over the tiles of the launch
        for (int64_t t = first + blockIdx.x; t < total; t += gridDim.x) batched_tile<BULK>(a, t, smem, bars, it);
        return;
    }
    __shared__ long long s_tile;                              // dynamic: tiles are claimed in order from a global counter
    for (int done = 0; a.tiles_per_cta == 0 || done < a.tiles_per_cta; ++done) {
        __syncthreads();                                      // the previous tile's readers of s_tile are done
        if (threadIdx.x == 0) s_tile = (long long)(first + (int64_t)atomicAdd(a.tile_counter, 1ull));
        __syncthreads();
        const int64_t t = s_tile;
        if (t >= total) break;
        batched_tile<BULK>(a, t, smem, bars, it);
    }
}

// Row statistics in a fixed order: lane l adds elements l, l+32, ... sequentially, then the
// xor butterfly 16, 8, 4, 2, 1 (oracle/mapper_oracle.py:_lane_sums).  One warp per row.
__device__ __forceinline__ double warp_tree(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__global__ void __launch_bounds__(256) row_normalize_kernel(const double* __restrict__ X, int64_t ldx, int N, int G,
                                                            int mode, double* __restrict__ Xn, int64_t ldxn,
                                                            double* __restrict__ sq) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= N) return;
    const double* x = X + (int64_t)row * ldx;
    double* y = Xn + (int64_t)row * ldxn;
    double mean = 0.0, inv_valid = 0.0, nrm = 1.0;
    if (mode == 1) {
        double s = 0.0;
        for (int k = lane; k < G; k += 32) s = __dadd_rn(s, x[k]);
        mean = __ddiv_rn(warp_tree(s), (double)G);
        double ss = 0.0;
        for (int k = lane; k < G; k += 32) {
            double c = __dsub_rn(x[k], mean);
            ss = __dadd_rn(ss, __dmul_rn(c, c));
        }
        nrm = sqrt(warp_tree(ss));
        inv_valid = nrm > 0.0 ? 1.0 : 0.0;
    }
    double s2 = 0.0;
    for (int k = lane; k < G; k += 32) {
        double v = x[k];
        if (mode == 1) v = inv_valid != 0.0 ? __ddiv_rn(__dsub_rn(v, mean), nrm) : 0.0;
        y[k] = v;
        s2 = __dadd_rn(s2, __dmul_rn(v, v));
    }
    for (int k = G + lane; k < ldxn; k += 32) y[k] = 0.0;    // zero the padding of the leading dimension
    s2 = warp_tree(s2);
    if (lane == 0 && sq) sq[row] = s2;
}

// bulk: the rows are zero-padded to a multiple of 16 columns (ldx >= roundup16(G)), every K stage is full
int launch_gemm(sctda_ctx* ctx, const GemmArgs& a, int64_t tiles_hint, cudaStream_t st, bool bulk, bool reuse_split = false) {
    if (ctx->precision == 1 && a.nrows > 0) return sctda_launch_gemm_split(ctx, a, st, reuse_split);   // opt-in 3xTF32 (gemm_split.cu)
    if (!(ctx->attr_done & 2u)) {
        SCTDA_CUDA(ctx, cudaFuncSetAttribute(gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
        SCTDA_CUDA(ctx, cudaFuncSetAttribute(gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
        ctx->attr_done |= 2u;
    }
    const int64_t resident = (int64_t)ctx->sm_count * GEMM_CTAS_PER_SM;
    int64_t grid = tiles_hint > 0 ? tiles_hint : resident;
    if (grid > resident) grid = resident;                    // persistent: every resident CTA slot
    if (a.grid_ctas > 0) grid = a.grid_ctas;
    if (grid < 1) grid = 1;
    if (a.untimed) {
        gemm_kernel<false><<<(unsigned)grid, GEMM_THREADS, GEMM_SMEM, st>>>(a);
        SCTDA_LAUNCH_CHECK(ctx, "gemm_kernel");
        return SCTDA_OK;
    }
    SCTDA_TIME_BEGIN(ctx, st);
    // Measured on B200 (config 2 / config 3 patch Gram matrices): the TMA bulk-copy staging is 22 %
    // SLOWER than 16-byte cp.async for these 128-byte gathered row segments (UBLKCP is a
    // warp-uniform instruction: 256 serialised issues per stage against 2,048 vector LDGSTS lanes),
    // so it is opt-in (SCTDA_GEMM_TMA_BULK=1) and cp.async stays the default.
    static const bool want_bulk = getenv("SCTDA_GEMM_TMA_BULK") && atoi(getenv("SCTDA_GEMM_TMA_BULK")) != 0;
    if (bulk && want_bulk) gemm_kernel<true><<<(unsigned)grid, GEMM_THREADS, GEMM_SMEM, st>>>(a);
    else gemm_kernel<false><<<(unsigned)grid, GEMM_THREADS, GEMM_SMEM, st>>>(a);
    SCTDA_TIME_END(ctx, st);
    SCTDA_LAUNCH_CHECK(ctx, "gemm_kernel");
    return SCTDA_OK;
}

}  // namespace

// Batched patch Gram matrices for the clustering stage (cluster.cu).
int sctda_internal_gram_bins(sctda_ctx* ctx, int G, const double* d_Xn, int64_t ldxn, int nbins,
                             const int32_t* d_prob_bin, const int32_t* d_bin_off, const int32_t* d_members,
                             const int64_t* d_tile_prefix, const int64_t* d_gram_off, double* d_gram, cudaStream_t st,
                             int64_t nrows, bool reuse_split, int prob_begin, int prob_end, unsigned long long* d_tile_counter,
                             int grid_ctas, bool untimed, int tiles_per_cta) {
    GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.X = d_Xn; a.ldx = ldxn; a.G = G; a.epilogue = EPI_GRAM; a.nrows = nrows;
    a.nprob = nbins; a.prob_bin = d_prob_bin; a.bin_off = d_bin_off; a.members = d_members;
    a.tile_prefix = d_tile_prefix; a.gram_off = d_gram_off; a.gram = d_gram;
    a.prob_begin = prob_begin; a.prob_end = prob_end; a.tile_counter = d_tile_counter;
    a.grid_ctas = grid_ctas; a.untimed = untimed ? 1 : 0; a.tiles_per_cta = tiles_per_cta;
    return launch_gemm(ctx, a, 0, st, ldxn >= ((int64_t)G + 15) / 16 * 16, reuse_split);
}

extern "C" int32_t sctda_gemm_tile_fp64(void) { return kGemmTileFp64; }

extern "C" int32_t sctda_row_normalize_f64(sctda_ctx* ctx, int32_t N, int32_t G, const double* d_X, int64_t ldx,
                                           int32_t mode, double* d_Xn, int64_t ldxn, double* d_sq, void* stream) {
    if (!ctx) return SCTDA_ERR_INVALID;
    if (N < 0 || G < 1 || !d_X || !d_Xn || ldx < G || ldxn < G || (mode != 0 && mode != 1))
        return sctda_fail(ctx, SCTDA_ERR_INVALID, "row_normalize: bad arguments%s");
    if (N == 0) return SCTDA_OK;
    SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
    row_normalize_kernel<<<(unsigned)cdiv64((int64_t)N * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        d_X, ldx, N, G, mode, d_Xn, ldxn, d_sq);
    SCTDA_LAUNCH_CHECK(ctx, "row_normalize_kernel");
    return SCTDA_OK;
}

extern "C" int32_t sctda_gram_gather_f64(sctda_ctx* ctx, int32_t G, const double* d_Xn, int64_t ldxn, int32_t na,
                                         const int32_t* d_rows_a, int32_t nb, const int32_t* d_rows_b, double* d_out,
                                         int64_t ldo, void* stream) {
    if (!ctx) return SCTDA_ERR_INVALID;
    SCTDA_TIME_RESET(ctx);
    if (G < 1 || na < 0 || nb < 0 || !d_Xn || !d_out || ldxn < G || ldo < nb)
        return sctda_fail(ctx, SCTDA_ERR_INVALID, "gram_gather: bad arguments%s");
    if ((ldxn & 1) || ((uintptr_t)d_Xn & 15))
        return sctda_fail(ctx, SCTDA_ERR_UNSUPPORTED, "gram_gather: rows must be 16-byte aligned (even ldxn)%s");
    if (na == 0 || nb == 0) return SCTDA_OK;
    SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
    GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.X = d_Xn; a.ldx = ldxn; a.G = G; a.epilogue = EPI_GRAM;
    a.single.rows_a = d_rows_a; a.single.rows_b = d_rows_b; a.single.na = na; a.single.nb = nb;
    a.single.out = d_out; a.single.ldo = ldo;
    a.symmetric = (d_rows_a == d_rows_b && na == nb) ? 1 : 0;    // the same rows on both sides: half the tiles, mirrored
    const int64_t q = cdiv64(nb, BN);
    return launch_gemm(ctx, a, a.symmetric ? q * (q + 1) / 2 : cdiv64(na, BM) * q, (cudaStream_t)stream,
                       ldxn >= ((int64_t)G + 15) / 16 * 16);
}

extern "C" int32_t sctda_pairdist_rows_f64(sctda_ctx* ctx, int32_t N, int32_t G, const double* d_Xn, int64_t ldxn,
                                           const double* d_sq, int32_t r0, int32_t r1, int32_t metric, double* d_D,
                                           int64_t ldd, void* stream) {
    if (!ctx) return SCTDA_ERR_INVALID;
    SCTDA_TIME_RESET(ctx);
    if (N < 1 || G < 1 || !d_Xn || !d_D || r0 < 0 || r1 < r0 || r1 > N || ldxn < G || ldd < N ||
        (metric != 0 && metric != 1) || (metric == 0 && !d_sq))
        return sctda_fail(ctx, SCTDA_ERR_INVALID, "pairdist_rows: bad arguments%s");
    if ((ldxn & 1) || ((uintptr_t)d_Xn & 15))
        return sctda_fail(ctx, SCTDA_ERR_UNSUPPORTED, "pairdist_rows: rows must be 16-byte aligned (even ldxn)%s");
    if (r1 == r0) return SCTDA_OK;
    SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
    GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.X = d_Xn; a.ldx = ldxn; a.G = G; a.sq = d_sq; a.nrows = N;
    a.epilogue = metric == 0 ? EPI_EUCLID : EPI_CORR;
    a.single.row0_a = r0; a.single.na = r1 - r0; a.single.row0_b = 0; a.single.nb = N;
    a.single.out = d_D; a.single.ldo = ldd;
    a.symmetric = (r0 == 0 && r1 == N) ? 1 : 0;              // the whole matrix: half the tiles, mirrored (bit-identical: products commute)
    const int64_t q = cdiv64(N, BN);
    return launch_gemm(ctx, a, a.symmetric ? q * (q + 1) / 2 : cdiv64(r1 - r0, BM) * q, (cudaStream_t)stream,
                       ldxn >= ((int64_t)G + 15) / 16 * 16);
}
