// Hot path (a), stage 1, OPT-IN precision: the cells x cells Gram / distance contraction as three
// half-precision tensor-core products on the 5th-generation tensor cores (tcgen05.mma kind::f16,
// fp32 accumulators in TMEM) -- the "3xTF32" idea with FP16 containers.
//
// BASELINE.json north_star (1): "... an opt-in FP32/3xTF32 mode at a stated <= 1e-5 relative
// tolerance".  The default stays the FP64 DMMA kernel of dist.cu (bit-exact against the oracle);
// sctda_set_precision(ctx, 1) routes the same problems (patch Gram tiles for the clustering of
// /root/reference/scTDA/main.py:768-771, distance row blocks for the MDS lens of :746-751) here.
//
// Arithmetic.  TF32 and FP16 both carry 11 significant bits; FP16 does it in half the bytes and at
// twice the tensor-core rate, and its narrow exponent range is handled by power-of-two scaling
// (exact).  With s = 2^e chosen so that max|x| s is in [2^14, 2^15), every value is split once into
//     hi = fp16(x s),   lo = fp16((x s - hi) 2^11)           (22 significant bits between them)
// and  x.y s^2 ~= hi.hi' + 2^-11 (hi.lo' + lo.hi')  -- three products, the dropped lo.lo' term is
// 2^-22 relative.  The large and the small products accumulate in SEPARATE fp32 TMEM accumulators
// and are combined in float64 in the epilogue, which then applies the same Gram / euclidean /
// correlation epilogue as dist.cu.
//
// fp32 accumulation in the tensor core truncates: over K = 5,000 a single accumulator drifts by
// 4e-5 (measured with TF32 operands).  The K loop is therefore cut into SEGMENTS of 512 features
// that alternate between the two TMEM accumulator stages; the epilogue warps drain each finished
// segment into float64 register sums while the next segment's MMAs run (measured error: see
// sctda_set_precision in the header).
//
// Kernel (sm_100a), persistent, one CTA per SM, 352 threads in three roles:
//   warps 8-9  producers: gather operand rows (index lists: a tiled TMA box cannot express them)
//              with 16-byte cp.async into the canonical K-major SWIZZLE_128B layout the UMMA
//              shared-memory descriptors expect, 3-stage ring of 64 KB (K = 64 per stage:
//              A_hi, A_lo, B_hi, B_lo blocks of 128 rows x 128 bytes);
//   warp 10    one elected thread issues tcgen05.mma: per K = 16 step
//                  D[:, 0:256)   += A_hi . [B_hi ; B_lo]^T     (M128 N256: hi.hi and hi.lo in one issue)
//                  D[:, 128:256) += A_lo . B_hi^T             (M128 N128)
//              and tcgen05.commit releases the stage / publishes the accumulator;
//   warps 0-7  epilogue: warp w owns TMEM lanes 32*(w%4).. and columns 64*(w/4)..; per segment
//              tcgen05.ld + float64 accumulation of both accumulators (64 sums per thread), at the
//              end of the tile the epilogue transform and direct + mirrored (symmetric) float64 stores.
// TMEM: 2 accumulator stages x 256 columns = all 512 columns, so draining segment s overlaps the
// MMAs of segment s+1 (also across tiles).  Every mbarrier wait is bounded (trap after ~4 s) so a
// protocol bug cannot hang the device.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "gemm.cuh"

namespace {

constexpr int TM = 128, TN = 128, KC = 64;                  // K elements per stage: 128-byte rows of fp16
constexpr int T_STAGES = 3;
constexpr int BLK_BYTES = TM * KC * 2;                  // 16 KB, one swizzled 128-row block
constexpr int STAGE_BYTES = 4 * BLK_BYTES;              // A_hi, A_lo, B_hi, B_lo
constexpr int EPI_THREADS = 256, PROD_THREADS = 64, MMA_WARP = (EPI_THREADS + PROD_THREADS) / 32;
constexpr int T_THREADS = EPI_THREADS + PROD_THREADS + 32;
constexpr int SEG = 8;                                  // K chunks (512 features) per fp32 accumulation segment
constexpr double LO_SCALE = 2048.0;                     // 2^11, applied to the residual before it is rounded to fp16
constexpr size_t T_SMEM = (size_t)T_STAGES * STAGE_BYTES + 1024;
constexpr int ACC_COLS = 256;                           // fp32 columns per accumulator stage

struct TArgs {
    GemmArgs g;
    const uint4* Xs;        // split operands: [row][kchunk][hi 64 halfs | lo 64 halfs]
    int kchunks;            // ceil(G / 64)
    const float* scale;     // device scalar: the power of two s applied before the split
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    const uint32_t addr = smem_u32(bar);
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > 8000000000LL) __trap();      // ~4 s: a protocol bug must not hang the device
    }
}
__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const void* gmem_src) {
    // .cg (L2 only).  Through L1 (.ca), which pays in the FP64 kernel (dist.cu), this kernel is slower: 10.4 -> 12.9 ms
    // on the config-2 patch tiles; its 192 KB of stages leave L1 too small to merge the two halves of a sector.
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor: 8-row groups of 1,024 bytes (SBO),
// descriptor version 1 (sm_100), start address in 16-byte units.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 with FP16 operands (format 0), fp32 accumulate, A and B K-major, M = 128
__device__ __forceinline__ constexpr uint32_t umma_idesc(int n) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct Tile {
    GemmProblem pr;
    int ti, tj;
    bool mirror;
    bool diag;      // diagonal tile of a patch Gram: the B rows ARE the A rows, only A is staged
};

// tile t of the launch -> problem and tile coordinates (same enumeration as dist.cu: a single
// rectangular problem, or the upper triangles of the patches' tile grids)
__device__ __forceinline__ Tile decode_tile(const GemmArgs& a, int64_t t) {
    Tile o;
    if (a.nprob == 0) {
        const int tn = (a.single.nb + TN - 1) / TN;
        o.pr = a.single;
        if (a.symmetric) {
            int64_t lt = t;
            int ti = 0;
            while (lt >= tn - ti) { lt -= tn - ti; ++ti; }
            o.ti = ti;
            o.tj = ti + (int)lt;
            o.mirror = lt != 0;
            o.diag = lt == 0;
            return o;
        }
        o.ti = (int)(t / tn);
        o.tj = (int)(t % tn);
        o.mirror = false;
        o.diag = false;
        return o;
    }
    int lo = 0, hi = a.nprob;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (a.tile_prefix[mid] <= t) lo = mid; else hi = mid;
    }
    const int bin = a.prob_bin[lo];
    const int m = a.bin_off[bin + 1] - a.bin_off[bin];
    const int tn = (m + TN - 1) / TN;
    int64_t lt = t - a.tile_prefix[lo];
    int ti = 0;
    while (lt >= tn - ti) { lt -= tn - ti; ++ti; }
    o.ti = ti;
    o.tj = ti + (int)lt;
    o.mirror = o.ti != o.tj;
    o.diag = !o.mirror;
    o.pr.rows_a = o.pr.rows_b = a.members + a.bin_off[bin];
    o.pr.row0_a = o.pr.row0_b = 0;
    o.pr.na = o.pr.nb = m;
    o.pr.out = a.gram + a.gram_off[lo];
    o.pr.ldo = m;
    return o;
}

__device__ __forceinline__ int64_t total_tiles(const GemmArgs& a) {
    if (a.nprob == 0) {
        const int64_t tn = (a.single.nb + TN - 1) / TN;
        return a.symmetric ? tn * (tn + 1) / 2 : (int64_t)((a.single.na + TM - 1) / TM) * tn;
    }
    return a.tile_prefix[a.nprob];
}

__global__ void __launch_bounds__(T_THREADS, 1) gemm_split3_kernel(TArgs ta) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[T_STAGES], bar_empty[T_STAGES], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_base_s;
    const GemmArgs& a = ta.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;     // SWIZZLE_128B blocks are 1,024-byte aligned

    if (threadIdx.x == 0) {
        for (int s = 0; s < T_STAGES; ++s) { mbar_init(bar_full + s, PROD_THREADS); mbar_init(bar_empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(bar_tfull + s, 1); mbar_init(bar_tempty + s, EPI_THREADS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const int64_t total = total_tiles(a);
    const int kchunks = ta.kchunks;

    if (warp >= EPI_THREADS / 32 && warp < MMA_WARP) {
        // ---------------- producers ----------------
        const int pt = threadIdx.x - EPI_THREADS;
        const int piece = pt & 15, rsub = pt >> 4;                     // 16 x 16-byte pieces of a row's 256-byte K chunk
        constexpr int RSTEP = PROD_THREADS / 16;                        // rows covered per pass of the producer threads
        constexpr int NR = 2 * TM / RSTEP;                              // copies per thread and stage (A rows, then B rows)
        const uint32_t dst_blk = (uint32_t)((piece >> 3) * BLK_BYTES);
        uint32_t it = 0;
        for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
            const Tile tl = decode_tile(a, t);
            const int m0 = tl.ti * TM, n0 = tl.tj * TN;
            uint32_t roff[NR];                                          // source offsets in 16-byte units
#pragma unroll
            for (int i = 0; i < NR / 2; ++i) {
                const int ra = min(m0 + rsub + RSTEP * i, tl.pr.na - 1), rb = min(n0 + rsub + RSTEP * i, tl.pr.nb - 1);
                const uint32_t ga = tl.pr.rows_a ? (uint32_t)tl.pr.rows_a[ra] : (uint32_t)(tl.pr.row0_a + ra);
                const uint32_t gb = tl.pr.rows_b ? (uint32_t)tl.pr.rows_b[rb] : (uint32_t)(tl.pr.row0_b + rb);
                roff[i] = ga * (uint32_t)(kchunks * 16) + piece;
                roff[NR / 2 + i] = gb * (uint32_t)(kchunks * 16) + piece;
            }
            for (int kb = 0; kb < kchunks; ++kb, ++it) {
                const uint32_t s = it % T_STAGES;
                mbar_wait(bar_empty + s, ((it / T_STAGES) & 1) ^ 1);
                const uint32_t dst = smem0 + s * STAGE_BYTES + dst_blk;
                const uint4* src = ta.Xs + kb * 16;
#pragma unroll
                for (int i = 0; i < NR; ++i) {
                    if (i >= NR / 2 && tl.diag) break;                  // diagonal tile: B = A
                    const int rr = rsub + RSTEP * (i % (NR / 2));       // row within the 128-row block
                    cp_async_16(dst + (i / (NR / 2)) * (2 * BLK_BYTES) + rr * 128 + (((piece & 7) ^ (rr & 7)) << 4), src + roff[i]);
                }
                // asynchronous arrival: the barrier is signalled when this thread's copies have landed, the
                // thread itself moves on to the next stage (waiting here would serialise issue and MMA)
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar_full + s)) : "memory");
            }
        }
        cp_async_commit();
        cp_async_wait<0>();
    } else if (warp == MMA_WARP) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            uint32_t it = 0, segcount = 0;
            const uint32_t idesc256 = umma_idesc(256), idesc128 = umma_idesc(128);
            for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
                const uint32_t b_off = decode_tile(a, t).diag ? 0u : 2u * BLK_BYTES;   // [B_hi ; B_lo] = [A_hi ; A_lo] on the diagonal
                for (int k0 = 0; k0 < kchunks; k0 += SEG, ++segcount) {
                    const uint32_t as = segcount & 1;
                    mbar_wait(bar_tempty + as, ((segcount >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d = tmem_base + as * ACC_COLS;
                    const int k1 = min(k0 + SEG, kchunks);
                    for (int kb = k0; kb < k1; ++kb, ++it) {
                        const uint32_t s = it % T_STAGES;
                        mbar_wait(bar_full + s, (it / T_STAGES) & 1);
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // cp.async wrote the stage through the generic proxy
                        tc_fence_after();
                        const uint32_t base = smem0 + s * STAGE_BYTES;
#pragma unroll
                        for (int k = 0; k < KC / 16; ++k) {
                            const uint64_t a_hi = umma_desc(base + k * 32);
                            const uint64_t a_lo = umma_desc(base + BLK_BYTES + k * 32);
                            const uint64_t b_hl = umma_desc(base + b_off + k * 32);
                            umma_f16(d, a_hi, b_hl, idesc256, (kb > k0 || k > 0) ? 1u : 0u);
                            umma_f16(d + 128, a_lo, b_hl, idesc128, 1u);
                        }
                        umma_commit(bar_empty + s);                     // the stage is free when these MMAs have read it
                    }
                    umma_commit(bar_tfull + as);                        // this K segment's accumulators are complete
                }
            }
        }
        __syncwarp();
    } else {
        // ---------------- epilogue: float64 running sums over the K segments ----------------
        uint32_t segcount = 0;
        const int quarter = warp & 3, half = warp >> 2;                 // TMEM lanes 32*quarter.., columns 64*half..
        const double sc = (double)*ta.scale;
        const double unscale = 1.0 / (sc * sc);                        // exact: s is a power of two
        for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
            const Tile tl = decode_tile(a, t);
            double acc[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) acc[j] = 0.0;
            for (int k0 = 0; k0 < kchunks; k0 += SEG, ++segcount) {
                const uint32_t as = segcount & 1;
                mbar_wait(bar_tfull + as, (segcount >> 1) & 1);
                __syncwarp();                                           // tcgen05.ld is warp-collective
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * ACC_COLS + half * 64;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t big[16], small[16];
                    tmem_ld16(taddr + c * 16, big);
                    tmem_ld16(taddr + 128 + c * 16, small);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        acc[c * 16 + j] += (double)__uint_as_float(big[j]) + (double)__uint_as_float(small[j]) * (1.0 / LO_SCALE);
                }
                tc_fence_before();
                mbar_arrive(bar_tempty + as);
            }
            const int r = tl.ti * TM + quarter * 32 + lane;
            if (r < tl.pr.na) {
                const int64_t gr = tl.pr.rows_a ? (int64_t)tl.pr.rows_a[r] : (int64_t)(tl.pr.row0_a + r);
                const double sqr = (a.epilogue == EPI_EUCLID) ? a.sq[gr] : 0.0;
                const int c0 = tl.tj * TN + half * 64;
                double* orow = tl.pr.out + (int64_t)r * tl.pr.ldo + c0;
#pragma unroll
                for (int j = 0; j < 64; ++j) {
                    if (c0 + j >= tl.pr.nb) continue;
                    double v = acc[j] * unscale;
                    if (a.epilogue != EPI_GRAM) {
                        const int64_t gc = tl.pr.rows_b ? (int64_t)tl.pr.rows_b[c0 + j] : (int64_t)(tl.pr.row0_b + c0 + j);
                        if (a.epilogue == EPI_EUCLID) v = sqrt(fmax(fma(-2.0, v, sqr + a.sq[gc]), 0.0));
                        else v = 1.0 - v;
                        if (gr == gc) v = 0.0;
                    }
                    orow[j] = v;
                    if (tl.mirror) tl.pr.out[(int64_t)(c0 + j) * tl.pr.ldo + r] = v;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// max |x| over the table as float bits (non-negative floats order like unsigned integers)
__global__ void __launch_bounds__(256) absmax_kernel(const double* __restrict__ X, int64_t ldx, int64_t nrows, int G,
                                                     unsigned* __restrict__ out) {
    const int64_t total = nrows * (int64_t)G;
    float m = 0.f;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / G;
        m = fmaxf(m, fabsf((float)X[row * ldx + (e - row * G)]));
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

// scale[0] = 2^e with max|x| 2^e in [2^14, 2^15)  (1 for an all-zero table)
__global__ void scale_kernel(const unsigned* __restrict__ absmax, float* __restrict__ scale) {
    const float m = __uint_as_float(*absmax);
    int e = 0;
    if (m > 0.f && isfinite(m)) {
        frexpf(m, &e);                                  // m = f 2^e, f in [0.5, 1)
        e = 15 - e;
    }
    scale[0] = ldexpf(1.f, max(min(e, 100), -100));
}

// x -> (fp16(x s), fp16((x s - hi) 2^11)), laid out [row][K chunk of 64][hi 64 | lo 64], zero padded in K
__global__ void __launch_bounds__(256) split_fp16_kernel(const double* __restrict__ X, int64_t ldx, int64_t nrows, int G,
                                                         int kchunks, const float* __restrict__ scale, __half* __restrict__ Xs) {
    const int64_t per_row = (int64_t)kchunks * KC;
    const int64_t total = nrows * per_row;
    const double s = (double)*scale;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / per_row;
        const int k = (int)(e - row * per_row);
        const double x = k < G ? X[row * ldx + k] * s : 0.0;
        const __half hi = __double2half(x);
        const __half lo = __double2half((x - (double)__half2float(hi)) * LO_SCALE);
        __half* dst = Xs + row * per_row * 2 + (int64_t)(k / KC) * (2 * KC) + (k % KC);
        dst[0] = hi;
        dst[KC] = lo;
    }
}

}  // namespace

int sctda_launch_gemm_split(sctda_ctx* ctx, const GemmArgs& a, cudaStream_t st, bool reuse_split) {
    if (a.nrows <= 0) return sctda_fail(ctx, SCTDA_ERR_UNSUPPORTED, "split-precision contraction: the row count of X is unknown%s");
    const int kchunks = (a.G + KC - 1) / KC;
    if ((double)a.nrows * kchunks * 16.0 >= 4294967296.0)
        return sctda_fail(ctx, SCTDA_ERR_UNSUPPORTED, "split-precision contraction: operand too large for 32-bit row offsets%s");
    if (!(ctx->attr_done & 64u)) {
        SCTDA_CUDA(ctx, cudaFuncSetAttribute(gemm_split3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T_SMEM));
        ctx->attr_done |= 64u;
    }
    // workspace: the split operands (4 bytes per element) followed by the scale and the max scratch
    const size_t xs_bytes = (size_t)a.nrows * kchunks * 2 * KC * sizeof(__half);
    unsigned char* wsp = (unsigned char*)sctda_ws(ctx, WS_SPLIT, xs_bytes + 256);
    if (!wsp) return SCTDA_ERR_NOMEM;
    __half* Xs = (__half*)wsp;
    float* d_scale = (float*)(wsp + xs_bytes);
    unsigned* d_absmax = (unsigned*)(wsp + xs_bytes + 64);
    const int64_t elems = a.nrows * (int64_t)kchunks * KC;
    int64_t sgrid = cdiv64(elems, 256);
    if (sgrid > (int64_t)ctx->sm_count * 16) sgrid = (int64_t)ctx->sm_count * 16;
    if (!reuse_split) {
        SCTDA_CUDA(ctx, cudaMemsetAsync(d_absmax, 0, 4, st));
        absmax_kernel<<<(unsigned)sgrid, 256, 0, st>>>(a.X, a.ldx, a.nrows, a.G, d_absmax);
        SCTDA_LAUNCH_CHECK(ctx, "absmax_kernel");
        scale_kernel<<<1, 1, 0, st>>>(d_absmax, d_scale);
        SCTDA_LAUNCH_CHECK(ctx, "scale_kernel");
        split_fp16_kernel<<<(unsigned)sgrid, 256, 0, st>>>(a.X, a.ldx, a.nrows, a.G, kchunks, d_scale, Xs);
        SCTDA_LAUNCH_CHECK(ctx, "split_fp16_kernel");
    }
    TArgs ta;
    ta.g = a;
    ta.Xs = (const uint4*)Xs;
    ta.kchunks = kchunks;
    ta.scale = d_scale;
    SCTDA_TIME_BEGIN(ctx, st);
    gemm_split3_kernel<<<(unsigned)ctx->sm_count, T_THREADS, T_SMEM, st>>>(ta);
    SCTDA_TIME_END(ctx, st);
    SCTDA_LAUNCH_CHECK(ctx, "gemm_split3_kernel");
    return SCTDA_OK;
}
