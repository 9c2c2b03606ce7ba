// Hot path (b), shared-permutation form of the connectivity permutation test
// (UnrootedGraph.connectivity_pvalue, /root/reference/scTDA/main.py:966-992, "ref:"): the kernel
// that serves sctda_conn_perm_test when one [n, S] permutation block is applied to every gene.
//
// Third layout of this kernel.  conn.cu gathers one dense 256-byte row [cell][32 genes] per node
// member; conn_tile64.cu gathers [cell][64 genes] rows and predicates every lane's 16-byte load on
// a per-(cell, sector) mask.  The profile of the latter (profiles/ncu_conn64_r02.txt) shows two
// walls: the LSU data pipe (5,800 wavefronts per gene and permutation: one per touched 128-byte
// line plus the shared-memory traffic of compacting member lists) and instruction issue (667
// warp instructions per node visit and 64 genes).  This kernel keeps the arithmetic and the
// summation order (node members in list order, ref:978-981) and removes instructions and
// wavefronts:
//   * tiles of 128 genes, ET[tile][cell][128]; a lane owns four neighbouring genes = one 32-byte
//     sector of a gathered row, fetched with ONE 256-bit load (LDG.E.256), so the per-member index
//     work, the node epilogue bookkeeping and the adjacency bookkeeping are shared by twice as many genes;
//   * the composed index perm[cols[m]] of every member slot is computed ONCE per permutation by a
//     small kernel (Lr[n][Mtot], 16-bit) instead of once per gene tile and permutation in the hot loop:
//     no permutation row in shared memory, one coalesced 64-byte load per 32 members;
//   * a skipped sector costs no zeroing: the landing registers keep their previous (finite) contents
//     and the sum is acc = fma(value, m, acc) with m = 1.0 for a fetched sector and 0.0 for a skipped one
//     (fma(v, 1.0, acc) == acc + v and fma(stale, 0.0, acc) == acc bit for bit; ptxas turns a predicated
//     DADD into DADD + 2 FSEL, and zero-initialising costs one instruction per double);
//   * no compaction pass: every member is one (mask test, address, predicated load, multiplier) quadruple
//     plus four DFMAs for 128 genes.
//   * the log table is replicated eight times so that the 16-byte lookups of a quarter warp fall into eight
//     different bank groups whatever the indices (4 wavefronts per lookup instead of ~9).
// Tried and dropped (bit-identical, measured on the B200, probe of 4,096 genes x 1,184 permutations = 112 ms):
// `prefetch.global.L1` of the next gather round's sectors (144 ms: the prefetches queue in the same LSU pipe);
// the next round's list entries read one round early (126 ms: spills at 80 registers); gather rounds of six or
// eight members with 20 / 16 warps (125 / 131 ms); a table-free log (116 ms: six more FP64 instructions per node
// value cost more than the 4 wavefronts of the lookup); two or three independent warp groups per CTA (named
// barriers) working on different permutation chunks so that one group's adjacency phase (20 % of the time, LSU
// pipe nearly idle) overlaps the others' gathers (+5 % / +35 % time: the phases share one L1/L2 path and the
// smaller groups have fewer rows in flight); `prefetch.global.L1` of the node values the NEXT adjacency row reads
// (+10 % time); rows stored COMPACTED to their non-empty sectors
// (a lane's sector at popc(mask & lanes below)), so that the fetched sectors are contiguous.  The LSU data pipe
// charges one wavefront per quad of lanes with an active lane, not per 128-byte line: wavefronts per
// gene*permutation did not move (3,900 -> 3,900) and the extra offset lookups made the kernel 40 % slower.
// The node values, the float32 rounding of ref:972, 0.693147, the fixed reduction order (node k -> warp
// k mod NW, partial sums combined in warp order) are those of the other two kernels.
#include <stdlib.h>

#include "common.cuh"
#include "conn_math.cuh"

namespace {

constexpr int kGPL = 4;               // genes per lane
constexpr int kT = 32 * kGPL;         // genes per tile
constexpr int kChunk = 8;             // permutations per work item
constexpr int kSeg = 64;              // members staged at a time (two chunks of 32)
#ifndef SCTDA_CONNW_GROUP
#define SCTDA_CONNW_GROUP 4
#endif
constexpr int kGroup = SCTDA_CONNW_GROUP;   // members per gather round (even)
constexpr int kList = kSeg + kGroup;  // list entries per buffer: up to 64 members + padding to a multiple of kGroup
constexpr int kTabRep = 8;             // copies of the log table in shared memory (one per bank group)
constexpr unsigned kGuard = 1u << 16; // half-width of the ambiguity band of pm_fast, in units of 2^-52 of the binade

__device__ double2 g_logtab_w[128];   // (1/c_j rounded, -log(1/c_j rounded)), c_j = 1 + (j + 1/2)/128

// float32(log1p(acc/q) / 0.693147), same bits as pm_log2_value(div_rn(acc, q, inv_q)); see conn_tile64.cu
// (pm_fast) for the error analysis: table-driven evaluation, and a fall-back to the full-accuracy path
// whenever the float32 rounding could depend on the last bits.
__device__ __forceinline__ float pm_fast_w(double acc, double q, double inv_q, const double2* __restrict__ tab) {
#ifdef SCTDA_CONNW_NOTABLE
    return pm_log2_value(div_rn(acc, q, inv_q));
#endif
    const double t = acc * inv_q;
    const double u = 1.0 + t;
    const int hu = __double2hiint(u);
    const int k = (hu >> 20) - 1023;
    const double2 tc = tab[((hu >> 13) & 127) * kTabRep];                 // tab is pre-offset by lane & 7
    const double m = __hiloint2double((hu & 0x000fffff) | 0x3ff00000, __double2loint(u));
    const double r = fma(m, tc.x, -1.0);
    double p = fma(r, 0.2, -0.25);
    p = fma(r, p, 0.33333333333333331);
    p = fma(r, p, -0.5);
    p = fma(r, p, 1.0);
    const double L = fma(r, p, fma((double)k, 0.69314718055994529, tc.y));
    const double y = L * (1.0 / 0.693147);
    const unsigned lo = (unsigned)__double2loint(y) & 0x1fffffffu;
    const bool clear = (lo - 0x10000000u + kGuard) >= 2u * kGuard;     // away from the rounding boundary
    float f = (float)y;
    if (acc == 0.0) f = 0.0f;                                          // an empty node: log1p(0) = 0
    else if (!(clear && t >= 0.00390625 && t <= 1e30)) f = pm_log2_value(div_rn(acc, q, inv_q));
    return f;
}

struct WideArgs {
    int V, S, Sp, G, n, log2_mode, nchunks, tiles, Mtot;
    int r0;                     // first permutation of this launch (row 0 of Lr)
    int nloc;                   // permutations of this launch
    long long items;
    const int32_t* node_off;
    const int32_t* adj_off;
    const int32_t* adj_idx;
    const double* adj_w;
    const double* ET;           // [tiles][S][kT]
    const uint32_t* sec;        // [tiles][Sp]: bit l set iff lane l's four genes hold a non-zero at this cell
    const uint16_t* Lr;         // [nloc][Mtot]: perm[cols[m]] of every member slot
    const int32_t* gmap;        // [tiles*kT]: gene of each tile slot, -1 = padding
    const double* inv_q;        // [V]
    const double* observed;
    double tie_rel;
    int32_t* exceed;
    int32_t* ties;
    double* scores;
    float* pm;                  // [grid][V][kT]
    unsigned long long* counter;
    const int* unit_w;          // 1 when all adjacency weights are the same number *w0
    const double* w0;
    const uint2* rowinfo;       // [V]: (first chunk, number of edges) of every adjacency row in ids16, or NULL
    const uint16_t* ids16;      // adjacency rows as 16-bit node ids, every row padded to whole chunks of 8
    const int* packed_ok;       // 1 when rowinfo / ids16 hold the whole adjacency
};

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// L2 eviction priorities (createpolicy): the gene tile is re-read by every CTA for every permutation and
// must stay resident; the node-value scratch (1 MB per CTA, written and read once per permutation) and the
// composed indices (read once) are marked evict-first so that they do not push the tile out.
__device__ __forceinline__ unsigned long long policy_evict_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned ld_u16_stream(const uint16_t* p, unsigned long long pol) {
    unsigned short c;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(c) : "l"(p), "l"(pol));
    return (unsigned)c;
}
__device__ __forceinline__ void st_pm(float* p, float4 v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ float4 ld_pm(const float* p, unsigned long long pol) {
    float4 v;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol) : "memory");
    return v;
}

// One member: the lane tests its bit of the member's mask, fetches its 32-byte sector of the row at `off`
// (32-byte units from the lane's base) under that predicate into registers that keep their contents
// otherwise, and returns the high word of the multiplier (1.0 when fetched, 0.0 when not).
__device__ __forceinline__ unsigned gather_one(double (&v)[kGPL], unsigned off, unsigned mask, unsigned lanebit,
                                               unsigned long long pol, const double* base) {
    unsigned mh;
    asm volatile(
        "{\n.reg .pred p;\n.reg .b32 t;\n.reg .b64 ad;\n"
        "and.b32 t, %6, %7;\n setp.ne.u32 p, t, 0;\n"
        "mad.wide.u32 ad, %5, 32, %9;\n"
        "@p ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f64 {%0, %1, %2, %3}, [ad], %8;\n"   // a gathered row is used once
        "selp.b32 %4, 0x3ff00000, 0, p;\n}\n"
        : "+d"(v[0]), "+d"(v[1]), "+d"(v[2]), "+d"(v[3]), "=r"(mh)
        : "r"(off), "r"(mask), "r"(lanebit), "l"(pol), "l"(base));
    return mh;
}

struct Landing {
    double v[kGroup][kGPL];
};

// Sequential sums over a staged member list (member order), kGroup members per gather round: the loads of a
// round are issued before its first add.
__device__ __forceinline__ void sum_list(double (&acc)[kGPL], Landing& ld, const uint2* list, int count, unsigned lanebit,
                                         unsigned long long pol, const double* base) {
    for (int g = 0; g < count; g += kGroup) {
        const uint4* q = (const uint4*)(list + g);                     // broadcast loads: (off, mask) pairs
        uint4 e[kGroup / 2];
#pragma unroll
        for (int u = 0; u < kGroup / 2; ++u) e[u] = q[u];
        unsigned mh[kGroup];
#pragma unroll
        for (int u = 0; u < kGroup / 2; ++u) {
            mh[2 * u] = gather_one(ld.v[2 * u], e[u].x, e[u].y, lanebit, pol, base);
            mh[2 * u + 1] = gather_one(ld.v[2 * u + 1], e[u].z, e[u].w, lanebit, pol, base);
        }
#pragma unroll
        for (int u = 0; u < kGroup; ++u) {
            const double m = __hiloint2double((int)mh[u], 0);
#pragma unroll
            for (int j = 0; j < kGPL; ++j) acc[j] = __fma_rn(ld.v[u][j], m, acc[j]);
        }
    }
    __syncwarp();
}

// Stages up to 64 members of a node, member order kept: entry = (row offset of the member's cell under this
// permutation in 32-byte units, its lane mask).  Members whose whole 128-gene row is zero (29 % of them on the
// bench table) are dropped here: adding +0.0 is the identity.  Returns the number of staged entries; the
// entries up to the next multiple of kGroup get an empty mask.
__device__ __forceinline__ int stage_list(uint2* list, const uint32_t* s_sec, unsigned c0, unsigned c1, int nmem, int lane) {
    const unsigned m0 = lane < nmem ? s_sec[c0] : 0u;
    const unsigned b0 = __ballot_sync(0xffffffffu, m0 != 0u);
    const unsigned below = (1u << lane) - 1u;
    if (m0) list[__popc(b0 & below)] = make_uint2(c0 * (unsigned)(kT / 4), m0);
    int count = __popc(b0);
    if (nmem > 32) {
        const unsigned m1 = lane + 32 < nmem ? s_sec[c1] : 0u;
        const unsigned b1 = __ballot_sync(0xffffffffu, m1 != 0u);
        if (m1) list[count + __popc(b1 & below)] = make_uint2(c1 * (unsigned)(kT / 4), m1);
        count += __popc(b1);
    }
    if (lane < kGroup) list[count + lane] = make_uint2(0u, 0u);
    __syncwarp();
    return count;
}

// s += sum over the first `cnt` (1..8) edges held by lanes 0..cnt-1 of w * pm[l], in edge order.
template <int C, bool UNIT>
__device__ __forceinline__ void spmv_rows_c(double (&s)[kGPL], const float* pm, int li, double lw, int lane,
                                            unsigned long long pol) {
    float4 v[C];
#pragma unroll
    for (int u = 0; u < C; ++u) {
        const int l = __shfl_sync(0xffffffffu, li, u);
        v[u] = ld_pm(pm + (size_t)l * kT + 4 * lane, pol);
    }
#pragma unroll
    for (int u = 0; u < C; ++u) {
        if (UNIT) {                                                // weight 1.0: w * x == x exactly
            s[0] += (double)v[u].x; s[1] += (double)v[u].y; s[2] += (double)v[u].z; s[3] += (double)v[u].w;
        } else {
            const double wt = __shfl_sync(0xffffffffu, lw, u);
            s[0] += wt * (double)v[u].x; s[1] += wt * (double)v[u].y;
            s[2] += wt * (double)v[u].z; s[3] += wt * (double)v[u].w;
        }
    }
}

template <bool UNIT>
__device__ __forceinline__ void spmv_rows(double (&s)[kGPL], const float* pm, int li, double lw, int cnt, int lane,
                                          unsigned long long pol) {
    switch (cnt) {
        case 1: spmv_rows_c<1, UNIT>(s, pm, li, lw, lane, pol); break;
        case 2: spmv_rows_c<2, UNIT>(s, pm, li, lw, lane, pol); break;
        case 3: spmv_rows_c<3, UNIT>(s, pm, li, lw, lane, pol); break;
        case 4: spmv_rows_c<4, UNIT>(s, pm, li, lw, lane, pol); break;
        case 5: spmv_rows_c<5, UNIT>(s, pm, li, lw, lane, pol); break;
        case 6: spmv_rows_c<6, UNIT>(s, pm, li, lw, lane, pol); break;
        case 7: spmv_rows_c<7, UNIT>(s, pm, li, lw, lane, pol); break;
        default: spmv_rows_c<8, UNIT>(s, pm, li, lw, lane, pol); break;
    }
}

// Unit weights, packed rows: s += sum over the first C (1..8) edges of a chunk of 8 node ids (four registers, the
// same in every lane) of pm[l], in edge order.  No shuffles: a SHFL costs ~1.8 wavefronts of the LSU data pipe
// (profiles/ncu_connwide_r02.txt), three per edge in the generic path; the chunk is one 16-byte load.
template <int C>
__device__ __forceinline__ void spmv_ids_c(double (&s)[kGPL], const float* pm, const uint4& ids, int lane,
                                           unsigned long long pol) {
    const unsigned wds[4] = {ids.x, ids.y, ids.z, ids.w};
    float4 v[C];
#pragma unroll
    for (int u = 0; u < C; ++u) {
        const unsigned l = (u & 1) ? (wds[u >> 1] >> 16) : (wds[u >> 1] & 0xffffu);
        v[u] = ld_pm(pm + (size_t)l * kT + 4 * lane, pol);
    }
#pragma unroll
    for (int u = 0; u < C; ++u) {                                  // weight 1.0: w * x == x exactly
        s[0] += (double)v[u].x; s[1] += (double)v[u].y; s[2] += (double)v[u].z; s[3] += (double)v[u].w;
    }
}

__device__ __forceinline__ void spmv_ids(double (&s)[kGPL], const float* pm, const uint4& ids, int cnt, int lane,
                                         unsigned long long pol) {
    switch (cnt) {
        case 1: spmv_ids_c<1>(s, pm, ids, lane, pol); break;
        case 2: spmv_ids_c<2>(s, pm, ids, lane, pol); break;
        case 3: spmv_ids_c<3>(s, pm, ids, lane, pol); break;
        case 4: spmv_ids_c<4>(s, pm, ids, lane, pol); break;
        case 5: spmv_ids_c<5>(s, pm, ids, lane, pol); break;
        case 6: spmv_ids_c<6>(s, pm, ids, lane, pol); break;
        case 7: spmv_ids_c<7>(s, pm, ids, lane, pol); break;
        default: spmv_ids_c<8>(s, pm, ids, lane, pol); break;
    }
}

template <int NW>
__global__ void __launch_bounds__(NW * 32, 1) conn_perm_wide_kernel(WideArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* s_sec = (uint32_t*)smem_raw;                        // [Sp]
    double2* s_tab = (double2*)(s_sec + a.Sp);                    // [128][kTabRep]
    double* s_tot = (double*)(s_tab + 128 * kTabRep);             // [NW][kT]
    double* s_part = s_tot + NW * kT;                             // [NW][kT]
    uint2* s_list = (uint2*)(s_part + NW * kT);                   // [NW][2][kList]
    __shared__ long long s_claim[2];
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const unsigned lanebit = 1u << lane;
    const unsigned long long pol_keep = policy_evict_last(), pol_stream = policy_evict_first();
    float* pm = a.pm + (size_t)blockIdx.x * a.V * kT;
    const double cor = (double)a.V / (double)(a.V - 1);           // ref:989
    const double w0 = *a.w0;
    const bool unit_w = *a.unit_w != 0 && w0 == 1.0;
    // one weight, 1.0 or 2.0: sum of w0 * x == w0 * (sum of x) bit for bit (scaling by a power of two is exact)
    const bool packed = *a.unit_w != 0 && (w0 == 1.0 || w0 == 2.0) && a.rowinfo != nullptr && *a.packed_ok != 0;
    const int chunks16 = a.Sp >> 2;                               // 16-byte pieces of the mask row
    for (int i = threadIdx.x; i < 128 * kTabRep; i += NW * 32) s_tab[i] = g_logtab_w[i / kTabRep];
    const double2* tab = s_tab + (lane & (kTabRep - 1));
    if (threadIdx.x == 0) {
        s_claim[0] = (long long)atomicAdd(a.counter, 1ull);
        s_claim[1] = (long long)atomicAdd(a.counter, 1ull);
    }
    __syncthreads();
    long long item = s_claim[0], next_item = s_claim[1];
    int cur_tile = -1;
    for (; item < a.items; item = next_item, next_item = s_claim[0]) {
        __syncthreads();                                           // s_claim[0] read, previous item finished
        if (threadIdx.x == 0) s_claim[0] = (long long)atomicAdd(a.counter, 1ull);
        const int tileIdx = (int)(item / a.nchunks);
        const int chunk = (int)(item % a.nchunks);
        if (tileIdx != cur_tile) {                                 // lane masks of the new tile
            const uint32_t* src = a.sec + (size_t)tileIdx * a.Sp;
            for (int i = threadIdx.x; i < chunks16; i += NW * 32) cp_async_16(s_sec + 4 * i, src + 4 * i);
            cp_async_wait_all();
            cur_tile = tileIdx;
            __syncthreads();
        }
        const double* base = a.ET + (size_t)tileIdx * a.S * kT + kGPL * lane;
        // warps 0..3 finish gene slot w of every lane (the statistic of gene 4*lane + w)
        const int gfin = w < kGPL ? a.gmap[tileIdx * kT + kGPL * lane + w] : -1;
        const double obfin = gfin >= 0 ? a.observed[gfin] : 0.0;
        int ex0 = 0, ti0 = 0;
        const int r_end = min(a.nloc, (chunk + 1) * kChunk);
        for (int r = chunk * kChunk; r < r_end; ++r) {
            const uint16_t* Lrow = a.Lr + (size_t)r * a.Mtot;
            // ---- node values (ref:978-987) ----
            // The list of node k+NW is staged while node k is summed: its composed member indices are fetched
            // before the sums and turned into list entries after them.
            Landing ld;                                            // landing registers: finite contents (this tile's genes) at all times
#pragma unroll
            for (int u = 0; u < kGroup; ++u)
#pragma unroll
                for (int j = 0; j < kGPL; ++j) ld.v[u][j] = 0.0;
            uint2* lst = s_list + (size_t)w * 2 * kList;
            int cur = 0;
            int k = w, beg = 0, end = 0, count = 0;
            if (k < a.V) {
                beg = __ldg(a.node_off + k);
                end = __ldg(a.node_off + k + 1);
                const int n0 = min(kSeg, end - beg);
                const unsigned c0 = lane < n0 ? ld_u16_stream(Lrow + beg + lane, pol_stream) : 0u;
                const unsigned c1 = lane + 32 < n0 ? ld_u16_stream(Lrow + beg + 32 + lane, pol_stream) : 0u;
                count = stage_list(lst, s_sec, c0, c1, n0, lane);
            }
            while (k < a.V) {
                const int len = end - beg;
                const int kn = k + NW;
                int begn = 0, endn = 0, nn = 0;
                unsigned cn0 = 0, cn1 = 0;
                if (kn < a.V) {                                    // next node's member cells, in flight under the sums
                    begn = __ldg(a.node_off + kn);
                    endn = __ldg(a.node_off + kn + 1);
                    nn = min(kSeg, endn - begn);
                    if (lane < nn) cn0 = ld_u16_stream(Lrow + begn + lane, pol_stream);
                    if (lane + 32 < nn) cn1 = ld_u16_stream(Lrow + begn + 32 + lane, pol_stream);
                }
                double acc[kGPL];
#pragma unroll
                for (int j = 0; j < kGPL; ++j) acc[j] = 0.0;
                sum_list(acc, ld, lst + cur * kList, count, lanebit, pol_keep, base);
                for (int sb = beg + kSeg; sb < end; sb += kSeg) {  // nodes above 64 members: further segments, in order
                    uint2* l2 = lst + cur * kList;
                    const int ns = min(kSeg, end - sb);
                    const unsigned c0 = lane < ns ? ld_u16_stream(Lrow + sb + lane, pol_stream) : 0u;
                    const unsigned c1 = lane + 32 < ns ? ld_u16_stream(Lrow + sb + 32 + lane, pol_stream) : 0u;
                    const int cs = stage_list(l2, s_sec, c0, c1, ns, lane);
                    sum_list(acc, ld, l2, cs, lanebit, pol_keep, base);
                }
                const double q = (double)len, iq = __ldg(a.inv_q + k);
                float4 pv;
                if (a.log2_mode) {                                 // ref:982-983, float32 store of ref:972
                    pv.x = pm_fast_w(acc[0], q, iq, tab);
                    pv.y = pm_fast_w(acc[1], q, iq, tab);
                    pv.z = pm_fast_w(acc[2], q, iq, tab);
                    pv.w = pm_fast_w(acc[3], q, iq, tab);
                } else {                                           // ref:986
                    pv.x = (float)div_rn(acc[0], q, iq);
                    pv.y = (float)div_rn(acc[1], q, iq);
                    pv.z = (float)div_rn(acc[2], q, iq);
                    pv.w = (float)div_rn(acc[3], q, iq);
                }
                st_pm(pm + (size_t)k * kT + 4 * lane, pv, pol_stream);
                if (kn < a.V) {
                    count = stage_list(lst + (cur ^ 1) * kList, s_sec, cn0, cn1, nn, lane);
                }
                cur ^= 1;
                k = kn; beg = begn; end = endn;
            }
            __syncthreads();
#ifndef SCTDA_CONNW_NO_PM_PREFETCH
            // The node values of this permutation (1 MB, written over the whole node phase) have mostly left L2 by now
            // (148 CTAs x 1 MB next to the tile): ask for all of them at once, so that the rows below are served from
            // L2 instead of paying a DRAM round trip each, one dependent row after the other.
            for (int i = threadIdx.x; i < a.V * (kT / 32); i += NW * 32)
                asm volatile("prefetch.global.L2 [%0];" :: "l"(pm + (size_t)i * 32));
#endif
            // ---- tot (ref:984, this warp's nodes in ascending order) and p'Ap on the un-normalised values; the
            // 1/tot^2 is applied once (ref:988-989) ----
            double part[kGPL], tot[kGPL];
#pragma unroll
            for (int j = 0; j < kGPL; ++j) part[j] = 0.0, tot[j] = 0.0;
            if (packed) {                                          // unit weights, rows as chunks of eight 16-bit ids
                int kk = w;
                uint2 ri = make_uint2(0u, 0u), rin = make_uint2(0u, 0u);
                uint4 ids = make_uint4(0u, 0u, 0u, 0u);
                if (kk < a.V) {
                    ri = __ldg(a.rowinfo + kk);
                    if (ri.y) ids = __ldg((const uint4*)(a.ids16 + (size_t)ri.x * 8));
                }
                if (kk + NW < a.V) rin = __ldg(a.rowinfo + kk + NW);
                while (kk < a.V) {
                    const int kn = kk + NW, k2 = kk + 2 * NW;
                    uint4 idsn = make_uint4(0u, 0u, 0u, 0u);
                    uint2 ri2 = make_uint2(0u, 0u);
                    if (kn < a.V && rin.y) idsn = __ldg((const uint4*)(a.ids16 + (size_t)rin.x * 8));   // next row's first chunk
                    if (k2 < a.V) ri2 = __ldg(a.rowinfo + k2);     // row header two rows ahead
                    const float4 vk = ld_pm(pm + (size_t)kk * kT + 4 * lane, pol_stream);
                    tot[0] += (double)vk.x;
                    tot[1] += (double)vk.y;
                    tot[2] += (double)vk.z;
                    tot[3] += (double)vk.w;
                    if (ri.y) {
                        double s[kGPL];
#pragma unroll
                        for (int j = 0; j < kGPL; ++j) s[j] = 0.0;
                        spmv_ids(s, pm, ids, min(8, (int)ri.y), lane, pol_stream);
                        for (int c = 1; 8 * c < (int)ri.y; ++c) {  // rows with more than eight edges
                            const uint4 more = __ldg((const uint4*)(a.ids16 + ((size_t)ri.x + c) * 8));
                            spmv_ids(s, pm, more, min(8, (int)ri.y - 8 * c), lane, pol_stream);
                        }
                        part[0] += (double)vk.x * (w0 * s[0]);
                        part[1] += (double)vk.y * (w0 * s[1]);
                        part[2] += (double)vk.z * (w0 * s[2]);
                        part[3] += (double)vk.w * (w0 * s[3]);
                    }
                    kk = kn; ri = rin; rin = ri2; ids = idsn;
                }
            } else {
                int kk = w, eb = 0, ee = 0, ebn = 0, een = 0, li = 0;
                double lw = 0.0;
                if (kk < a.V) {
                    eb = __ldg(a.adj_off + kk);
                    ee = __ldg(a.adj_off + kk + 1);
                    if (lane < min(8, ee - eb)) {
                        li = __ldg(a.adj_idx + eb + lane);
                        if (!unit_w) lw = __ldg(a.adj_w + eb + lane);
                    }
                }
                if (kk + NW < a.V) {
                    ebn = __ldg(a.adj_off + kk + NW);
                    een = __ldg(a.adj_off + kk + NW + 1);
                }
                while (kk < a.V) {
                    const int kn = kk + NW, k2 = kk + 2 * NW;
                    int lin = 0, eb2 = 0, ee2 = 0;
                    double lwn = 0.0;
                    if (kn < a.V && lane < min(8, een - ebn)) {    // next row's first edges
                        lin = __ldg(a.adj_idx + ebn + lane);
                        if (!unit_w) lwn = __ldg(a.adj_w + ebn + lane);
                    }
                    if (k2 < a.V) {                                // offsets two rows ahead
                        eb2 = __ldg(a.adj_off + k2);
                        ee2 = __ldg(a.adj_off + k2 + 1);
                    }
                    const float4 vk = ld_pm(pm + (size_t)kk * kT + 4 * lane, pol_stream);
                    tot[0] += (double)vk.x;
                    tot[1] += (double)vk.y;
                    tot[2] += (double)vk.z;
                    tot[3] += (double)vk.w;
                    if (ee > eb) {
                        double s[kGPL];
#pragma unroll
                        for (int j = 0; j < kGPL; ++j) s[j] = 0.0;
                        if (unit_w) spmv_rows<true>(s, pm, li, lw, min(8, ee - eb), lane, pol_stream);
                        else spmv_rows<false>(s, pm, li, lw, min(8, ee - eb), lane, pol_stream);
                        for (int e0 = eb + 8; e0 < ee; e0 += 8) {  // rows with more than eight edges
                            const int cnt = min(8, ee - e0);
                            int l2 = 0;
                            double w2 = 0.0;
                            if (lane < cnt) {
                                l2 = __ldg(a.adj_idx + e0 + lane);
                                if (!unit_w) w2 = __ldg(a.adj_w + e0 + lane);
                            }
                            if (unit_w) spmv_rows<true>(s, pm, l2, w2, cnt, lane, pol_stream);
                            else spmv_rows<false>(s, pm, l2, w2, cnt, lane, pol_stream);
                        }
                        part[0] += (double)vk.x * s[0];
                        part[1] += (double)vk.y * s[1];
                        part[2] += (double)vk.z * s[2];
                        part[3] += (double)vk.w * s[3];
                    }
                    kk = kn; eb = ebn; ee = een; ebn = eb2; een = ee2; li = lin; lw = lwn;
                }
            }
#pragma unroll
            for (int j = 0; j < kGPL; ++j) {
                s_part[w * kT + kGPL * lane + j] = part[j];
                s_tot[w * kT + kGPL * lane + j] = tot[j];
            }
            __syncthreads();
            if (w < kGPL) {                                        // warp j finishes gene slot j of every lane
                const int j = w;
                double t0 = 0.0, p0 = 0.0;
#pragma unroll 8
                for (int i = 0; i < NW; ++i) {
                    t0 += s_tot[i * kT + kGPL * lane + j];
                    p0 += s_part[i * kT + kGPL * lane + j];
                }
                const double c0 = cor * (p0 / (t0 * t0));
                if (gfin >= 0) {
                    const bool tied = fabs(c0 - obfin) <= a.tie_rel * fabs(obfin);   // the same number up to rounding
                    if (c0 > obfin && !tied) ++ex0;                // ref:990, strict: a tie does not exceed
                    if (tied) ++ti0;
                    if (a.scores) a.scores[(size_t)gfin * a.n + a.r0 + r] = c0;
                }
            }
            __syncthreads();
        }
        if (gfin >= 0) {
            if (ex0) atomicAdd(a.exceed + gfin, ex0);
            if (a.ties && ti0) atomicAdd(a.ties + gfin, ti0);
        }
    }
}

// ET[tile][s][slot] = f(Y[gmap[tile*128+slot]][samples[s]]), f(x) = 2^x - 1 (log2 mode) or x; padding
// slots are zeros.  sec[tile][s]: bit l set iff one of the slots 4l..4l+3 holds anything but a zero.
__global__ void __launch_bounds__(256) et_prep_wide_kernel(const double* __restrict__ Y, int64_t ldy, int S, int Sp,
                                                           const int32_t* __restrict__ samples,
                                                           const int32_t* __restrict__ gmap, int log2_mode,
                                                           double* __restrict__ ET, uint32_t* __restrict__ sec) {
    __shared__ double tile[kT][33];
    const int tileIdx = blockIdx.y;
    const int s0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;        // 32 x 8
    const int s = s0 + tx;
    int col = 0;
    if (s < S) col = samples ? samples[s] : s;
#pragma unroll
    for (int r = 0; r < kT / 8; ++r) {
        const int gl = ty + 8 * r;
        const int g = gmap[tileIdx * kT + gl];
        double v = 0.0;
        if (g >= 0 && s < S) {
            const double x = Y[(int64_t)g * ldy + col];
            v = log2_mode ? (exp2(x) - 1.0) : x;
        }
        tile[gl][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int sl = ty + 8 * r;                                  // cell of this warp
        double v[kGPL];
        bool any = false;
#pragma unroll
        for (int j = 0; j < kGPL; ++j) {
            v[j] = tile[kGPL * tx + j][sl];
            any = any || !(v[j] == 0.0);
        }
        const unsigned m = __ballot_sync(0xffffffffu, any);
        if (s0 + sl < S) {
            double* row = ET + ((int64_t)tileIdx * S + s0 + sl) * kT + kGPL * tx;
            *(double2*)row = make_double2(v[0], v[1]);
            *(double2*)(row + 2) = make_double2(v[2], v[3]);
            if (tx == 0) sec[(int64_t)tileIdx * Sp + s0 + sl] = m;
        }
    }
}

// Lr[r][m] = perm[r0 + r][cols[m]]: the cell that sits at member slot m under permutation r0 + r.
__global__ void __launch_bounds__(256) compose_kernel(const int32_t* __restrict__ perm, int S, int r0, int nloc,
                                                      const int32_t* __restrict__ cols, int Mtot,
                                                      uint16_t* __restrict__ Lr) {
    const int r = blockIdx.y;
    const int32_t* prow = perm + (size_t)(r0 + r) * S;
    uint16_t* out = Lr + (size_t)r * Mtot;
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < Mtot; m += gridDim.x * blockDim.x)
        out[m] = (uint16_t)__ldg(prow + __ldg(cols + m));
    (void)nloc;
}

__global__ void gmap_wide_kernel(int32_t* gmap, const int32_t* __restrict__ order, int G, int slots) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < slots) gmap[i] = i < G ? (order ? order[i] : i) : -1;
}

__global__ void inv_size_wide_kernel(const int32_t* __restrict__ off, int V, double* __restrict__ inv_q) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < V) inv_q[k] = 1.0 / (double)(off[k + 1] - off[k]);
}

// *flag (preset to 1) is cleared when the edge weights are not all the same number; *w0 = that number (the first
// weight).  graph.DeviceGraph hands over triu(A + A', 1) + diag(A): 2.0 everywhere for a 0/1 adjacency without loops.
__global__ void unit_weight_wide_kernel(const int32_t* __restrict__ adj_off, int V, const double* __restrict__ w, int* flag,
                                        double* w0) {
    const int nnz = adj_off[V];
    const double first = nnz > 0 ? w[0] : 1.0;
    if (blockIdx.x == 0 && threadIdx.x == 0) *w0 = first;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += gridDim.x * blockDim.x)
        if (w[e] != first) *flag = 0;
}

// Adjacency rows as 16-bit node ids in chunks of eight (the unit-weight path of the kernel): rowinfo[k] =
// (first chunk, number of edges).  One CTA; *ok (preset to 1) is cleared when `cap` chunks do not hold the rows.
__global__ void __launch_bounds__(1024) pack_adj_wide_kernel(const int32_t* __restrict__ adj_off,
                                                             const int32_t* __restrict__ adj_idx, int V, unsigned cap,
                                                             uint2* __restrict__ rowinfo, uint16_t* __restrict__ ids16,
                                                             int* ok) {
    __shared__ unsigned tot[1024];
    const int t = threadIdx.x;
    const int per = (V + 1023) / 1024;
    const int k0 = min(V, t * per), k1 = min(V, k0 + per);
    unsigned run = 0;
    for (int k = k0; k < k1; ++k) run += (unsigned)(adj_off[k + 1] - adj_off[k] + 7) >> 3;
    tot[t] = run;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {                            // inclusive scan of the per-thread chunk counts
        const unsigned add = t >= d ? tot[t - d] : 0u;
        __syncthreads();
        tot[t] += add;
        __syncthreads();
    }
    const bool fits = tot[1023] <= cap;
    if (t == 0 && !fits) *ok = 0;
    if (!fits) return;
    unsigned c = tot[t] - run;
    for (int k = k0; k < k1; ++k) {
        const int eb = adj_off[k], deg = adj_off[k + 1] - eb;
        rowinfo[k] = make_uint2(c, (unsigned)deg);
        for (int e = 0; e < ((deg + 7) & ~7); ++e) ids16[(size_t)c * 8 + e] = e < deg ? (uint16_t)adj_idx[eb + e] : (uint16_t)0;
        c += (unsigned)(deg + 7) >> 3;
    }
}

int ensure_logtab_w(sctda_ctx* ctx) {
    if (ctx->attr_done & 0x400u) return SCTDA_OK;
    double2 h[128];
    for (int j = 0; j < 128; ++j) {
        const long double c = 1.0L + ((long double)j + 0.5L) / 128.0L;
        const double invc = (double)(1.0L / c);
        h[j].x = invc;
        h[j].y = (double)(-logl((long double)invc));
    }
    SCTDA_CUDA(ctx, cudaMemcpyToSymbol(g_logtab_w, h, sizeof(h)));
    ctx->attr_done |= 0x400u;
    return SCTDA_OK;
}

template <int NW>
int launch_wide(sctda_ctx* ctx, const WideArgs& a, int grid, size_t smem, cudaStream_t st) {
    auto kern = conn_perm_wide_kernel<NW>;
    SCTDA_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, NW * 32, smem, st>>>(a);
    return SCTDA_OK;
}

}  // namespace

static size_t wide_smem(int Sp, int NW) {
    return (size_t)Sp * 4 + 128 * kTabRep * sizeof(double2) + (size_t)2 * NW * kT * sizeof(double) + (size_t)NW * 2 * kList * sizeof(uint2);
}

static int wide_warps() {
    static int nw = 0;
    if (!nw) {
        const char* e = getenv("SCTDA_CONNW_WARPS");
        nw = e ? atoi(e) : 24;
        if (nw != 8 && nw != 16 && nw != 20 && nw != 24 && nw != 32) nw = 24;
    }
    return nw;
}

// True when sctda_conn_perm_test's shared-permutation call can run on the 128-gene kernel.
bool sctda_wide_eligible(const sctda_graph* gr) {
    const int Sp = (gr->S + 3) & ~3;
    return gr->S <= 65535 && gr->V >= 2 && wide_smem(Sp, wide_warps()) <= (size_t)220 * 1024;
}

// The shared-permutation test on 128-gene tiles.  d_order [G] (or NULL): gene of each tile slot.
int sctda_conn_perm_wide(sctda_ctx* ctx, const sctda_graph* gr, int G, const double* d_Y, int64_t ldy, int log2_mode,
                         int n, const int32_t* d_perm, const int32_t* d_order, const double* d_observed,
                         double tie_rel, int32_t* d_exceed, int32_t* d_ties, double* d_scores, cudaStream_t st) {
    int rc = ensure_logtab_w(ctx);
    if (rc) return rc;
    const int tiles = (G + kT - 1) / kT;
    const int Sp = (gr->S + 3) & ~3;
    int Mtot = gr->Mtot;
    if (Mtot <= 0) {                                               // not given: read node_off[V] (one synchronising 4-byte copy)
        SCTDA_CUDA(ctx, cudaMemcpyAsync(&Mtot, gr->d_node_off + gr->V, sizeof(int), cudaMemcpyDeviceToHost, st));
        SCTDA_CUDA(ctx, cudaStreamSynchronize(st));
    }
    if (Mtot <= 0) return sctda_fail(ctx, SCTDA_ERR_INVALID, "conn_perm_test: the node member lists are empty%s");
    // workspace: ET | sec | gmap in WS_ET, Lr in WS_MISC2, inv_q + queue heads in WS_MISC
    const size_t et_bytes = (size_t)tiles * gr->S * kT * sizeof(double);
    const size_t sec_bytes = ((size_t)tiles * Sp * sizeof(uint32_t) + 255) & ~(size_t)255;
    const size_t gmap_bytes = (size_t)tiles * kT * sizeof(int32_t);
    unsigned char* ws = (unsigned char*)sctda_ws(ctx, WS_ET, et_bytes + sec_bytes + gmap_bytes);
    if (!ws) return SCTDA_ERR_NOMEM;
    double* ET = (double*)ws;
    uint32_t* sec = (uint32_t*)(ws + et_bytes);
    int32_t* gmap = (int32_t*)(ws + et_bytes + sec_bytes);
    // composed indices: all permutations at once when they fit a bounded workspace, else in blocks
    size_t lr_cap = (size_t)4 << 30;
    if (const char* e = getenv("SCTDA_CONNW_LR_MB")) lr_cap = (size_t)atoll(e) << 20;
    int nblk = n;
    if ((size_t)n * Mtot * sizeof(uint16_t) > lr_cap) {
        nblk = (int)(lr_cap / ((size_t)Mtot * sizeof(uint16_t)));
        nblk = nblk / kChunk * kChunk;
        if (nblk < kChunk) nblk = kChunk;
    }
    if (nblk > 65528) nblk = 65528;                                // grid.y of the composing kernel
    const int nlaunch = (n + nblk - 1) / nblk;
    uint16_t* Lr = (uint16_t*)sctda_ws(ctx, WS_MISC2, (size_t)nblk * Mtot * sizeof(uint16_t) + 64);
    if (!Lr) return SCTDA_ERR_NOMEM;
    // WS_MISC: inv_q | flags | queue heads | row headers | packed adjacency
    const size_t off_flags = ((size_t)gr->V * sizeof(double) + 63) & ~(size_t)63;
    const size_t off_counters = off_flags + 64;
    const size_t off_rowinfo = (off_counters + (size_t)nlaunch * sizeof(unsigned long long) + 63) & ~(size_t)63;
    const size_t off_ids = (off_rowinfo + (size_t)gr->V * sizeof(uint2) + 255) & ~(size_t)255;
    size_t cap_chunks = 0;                                         // chunks of 8 ids the packed adjacency may take
    if (gr->V <= 65535) {
        cap_chunks = ((size_t)gr->V * ((size_t)gr->V + 1) / 2 + 7) / 8 + (size_t)gr->V;
        if (cap_chunks > ((size_t)8 << 20)) cap_chunks = (size_t)8 << 20;   // 128 MB of ids at most
    }
    unsigned char* misc = (unsigned char*)sctda_ws(ctx, WS_MISC, off_ids + cap_chunks * 16 + 64);
    if (!misc) return SCTDA_ERR_NOMEM;
    double* inv_q = (double*)misc;
    int* unit_w = (int*)(misc + off_flags);
    int* packed_ok = unit_w + 1;
    unsigned long long* counters = (unsigned long long*)(misc + off_counters);
    uint2* rowinfo = (uint2*)(misc + off_rowinfo);
    uint16_t* ids16 = (uint16_t*)(misc + off_ids);
    const int ones[2] = {1, 1};
    SCTDA_CUDA(ctx, cudaMemsetAsync(counters, 0, (size_t)nlaunch * sizeof(unsigned long long), st));
    SCTDA_CUDA(ctx, cudaMemcpyAsync(unit_w, ones, sizeof(ones), cudaMemcpyHostToDevice, st));
    unit_weight_wide_kernel<<<64, 256, 0, st>>>(gr->d_adj_off, gr->V, gr->d_adj_w, unit_w, (double*)(unit_w + 2));
    SCTDA_LAUNCH_CHECK(ctx, "unit_weight_wide_kernel");
    if (cap_chunks) {
        pack_adj_wide_kernel<<<1, 1024, 0, st>>>(gr->d_adj_off, gr->d_adj_idx, gr->V, (unsigned)cap_chunks, rowinfo, ids16, packed_ok);
        SCTDA_LAUNCH_CHECK(ctx, "pack_adj_wide_kernel");
    }
    gmap_wide_kernel<<<(tiles * kT + 255) / 256, 256, 0, st>>>(gmap, d_order, G, tiles * kT);
    SCTDA_LAUNCH_CHECK(ctx, "gmap_wide_kernel");
    if (Sp != gr->S) SCTDA_CUDA(ctx, cudaMemsetAsync(sec, 0, (size_t)tiles * Sp * sizeof(uint32_t), st));
    dim3 grid_p((gr->S + 31) / 32, tiles);
    et_prep_wide_kernel<<<grid_p, 256, 0, st>>>(d_Y, ldy, gr->S, Sp, gr->d_samples, gmap, log2_mode, ET, sec);
    SCTDA_LAUNCH_CHECK(ctx, "et_prep_wide_kernel");
    inv_size_wide_kernel<<<(gr->V + 255) / 256, 256, 0, st>>>(gr->d_node_off, gr->V, inv_q);
    SCTDA_LAUNCH_CHECK(ctx, "inv_size_wide_kernel");
    WideArgs a;
    a.V = gr->V; a.S = gr->S; a.Sp = Sp; a.G = G; a.n = n; a.log2_mode = log2_mode; a.Mtot = Mtot;
    a.tiles = tiles;
    a.node_off = gr->d_node_off;
    a.adj_off = gr->d_adj_off; a.adj_idx = gr->d_adj_idx; a.adj_w = gr->d_adj_w;
    a.ET = ET; a.sec = sec; a.Lr = Lr; a.gmap = gmap; a.inv_q = inv_q;
    a.observed = d_observed; a.tie_rel = tie_rel > 0.0 ? tie_rel : 1e-11;
    a.exceed = d_exceed; a.ties = d_ties; a.scores = d_scores; a.unit_w = unit_w; a.w0 = (const double*)(unit_w + 2);
    a.rowinfo = cap_chunks ? rowinfo : nullptr; a.ids16 = ids16; a.packed_ok = packed_ok;
    const int NW = wide_warps();
    const size_t smem = wide_smem(Sp, NW);
    long long grid_max = ctx->sm_count;                            // one CTA per SM
    if (const char* e = getenv("SCTDA_CONNW_GRID")) grid_max = atoi(e) > 0 ? atoi(e) : grid_max;
    a.pm = (float*)sctda_ws(ctx, WS_PM, (size_t)grid_max * gr->V * kT * sizeof(float));
    if (!a.pm) return SCTDA_ERR_NOMEM;
    for (int b = 0; b < nlaunch; ++b) {
        a.r0 = b * nblk;
        a.nloc = n - a.r0 < nblk ? n - a.r0 : nblk;
        a.nchunks = (a.nloc + kChunk - 1) / kChunk;
        a.items = (long long)tiles * a.nchunks;
        a.counter = counters + b;
        dim3 grid_c((unsigned)((Mtot + 255) / 256 < 64 ? (Mtot + 255) / 256 : 64), (unsigned)a.nloc);
        compose_kernel<<<grid_c, 256, 0, st>>>(d_perm, gr->S, a.r0, a.nloc, gr->d_node_cols, Mtot, Lr);
        SCTDA_LAUNCH_CHECK(ctx, "compose_kernel");
        const int grid = (int)(grid_max < a.items ? grid_max : a.items);
        SCTDA_TIME_BEGIN(ctx, st);
        switch (NW) {
            case 8: rc = launch_wide<8>(ctx, a, grid, smem, st); break;
            case 16: rc = launch_wide<16>(ctx, a, grid, smem, st); break;
            case 20: rc = launch_wide<20>(ctx, a, grid, smem, st); break;
            case 32: rc = launch_wide<32>(ctx, a, grid, smem, st); break;
            default: rc = launch_wide<24>(ctx, a, grid, smem, st); break;
        }
        SCTDA_TIME_END(ctx, st);
        if (rc) return rc;
        SCTDA_LAUNCH_CHECK(ctx, "conn_perm_wide_kernel");
    }
    return SCTDA_OK;
}
