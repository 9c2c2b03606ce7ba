// Shared internals of libsctda_b200.so: the context, the workspace and the error convention
// of include/sctda_b200.h.  Nothing here is visible through the C-ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/sctda_b200.h"

#define SCTDA_NUM_SLOTS 16

struct sctda_ctx {
    int device;
    int sm_count;
    char err[512];
    void* slot_ptr[SCTDA_NUM_SLOTS];     // grow-only device workspace, one buffer per slot
    size_t slot_bytes[SCTDA_NUM_SLOTS];
    int64_t launches;
    int timing;                          // sctda_set_kernel_timing
    int timed;                           // number of (begin, end) event pairs recorded by the last call
    cudaEvent_t tev[2 * 64];             // up to 64 launches of the dominant kernel per call
    unsigned attr_done;                  // per-context (= per-device) one-time kernel attribute setup, one bit per file
    cudaStream_t aux_stream;             // second stream for the patch pipeline (cluster.cu)
    cudaEvent_t pipe_ev[4];
    cudaStream_t fork_stream;            // MST size classes run side by side (cluster.cu)
    cudaEvent_t fork_ev[2];
    int precision;                       // sctda_set_precision: 0 = FP64 (bit-exact, default), 1 = 3xTF32 contraction
    int linkage;                         // sctda_set_linkage: 0 single (default), 1 average, 2 ward, 3 complete
};

// Workspace slots (one live use per call; calls on one context are serialised by contract).
enum {
    WS_ET = 0,        // transformed, transposed expression tiles
    WS_PM = 1,        // per-CTA node value scratch
    WS_MISC = 2,
    WS_COVER = 4,
    WS_CLUSTER = 5,
    WS_NERVE = 6,
    WS_CLUSTER2 = 7,
    WS_CLUSTER3 = 8,
    WS_NODES = 9,
    WS_CLUSTER4 = 10,   // second Gram buffer of the patch pipeline
    WS_CLUSTER5 = 11,   // second batch-metadata buffer
    WS_MISC2 = 13,
    WS_SPLIT = 12,      // (hi, lo) TF32 operands of the opt-in 3xTF32 contraction
    WS_LINK = 14,       // working distance matrices of the non-single linkages (shared with the PCA lens' Gram matrix:
                        // the two never run inside one call)
};

static inline int sctda_fail(sctda_ctx* ctx, int code, const char* fmt, const char* a = "", long long b = 0) {
    if (ctx) snprintf(ctx->err, sizeof(ctx->err), fmt, a, b);
    return code;
}

#define SCTDA_CUDA(ctx, call)                                                              \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s failed at %s:%d: %s", #call,      \
                     __FILE__, __LINE__, cudaGetErrorString(e_));                          \
            return SCTDA_ERR_CUDA;                                                         \
        }                                                                                  \
    } while (0)

#define SCTDA_LAUNCH_CHECK(ctx, name)                                                      \
    do {                                                                                   \
        (ctx)->launches++;                                                                 \
        cudaError_t e_ = cudaGetLastError();                                               \
        if (e_ != cudaSuccess) {                                                           \
            snprintf((ctx)->err, sizeof((ctx)->err), "launch of %s failed: %s", name,      \
                     cudaGetErrorString(e_));                                              \
            return SCTDA_ERR_CUDA;                                                         \
        }                                                                                  \
    } while (0)

// Returns a device buffer of at least `bytes` in `slot` (contents undefined), or NULL.
static inline void* sctda_ws(sctda_ctx* ctx, int slot, size_t bytes) {
    if (bytes == 0) bytes = 256;
    if (ctx->slot_bytes[slot] >= bytes) return ctx->slot_ptr[slot];
    if (ctx->slot_ptr[slot]) {
        cudaDeviceSynchronize();          // earlier stream work may still read the old buffer
        cudaFree(ctx->slot_ptr[slot]);
        ctx->slot_ptr[slot] = nullptr;
        ctx->slot_bytes[slot] = 0;
    }
    size_t want = bytes + bytes / 8;      // slack so that slowly growing requests do not thrash
    void* p = nullptr;
    if (cudaMalloc(&p, want) != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            cudaGetLastError();
            snprintf(ctx->err, sizeof(ctx->err), "workspace slot %d: cudaMalloc of %zu bytes failed", slot, bytes);
            return nullptr;
        }
    }
    ctx->slot_ptr[slot] = p;
    ctx->slot_bytes[slot] = want;
    return p;
}

static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// conn_tile64.cu: the shared-permutation kernel on 64-gene tiles (internal entry, see conn.cu)
bool sctda_tile64_eligible(const sctda_graph* gr);
int sctda_conn_perm_tile64(sctda_ctx* ctx, const sctda_graph* gr, int G, const double* d_Y, int64_t ldy, int log2_mode,
                           int n, const int32_t* d_perm, const int32_t* d_order, const double* d_observed,
                           double tie_rel, int32_t* d_exceed, int32_t* d_ties, double* d_scores, cudaStream_t st);

// conn_wide.cu: the shared-permutation kernel on 128-gene tiles with per-permutation composed member indices
bool sctda_wide_eligible(const sctda_graph* gr);
int sctda_conn_perm_wide(sctda_ctx* ctx, const sctda_graph* gr, int G, const double* d_Y, int64_t ldy, int log2_mode,
                         int n, const int32_t* d_perm, const int32_t* d_order, const double* d_observed,
                         double tie_rel, int32_t* d_exceed, int32_t* d_ties, double* d_scores, cudaStream_t st);

// Brackets the dominant kernel of a call with events when timing is on (see sctda_set_kernel_timing).
#define SCTDA_TIME_RESET(ctx) do { if ((ctx)->timing) (ctx)->timed = 0; } while (0)
#define SCTDA_TIME_BEGIN(ctx, st) do { if ((ctx)->timing && (ctx)->timed < 64) cudaEventRecord((ctx)->tev[2 * (ctx)->timed], st); } while (0)
#define SCTDA_TIME_END(ctx, st) do { if ((ctx)->timing && (ctx)->timed < 64) { cudaEventRecord((ctx)->tev[2 * (ctx)->timed + 1], st); (ctx)->timed++; } } while (0)
