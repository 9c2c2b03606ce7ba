// Context management of the C-ABI (include/sctda_b200.h).
#include <stdlib.h>

#include "common.cuh"

extern "C" int32_t sctda_abi_version(void) { return SCTDA_ABI_VERSION; }

extern "C" int32_t sctda_ctx_create(int32_t device, sctda_ctx** out_ctx) {
    if (!out_ctx) return SCTDA_ERR_INVALID;
    *out_ctx = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        cudaGetLastError();
        return SCTDA_ERR_CUDA;
    }
    if (cudaSetDevice(device) != cudaSuccess) return SCTDA_ERR_CUDA;
    sctda_ctx* ctx = (sctda_ctx*)calloc(1, sizeof(sctda_ctx));
    if (!ctx) return SCTDA_ERR_NOMEM;
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        free(ctx);
        return SCTDA_ERR_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) {
        // sm_100a only: there is no other code path in this library.
        free(ctx);
        return SCTDA_ERR_UNSUPPORTED;
    }
    *out_ctx = ctx;
    return SCTDA_OK;
}

extern "C" int32_t sctda_ctx_destroy(sctda_ctx* ctx) {
    if (!ctx) return SCTDA_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < SCTDA_NUM_SLOTS; ++i)
        if (ctx->slot_ptr[i]) cudaFree(ctx->slot_ptr[i]);
    for (int i = 0; i < 128; ++i)
        if (ctx->tev[i]) cudaEventDestroy(ctx->tev[i]);
    for (int i = 0; i < 4; ++i)
        if (ctx->pipe_ev[i]) cudaEventDestroy(ctx->pipe_ev[i]);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->fork_stream) {
        cudaStreamDestroy(ctx->fork_stream);
        for (int i = 0; i < 2; ++i) cudaEventDestroy(ctx->fork_ev[i]);
    }
    free(ctx);
    return SCTDA_OK;
}

extern "C" const char* sctda_last_error(const sctda_ctx* ctx) { return ctx ? ctx->err : "null context"; }

extern "C" int64_t sctda_workspace_bytes(const sctda_ctx* ctx) {
    if (!ctx) return 0;
    int64_t t = 0;
    for (int i = 0; i < SCTDA_NUM_SLOTS; ++i) t += (int64_t)ctx->slot_bytes[i];
    return t;
}

extern "C" int64_t sctda_launch_count(const sctda_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int32_t sctda_set_precision(sctda_ctx* ctx, int32_t mode) {
    if (!ctx) return SCTDA_ERR_INVALID;
    if (mode != 0 && mode != 1) return sctda_fail(ctx, SCTDA_ERR_UNSUPPORTED, "set_precision: 0 (FP64) or 1 (3xTF32)%s");
    ctx->precision = mode;
    return SCTDA_OK;
}

extern "C" int32_t sctda_get_precision(const sctda_ctx* ctx) { return ctx ? ctx->precision : SCTDA_ERR_INVALID; }

extern "C" int32_t sctda_set_linkage(sctda_ctx* ctx, int32_t mode) {
    if (!ctx) return SCTDA_ERR_INVALID;
    if (mode < 0 || mode > 3) return sctda_fail(ctx, SCTDA_ERR_UNSUPPORTED, "set_linkage: 0 single, 1 average, 2 ward, 3 complete%s");
    ctx->linkage = mode;
    return SCTDA_OK;
}

extern "C" int32_t sctda_get_linkage(const sctda_ctx* ctx) { return ctx ? ctx->linkage : SCTDA_ERR_INVALID; }

extern "C" int32_t sctda_set_kernel_timing(sctda_ctx* ctx, int32_t enabled) {
    if (!ctx) return SCTDA_ERR_INVALID;
    if (enabled && !ctx->tev[0]) {
        SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
        for (int i = 0; i < 128; ++i) SCTDA_CUDA(ctx, cudaEventCreate(&ctx->tev[i]));
    }
    ctx->timing = enabled ? 1 : 0;
    return SCTDA_OK;
}

extern "C" int32_t sctda_last_kernel_ms(sctda_ctx* ctx, float* out_ms) {
    if (!ctx || !out_ms) return SCTDA_ERR_INVALID;
    if (!ctx->timed) return sctda_fail(ctx, SCTDA_ERR_INVALID, "no timed kernel yet (sctda_set_kernel_timing)%s");
    float total = 0.f;
    for (int i = 0; i < ctx->timed; ++i) {
        float ms = 0.f;
        SCTDA_CUDA(ctx, cudaEventSynchronize(ctx->tev[2 * i + 1]));
        SCTDA_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->tev[2 * i], ctx->tev[2 * i + 1]));
        total += ms;
    }
    *out_ms = total;
    return SCTDA_OK;
}

// ---- yard-stick: L2 -> SM read bandwidth --------------------------------------------------------
// Every CTA streams the same L2-resident buffer with 16-byte loads (no reuse in L1: ld.global.cg).
__global__ void __launch_bounds__(1024) l2_probe_kernel(const uint4* __restrict__ buf, size_t n16, int iters,
                                                        unsigned* __restrict__ sink) {
    unsigned acc = 0;
    for (int it = 0; it < iters; ++it) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x * 4) {
            uint4 v0 = __ldcg(buf + i);
            size_t i1 = i + (size_t)gridDim.x * blockDim.x, i2 = i1 + (size_t)gridDim.x * blockDim.x,
                   i3 = i2 + (size_t)gridDim.x * blockDim.x;
            uint4 v1 = i1 < n16 ? __ldcg(buf + i1) : make_uint4(0, 0, 0, 0);
            uint4 v2 = i2 < n16 ? __ldcg(buf + i2) : make_uint4(0, 0, 0, 0);
            uint4 v3 = i3 < n16 ? __ldcg(buf + i3) : make_uint4(0, 0, 0, 0);
            acc ^= v0.x ^ v1.y ^ v2.z ^ v3.w;
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

extern "C" int32_t sctda_probe_l2_read_gbs(sctda_ctx* ctx, double* out_gbs) {
    if (!ctx || !out_gbs) return SCTDA_ERR_INVALID;
    SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)32 << 20;                       // resident in the 126 MB L2
    unsigned char* buf = (unsigned char*)sctda_ws(ctx, WS_MISC, bytes + 64);
    if (!buf) return SCTDA_ERR_NOMEM;
    SCTDA_CUDA(ctx, cudaMemset(buf, 1, bytes + 64));
    cudaEvent_t e0, e1;
    SCTDA_CUDA(ctx, cudaEventCreate(&e0));
    SCTDA_CUDA(ctx, cudaEventCreate(&e1));
    const int iters = 40;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        SCTDA_CUDA(ctx, cudaEventRecord(e0));
        l2_probe_kernel<<<ctx->sm_count * 2, 1024>>>((const uint4*)buf, bytes / 16, iters, (unsigned*)(buf + bytes));
        SCTDA_CUDA(ctx, cudaEventRecord(e1));
        SCTDA_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0.f;
        SCTDA_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        const double gbs = (double)bytes * iters / (ms * 1e-3) / 1e9;
        if (rep > 0 && gbs > best) best = gbs;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    SCTDA_LAUNCH_CHECK(ctx, "l2_probe_kernel");
    *out_gbs = best;
    return SCTDA_OK;
}
