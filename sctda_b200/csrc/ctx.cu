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
    for (int i = 0; i < 32; ++i)
        if (ctx->tev[i]) cudaEventDestroy(ctx->tev[i]);
    for (int i = 0; i < 4; ++i)
        if (ctx->pipe_ev[i]) cudaEventDestroy(ctx->pipe_ev[i]);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
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

extern "C" int32_t sctda_set_kernel_timing(sctda_ctx* ctx, int32_t enabled) {
    if (!ctx) return SCTDA_ERR_INVALID;
    if (enabled && !ctx->tev[0]) {
        SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
        for (int i = 0; i < 32; ++i) SCTDA_CUDA(ctx, cudaEventCreate(&ctx->tev[i]));
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
