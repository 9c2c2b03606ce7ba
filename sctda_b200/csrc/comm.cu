// NCCL glue of the C-ABI: the two exchanges of the multi-GPU paths, on DEVICE buffers.
//
//   sctda_allgather_csr     the one exchange of the sharded Mapper build: every rank has clustered its own
//                           patches (sctda_bin_cluster on the sub-CSR sctda_csr_select cut out); the label
//                           fragments and the per-patch K meet on every rank through ncclAllGather and are
//                           placed at their positions of the full cover CSR by a kernel;
//   sctda_gather_results    the final gather of the permutation test: the per-gene counts of all ranks.
//
// The communicator is created by the HOST (one rank per GPU; sctda_b200/parallel.py does it with ctypes on
// the process's own libnccl.so.2, torch.distributed only carries the 128-byte unique id); this library never
// initialises NCCL.  NCCL is resolved at run time (dlopen of libnccl.so.2: the copy already loaded in the
// process, if any), so the library itself links against nothing but the CUDA runtime.
#include <dlfcn.h>

#include "common.cuh"

namespace {

typedef int (*nccl_allgather_fn)(const void*, void*, size_t, int /*ncclDataType_t*/, void* /*ncclComm_t*/, cudaStream_t);
typedef const char* (*nccl_errstr_fn)(int);
constexpr int kNcclInt32 = 2;     // ncclInt32 in nccl.h (stable across NCCL 2.x)

nccl_allgather_fn g_allgather = nullptr;
nccl_errstr_fn g_errstr = nullptr;

int resolve_nccl(sctda_ctx* ctx) {
    if (g_allgather) return SCTDA_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return sctda_fail(ctx, SCTDA_ERR_UNSUPPORTED, "libnccl.so.2 cannot be loaded: %s", dlerror());
    g_allgather = (nccl_allgather_fn)dlsym(h, "ncclAllGather");
    g_errstr = (nccl_errstr_fn)dlsym(h, "ncclGetErrorString");
    if (!g_allgather) return sctda_fail(ctx, SCTDA_ERR_UNSUPPORTED, "ncclAllGather not found in libnccl.so.2%s");
    return SCTDA_OK;
}

int nccl_check(sctda_ctx* ctx, int rc, const char* what) {
    if (rc == 0) return SCTDA_OK;
    snprintf(ctx->err, sizeof(ctx->err), "%s failed: %s", what, g_errstr ? g_errstr(rc) : "NCCL error");
    return SCTDA_ERR_CUDA;
}

// sub_members[sub_off[i] + t] = members[bin_off[sel[i]] + t]; one CTA per selected patch
__global__ void csr_select_kernel(const int32_t* __restrict__ sel, const int32_t* __restrict__ bin_off,
                                  const int32_t* __restrict__ members, const int32_t* __restrict__ sub_off,
                                  int32_t* __restrict__ sub_members) {
    const int i = blockIdx.x;
    const int b = sel[i];
    const int src = bin_off[b], m = bin_off[b + 1] - src, dst = sub_off[i];
    for (int t = threadIdx.x; t < m; t += blockDim.x) sub_members[dst + t] = members[src + t];
}

// labels[bin_off[b] + t] = gathered[owner[b]][frag_off[b] + t], bin_K[b] = gatheredK[owner[b]][frag_idx[b]]
__global__ void place_fragments_kernel(const int32_t* __restrict__ gathered, int64_t max_frag,
                                       const int32_t* __restrict__ gatheredK, int max_nb,
                                       const int32_t* __restrict__ bin_off, const int32_t* __restrict__ owner,
                                       const int32_t* __restrict__ frag_off, const int32_t* __restrict__ frag_idx,
                                       int32_t* __restrict__ labels, int32_t* __restrict__ bin_K) {
    const int b = blockIdx.x;
    const int dst = bin_off[b], m = bin_off[b + 1] - dst;
    const int r = owner[b];
    const int32_t* src = gathered + (int64_t)r * max_frag + frag_off[b];
    for (int t = threadIdx.x; t < m; t += blockDim.x) labels[dst + t] = src[t];
    if (threadIdx.x == 0) bin_K[b] = gatheredK[(int64_t)r * max_nb + frag_idx[b]];
}

__global__ void pad_copy_kernel(const int32_t* __restrict__ src, int64_t n, int32_t* __restrict__ dst, int64_t n_padded) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_padded) dst[i] = i < n ? src[i] : 0;
}

}  // namespace

extern "C" int32_t sctda_csr_select(sctda_ctx* ctx, int32_t nsel, const int32_t* d_sel, const int32_t* d_bin_off,
                                    const int32_t* d_members, const int32_t* d_sub_off, int32_t* d_sub_members,
                                    void* stream) {
    if (!ctx) return SCTDA_ERR_INVALID;
    if (nsel < 0 || (nsel > 0 && (!d_sel || !d_bin_off || !d_members || !d_sub_off || !d_sub_members)))
        return sctda_fail(ctx, SCTDA_ERR_INVALID, "csr_select: bad arguments%s");
    if (nsel == 0) return SCTDA_OK;
    SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
    csr_select_kernel<<<nsel, 256, 0, (cudaStream_t)stream>>>(d_sel, d_bin_off, d_members, d_sub_off, d_sub_members);
    SCTDA_LAUNCH_CHECK(ctx, "csr_select_kernel");
    return SCTDA_OK;
}

extern "C" int32_t sctda_allgather_csr(sctda_ctx* ctx, void* nccl_comm, int32_t nranks, const int32_t* d_frag_labels,
                                       int64_t frag_len, int64_t max_frag, const int32_t* d_frag_K, int32_t frag_nb,
                                       int32_t max_nb, int32_t nbins, const int32_t* d_bin_off, const int32_t* d_owner,
                                       const int32_t* d_frag_off, const int32_t* d_frag_idx, int32_t* d_labels,
                                       int32_t* d_bin_K, void* stream) {
    if (!ctx) return SCTDA_ERR_INVALID;
    if ((!nccl_comm && nranks != 1) || nranks < 1 || frag_len < 0 || max_frag < frag_len || frag_nb < 0 || max_nb < frag_nb ||
        nbins < 0 || !d_bin_off || !d_owner || !d_frag_off || !d_frag_idx || !d_labels || !d_bin_K)
        return sctda_fail(ctx, SCTDA_ERR_INVALID, "allgather_csr: bad arguments (a communicator is needed for more than one rank)%s");
    int rc = nccl_comm ? resolve_nccl(ctx) : SCTDA_OK;
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t mf = max_frag > 0 ? max_frag : 1, mk = max_nb > 0 ? max_nb : 1;
    // workspace: [send labels mf | send K mk | recv labels nranks*mf | recv K nranks*mk]
    int32_t* ws = (int32_t*)sctda_ws(ctx, WS_NODES, (size_t)((nranks + 1) * (mf + mk)) * sizeof(int32_t));
    if (!ws) return SCTDA_ERR_NOMEM;
    int32_t* send_l = ws;
    int32_t* send_k = send_l + mf;
    int32_t* recv_l = send_k + mk;
    int32_t* recv_k = recv_l + (int64_t)nranks * mf;
    pad_copy_kernel<<<(unsigned)cdiv64(mf, 256), 256, 0, st>>>(d_frag_labels, frag_len, send_l, mf);
    pad_copy_kernel<<<(unsigned)cdiv64(mk, 256), 256, 0, st>>>(d_frag_K, frag_nb, send_k, mk);
    SCTDA_LAUNCH_CHECK(ctx, "pad_copy_kernel");
    if (nccl_comm) {
        rc = nccl_check(ctx, g_allgather(send_l, recv_l, (size_t)mf, kNcclInt32, nccl_comm, st), "ncclAllGather (labels)");
        if (rc) return rc;
        rc = nccl_check(ctx, g_allgather(send_k, recv_k, (size_t)mk, kNcclInt32, nccl_comm, st), "ncclAllGather (K)");
        if (rc) return rc;
    } else {                                                       // one rank, no communicator: the gather is a copy
        SCTDA_CUDA(ctx, cudaMemcpyAsync(recv_l, send_l, (size_t)mf * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        SCTDA_CUDA(ctx, cudaMemcpyAsync(recv_k, send_k, (size_t)mk * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    }
    if (nbins > 0) {
        place_fragments_kernel<<<nbins, 256, 0, st>>>(recv_l, mf, recv_k, (int)mk, d_bin_off, d_owner, d_frag_off, d_frag_idx,
                                                      d_labels, d_bin_K);
        SCTDA_LAUNCH_CHECK(ctx, "place_fragments_kernel");
    }
    return SCTDA_OK;
}

extern "C" int32_t sctda_gather_results(sctda_ctx* ctx, void* nccl_comm, int32_t nranks, const int32_t* d_mine,
                                        int64_t n_mine, int64_t max_n, int32_t* d_all, void* stream) {
    if (!ctx) return SCTDA_ERR_INVALID;
    if (!nccl_comm || nranks < 1 || n_mine < 0 || max_n < n_mine || max_n < 1 || !d_all || (n_mine > 0 && !d_mine))
        return sctda_fail(ctx, SCTDA_ERR_INVALID, "gather_results: bad arguments%s");
    int rc = resolve_nccl(ctx);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    SCTDA_CUDA(ctx, cudaSetDevice(ctx->device));
    int32_t* send = (int32_t*)sctda_ws(ctx, WS_NODES, (size_t)max_n * sizeof(int32_t));
    if (!send) return SCTDA_ERR_NOMEM;
    pad_copy_kernel<<<(unsigned)cdiv64(max_n, 256), 256, 0, st>>>(d_mine, n_mine, send, max_n);
    SCTDA_LAUNCH_CHECK(ctx, "pad_copy_kernel");
    return nccl_check(ctx, g_allgather(send, d_all, (size_t)max_n, kNcclInt32, nccl_comm, st), "ncclAllGather (counts)");
}
