// Multi-GPU inside ONE process: the collective of the hot path — the three MPI_Allreduce calls of mcrg.cpp:101-103 — as one
// ncclAllReduce(ncclInt64, ncclSum) per device over the accumulator totals written as 32-bit limbs (exact and order
// independent, see k_total_limbs).  NCCL is loaded with dlopen at the first mcrg_comm_init_all, so the library itself has
// no link-time dependency on it: single-GPU users never touch NCCL.  (bench.py runs one process per GPU and lets
// torch.distributed own the communicator; this entry point is for C/C++ hosts such as the drop-in layer.)
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only

#include <vector>

#include "capi_internal.cuh"

using namespace mcrg;

namespace {

struct NcclApi {
    void *handle = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

NcclApi g_nccl;

const char *load_nccl() {
    if (g_nccl.handle) return nullptr;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return dlerror();
#define MCRG_SYM(field, name)                                              \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name)); \
    if (!g_nccl.field) return "symbol " name " missing from libnccl";
    MCRG_SYM(CommInitAll, "ncclCommInitAll")
    MCRG_SYM(CommDestroy, "ncclCommDestroy")
    MCRG_SYM(AllReduce, "ncclAllReduce")
    MCRG_SYM(GroupStart, "ncclGroupStart")
    MCRG_SYM(GroupEnd, "ncclGroupEnd")
    MCRG_SYM(GetErrorString, "ncclGetErrorString")
#undef MCRG_SYM
    g_nccl.handle = h;
    return nullptr;
}

#define NK(call)                                                                                                   \
    do {                                                                                                           \
        ncclResult_t r_ = (call);                                                                                  \
        if (r_ != ncclSuccess) return mcrg_fail(MCRG_ERR_CUDA, "%s failed: %s", #call, g_nccl.GetErrorString(r_)); \
    } while (0)

#define CKC(call)                                                                                                  \
    do {                                                                                                           \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess) return mcrg_fail(MCRG_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));    \
    } while (0)

}  // namespace

extern "C" {

int mcrg_comm_init_all(int n, mcrg_ctx **ctxs) {
    if (n < 1 || !ctxs) return mcrg_fail(MCRG_ERR_ARG, "need n >= 1 contexts");
    std::vector<int> devs(n);
    for (int i = 0; i < n; ++i) {
        if (!ctxs[i]) return mcrg_fail(MCRG_ERR_ARG, "context %d is null", i);
        if (mcrg_ctx_comm(ctxs[i])) return mcrg_fail(MCRG_ERR_STATE, "context %d already has a communicator", i);
        devs[i] = mcrg_ctx_device(ctxs[i]);
        for (int j = 0; j < i; ++j)
            if (devs[j] == devs[i]) return mcrg_fail(MCRG_ERR_ARG, "contexts %d and %d share device %d: one context per device", j, i, devs[i]);
    }
    if (const char *err = load_nccl()) return mcrg_fail(MCRG_ERR_CUDA, "cannot load NCCL: %s", err);
    std::vector<ncclComm_t> comms(n);
    NK(g_nccl.CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; ++i) {
        CKC(cudaSetDevice(devs[i]));
        void *limbs = nullptr;
        CKC(cudaMalloc(&limbs, (size_t)4 * N_SLOTS * sizeof(long long)));
        mcrg_ctx_set_comm(ctxs[i], comms[i], limbs);
    }
    return 0;
}

int mcrg_comm_destroy_all(int n, mcrg_ctx **ctxs) {
    if (n < 1 || !ctxs) return mcrg_fail(MCRG_ERR_ARG, "need n >= 1 contexts");
    for (int i = 0; i < n; ++i) {
        if (!ctxs[i] || !mcrg_ctx_comm(ctxs[i])) continue;
        cudaSetDevice(mcrg_ctx_device(ctxs[i]));
        cudaStreamSynchronize(mcrg_ctx_stream(ctxs[i]));
        g_nccl.CommDestroy(static_cast<ncclComm_t>(mcrg_ctx_comm(ctxs[i])));
        cudaFree(mcrg_ctx_limbs(ctxs[i]));
        mcrg_ctx_set_comm(ctxs[i], nullptr, nullptr);
    }
    return 0;
}

int mcrg_allreduce_accumulators(int n, mcrg_ctx **ctxs, int64_t *hi, uint64_t *lo) {
    if (n < 1 || !ctxs) return mcrg_fail(MCRG_ERR_ARG, "need n >= 1 contexts");
    for (int i = 0; i < n; ++i)
        if (!ctxs[i] || !mcrg_ctx_comm(ctxs[i])) return mcrg_fail(MCRG_ERR_STATE, "context %d has no communicator: call mcrg_comm_init_all", i);
    // every device: totals of its own replicas and bins as limbs, on its own stream, behind the work already enqueued
    for (int i = 0; i < n; ++i) {
        int rc = mcrg_accumulators_total_limbs_device(ctxs[i], mcrg_ctx_limbs(ctxs[i]));
        if (rc) return rc;
    }
    NK(g_nccl.GroupStart());
    for (int i = 0; i < n; ++i)
        NK(g_nccl.AllReduce(mcrg_ctx_limbs(ctxs[i]), mcrg_ctx_limbs(ctxs[i]), (size_t)4 * N_SLOTS, ncclInt64, ncclSum,
                            static_cast<ncclComm_t>(mcrg_ctx_comm(ctxs[i])), mcrg_ctx_stream(ctxs[i])));
    NK(g_nccl.GroupEnd());
    std::vector<long long> limbs((size_t)4 * N_SLOTS);
    for (int i = 0; i < n; ++i) {
        CKC(cudaSetDevice(mcrg_ctx_device(ctxs[i])));
        if (i == 0) CKC(cudaMemcpyAsync(limbs.data(), mcrg_ctx_limbs(ctxs[0]), limbs.size() * sizeof(long long), cudaMemcpyDeviceToHost, mcrg_ctx_stream(ctxs[0])));
        CKC(cudaStreamSynchronize(mcrg_ctx_stream(ctxs[i])));
    }
    for (int s = 0; s < N_SLOTS; ++s) {  // limbs -> one exact 128-bit integer per slot (every limb may have carried)
        __int128 v = 0;
        for (int k = 3; k >= 0; --k) v = (v << 32) + (__int128)limbs[(size_t)4 * s + k];
        if (hi) hi[s] = (int64_t)(v >> 64);
        if (lo) lo[s] = (uint64_t)v;
    }
    return 0;
}

}  // extern "C"
