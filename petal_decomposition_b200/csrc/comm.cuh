// Row-sharded multi-GPU support: one process per GPU, NCCL all-reduce / all-gather of the small
// replicated quantities (column sums, Gram matrices, X^T*Q partials, FastICA k x k sums) over
// NVLink 5 / NVSwitch.  The reference is single-host; this is the one
// data-parallel axis the path has: rows (samples).
//
// NCCL is loaded lazily with dlopen so that single-GPU use never touches it; inside a torch
// process the already-loaded torch-bundled libnccl.so.2 is reused (same soname).
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace petal {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

inline NcclApi& nccl_api() {
    static NcclApi api;
    if (api.handle) return api;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) linalg_error(std::string("cannot load NCCL: ") + dlerror());
    auto sym = [&](const char* name) {
        void* s = dlsym(h, name);
        if (!s) linalg_error(std::string("NCCL symbol missing: ") + name);
        return s;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.handle = h;
    return api;
}

#define PETAL_NCCL(expr)                                                                         \
    do {                                                                                         \
        ncclResult_t _r = (expr);                                                                \
        if (_r != ncclSuccess)                                                                   \
            ::petal::linalg_error(std::string("NCCL error: ") + ::petal::nccl_api().GetErrorString(_r) + \
                                  " (" #expr ")");                                               \
    } while (0)

struct Comm {
    ncclComm_t comm = nullptr;
    ~Comm() {
        if (comm) nccl_api().CommDestroy(comm);
    }
};

static_assert(sizeof(ncclUniqueId) == PETAL_COMM_ID_BYTES, "NCCL unique id size");

// In-place sum over ranks of a device f64 buffer; no-op on a single rank.
inline void allreduce_sum(petal_ctx* ctx, double* buf, size_t count) {
    if (ctx->world <= 1 || count == 0) return;
    PETAL_NCCL(nccl_api().AllReduce(buf, buf, count, ncclFloat64, ncclSum, ctx->comm->comm, ctx->stream));
}

// dst <- sum over ranks of src (src is left untouched, so repeating the call with an unchanged src is idempotent);
// a plain device copy on a single rank.
inline void allreduce_sum_to(petal_ctx* ctx, const double* src, double* dst, size_t count) {
    if (count == 0) return;
    if (ctx->world <= 1) {
        if (src != dst) PETAL_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        return;
    }
    PETAL_NCCL(nccl_api().AllReduce(src, dst, count, ncclFloat64, ncclSum, ctx->comm->comm, ctx->stream));
}

// recv[world * count] <- concatenation over ranks of send[count].
inline void allgather(petal_ctx* ctx, const double* send, double* recv, size_t count) {
    if (ctx->world <= 1) {
        PETAL_CUDA(cudaMemcpyAsync(recv, send, count * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        return;
    }
    PETAL_NCCL(nccl_api().AllGather(send, recv, count, ncclFloat64, ctx->comm->comm, ctx->stream));
}

}  // namespace petal
