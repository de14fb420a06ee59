// Shared host/device plumbing for libpetal_b200: error type, context, device buffers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/petal_b200.h"

namespace petal {

struct Error {
    int code;
    std::string msg;
};

#define PETAL_CUDA(expr)                                                                        \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            throw ::petal::Error{PETAL_LINALG_ERROR, std::string("CUDA error: ") +               \
                                                         cudaGetErrorString(_e) + " (" #expr    \
                                                         ") at " __FILE__ ":" +                 \
                                                         std::to_string(__LINE__)};             \
        }                                                                                       \
    } while (0)

inline void invalid_input(const std::string& m) { throw Error{PETAL_INVALID_INPUT, m}; }
inline void linalg_error(const std::string& m) { throw Error{PETAL_LINALG_ERROR, m}; }

struct Comm;  // comm.cuh

}  // namespace petal

// The opaque context of the C ABI.
struct petal_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    int sm_count = 148;
    int64_t launches = 0;
    int f32_engine = 1;  // 0 = SIMT FFMA, 1 = tcgen05 3xTF32 where supported
    int f64_engine = 1;  // 0 = SIMT DFMA, 1 = DMMA (mma.sync f64) for Gram-shaped contractions
    bool tc_precise = true;  // tcgen05 engine: cut the TMEM accumulation chains (fp32-accurate) unless a caller opts out
    std::string last_error;
    // optional per-kernel timing (CUDA events on the launch stream), see petal_ctx_profile_json
    bool profiling = false;
    struct ProfEntry {
        const char* name;
        cudaEvent_t a, b;
        double work;  // caller-defined work units (bytes) for this launch
    };
    std::vector<ProfEntry> prof;
    petal::Comm* comm = nullptr;
    int rank = 0;
    int world = 1;
    // one call at a time per context (the reference's `&mut self` on fit, src/pca.rs:116): entry points lock this
    std::mutex mu;
    // per-context (= per-device) launch state: opted-in dynamic shared memory per kernel, cooperative-launch limits
    std::vector<std::pair<const void*, size_t>> smem_optin;
    int coop_ok = -1;
    int coop_max_ctas = 0;
    // device-side status word raised by kernels that cannot report through the host (Jacobi sweeps exhausted)
    int* dev_status = nullptr;
    bool status_armed = false;  // a kernel that may raise dev_status was launched during this call
    // host-resident X (row_stream.cuh): 0 = auto (resident copy when it fits, else out-of-core ring), 1 = always a
    // resident copy, 2 = always the two-slot ring; rows of X per H2D chunk are derived from host_chunk_bytes
    int host_gram = 1;    // randomized PCA on a host-fed X: power iterations on the Gram matrix taken during the ingest
    int host_staging = 0;
    int64_t host_chunk_bytes = (int64_t)1 << 30;
    cudaStream_t copy_stream = nullptr;   // created on first use
    cudaEvent_t copy_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // ready[2], freed[2], alloc
    // statistics of the last call that streamed a host X (petal_ctx_host_stream_stats)
    int64_t last_h2d_bytes = 0;
    int64_t last_traversals = 0;
    int last_ring = 0;
};

namespace petal {

// 16-byte vector of T (float4 / double2 worth) that the compiler loads/stores as one LDG/STS.128.
template <typename T>
struct alignas(16) Pack {
    static constexpr int N = 16 / sizeof(T);
    T v[N];
};

inline bool is_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

inline bool is_device_pointer(const void* p) {
    if (p == nullptr) return false;
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

// Stream-ordered device buffer.
template <typename T>
struct DBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaStream_t s = nullptr;
    DBuf() {}
    DBuf(petal_ctx* ctx, size_t count) { alloc(ctx, count); }
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    DBuf(DBuf&& o) noexcept : p(o.p), n(o.n), s(o.s) { o.p = nullptr; }
    void alloc(petal_ctx* ctx, size_t count) {
        release();
        s = ctx->stream;
        n = count;
        if (count == 0) return;
        PETAL_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&p), count * sizeof(T), s));
    }
    void zero() {
        if (n) PETAL_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
    }
    void release() {
        if (p) cudaFreeAsync(p, s);
        p = nullptr;
    }
    ~DBuf() { release(); }
};

// Input that may live on the host: staged into HBM when needed.
template <typename T>
struct DevIn {
    const T* p = nullptr;
    DBuf<T> staged;
    size_t bytes_h2d = 0;
    DevIn() {}
    DevIn(petal_ctx* ctx, const T* user, size_t count) { set(ctx, user, count); }
    void set(petal_ctx* ctx, const T* user, size_t count) {
        if (user == nullptr || count == 0) {
            p = user;
            return;
        }
        if (is_device_pointer(user)) {
            p = user;
        } else {
            staged.alloc(ctx, count);
            PETAL_CUDA(cudaMemcpyAsync(staged.p, user, count * sizeof(T), cudaMemcpyHostToDevice,
                                       ctx->stream));
            p = staged.p;
            bytes_h2d = count * sizeof(T);
        }
    }
};

// Output that may live on the host: written to HBM, copied back by commit().
template <typename T>
struct DevOut {
    T* p = nullptr;  // device pointer to write
    T* user = nullptr;
    size_t n = 0;
    bool to_host = false;
    DBuf<T> staged;
    DevOut() {}
    DevOut(petal_ctx* ctx, T* user_ptr, size_t count) { set(ctx, user_ptr, count); }
    void set(petal_ctx* ctx, T* user_ptr, size_t count) {
        user = user_ptr;
        n = count;
        if (user == nullptr || count == 0) {
            p = nullptr;
            return;
        }
        if (is_device_pointer(user)) {
            p = user;
        } else {
            staged.alloc(ctx, count);
            p = staged.p;
            to_host = true;
        }
    }
    explicit operator bool() const { return p != nullptr; }
    void commit(petal_ctx* ctx) {
        if (to_host && n)
            PETAL_CUDA(cudaMemcpyAsync(user, p, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    }
};

// RAII event pair around one kernel launch (or a launch sequence) when profiling is on.
struct KTimer {
    petal_ctx* ctx;
    int idx = -1;
    KTimer(petal_ctx* c, const char* name, double work = 0.0) : ctx(c) {
        if (!c->profiling) return;
        petal_ctx::ProfEntry e;
        e.name = name;
        e.work = work;
        if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess) return;
        cudaEventRecord(e.a, c->stream);
        c->prof.push_back(e);
        idx = (int)c->prof.size() - 1;
    }
    ~KTimer() {
        if (idx >= 0) cudaEventRecord(ctx->prof[idx].b, ctx->stream);
    }
};

template <typename T>
inline const char* kname(const char* f32, const char* f64) {
    return sizeof(T) == 4 ? f32 : f64;
}

inline void check_launch(petal_ctx* ctx) {
    ctx->launches++;
    PETAL_CUDA(cudaGetLastError());
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: remembered per context, never in a
// process-wide static (several contexts / devices per process are allowed).
template <typename K>
inline void ensure_dynamic_smem(petal_ctx* ctx, K kernel, size_t bytes) {
    const void* key = reinterpret_cast<const void*>(kernel);
    for (auto& e : ctx->smem_optin)
        if (e.first == key) {
            if (e.second >= bytes) return;
            PETAL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            e.second = bytes;
            return;
        }
    PETAL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    ctx->smem_optin.emplace_back(key, bytes);
}

// status bits of petal_ctx::dev_status
constexpr int kStatusJacobiNotConverged = 1;

}  // namespace petal
