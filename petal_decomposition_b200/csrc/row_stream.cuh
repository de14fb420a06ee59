// Rows of the data matrix X as the streaming passes of a fit see them (out-of-core ingest).
//
// The reference keeps X (and one to three centred copies of it, src/pca.rs:217,531, src/ica.rs:178-189) in host
// memory; here the fit flows only ever ask for "every row of X once, in row chunks" (`traverse`), so the same flow
// serves three situations:
//   device   : X already lives in HBM (device pointer): one chunk = all rows, nothing is copied.
//   resident : X is a host buffer that fits in HBM: the FIRST traversal copies it chunk by chunk into its final place
//              on a dedicated copy stream and hands each chunk to the compute stream as soon as it has landed, so
//              the first streaming pass(es) of the fit run underneath the PCIe transfer; later traversals see one
//              resident chunk.
//   ring     : X is a host buffer larger than the HBM that is left (or the caller forces it): every traversal
//              re-streams X through a two-slot ring (double-buffered H2D from the caller's - ideally pinned - memory;
//              slot reuse is ordered with events, the host never blocks).  The flows are arranged so that both
//              contractions of a range-finder iteration (X_c B and X_c^T Y_c) run on a chunk while it is resident:
//              q power iterations cost q + 1 trips over PCIe, not 2q + 2.
// Chunks start at multiples of 32 rows (panel-major Y blocks) and never have fewer than 1024 rows unless X has
// (the tcgen05 kernels' minimum): a short tail is absorbed by the chunk before it.
#pragma once
#include <algorithm>

#include "common.cuh"

namespace petal {

constexpr int64_t kMinChunkRows = 1024;

inline void ensure_copy_stream(petal_ctx* ctx) {
    if (ctx->copy_stream != nullptr) return;
    PETAL_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (auto& e : ctx->copy_ev) PETAL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
}

template <typename T>
struct RowStream {
    petal_ctx* ctx = nullptr;
    const T* user = nullptr;
    int64_t n = 0, d = 0;
    bool host = false;    // user buffer is host memory
    bool ring = false;    // out-of-core: X is never fully resident
    bool loaded = false;  // resident copy complete
    DBuf<T> full;         // resident copy of a host X
    DBuf<T> slots;        // ring: two slots of slot_rows rows
    int64_t rows_each = 0, slot_rows = 0;
    bool slot_used[2] = {false, false};
    int64_t h2d_bytes = 0, traversals = 0;

    RowStream() {}
    RowStream(const RowStream&) = delete;
    RowStream& operator=(const RowStream&) = delete;
    ~RowStream() {
        // error paths may leave copies in flight that target buffers about to be freed on the compute stream
        if (host && ctx && ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
        if (host && ctx) {
            ctx->last_h2d_bytes = h2d_bytes;
            ctx->last_traversals = traversals;
            ctx->last_ring = ring ? 1 : 0;
        }
    }

    // `reserve_bytes`: HBM the flow will need besides X (the resident / ring decision in auto mode).
    void open(petal_ctx* c, const T* x, int64_t rows, int64_t cols, size_t reserve_bytes) {
        ctx = c;
        user = x;
        n = rows;
        d = cols;
        host = (x != nullptr) && rows > 0 && cols > 0 && !is_device_pointer(x);
        if (!host) return;
        int mode = ctx->host_staging;
        if (const char* e = getenv("PETAL_HOST_STAGING")) mode = atoi(e);
        int64_t chunk_bytes = ctx->host_chunk_bytes;
        if (const char* e = getenv("PETAL_HOST_CHUNK_BYTES")) chunk_bytes = atoll(e);
        const size_t bytes = (size_t)n * (size_t)d * sizeof(T);
        if (mode == 2) {
            ring = true;
        } else if (mode == 1) {
            ring = false;
        } else {
            size_t free_b = 0, total_b = 0;
            PETAL_CUDA(cudaMemGetInfo(&free_b, &total_b));
            // workspaces of earlier calls stay in the stream-ordered pool (release threshold = infinity): idle pool
            // memory is as good as free memory for the allocations of this call
            cudaMemPool_t pool;
            uint64_t reserved = 0, used = 0;
            if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess &&
                cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
                cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
                free_b += (size_t)(reserved - used);
            const double budget = 0.92 * (double)free_b - (double)reserve_bytes - 2.0 * (double)chunk_bytes;
            ring = (double)bytes > budget;
        }
        const int64_t row_bytes = d * (int64_t)sizeof(T);
        int64_t want = std::max<int64_t>(kMinChunkRows, chunk_bytes / std::max<int64_t>(row_bytes, 1));
        want = std::min<int64_t>(want, n);
        rows_each = std::max<int64_t>(32, ((want + 31) / 32) * 32);
        slot_rows = rows_each + kMinChunkRows;
        ensure_copy_stream(ctx);
        if (!ring) {
            try {
                full.alloc(ctx, (size_t)(n * d));
            } catch (const Error&) {
                // the estimate of what fits was too optimistic (fragmentation, another process on the device): go out
                // of core instead of failing, unless the caller insisted on a resident copy
                cudaGetLastError();
                full.p = nullptr;
                if (mode == 1) throw;
                ring = true;
            }
        }
        if (ring) slots.alloc(ctx, (size_t)(2 * slot_rows * d));
        // the buffers are stream-ordered allocations of the compute stream: the copy stream may touch them only after
        // the allocation point
        PETAL_CUDA(cudaEventRecord(ctx->copy_ev[4], ctx->stream));
        PETAL_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev[4], 0));
    }

    // all rows are (or will be after the first traversal) addressable through one device pointer
    bool is_resident() const { return !ring; }
    // one device pointer to all of X; only for device inputs or after the first traversal of a resident stream
    const T* resident_ptr() const {
        if (!host) return user;
        if (ring || !loaded) linalg_error("internal: resident pointer of a row stream that is not loaded");
        return full.p;
    }
    // alignment the chunks will have on the device (host inputs are staged into fresh 256 B-aligned allocations)
    bool chunks_aligned16() const { return host ? ((d * (int64_t)sizeof(T)) % 16 == 0) : is_aligned16(user); }
    // smallest chunk a traversal can hand out (the kernels' minimum-row conditions must hold for every chunk)
    int64_t min_chunk_rows() const { return (!host || loaded) ? n : std::min<int64_t>(n, rows_each); }

    // device pointer to the first `rows` rows (the provisional-mean sample) without a traversal
    const T* head(int64_t rows, DBuf<T>& tmp) {
        if (!host) return user;
        if (loaded) return full.p;
        tmp.alloc(ctx, (size_t)(rows * d));
        PETAL_CUDA(cudaMemcpyAsync(tmp.p, user, (size_t)(rows * d) * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        h2d_bytes += rows * d * (int64_t)sizeof(T);
        return tmp.p;
    }

    // makes a resident stream addressable through resident_ptr() (no-op for device inputs and rings)
    void load() {
        if (host && !ring && !loaded) traverse([](const T*, int64_t, int64_t) {});
    }

    // fn(const T* chunk, int64_t first_row, int64_t rows), called in row order with the work queued on ctx->stream
    template <typename F>
    void traverse(F&& fn) {
        ++traversals;
        if (!host || loaded) {
            fn(host ? (const T*)full.p : user, (int64_t)0, n);
            return;
        }
        cudaStream_t cs = ctx->copy_stream;
        int64_t r0 = 0;
        for (int64_t c = 0; r0 < n; ++c) {
            int64_t rows = std::min<int64_t>(rows_each, n - r0);
            if (n - (r0 + rows) < kMinChunkRows) rows = n - r0;  // absorb a short tail
            const int slot = (int)(c & 1);
            T* dst = ring ? slots.p + (size_t)slot * (size_t)(slot_rows * d) : full.p + (size_t)r0 * (size_t)d;
            if (ring && slot_used[slot]) PETAL_CUDA(cudaStreamWaitEvent(cs, ctx->copy_ev[2 + slot], 0));
            PETAL_CUDA(cudaMemcpyAsync(dst, user + (size_t)r0 * (size_t)d, (size_t)(rows * d) * sizeof(T), cudaMemcpyHostToDevice, cs));
            h2d_bytes += rows * d * (int64_t)sizeof(T);
            PETAL_CUDA(cudaEventRecord(ctx->copy_ev[slot], cs));
            PETAL_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[slot], 0));
            fn((const T*)dst, r0, rows);
            if (ring) {
                PETAL_CUDA(cudaEventRecord(ctx->copy_ev[2 + slot], ctx->stream));
                slot_used[slot] = true;
            }
            r0 += rows;
        }
        if (!ring) loaded = true;
    }
};

}  // namespace petal
