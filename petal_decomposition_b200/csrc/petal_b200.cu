// libpetal_b200 - C ABI implementation: the fit / transform flows of exact PCA, randomized PCA
// and FastICA, assembled from the streaming kernels (stream_kernels.cuh, tc_kernels.cuh) and the
// small on-device factorizations (small_linalg.cuh).  See include/petal_b200.h for the contract
// and DESIGN.md for the data layout and per-kernel rooflines.
#include <algorithm>
#include <mutex>

#include "comm.cuh"
#include "common.cuh"
#include "row_stream.cuh"
#include "small_linalg.cuh"
#include "stream_kernels.cuh"
#include "dense_f64.cuh"
#include "tc_kernels.cuh"
#include "ica_kernels.cuh"
#include "ica_deflation.cuh"

using namespace petal;

namespace {

// relative entry noise of a Gram matrix accumulated in f64 (orthogonality floor of the Jacobi solver)
const double kGramNoise = 4.0 * 2.220446049250313e-16;

std::string g_global_error;
std::mutex g_global_mutex;

// Device-side failures that have no host round trip of their own (Jacobi sweeps exhausted) are collected in
// ctx->dev_status and turned into the reference's error here (src/linalg.rs:84,115: "did not converge").
void check_device_status(petal_ctx* ctx) {
    if (!ctx->status_armed || ctx->dev_status == nullptr) return;
    ctx->status_armed = false;
    int h = 0;
    PETAL_CUDA(cudaMemcpyAsync(&h, ctx->dev_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PETAL_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h != 0) {
        PETAL_CUDA(cudaMemsetAsync(ctx->dev_status, 0, sizeof(int), ctx->stream));
        if (h & kStatusJacobiNotConverged) linalg_error("did not converge");
        linalg_error("device-side failure " + std::to_string(h));
    }
}

template <typename F>
int guarded(petal_ctx* ctx, F&& f) {
    if (!ctx) return PETAL_INVALID_INPUT;
    std::lock_guard<std::mutex> lock(ctx->mu);  // one call at a time per context
    try {
        PETAL_CUDA(cudaSetDevice(ctx->device));
        f();
        check_device_status(ctx);
        return PETAL_OK;
    } catch (const Error& e) {
        ctx->last_error = e.msg;
        cudaGetLastError();
        return e.code;
    } catch (const std::exception& e) {
        ctx->last_error = std::string("internal error: ") + e.what();
        return PETAL_LINALG_ERROR;
    }
}

// ---------------------------------------------------------------------------------------
// tiny kernels used by the flows
// ---------------------------------------------------------------------------------------
__global__ void set_value_kernel(double* p, double v) { *p = v; }

template <typename T>
__global__ void finish_mean_kernel(const double* __restrict__ sum, double inv_n, int64_t d,
                                   double* __restrict__ mean_d, T* __restrict__ mean_t) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d) return;
    double m = sum[j] * inv_n;
    T mt = (T)m;
    mean_t[j] = mt;
    mean_d[j] = (double)mt;  // the mean that is actually subtracted (type A, like the reference)
}

// ---- column means folded into the range finder's first two passes (randomized PCA, panel path) ----
// provisional mean mu~ = (T)(sum / count) from a row sample; sum[d] carries the sample's row count
template <typename T>
__global__ void provisional_mean_kernel(const double* __restrict__ sum, int64_t d, double* __restrict__ mean_d,
                                        T* __restrict__ mean_t) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d) return;
    const double cnt = sum[d];
    const T mt = (T)(cnt > 0.0 ? sum[j] / cnt : 0.0);
    mean_t[j] = mt;
    mean_d[j] = (double)mt;
}
// c = column sums of X - mu~ (last column of Zt [d x (l+1)]), delta = (T)(mu~ + c / n) - mu~ (the mean that will be
// subtracted from now on, minus the provisional one);  wu[0..l) = Omega^T delta, wu[l..2l) = Omega^T c;
// sc[0] = delta . c, sc[1] = |delta|^2.  One block per Omega column (plus one for the scalars).
template <typename T>
__global__ void __launch_bounds__(256)
fold_mean_vectors_kernel(const double* __restrict__ Zt, const T* __restrict__ Omega, int64_t ldo, int64_t d, int64_t l,
                         double inv_n, const double* __restrict__ mean_d, double* __restrict__ wu, double* __restrict__ sc) {
    const int64_t col = blockIdx.x;
    double a = 0.0, b = 0.0;
    for (int64_t j = threadIdx.x; j < d; j += 256) {
        const double c = Zt[j * (l + 1) + l];
        const double delta = (double)(T)(mean_d[j] + c * inv_n) - mean_d[j];
        if (col < l) {
            const double om = (double)Omega[j * ldo + col];
            a += om * delta;
            b += om * c;
        } else {
            a += delta * c;
            b += delta * delta;
        }
    }
    __shared__ double ra[256], rb[256];
    ra[threadIdx.x] = a;
    rb[threadIdx.x] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        double sa = 0.0, sb = 0.0;
        for (int i = 0; i < 256; ++i) {
            sa += ra[i];
            sb += rb[i];
        }
        if (col < l) {
            wu[col] = sa;
            wu[l + col] = sb;
        } else {
            sc[0] = sa;
            sc[1] = sb;
        }
    }
}
// Z[j][c] = Zt[j][c] - c_j w_c - delta_j u_c + n delta_j w_c   (= (X - mu)^T ((X - mu) Omega) from the products taken
// with the provisional mean);  mean <- mu~ + delta;  tv <- tv - 2 delta.c + n |delta|^2 (thread 0)
// (tv is still this rank's partial sum and is all-reduced later: each rank applies its 1/world share of the correction)
template <typename T>
__global__ void fold_mean_fix_kernel(const double* __restrict__ Zt, int64_t d, int64_t l, double inv_n, double n_total,
                                     const double* __restrict__ wu, const double* __restrict__ sc, const double* __restrict__ mean_d,
                                     double inv_world, double* __restrict__ Z, double* __restrict__ tv) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0) *tv = *tv + inv_world * (-2.0 * sc[0] + n_total * sc[1]);
    if (idx >= d * l) return;
    const int64_t j = idx / l, c = idx % l;
    const double cs = Zt[j * (l + 1) + l];
    const double mu0 = mean_d[j];
    const T mt = (T)(mu0 + cs * inv_n);
    const double delta = (double)mt - mu0;
    Z[idx] = Zt[j * (l + 1) + c] - cs * wu[c] - delta * wu[l + c] + n_total * delta * wu[c];
}
template <typename T>
__global__ void fold_mean_commit_kernel(const double* __restrict__ Zt, int64_t d, int64_t l, double inv_n,
                                        double* __restrict__ mean_d, T* __restrict__ mean_t) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d) return;
    const T mt = (T)(mean_d[j] + Zt[j * (l + 1) + l] * inv_n);
    mean_t[j] = mt;
    mean_d[j] = (double)mt;
}

__global__ void trace_kernel(const double* __restrict__ G, int64_t d, double* __restrict__ out, double scale = 1.0) {
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < d; i += blockDim.x) s += G[i * d + i];
    __shared__ double red[256];
    red[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < blockDim.x; ++i) t += red[i];
        *out = t * scale;
    }
}

// out[j] = (T) sqrt(max(lambda[j], 0))
template <typename T>
__global__ void sqrt_cast_kernel(const double* __restrict__ lam, int64_t k, T* __restrict__ out) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < k) out[j] = (T)sqrt(fmax(lam[j], 0.0));
}

// combine per-rank (absmax, idx, sign) triples: first rank with the strictly larger |.| wins
__global__ void combine_absmax_kernel(const double* __restrict__ gathered, int world, int64_t k,
                                      double* __restrict__ out3) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    double ba = gathered[j * 3 + 0], bi = gathered[j * 3 + 1], bs = gathered[j * 3 + 2];
    for (int r = 1; r < world; ++r) {
        const double* g = gathered + ((int64_t)r * k + j) * 3;
        if (g[0] > ba) {
            ba = g[0];
            bi = g[1];
            bs = g[2];
        }
    }
    out3[j * 3 + 0] = ba;
    out3[j * 3 + 1] = bi;
    out3[j * 3 + 2] = bs;
}

// K[i][j] = Jt[i][j] / sqrt(lambda[i]) * scale   (whitening matrix, reference src/ica.rs:190-203)
// Directions whose eigenvalue is below cutoff * lambda_max are numerically zero in the Gram matrix (its entries carry
// eps * lambda_max of noise): they are dropped, not amplified by 1/sqrt(lambda).
__global__ void whitening_kernel(const double* __restrict__ Jt, const double* __restrict__ lam, int64_t nc,
                                 int64_t d, double scale, double cutoff, double* __restrict__ K) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nc * d) return;
    int64_t i = idx / d;
    double l = lam[i];
    K[idx] = (l > cutoff * lam[0] && l > 0.0) ? Jt[idx] * rsqrt(l) * scale : 0.0;
}

// Gd[i][j] = HK[i][j] * inv_n - gp[i] * inv_n * W[i][j]     (reference src/ica.rs:334-342)
__global__ void ica_gd_kernel(const double* __restrict__ HK, const double* __restrict__ gp,
                              const double* __restrict__ W, int64_t nc, double inv_n, double* __restrict__ Gd) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nc * nc) return;
    int64_t i = idx / nc;
    Gd[idx] = HK[idx] * inv_n - gp[i] * inv_n * W[idx];
}

// lim = max_i | |sum_j W1[i][j] * (variant ? W[j][i] : W[i][j])| - 1 |   (reference src/ica.rs:344-354)
__global__ void ica_lim_kernel(const double* __restrict__ W1, const double* __restrict__ W, int nc, int variant,
                               double* __restrict__ lim) {
    __shared__ double red[256];
    double best = 0.0;
    for (int i = threadIdx.x; i < nc; i += blockDim.x) {
        double s = 0.0;
        for (int j = 0; j < nc; ++j) s += W1[(int64_t)i * nc + j] * (variant ? W[(int64_t)j * nc + i] : W[(int64_t)i * nc + j]);
        double v = fabs(fabs(s) - 1.0);
        if (v > best || v != v) best = v;
    }
    red[threadIdx.x] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = 0.0;
        for (int i = 0; i < blockDim.x; ++i)
            if (red[i] > b || red[i] != red[i]) b = red[i];
        *lim = b;
    }
}

__global__ void scale_cols_kernel(double* __restrict__ S, int64_t rows, int64_t cols,
                                  const double* __restrict__ sig) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < rows * cols) S[idx] *= sig[idx % cols];
}

// ---------------------------------------------------------------------------------------
// shared pieces of the flows
// ---------------------------------------------------------------------------------------
inline void launch1(petal_ctx* ctx) { check_launch(ctx); }

// Total row count over all ranks plus rank-uniform capability bits: every decision that changes which
// collectives a flow issues (panel layout, one-pass FastICA kernel, tcgen05 engine) must be the same on every
// rank, so each rank contributes the capabilities it LACKS and a path is taken only when no rank lacks it.
// One small all-reduce + host read per call (shapes are needed on the host anyway).
struct GlobalInfo {
    int64_t n_total;
    bool cap[4];
};
__global__ void set_info_kernel(double* p, double n, int lack0, int lack1, int lack2, int lack3) {
    p[0] = n;
    p[1] = lack0;
    p[2] = lack1;
    p[3] = lack2;
    p[4] = lack3;
}
GlobalInfo global_info(petal_ctx* ctx, int64_t n, bool c0 = true, bool c1 = true, bool c2 = true, bool c3 = true) {
    GlobalInfo gi{n, {c0, c1, c2, c3}};
    if (ctx->world <= 1) return gi;
    DBuf<double> buf(ctx, 5);
    set_info_kernel<<<1, 1, 0, ctx->stream>>>(buf.p, (double)n, c0 ? 0 : 1, c1 ? 0 : 1, c2 ? 0 : 1, c3 ? 0 : 1);
    launch1(ctx);
    allreduce_sum(ctx, buf.p, 5);
    double h[5] = {0, 0, 0, 0, 0};
    PETAL_CUDA(cudaMemcpyAsync(h, buf.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    PETAL_CUDA(cudaStreamSynchronize(ctx->stream));
    gi.n_total = (int64_t)(h[0] + 0.5);
    for (int i = 0; i < 4; ++i) gi.cap[i] = (h[1 + i] == 0.0);
    return gi;
}
int64_t global_rows(petal_ctx* ctx, int64_t n) { return global_info(ctx, n).n_total; }

// would a user buffer be 16 B aligned once it is on the device?  (host inputs are staged into fresh allocations)
inline bool aligned_on_device(const void* user) { return !is_device_pointer(user) || is_aligned16(user); }

template <typename T>
struct ColMean {
    DBuf<double> mean_d;
    DBuf<T> mean_t;
    const T* mu = nullptr;  // nullptr when centering is off
};

// reference: input.mean_axis(Axis(0)) (src/pca.rs:207,521; src/ica.rs:174)
template <typename T>
void compute_mean(petal_ctx* ctx, RowStream<T>& X, int64_t d, int64_t n_total, bool centering, ColMean<T>& out) {
    out.mean_d.alloc(ctx, (size_t)d);
    out.mean_t.alloc(ctx, (size_t)d);
    out.mean_d.zero();
    out.mean_t.zero();
    out.mu = nullptr;
    if (!centering || d == 0) return;
    DBuf<double> sum(ctx, (size_t)d);
    sum.zero();
    X.traverse([&](const T* Xc, int64_t, int64_t rows) { launch_colsum<T>(ctx, Xc, rows, d, d, sum.p); });
    allreduce_sum(ctx, sum.p, (size_t)d);
    finish_mean_kernel<T><<<(unsigned)ceil_div(d, 256), 256, 0, ctx->stream>>>(sum.p, 1.0 / (double)n_total, d,
                                                                               out.mean_d.p, out.mean_t.p);
    launch1(ctx);
    out.mu = out.mean_t.p;
}

// G[d x d] (f64) += (X - mu)^T (X - mu) over `n` rows (this rank's partial; upper tiles only on the symmetric engines)
template <typename T>
void centered_gram_acc(petal_ctx* ctx, const T* X, int64_t n, int64_t d, int64_t ld, const T* mu, double* G,
                       bool allow_tc = true) {
    if constexpr (sizeof(T) == 4) {
        // narrow f32 Gram (d <= 128): one tcgen05 pass, both operands centred on load.  Above 80 columns that pass
        // runs with long accumulation chains, whose truncation bias shows on sums of squares (the diagonal, i.e. the
        // total variance) at the 1e-5 level: callers that report trace(G) ask for the SIMT kernel (fp32-exact
        // products, f64 flush every 8192 rows) instead.
        if (allow_tc && ctx->f32_engine == 1 && tc::atb_supported(X, ld, d, X, ld, d, n) && is_aligned16(mu)) {
            tc::launch_tc_atb(ctx, X, ld, d, mu, X, ld, d, n, G, d, false, nullptr, mu);
            return;
        }
    }
    AtbParams<T> p{};
    p.A = X; p.lda = ld; p.da = d; p.mua = mu;
    p.B = X; p.ldb = ld; p.db = d; p.mub = mu;
    p.n = n; p.C = G; p.ldc = d; p.symmetric = 1;
    launch_atb<T>(ctx, p);
}
// G[d x d] (f64) = (X - mu)^T (X - mu), all-reduced over ranks and mirrored.
template <typename T>
void centered_gram(petal_ctx* ctx, const T* X, int64_t n, int64_t d, int64_t ld, const T* mu, double* G) {
    PETAL_CUDA(cudaMemsetAsync(G, 0, (size_t)(d * d) * sizeof(double), ctx->stream));
    centered_gram_acc<T>(ctx, X, n, d, ld, mu, G);
    allreduce_sum(ctx, G, (size_t)(d * d));
    launch_symmetrize(ctx, G, d);
}

// Column means and centred Gram matrix of a row stream (exact PCA pass 1, FastICA whitening).
//  X in HBM: the two passes of the reference order (mean, then the Gram of the centred rows).
//  X on the host: ONE traversal (one trip over PCIe, and the Gram runs underneath the transfer): the rows are
//  centred with a provisional mean mu~ (first rows of every rank's shard, all-reduced), the same traversal takes the
//  column sums c = sum(x - mu~), and with delta = mu - mu~ (mu = the type-T mean the later passes subtract)
//      (X - mu)^T (X - mu) = G~ - c delta^T - delta c^T + n delta delta^T.
//  delta is of the order sigma / sqrt(sample), so nothing cancels.
__global__ void gram_shift_kernel(double* __restrict__ G, int64_t d, const double* __restrict__ c,
                                  const double* __restrict__ delta, double n_total) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= d * d) return;
    const int64_t i = idx / d, j = idx % d;
    G[idx] += -c[i] * delta[j] - delta[i] * c[j] + n_total * delta[i] * delta[j];
}
// sum[j] (column sums of x over all ranks) -> mean (type T and its f64 image), c = sum - n mu~, delta = mu - mu~
template <typename T>
__global__ void shifted_mean_kernel(const double* __restrict__ sum, const double* __restrict__ mu0, double n_total, int64_t d,
                                    double* __restrict__ mean_d, T* __restrict__ mean_t, double* __restrict__ c,
                                    double* __restrict__ delta) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d) return;
    const T mt = (T)(sum[j] / n_total);
    c[j] = sum[j] - n_total * mu0[j];
    delta[j] = (double)mt - mu0[j];
    mean_t[j] = mt;
    mean_d[j] = (double)mt;
}
// `all_ranks_one_trip`: the caller has established (global_info) that EVERY rank is host-fed, so all of them take the
// one-trip route and issue its collectives; without that, several ranks keep the plain order (host-fed and device-fed
// ranks then issue identical collectives).
template <typename T>
void mean_and_gram(petal_ctx* ctx, RowStream<T>& X, int64_t d, int64_t n_total, bool centering, ColMean<T>& cm, double* G,
                   bool all_ranks_one_trip = false, bool allow_tc = true) {
    const bool one_trip = all_ranks_one_trip || (ctx->world == 1 && X.host && !X.loaded);
    if (!one_trip || !centering) {
        compute_mean<T>(ctx, X, d, n_total, centering, cm);
        PETAL_CUDA(cudaMemsetAsync(G, 0, (size_t)(d * d) * sizeof(double), ctx->stream));
        X.traverse([&](const T* Xc, int64_t, int64_t rows) { centered_gram_acc<T>(ctx, Xc, rows, d, d, cm.mu, G, allow_tc); });
        allreduce_sum(ctx, G, (size_t)(d * d));
        launch_symmetrize(ctx, G, d);
        return;
    }
    cm.mean_d.alloc(ctx, (size_t)d);
    cm.mean_t.alloc(ctx, (size_t)d);
    DBuf<double> sum(ctx, (size_t)d + 1), mu0(ctx, (size_t)d), c(ctx, (size_t)d), delta(ctx, (size_t)d);
    DBuf<T> head_tmp;
    sum.zero();
    const int64_t rows_s = std::min<int64_t>(X.n, 8192);
    if (rows_s > 0) launch_colsum<T>(ctx, X.head(rows_s, head_tmp), rows_s, d, d, sum.p);
    set_value_kernel<<<1, 1, 0, ctx->stream>>>(sum.p + d, (double)rows_s);
    launch1(ctx);
    allreduce_sum(ctx, sum.p, (size_t)d + 1);  // the same provisional mean on every rank
    provisional_mean_kernel<T><<<(unsigned)ceil_div(d, 256), 256, 0, ctx->stream>>>(sum.p, d, mu0.p, cm.mean_t.p);
    launch1(ctx);
    sum.zero();
    PETAL_CUDA(cudaMemsetAsync(G, 0, (size_t)(d * d) * sizeof(double), ctx->stream));
    X.traverse([&](const T* Xc, int64_t, int64_t rows) {
        launch_colsum<T>(ctx, Xc, rows, d, d, sum.p);
        centered_gram_acc<T>(ctx, Xc, rows, d, d, cm.mean_t.p, G, allow_tc);
    });
    allreduce_sum(ctx, sum.p, (size_t)d);
    allreduce_sum(ctx, G, (size_t)(d * d));
    launch_symmetrize(ctx, G, d);
    shifted_mean_kernel<T><<<(unsigned)ceil_div(d, 256), 256, 0, ctx->stream>>>(sum.p, mu0.p, (double)n_total, d, cm.mean_d.p,
                                                                               cm.mean_t.p, c.p, delta.p);
    launch1(ctx);
    gram_shift_kernel<<<(unsigned)ceil_div(d * d, 256), 256, 0, ctx->stream>>>(G, d, c.p, delta.p, (double)n_total);
    launch1(ctx);
    cm.mu = cm.mean_t.p;
}

template <typename T>
void gemm_xb(petal_ctx* ctx, const T* A, int64_t lda, int64_t n, int64_t K, const T* B, int64_t ldb,
             bool b_trans, int64_t L, const T* mu, const T* bias, T* Y, int64_t ldy, double* sumsq = nullptr) {
    if constexpr (sizeof(T) == 4) {
        if (ctx->f32_engine == 1 && bias == nullptr && tc::xb_supported(A, lda, n, K, L) && is_aligned16(mu)) {
            tc::launch_tc_xb<float>(ctx, A, lda, n, K, B, ldb, b_trans, L, mu, Y, ldy, sumsq);
            return;
        }
        // wide outputs (inverse_transform: L = d) and biased ones: 128-column windows of Y, one tcgen05 pass over A per
        // window (A is the narrow side there), the bias added in the epilogue.  PETAL_XB_WIDE=1 walks the windows inside
        // ONE launch instead (a CTA produces all windows of a 256-row super-tile back to back, the A tile re-read from L2
        // rather than HBM: 30 % less DRAM traffic) - measured SLOWER on B200 (4M x 64 -> 1024: 7.40 ms against 6.20 ms,
        // same box, r02): the pass is bound by the epilogue's store stream, not by the A reads; kept for experiments.
        const char* wide_env = getenv("PETAL_XB_WIDE");
        const int wide_mode = wide_env ? atoi(wide_env) : 2;
        if (wide_mode != 0 && ctx->f32_engine == 1 && sumsq == nullptr && (L > 128 || bias != nullptr) &&
            tc::xb_supported(A, lda, n, K, std::min<int64_t>(L, 128)) && is_aligned16(mu) && (ldy % 4 == 0) &&
            is_aligned16(Y)) {
            if (wide_mode == 2) {
                for (int64_t j0 = 0; j0 < L; j0 += 128) {
                    const int64_t lb = std::min<int64_t>(128, L - j0);
                    const float* Bw = b_trans ? B + j0 * ldb : B + j0;
                    tc::launch_tc_xb<float>(ctx, A, lda, n, K, Bw, ldb, b_trans, lb, mu, Y + j0, ldy, nullptr, false, nullptr, -1,
                                            bias ? bias + j0 : nullptr, lb);
                }
            } else {
                tc::launch_tc_xb<float>(ctx, A, lda, n, K, B, ldb, b_trans, L, mu, Y, ldy, nullptr, false, nullptr, -1, bias);
            }
            return;
        }
    }
    XbParams<T> p{};
    p.A = A; p.lda = lda; p.n = n; p.K = K; p.B = B; p.ldb = ldb; p.b_trans = b_trans ? 1 : 0; p.L = L;
    p.mu = mu; p.bias = bias; p.Y = Y; p.ldy = ldy; p.sumsq = sumsq;
    launch_xb<T>(ctx, p);
}

// Y = (A - mu) * B with B given in f64 (small replicated matrix): the tcgen05 engine splits B into
// hi/lo straight from the f64 values; the SIMT engine gets a rounded copy.
template <typename T>
void gemm_xb_b64(petal_ctx* ctx, const T* A, int64_t lda, int64_t n, int64_t K, const double* B, int64_t ldb,
                 int64_t L, const T* mu, T* Y, int64_t ldy) {
    if constexpr (sizeof(T) == 4) {
        if (ctx->f32_engine == 1 && tc::xb_supported(A, lda, n, K, L) && is_aligned16(mu)) {
            tc::launch_tc_xb<double>(ctx, A, lda, n, K, B, ldb, false, L, mu, Y, ldy, nullptr);
            return;
        }
    }
    DBuf<T> Bt(ctx, (size_t)(K * ldb));
    launch_cast<double, T>(ctx, B, Bt.p, K * ldb);
    gemm_xb<T>(ctx, A, lda, n, K, Bt.p, ldb, false, L, mu, nullptr, Y, ldy);
}

// C[da x db] (f64, zeroed here unless !zero: then accumulated) = (A - mua)^T (B - mub)
template <typename T>
void gemm_atb(petal_ctx* ctx, const T* A, int64_t lda, int64_t da, const T* mua, const T* B, int64_t ldb,
              int64_t db, const T* mub, int64_t n, double* C, bool zero = true) {
    if (zero) PETAL_CUDA(cudaMemsetAsync(C, 0, (size_t)(da * db) * sizeof(double), ctx->stream));
    if constexpr (sizeof(T) == 4) {
        if (ctx->f32_engine == 1 && mub == nullptr && tc::atb_supported(A, lda, da, B, ldb, db, n) &&
            is_aligned16(mua)) {
            tc::launch_tc_atb(ctx, A, lda, da, mua, B, ldb, db, n, C, db);
            return;
        }
    }
    AtbParams<T> p{};
    p.A = A; p.lda = lda; p.da = da; p.mua = mua; p.B = B; p.ldb = ldb; p.db = db; p.mub = mub;
    p.n = n; p.C = C; p.ldc = db; p.symmetric = 0;
    launch_atb<T>(ctx, p);
}

template <typename T>
double rank_cutoff() {
    // eigenvalues of a Gram matrix below cutoff * lambda_max are treated as numerically zero
    double eps = (sizeof(T) == 4) ? 1.1920929e-07 : 2.220446049250313e-16;
    double c = 8.0 * eps;
    return std::max(c * c, 1e-13);
}

// Orthonormal basis (f64) of the range of Z[rows x l] via two rounds of Gram + Jacobi eigh
// (robust to rank deficiency: null directions become zero columns).  In place.
// Plays the role of the reference's re-normalisation between power iterations
// (lu::Factorized::into_pl, src/pca.rs:709-713) - same range, better conditioned.
void orthonormalize_columns(petal_ctx* ctx, double* Z, int64_t rows, int64_t l, double cutoff) {
    DBuf<double> G(ctx, (size_t)(l * l)), P(ctx, (size_t)(l * l));
    DBuf<double> Z1(ctx, (size_t)(rows * l));
    for (int round = 0; round < 2; ++round) {
        gemm_atb<double>(ctx, Z, l, l, nullptr, Z, l, l, nullptr, rows, G.p);
        gram_to_orthonormalizer(ctx, G.p, l, round == 0 ? cutoff : 1e-6, kGramNoise, P.p);
        gemm_xb<double>(ctx, Z, l, rows, l, P.p, l, false, l, nullptr, nullptr, Z1.p, l);
        PETAL_CUDA(cudaMemcpyAsync(Z, Z1.p, (size_t)(rows * l) * sizeof(double), cudaMemcpyDeviceToDevice,
                                   ctx->stream));
    }
}

// svd_flip (reference src/pca.rs:815-850) on the k leading columns: decides each sign from the
// max-|.| entry of the score column (same sign as the U column since sigma >= 0), first row
// wins, across all ranks; flips score columns and component rows.
template <typename T>
void flip_signs(petal_ctx* ctx, T* scores, int64_t n, int64_t k, T* comps, int64_t d, bool scores_wanted = true,
                double* have_local3 = nullptr) {
    if (k == 0) return;
    DBuf<double> local3;
    if (have_local3 == nullptr) {
        local3.alloc(ctx, (size_t)(k * 3));
        launch_colabsmax<T>(ctx, scores, n, k, k, local3.p);
        have_local3 = local3.p;
    }
    double* flip3 = have_local3;
    DBuf<double> gathered, out3;
    if (ctx->world > 1) {
        gathered.alloc(ctx, (size_t)(ctx->world * k * 3));
        out3.alloc(ctx, (size_t)(k * 3));
        allgather(ctx, have_local3, gathered.p, (size_t)(k * 3));
        combine_absmax_kernel<<<(unsigned)ceil_div(k, 128), 128, 0, ctx->stream>>>(gathered.p, ctx->world, k, out3.p);
        launch1(ctx);
        flip3 = out3.p;
    }
    launch_apply_flip<T>(ctx, scores_wanted ? scores : nullptr, n, k, k, comps, d, flip3);
}

void finish_call(petal_ctx* ctx, bool any_host_output) {
    if (any_host_output) PETAL_CUDA(cudaStreamSynchronize(ctx->stream));
}

// debug aid: PETAL_PHASES=1 prints host wall time per phase (synchronising after each)
struct PhaseClock {
    petal_ctx* ctx;
    bool on;
    double t0;
    static double now() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    }
    explicit PhaseClock(petal_ctx* c) : ctx(c), on(getenv("PETAL_PHASES") != nullptr), t0(0) {
        if (on) { cudaStreamSynchronize(ctx->stream); t0 = now(); }
    }
    void mark(const char* label) {
        if (!on) return;
        double t1 = now();
        cudaStreamSynchronize(ctx->stream);
        double t2 = now();
        fprintf(stderr, "[petal phase] %-28s host %8.3f ms  (+%8.3f ms waiting for the device)\n", label, t1 - t0, t2 - t1);
        t0 = now();
    }
};

std::string dim_message(int64_t k) { return "every dimension should be at least " + std::to_string(k); }

// ---------------------------------------------------------------------------------------
// exact PCA  (reference Pca::inner_fit, src/pca.rs:195-231)
// ---------------------------------------------------------------------------------------
template <typename T>
void pca_fit(petal_ctx* ctx, const T* x_user, int64_t n, int64_t d, int64_t k, bool centering, T* comps_u,
             T* mean_u, T* sing_u, T* tv_u, T* scores_u) {
    if (n < 0 || d < 0 || k < 0) invalid_input("negative dimension");
    const int64_t n_total = global_rows(ctx, n);
    if (n_total < k || d < k) invalid_input(dim_message(k));  // src/pca.rs:199-204
    if (n_total == 0) return;                                  // src/pca.rs:207-211
    if (d == 0) return;

    RowStream<T> X;
    X.open(ctx, x_user, n, d, (size_t)(10 * d * d) * sizeof(double) + ((size_t)1 << 30) + (size_t)(n * k) * sizeof(T));
    DevOut<T> comps(ctx, comps_u, (size_t)(k * d)), mean(ctx, mean_u, (size_t)d), sing(ctx, sing_u, (size_t)k),
        tv(ctx, tv_u, 1), scores(ctx, scores_u, (size_t)(n * k));

    ColMean<T> cm;

    // sigma and Vt of the centred data (what the reference takes from gesvd, src/pca.rs:216-220); the n x n U is
    // never formed.
    //  f64: CholeskyQR2 of Xc - pass 1: G1 = Xc^T Xc, R1 = chol(G1); pass 2: G2 = (Xc R1^-1)^T (Xc R1^-1) with the
    //       triangular solve applied on the fly chunk by chunk, R2 = chol(G2); R = R2 R1 - followed by a one-sided
    //       Jacobi SVD of R (rows): singular values to eps * sigma_1 like a backward-stable SVD of Xc, instead of the
    //       eps * sigma_1^2 / sigma_j a Gram eigen-decomposition gives.  Rank-deficient data (a Cholesky pivot
    //       below 1e-12 of the largest diagonal entry, or fewer rows than columns) takes the second route.
    //  f32, and the fallback: eigen-decomposition of G1 (f64 accumulation: far inside the f32 tolerance).
    DBuf<double> G(ctx, (size_t)(d * d)), Jt(ctx, (size_t)(d * d)), lam(ctx, (size_t)d), tvd(ctx, 1);
    mean_and_gram<T>(ctx, X, d, n_total, centering, cm, G.p);
    trace_kernel<<<1, 256, 0, ctx->stream>>>(G.p, d, tvd.p);  // sum of all sigma^2, src/pca.rs:224
    launch1(ctx);
    bool have_sigma = false;  // lam holds sigma (true) or sigma^2 (false)
    if constexpr (sizeof(T) == 8) {
        bool qr2 = n_total > d && d >= 2;
        if (const char* e = getenv("PETAL_PCA_QR2")) qr2 = qr2 && atoi(e) != 0;
        if (qr2) {
            const double piv = 1e-12;
            DBuf<double> R1(ctx, (size_t)(d * d)), P1(ctx, (size_t)(d * d));
            DBuf<int> fail(ctx, 1);
            fail.zero();
            PETAL_CUDA(cudaMemcpyAsync(R1.p, G.p, (size_t)(d * d) * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
            chol_blocked(ctx, R1.p, d, piv, P1.p, fail.p);
            int hfail = 0;
            PETAL_CUDA(cudaMemcpyAsync(&hfail, fail.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            PETAL_CUDA(cudaStreamSynchronize(ctx->stream));
            if (!hfail) {
                // pass 2: Q1 = Xc P1 in row chunks (workspace <= 1 GiB), G2 += Q1^T Q1
                DBuf<double> G2(ctx, (size_t)(d * d));
                G2.zero();
                int64_t rows_c = std::max<int64_t>(128, ((int64_t)1 << 30) / (d * (int64_t)sizeof(double)));
                rows_c = std::min<int64_t>(std::max<int64_t>(n, 1), (rows_c / 128) * 128);
                DBuf<double> Q1(ctx, (size_t)(rows_c * d));
                X.traverse([&](const T* Xc, int64_t, int64_t rows_x) {
                    for (int64_t r0 = 0; r0 < rows_x; r0 += rows_c) {
                        const int64_t rows = std::min<int64_t>(rows_c, rows_x - r0);
                        gemm_nn(ctx, reinterpret_cast<const double*>(Xc) + r0 * d, d, P1.p, d, Q1.p, d, rows, d, d, 1.0, false, false,
                                /*b_upper=*/true, false, reinterpret_cast<const double*>(cm.mu));
                        AtbParams<double> ap{};
                        ap.A = Q1.p; ap.lda = d; ap.da = d; ap.B = Q1.p; ap.ldb = d; ap.db = d; ap.n = rows; ap.C = G2.p; ap.ldc = d;
                        ap.symmetric = 1;
                        launch_atb<double>(ctx, ap);
                    }
                });
                allreduce_sum(ctx, G2.p, (size_t)(d * d));
                launch_symmetrize(ctx, G2.p, d);
                DBuf<double> P2(ctx, (size_t)(d * d));
                chol_blocked(ctx, G2.p, d, 1e-6, P2.p, fail.p);  // G2 ~ I: anything near-singular here means round 1 failed
                PETAL_CUDA(cudaMemcpyAsync(&hfail, fail.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
                PETAL_CUDA(cudaStreamSynchronize(ctx->stream));
                if (!hfail) {
                    // R = R2 R1 (upper x upper), then SVD of R by one-sided Jacobi on its rows: R = U diag(s) N, N = Vt
                    DBuf<double> R(ctx, (size_t)(d * d)), Aout(ctx, (size_t)(d * d));
                    R.zero();
                    gemm_nn(ctx, G2.p, d, R1.p, d, R.p, d, d, d, d, 1.0, false, /*a_upper=*/true, /*b_upper=*/true, /*c_upper=*/true);
                    jacobi_rows(ctx, R.p, d, d, Aout.p, nullptr, lam.p);
                    launch_normalize_rows(ctx, Aout.p, lam.p, d, d, 0.0, Jt.p);
                    have_sigma = true;
                }
            }
        }
    }
    if (!have_sigma) jacobi_rows(ctx, G.p, d, d, nullptr, Jt.p, lam.p, kGramNoise);

    DBuf<T> comps_tmp;
    T* comps_dev = comps.p;
    if (!comps_dev) {
        comps_tmp.alloc(ctx, (size_t)(k * d));
        comps_dev = comps_tmp.p;
    }
    launch_cast<double, T>(ctx, Jt.p, comps_dev, k * d);  // components = vt[0..k] (src/pca.rs:225)

    if (k > 0) {
        // scores = Xc * V_k^T = U_k * sigma_k (transform_with_u, src/pca.rs:758-779); also the
        // carrier of the u-based sign decision of svd_flip.
        DBuf<T> scores_tmp;
        T* scores_dev = scores.p;
        if (!scores_dev) {
            scores_tmp.alloc(ctx, (size_t)(n * k));
            scores_dev = scores_tmp.p;
        }
        X.traverse([&](const T* Xc, int64_t r0, int64_t rows) {
            gemm_xb<T>(ctx, Xc, d, rows, d, comps_dev, d, true, k, cm.mu, nullptr, scores_dev + r0 * k, k);
        });
        flip_signs<T>(ctx, scores_dev, n, k, comps_dev, d, (bool)scores);
        if (sing) {
            if (have_sigma) {
                launch_cast<double, T>(ctx, lam.p, sing.p, k);
            } else {
                sqrt_cast_kernel<T><<<(unsigned)ceil_div(k, 128), 128, 0, ctx->stream>>>(lam.p, k, sing.p);
                launch1(ctx);
            }
        }
    }
    if (mean) launch_cast<T, T>(ctx, cm.mean_t.p, mean.p, d);
    if (tv) launch_cast<double, T>(ctx, tvd.p, tv.p, 1);

    comps.commit(ctx); mean.commit(ctx); sing.commit(ctx); tv.commit(ctx); scores.commit(ctx);
    finish_call(ctx, comps.to_host || mean.to_host || sing.to_host || tv.to_host || scores.to_host);
}

// ---------------------------------------------------------------------------------------
// randomized PCA  (reference RandomizedPca::inner_fit / randomized_svd / randomized_range_finder,
//                  src/pca.rs:509-550, 668-718)
// ---------------------------------------------------------------------------------------
template <typename T>
void rpca_fit(petal_ctx* ctx, const T* x_user, int64_t n, int64_t d, int64_t k, bool centering,
              int64_t n_over, int64_t n_iter, const T* omega_user, T* comps_u, T* mean_u, T* sing_u, T* tv_u,
              T* scores_u) {
    if (n < 0 || d < 0 || k < 0 || n_over < 0 || n_iter < 0) invalid_input("negative dimension");
    // local capability for the panel-major tcgen05 path (decided for all ranks together, see global_info)
    bool panel_local = false;
    if constexpr (sizeof(T) == 4) {
        const int64_t l_guess = std::min<int64_t>(k + n_over, d);
        panel_local = ctx->f32_engine == 1 && aligned_on_device(x_user) && tc::xb_supported(nullptr, d, n, d, std::max<int64_t>(l_guess, 1)) &&
                      n >= 1024 && d >= 32;
        if (const char* e = getenv("PETAL_PANEL")) panel_local = panel_local && atoi(e) != 0;
    }
    // rank-uniform "every rank's shard is a non-empty host buffer" (the Gram-mode decision below changes the collectives)
    const bool host_fed_local = x_user != nullptr && n > 0 && d > 0 && !is_device_pointer(x_user) && ctx->host_gram != 0;
    const bool dev_fed_local = x_user != nullptr && n > 0 && d > 0 && is_device_pointer(x_user) && ctx->host_gram != 0;
    const GlobalInfo ginfo = global_info(ctx, n, panel_local, host_fed_local, dev_fed_local);
    const int64_t n_total = ginfo.n_total;
    if (n_total < k || d < k) invalid_input(dim_message(k));  // src/pca.rs:513-518
    if (n_total == 0 || d == 0) return;                        // src/pca.rs:521-525
    if (omega_user == nullptr) invalid_input("omega (d x (k + n_oversamples) test matrix) is required");
    const int64_t l_full = k + n_over;                         // src/pca.rs:679
    // working width: the reference shrinks to min(rows, cols) after the first product
    // (src/pca.rs:710,713); we use the leading l columns of Omega from the start.
    const int64_t l = std::min<int64_t>(l_full, std::min<int64_t>(n_total, d));
    if (l == 0) return;

    PhaseClock pc(ctx);
    const int64_t ly = ((l + 15) / 16) * 16;  // row pitch of Y: whole 64 B chunks (TMA reads, vector epilogue stores)
    RowStream<T> X;
    X.open(ctx, x_user, n, d, (size_t)(2 * n * ly + n * k) * sizeof(T) + (size_t)(8 * d * l) * sizeof(double));
    DevIn<T> Omega(ctx, omega_user, (size_t)(d * l_full));
    DevOut<T> comps(ctx, comps_u, (size_t)(k * d)), mean(ctx, mean_u, (size_t)d), sing(ctx, sing_u, (size_t)k),
        tv(ctx, tv_u, 1), scores(ctx, scores_u, (size_t)(n * k));

    // f32 on the tcgen05 engine keeps Y panel-major ([row block of 32][ly][32]): tc_xb then writes whole 128 B
    // lines and tc_atb reads its B tiles by TMA without transposing (see tc_kernels.cuh).
    // rank-uniform: every rank can run the panel path on every chunk of its shard (otherwise ranks would issue
    // different collectives: the panel path reduces C' = Xc^T Y, the row-major path G1, then [G2 | C'])
    bool panel = false;
    if constexpr (sizeof(T) == 4) {
        panel = ginfo.cap[0] && X.chunks_aligned16() && tc::xb_supported(nullptr, d, X.min_chunk_rows(), d, l) &&
                X.min_chunk_rows() >= 1024;
        if (ginfo.cap[0] && !panel && ctx->world > 1) linalg_error("inconsistent panel-path decision across ranks");
    }

    // Column means.  On the panel path (f32, tcgen05) with at least one power iteration and a spare padding column
    // in the Y panel, the separate pass over X is saved: X Omega runs with a provisional mean mu~ taken from a row
    // sample, its panel carries a column of ones, so the X^T Y pass that follows also returns the column sums of
    // X - mu~; the exact mean mu = mu~ + delta and the exact Z = Xc^T (Xc Omega), ||Xc||_F^2 follow by rank-one
    // corrections of the small side (delta ~ sigma / sqrt(sample): no cancellation).
    // Host-fed X (on every rank): while the rows cross PCIe the GPU is idle - the transfer of a chunk takes several times
    // as long as any pass over it - so the ingest traversal also accumulates the d x d Gram matrix G = Xc^T Xc (f64
    // accumulation; `mean_and_gram`, one trip).  The power iterations Z <- Xc^T (Xc B) = G B (src/pca.rs:708-715) then
    // run on the replicated small side without touching X again; only the last pair - Y = Xc B_q and C' = Xc^T Y, the
    // products that define the result - is taken from X itself, so the singular values keep the accuracy of a direct
    // pass (G only has to preserve the dominant subspace, like every intermediate iterate).  Resident copy: q of the
    // q + 1 pass pairs disappear behind the transfer; out of core: 2 trips over PCIe instead of q + 1.
    // For X already in HBM the route is taken only when it is cheaper in passes over X (gram_dev below).
    bool gram_mode = ginfo.cap[1] && host_fed_local && n_iter >= 1 && d <= 2048 && n_total >= 2 * d;
    // X already in HBM (on every rank), f32 tcgen05 path: the same route pays when G costs fewer passes over X than the
    // power iterations it replaces.  G is taken in 64-column windows, G[:, w] = Xc^T Xc[:, w], each one precise tc_atb
    // pass with the window of X itself as the (row-major, centred-on-load) Y operand - the kernel quality of the X^T Y
    // passes they stand in for; a window pass costs ~1.25 plain passes (the Y tile is transposed by the transform warps),
    // the mean one 0.6-pass column sum.  d = 256 (c5), q = 4: 4 windows + final pair against 5 pairs.
    bool gram_dev = false;
    if constexpr (sizeof(T) == 4) {
        const double gwin = (double)ceil_div(d, 64);
        gram_dev = ginfo.cap[2] && dev_fed_local && panel && n_iter >= 2 && n_total >= 2 * d && d % 4 == 0 &&
                   1.25 * gwin < 2.0 * (double)n_iter - 1.1 && tc::atb_supported(x_user, d, d, x_user, d, 64, n);
        // (rank-uniform: `panel` already says that every rank's shard has >= 1024 rows and is 16 B aligned)
    }
    if (const char* e = getenv("PETAL_RPCA_GRAM")) {  // (set it on every rank)
        gram_mode = gram_mode && atoi(e) != 0;
        gram_dev = gram_dev && atoi(e) != 0;
    }
    const bool gram_any = gram_mode || gram_dev;
    DBuf<double> Gm(ctx, gram_any ? (size_t)(d * d) : 0);

    ColMean<T> cm;
    bool fold_mean = false;
    if constexpr (sizeof(T) == 4) {
        fold_mean = !gram_any && panel && centering && n_iter >= 1 && (l % 16) != 0;
        if (const char* e = getenv("PETAL_FOLD_MEAN")) fold_mean = fold_mean && atoi(e) != 0;
    }
    if (gram_mode) {
        mean_and_gram<T>(ctx, X, d, n_total, centering, cm, Gm.p, /*all_ranks_one_trip=*/true, /*allow_tc=*/false);
    } else if (gram_dev) {
        compute_mean<T>(ctx, X, d, n_total, centering, cm);
        Gm.zero();
        if constexpr (sizeof(T) == 4) {
            X.traverse([&](const T* Xc, int64_t, int64_t rows) {
                for (int64_t c0 = 0; c0 < d; c0 += 64) {
                    const int64_t wc = std::min<int64_t>(64, d - c0);
                    tc::launch_tc_atb(ctx, Xc, d, d, cm.mu, Xc + c0, d, wc, rows, Gm.p + c0, d, false, nullptr,
                                      cm.mu ? cm.mu + c0 : nullptr, -1);
                }
            });
        }
        allreduce_sum(ctx, Gm.p, (size_t)(d * d));
    } else if (fold_mean) {
        cm.mean_d.alloc(ctx, (size_t)d);
        cm.mean_t.alloc(ctx, (size_t)d);
        DBuf<double> sum(ctx, (size_t)d + 1);
        DBuf<T> head_tmp;
        sum.zero();
        const int64_t rows_s = std::min<int64_t>(n, 8192);
        launch_colsum<T>(ctx, X.head(rows_s, head_tmp), rows_s, d, d, sum.p);
        set_value_kernel<<<1, 1, 0, ctx->stream>>>(sum.p + d, (double)rows_s);
        launch1(ctx);
        allreduce_sum(ctx, sum.p, (size_t)d + 1);
        provisional_mean_kernel<T><<<(unsigned)ceil_div(d, 256), 256, 0, ctx->stream>>>(sum.p, d, cm.mean_d.p, cm.mean_t.p);
        launch1(ctx);
        cm.mu = cm.mean_t.p;
    } else {
        compute_mean<T>(ctx, X, d, n_total, centering, cm);
    }
    const double cutoff = rank_cutoff<T>();
    pc.mark("stage inputs + mean");

    const int64_t nblk = ceil_div(n, 32);
    DBuf<T> Y(ctx, panel ? (size_t)(nblk * ly * 32) : (size_t)(n * ly));
    // y - tf32(y), the B_lo operand of the X^T Y passes, is derived inside tc_atb by default; PETAL_YLO=1 keeps the
    // older second panel in HBM (written by tc_xb, TMA-loaded by tc_atb) for comparison
    const bool ylo_panel = panel && getenv("PETAL_YLO") != nullptr && atoi(getenv("PETAL_YLO")) != 0;
    DBuf<T> Ylo(ctx, ylo_panel ? (size_t)(nblk * ly * 32) : 0);
    DBuf<double> small(ctx, (size_t)(l * l + d * l + 1));  // [Gram of Y | C' | tv] reduced together
    double* G2 = small.p;
    double* Cp = small.p + l * l;
    double* tvd = small.p + l * l + d * l;
    DBuf<double> P(ctx, (size_t)(l * l));
    DBuf<T> Y1;
    PETAL_CUDA(cudaMemsetAsync(tvd, 0, sizeof(double), ctx->stream));

    // The range finder (src/pca.rs:707-715) as q + 1 traversals of X.  Traversal j computes, chunk by chunk,
    //     Y_c = Xc_c B_j     (B_0 = Omega: src/pca.rs:707, fused with ||Xc||_F^2, src/pca.rs:533; B_j = orth(Z_j): :713)
    //     Z_{j+1} += Xc_c^T Y_c   (src/pca.rs:711; the last one is C' = Xc^T Y of the projection, src/pca.rs:681)
    // Both contractions of an iteration only need the chunk that is resident, so a host X that does not fit in HBM
    // crosses PCIe q + 1 times instead of 2q + 2, and a host X that does fit is consumed while it arrives.
    // With X in HBM there is one chunk and this is the plain pass order.
    // Only the passes that define the result (the last Y and C' = Xc^T Y) run with cut accumulation chains; the
    // X Z passes of the power iterations in between only have to keep the dominant subspace and use the faster long
    // chains.  X^T Y: its truncation bias (unlike X Z's, which mostly rescales columns) tilts the subspace - measured:
    // even two long-chain passes at the start cost two digits on the trailing components, so all of them run with cut
    // chains (PETAL_ATB_FAST_ITERS = number of leading iterations on the fast path, for experiments).
    int64_t fast_atb_iters = 0;
    if (const char* e = getenv("PETAL_ATB_FAST_ITERS")) fast_atb_iters = atoi(e);
    DBuf<double> Zd(ctx, (size_t)(d * l)), Zacc(ctx, (size_t)(d * (l + 1))), wu(ctx, (size_t)(2 * l)), sc(ctx, 2);
    if (fold_mean && !panel) linalg_error("inconsistent panel-path decision (folded mean)");
    int64_t j_first = 0;
    if (gram_any) {
        // ||Xc||_F^2 = trace(G) (src/pca.rs:533); Z_1 = G Omega, Z_{i+1} = G orth(Z_i)
        // (G is already summed over the ranks; tv is all-reduced with C' later, so each rank carries its 1/world share)
        trace_kernel<<<1, 256, 0, ctx->stream>>>(Gm.p, d, tvd, 1.0 / (double)ctx->world);
        launch1(ctx);
        DBuf<double> Om(ctx, (size_t)(d * l_full));
        launch_cast<T, double>(ctx, Omega.p, Om.p, d * l_full);
        const double* Bcur = Om.p;
        int64_t ldb = l_full;
        for (int64_t it = 0; it < n_iter; ++it) {
            gemm_xb<double>(ctx, Gm.p, d, d, d, Bcur, ldb, false, l, nullptr, nullptr, Zacc.p, l);
            PETAL_CUDA(cudaMemcpyAsync(Zd.p, Zacc.p, (size_t)(d * l) * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
            orthonormalize_columns(ctx, Zd.p, d, l, cutoff);
            Bcur = Zd.p;
            ldb = l;
        }
        pc.mark("power iterations on G");
        j_first = n_iter;
    }
    for (int64_t j = j_first; j <= n_iter; ++j) {
        const bool last = (j == n_iter);
        const bool folded = fold_mean && j == 0;
        // accumulator of this traversal's X^T Y: Zt [d x (l + 1)] with the ones column when the mean is folded in,
        // C' for the last one, otherwise the next Z (never the buffer the X B products of the same traversal read)
        const bool do_atb = panel || !last;
        double* acc = last ? Cp : Zacc.p;
        const int64_t acc_cols = folded ? l + 1 : l;
        if (do_atb) PETAL_CUDA(cudaMemsetAsync(acc, 0, (size_t)(d * acc_cols) * sizeof(double), ctx->stream));
        const int xb_precise = last ? -1 : 0;
        const int atb_precise = (!last && j < fast_atb_iters) ? 0 : -1;
        X.traverse([&](const T* Xc, int64_t r0, int64_t rows) {
            bool done = false;
            if constexpr (sizeof(T) == 4) {
                if (panel) {
                    float* Yc = Y.p + (r0 / 32) * ly * 32;
                    float* Yloc = ylo_panel ? Ylo.p + (r0 / 32) * ly * 32 : nullptr;
                    if (j == 0)
                        tc::launch_tc_xb<float>(ctx, Xc, d, rows, d, Omega.p, l_full, false, l, cm.mu, Yc, ly, tvd, true, Yloc,
                                                xb_precise, nullptr, -1, fold_mean ? (int)l : -1);
                    else
                        tc::launch_tc_xb<double>(ctx, Xc, d, rows, d, Zd.p, l, false, l, cm.mu, Yc, ly, nullptr, true, Yloc,
                                                 xb_precise);
                    tc::launch_tc_atb(ctx, Xc, d, d, cm.mu, Yc, ly, acc_cols, rows, acc, acc_cols, true, Yloc, nullptr,
                                      atb_precise);
                    done = true;
                }
            }
            if (!done) {
                T* Yc = Y.p + r0 * ly;
                if (j == 0) gemm_xb<T>(ctx, Xc, d, rows, d, Omega.p, l_full, false, l, cm.mu, nullptr, Yc, ly, tvd);
                else gemm_xb_b64<T>(ctx, Xc, d, rows, d, Zd.p, l, l, cm.mu, Yc, ly);
                if (do_atb) gemm_atb<T>(ctx, Xc, d, d, cm.mu, Yc, ly, l, nullptr, rows, acc, /*zero=*/false);
            }
        });
        pc.mark(last ? "Y = Xc Z, C' = Xc^T Y" : "  Y = Xc B, Z = Xc^T Y");
        if (last) break;
        if (folded) {
            if constexpr (sizeof(T) == 4) {
                // Zt = (X - mu~)^T [Y~ | 1], then the rank-one corrections
                allreduce_sum(ctx, Zacc.p, (size_t)(d * (l + 1)));
                const double inv_n = 1.0 / (double)n_total;
                fold_mean_vectors_kernel<T><<<(unsigned)(l + 1), 256, 0, ctx->stream>>>(Zacc.p, Omega.p, l_full, d, l, inv_n,
                                                                                       cm.mean_d.p, wu.p, sc.p);
                launch1(ctx);
                fold_mean_fix_kernel<T><<<(unsigned)ceil_div(d * l, 256), 256, 0, ctx->stream>>>(
                    Zacc.p, d, l, inv_n, (double)n_total, wu.p, sc.p, cm.mean_d.p, 1.0 / (double)ctx->world, Zd.p, tvd);
                launch1(ctx);
                fold_mean_commit_kernel<T><<<(unsigned)ceil_div(d, 256), 256, 0, ctx->stream>>>(Zacc.p, d, l, inv_n, cm.mean_d.p,
                                                                                             cm.mean_t.p);
                launch1(ctx);
            }
        } else {
            allreduce_sum(ctx, Zacc.p, (size_t)(d * l));
            PETAL_CUDA(cudaMemcpyAsync(Zd.p, Zacc.p, (size_t)(d * l) * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        }
        orthonormalize_columns(ctx, Zd.p, d, l, cutoff);
        pc.mark("  orth(Z)");
    }

    if (panel) {
        // thin QR of Y (src/pca.rs:716), implicit: G = Y^T Y accumulated in f64 from the exact fp32 products,
        // P = R^-1 (Cholesky; Jacobi when rank deficient) so that Q = Y P is orthonormal to eps64 * cond(Y)^2.
        // B = Q^T Xc (src/pca.rs:681) = P^T (Y^T Xc): C' = Xc^T Y came out of the last traversal.
        // After a power iteration Y = Xc Z with the replicated Z still in Zd, so G = Y^T Y = Z^T (Xc^T Y) = Z^T C':
        // a small replicated product instead of another pass over Y (same ~eps32 relative accuracy as the
        // fp32-accumulated pass over the panels it replaces; that pass remains for n_iter == 0).
        bool gram_from_c = n_iter > 0;
        if (const char* e = getenv("PETAL_GRAM_FROM_C")) gram_from_c = gram_from_c && atoi(e) != 0;
        if constexpr (sizeof(T) == 4) {
            if (!gram_from_c) {
                PETAL_CUDA(cudaMemsetAsync(G2, 0, (size_t)(l * l) * sizeof(double), ctx->stream));
                DBuf<double> Gp(ctx, (size_t)(ly * ly));
                Gp.zero();
                launch_panel_gram(ctx, Y.p, n, (int)ly, Gp.p);
                // compact ly x ly -> l x l
                PETAL_CUDA(cudaMemcpy2DAsync(G2, (size_t)l * sizeof(double), Gp.p, (size_t)ly * sizeof(double),
                                             (size_t)l * sizeof(double), (size_t)l, cudaMemcpyDeviceToDevice, ctx->stream));
            }
        }
        if (gram_from_c) {
            allreduce_sum(ctx, Cp, (size_t)(d * l + 1));
            gemm_atb<double>(ctx, Zd.p, l, l, nullptr, Cp, l, l, nullptr, d, G2);
            launch_symmetrize(ctx, G2, l);
        } else {
            allreduce_sum(ctx, small.p, (size_t)(l * l + d * l + 1));
        }
        pc.mark("G = Y^T Y");
        gram_to_orthonormalizer(ctx, G2, l, cutoff, kGramNoise, P.p);
    } else {
        // thin QR of Y (src/pca.rs:716) done implicitly in two Gram rounds (CholeskyQR2-style, with a
        // Jacobi eigensolver instead of Cholesky when Y is rank deficient):
        //   round 1: Y1 = Y * P1 (materialised),  round 2: Q = Y1 * P2 (implicit)
        DBuf<double> G1(ctx, (size_t)(l * l));
        Y1.alloc(ctx, (size_t)(n * ly));
        gemm_atb<T>(ctx, Y.p, ly, l, nullptr, Y.p, ly, l, nullptr, n, G1.p);
        allreduce_sum(ctx, G1.p, (size_t)(l * l));
        gram_to_orthonormalizer(ctx, G1.p, l, cutoff, kGramNoise, P.p);
        pc.mark("G1 = Y^T Y, P1");
        gemm_xb_b64<T>(ctx, Y.p, ly, n, l, P.p, l, l, nullptr, Y1.p, ly);
        pc.mark("Y1 = Y P1");
        // B = Q^T Xc (src/pca.rs:681) = P2^T (Y1^T Xc):  C' = Xc^T Y1 (d x l) in one pass over X
        gemm_atb<T>(ctx, Y1.p, ly, l, nullptr, Y1.p, ly, l, nullptr, n, G2);
        PETAL_CUDA(cudaMemsetAsync(Cp, 0, (size_t)(d * l) * sizeof(double), ctx->stream));
        X.traverse([&](const T* Xc, int64_t r0, int64_t rows) {
            gemm_atb<T>(ctx, Xc, d, d, cm.mu, Y1.p + r0 * ly, ly, l, nullptr, rows, Cp, /*zero=*/false);
        });
        allreduce_sum(ctx, small.p, (size_t)(l * l + d * l + 1));
        pc.mark("G2, C' = Xc^T Y1");
        gram_to_orthonormalizer(ctx, G2, l, 1e-6, kGramNoise, P.p);  // P2 (l x l)
    }
    DBuf<double> M1(ctx, (size_t)(d * l)), Bm(ctx, (size_t)(l * d));
    gemm_xb<double>(ctx, Cp, l, d, l, P.p, l, false, l, nullptr, nullptr, M1.p, l);
    launch_transpose(ctx, M1.p, d, l, Bm.p);  // B (l x d)

    // SVD of B (src/pca.rs:682, gesdd).  B is l x d with l << d: first diagonalise the small Gram
    // B B^T = W diag(s^2) W^T (single-CTA Jacobi), rotate B' = W^T B (rows now orthogonal up to
    // eps * cond^2), then let the one-sided Jacobi on B' clean up (f64 API; it converges in a
    // sweep or two from there).  For the f32 API the Gram route is already far below the 1e-4
    // tolerance and B' is final.
    DBuf<double> Bout(ctx, (size_t)(l * d)), JtB(ctx, (size_t)(l * l)), sigB(ctx, (size_t)l), Vt(ctx, (size_t)(l * d));
    {
        DBuf<double> GB(ctx, (size_t)(l * l)), W1(ctx, (size_t)(l * l)), lamB(ctx, (size_t)l), Bp(ctx, (size_t)(l * d));
        gemm_xb<double>(ctx, Bm.p, d, l, d, Bm.p, d, true, l, nullptr, nullptr, GB.p, l);   // B B^T
        jacobi_rows(ctx, GB.p, l, l, nullptr, W1.p, lamB.p, kGramNoise);                               // rows of W1 = eigenvectors
        gemm_xb<double>(ctx, W1.p, l, l, l, Bm.p, d, false, d, nullptr, nullptr, Bp.p, d);  // B' = W1 B
        if (sizeof(T) == 4) {
            row_norm_kernel<<<(unsigned)l, 256, 0, ctx->stream>>>(Bp.p, (int)l, (int)d, sigB.p);
            launch1(ctx);
            PETAL_CUDA(cudaMemcpyAsync(Bout.p, Bp.p, (size_t)(l * d) * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
            PETAL_CUDA(cudaMemcpyAsync(JtB.p, W1.p, (size_t)(l * l) * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        } else {
            DBuf<double> J2(ctx, (size_t)(l * l));
            jacobi_rows(ctx, Bp.p, l, d, Bout.p, J2.p, sigB.p);
            gemm_xb<double>(ctx, J2.p, l, l, l, W1.p, l, false, l, nullptr, nullptr, JtB.p, l);  // Jt = J2 W1
        }
    }
    launch_normalize_rows(ctx, Bout.p, sigB.p, l, d, 0.0, Vt.p);
    pc.mark("small SVD of B");

    DBuf<T> comps_tmp;
    T* comps_dev = comps.p;
    if (!comps_dev) {
        comps_tmp.alloc(ctx, (size_t)(k * d));
        comps_dev = comps_tmp.p;
    }
    launch_cast<double, T>(ctx, Vt.p, comps_dev, k * d);  // components = vt[0..k] (src/pca.rs:544)

    if (k > 0) {
        // U * Sigma = Q * U_B * Sigma (src/pca.rs:683, transform_with_u) = Y1 * (P2 * U_B[:, :k] * Sigma_k)
        DBuf<double> S(ctx, (size_t)(l * k));
        gemm_xb<double>(ctx, P.p, l, l, l, JtB.p, l, true, k, nullptr, nullptr, S.p, k);
        scale_cols_kernel<<<(unsigned)ceil_div(l * k, 256), 256, 0, ctx->stream>>>(S.p, l, k, sigB.p);
        launch1(ctx);
        DBuf<T> scores_tmp;
        T* scores_dev = scores.p;
        DBuf<double> absmax3;
        bool done = false;
        if constexpr (sizeof(T) == 4) {
            if (panel) {
                // scores from the panels; the column |max| / row / sign triples of svd_flip come out of the same
                // kernel, and when the caller did not ask for the scores they are never written
                DBuf<float> Sf(ctx, (size_t)(l * k));
                launch_cast<double, float>(ctx, S.p, Sf.p, l * k);
                absmax3.alloc(ctx, (size_t)(k * 3));
                launch_panel_xb(ctx, Y.p, n, (int)ly, (int)l, Sf.p, (int)k, scores_dev, k, absmax3.p);
                done = true;
            }
        }
        if (!done) {
            if (!scores_dev) {
                scores_tmp.alloc(ctx, (size_t)(n * k));
                scores_dev = scores_tmp.p;
            }
            gemm_xb_b64<T>(ctx, Y1.p, ly, n, l, S.p, k, k, nullptr, scores_dev, k);
        }
        pc.mark("scores = Y1 S");
        flip_signs<T>(ctx, scores_dev, n, k, comps_dev, d, (bool)scores, absmax3.p);  // svd_flip, src/pca.rs:684
        pc.mark("svd_flip");
        if (sing) launch_cast<double, T>(ctx, sigB.p, sing.p, k);
    }
    if (mean) launch_cast<T, T>(ctx, cm.mean_t.p, mean.p, d);
    if (tv) launch_cast<double, T>(ctx, tvd, tv.p, 1);

    comps.commit(ctx); mean.commit(ctx); sing.commit(ctx); tv.commit(ctx); scores.commit(ctx);
    finish_call(ctx, comps.to_host || mean.to_host || sing.to_host || tv.to_host || scores.to_host);
}

// ---------------------------------------------------------------------------------------
// transform / inverse_transform (reference src/pca.rs:726-750, 788-811; src/ica.rs:120-131)
// ---------------------------------------------------------------------------------------
template <typename T>
void transform(petal_ctx* ctx, const T* x_user, int64_t n, int64_t d, const T* comps_user, int64_t k,
               const T* mean_user, T* out_user) {
    if (n < 0 || d < 0 || k < 0) invalid_input("negative dimension");
    if (n == 0 || k == 0) return;
    RowStream<T> X;
    X.open(ctx, x_user, n, d, (size_t)(n * k) * sizeof(T));
    DevIn<T> C(ctx, comps_user, (size_t)(k * d)), mu(ctx, mean_user, (size_t)d);
    DevOut<T> out(ctx, out_user, (size_t)(n * k));
    if (!out) invalid_input("output buffer is null");
    X.traverse([&](const T* Xc, int64_t r0, int64_t rows) {
        gemm_xb<T>(ctx, Xc, d, rows, d, C.p, d, true, k, mu.p, nullptr, out.p + r0 * k, k);
    });
    out.commit(ctx);
    finish_call(ctx, out.to_host);
}

template <typename T>
void inverse_transform(petal_ctx* ctx, const T* y_user, int64_t n, int64_t k, const T* comps_user, int64_t d,
                       const T* mean_user, T* out_user) {
    if (n < 0 || d < 0 || k < 0) invalid_input("negative dimension");
    if (n == 0 || d == 0) return;
    DevIn<T> Y(ctx, y_user, (size_t)(n * k)), C(ctx, comps_user, (size_t)(k * d)), mu(ctx, mean_user, (size_t)d);
    if (out_user == nullptr) invalid_input("output buffer is null");
    // host output (n x d, the large side here): computed in row chunks into a two-slot device buffer and drained by the
    // copy stream while the next chunk is computed - the reconstruction never has to fit in HBM
    int64_t rows_each = std::max<int64_t>(kMinChunkRows, ctx->host_chunk_bytes / std::max<int64_t>(d * (int64_t)sizeof(T), 1));
    if (const char* e = getenv("PETAL_HOST_CHUNK_BYTES")) rows_each = std::max<int64_t>(kMinChunkRows, atoll(e) / (d * (int64_t)sizeof(T)));
    rows_each = (rows_each / 32) * 32;
    if (!is_device_pointer(out_user) && n > rows_each + kMinChunkRows) {
        ensure_copy_stream(ctx);
        DBuf<T> slots(ctx, (size_t)(2 * (rows_each + kMinChunkRows) * d));
        bool used[2] = {false, false};
        int64_t r0 = 0;
        for (int64_t c = 0; r0 < n; ++c) {
            int64_t rows = std::min<int64_t>(rows_each, n - r0);
            if (n - (r0 + rows) < kMinChunkRows) rows = n - r0;
            const int slot = (int)(c & 1);
            T* dst = slots.p + (size_t)slot * (size_t)((rows_each + kMinChunkRows) * d);
            if (used[slot]) PETAL_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[2 + slot], 0));  // slot drained
            gemm_xb<T>(ctx, Y.p + r0 * k, k, rows, k, C.p, d, false, d, nullptr, mu.p, dst, d);
            PETAL_CUDA(cudaEventRecord(ctx->copy_ev[slot], ctx->stream));
            PETAL_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev[slot], 0));
            PETAL_CUDA(cudaMemcpyAsync(out_user + (size_t)r0 * (size_t)d, dst, (size_t)(rows * d) * sizeof(T), cudaMemcpyDeviceToHost,
                                       ctx->copy_stream));
            PETAL_CUDA(cudaEventRecord(ctx->copy_ev[2 + slot], ctx->copy_stream));
            used[slot] = true;
            r0 += rows;
        }
        PETAL_CUDA(cudaStreamSynchronize(ctx->copy_stream));
        PETAL_CUDA(cudaStreamSynchronize(ctx->stream));
        return;
    }
    DevOut<T> out(ctx, out_user, (size_t)(n * d));
    gemm_xb<T>(ctx, Y.p, k, n, k, C.p, d, false, d, nullptr, mu.p, out.p, d);
    out.commit(ctx);
    finish_call(ctx, out.to_host);
}

// ---------------------------------------------------------------------------------------
// FastICA
// ---------------------------------------------------------------------------------------
// symmetric_decorrelation (reference src/ica.rs:363-381): (W W^T)^-1/2 W = polar factor of W,
// from the one-sided Jacobi of W's rows:  Jt W = diag(s) N  ->  result = Jt^T N.
void symmetric_decorrelation(petal_ctx* ctx, const double* W, int64_t m, double* out) {
    DBuf<double> Aout(ctx, (size_t)(m * m)), Jt(ctx, (size_t)(m * m)), sig(ctx, (size_t)m), N(ctx, (size_t)(m * m));
    jacobi_rows(ctx, W, m, m, Aout.p, Jt.p, sig.p);
    launch_normalize_rows(ctx, Aout.p, sig.p, m, m, 0.0, N.p);
    gemm_atb<double>(ctx, Jt.p, m, m, nullptr, N.p, m, m, nullptr, m, out);
}


// ---------------------------------------------------------------------------------------
// Fused small-side update of one FastICA fixed-point iteration (single CTA, f64, nc <= kIcaFusedMax):
//   Gd = (H K1^T)/n - diag(mean g') W            (reference src/ica.rs:334-342)
//   W1 = (Gd Gd^T)^-1/2 Gd                        (symmetric_decorrelation, src/ica.rs:363-381) computed as
//        the polar factor of Gd by the Newton-Schulz iteration X <- X (3I - X^T X)/2, X0 = Gd / sqrt(|Gd|_1 |Gd|_inf)
//   lim = max_i | |<w1_i, w_i>| - 1 |             (src/ica.rs:344-354; variant 1: rows of w1 with columns of w)
//   W <- W1,  W~ = W1 K1 (cast to T) for the next streaming pass
// status[0] = 1 when Newton-Schulz did not converge (singular Gd): the host then redoes the iteration with the
// Jacobi-based path.
// ---------------------------------------------------------------------------------------
constexpr int kIcaFusedMax = 64;
constexpr int kIcaThreads = 256;
constexpr int kIcaLdPad = 4;  // row pitch nc + 4 doubles: conflict-free DMMA fragment loads (see smem_gemm_dmma)

// C = op(A) * B for nc x nc matrices in shared memory (row pitch ld).  256 threads as a 16 x 16 grid, each
// owning the interleaved 4 x 4 outputs (ty + 16u, tx + 16v): B reads are 16 consecutive words per half-warp
// (conflict-free), A reads are warp broadcasts.  TRANS_A: C = A^T B.  nc <= 64.
template <bool TRANS_A>
__device__ __forceinline__ void smem_gemm_k(const double* A, const double* B, double* C, int nc, int kdim, int ld) {
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    double acc[4][4] = {};
    for (int k = 0; k < kdim; ++k) {
        double a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = min(ty + 16 * u, nc - 1), j = min(tx + 16 * u, nc - 1);
            a[u] = TRANS_A ? A[k * ld + i] : A[i * ld + k];
            b[u] = B[k * ld + j];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[u][v] += a[u] * b[v];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v)
            if (ty + 16 * u < nc && tx + 16 * v < nc) C[(ty + 16 * u) * ld + tx + 16 * v] = acc[u][v];
}
// Same product on the FP64 tensor path (mma.sync.m8n8k4.f64) when nc is a multiple of 8: warp w owns the tile rows
// w, w + 8, ...; per k-step of 4 one A fragment and nc/8 B fragments.  With ld = nc + 4 (ld mod 16 in {4, 12}) every
// fragment load of a half-warp touches 16 distinct 8-byte bank pairs.  ~2x the DFMA version at nc = 64.
template <bool TRANS_A>
__device__ __forceinline__ void smem_gemm_dmma(const double* A, const double* B, double* C, int nc, int ld) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, kq = lane & 3, rq = lane >> 2;
    const int nt = nc >> 3;
    for (int ti = warp; ti < nt; ti += kIcaThreads / 32) {
        double acc[8][2];
#pragma unroll
        for (int tj = 0; tj < 8; ++tj) acc[tj][0] = acc[tj][1] = 0.0;
        for (int k4 = 0; k4 < (nc >> 2); ++k4) {
            const int k = k4 * 4 + kq;
            const double a = TRANS_A ? A[k * ld + ti * 8 + rq] : A[(ti * 8 + rq) * ld + k];
#pragma unroll
            for (int tj = 0; tj < 8; ++tj)
                if (tj < nt) dmma_m8n8k4(acc[tj][0], acc[tj][1], a, B[k * ld + tj * 8 + rq]);
        }
#pragma unroll
        for (int tj = 0; tj < 8; ++tj)
            if (tj < nt) {
                C[(ti * 8 + rq) * ld + tj * 8 + 2 * kq] = acc[tj][0];
                C[(ti * 8 + rq) * ld + tj * 8 + 2 * kq + 1] = acc[tj][1];
            }
    }
}

template <bool TRANS_A>
__device__ __forceinline__ void smem_gemm(const double* A, const double* B, double* C, int nc, int ld) {
    if ((nc & 7) == 0) smem_gemm_dmma<TRANS_A>(A, B, C, nc, ld);
    else smem_gemm_k<TRANS_A>(A, B, C, nc, nc, ld);
}

__device__ __forceinline__ double block_reduce_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    return s;
}
__device__ __forceinline__ double block_reduce_max(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s = fmax(s, red[w]);
    return s;
}

template <typename T>
__global__ void __launch_bounds__(kIcaThreads)
ica_update_kernel(const double* __restrict__ Ht /* d x nc */, const double* __restrict__ gp, double* __restrict__ W,
                  const double* __restrict__ K1 /* nc x d or null */, int nc, int d, double inv_n, int lim_variant,
                  T* __restrict__ Wt_out /* nc x d */, double* __restrict__ out2 /* state, see below */, double tol,
                  float* __restrict__ whi, float* __restrict__ wlo /* [64][64] operand arrays of the one-pass kernel or null */,
                  double* __restrict__ consumed, int consumed_n /* zeroed after a successful update (Ht | gp) */) {
    // state: [0] lim, [1] 1 = Newton-Schulz failed, [2] NS steps, [3..5] cycle counters, [6] done (1 converged,
    // 2 failed: the host redoes this iteration with the Jacobi path), [7] completed fixed-point iterations
    if (out2[6] != 0.0) return;
    extern __shared__ double sm[];
    const int brows = max(nc, d);  // the staging of H^T / K1^T uses d rows, that of K1 d columns
    const int ld = brows + kIcaLdPad;
    double* X = sm;
    double* Tm = sm + (size_t)brows * ld;
    double* Y = sm + 2 * (size_t)brows * ld;
    __shared__ double red[kIcaThreads / 32];
    const int tid = threadIdx.x;
    const long long c0 = clock64();

    // Gd -> X.  With whitening: HK = H K1^T = (Ht)^T (K1^T): stage Ht (d x nc) in Tm and K1^T (d x nc) in Y
    // (both need d <= nc-sized buffers: d <= kIcaFusedMax is checked by the host) and multiply in shared memory.
    if (Ht == nullptr) {
        // initial symmetric decorrelation of w_init (src/ica.rs:329): polar factor of W itself
        for (int e = tid; e < nc * nc; e += kIcaThreads) X[(e / nc) * ld + e % nc] = W[e];
    } else if (K1) {
        for (int e = tid; e < d * nc; e += kIcaThreads) {
            const int f = e / nc, i = e % nc;
            Tm[f * ld + i] = Ht[e];                    // Ht[f][i]
            Y[f * ld + i] = K1[(size_t)i * d + f];     // K1^T[f][i]
        }
        __syncthreads();
        smem_gemm_k<true>(Tm, Y, X, nc, d, ld);        // X = Ht^T K1^T  (nc x nc), reduction over d
        __syncthreads();
        for (int e = tid; e < nc * nc; e += kIcaThreads) {
            const int i = e / nc, j = e % nc;
            X[i * ld + j] = X[i * ld + j] * inv_n - gp[i] * inv_n * W[e];
        }
    } else {
        for (int e = tid; e < nc * nc; e += kIcaThreads) {
            const int i = e / nc, j = e % nc;
            X[i * ld + j] = Ht[(size_t)j * nc + i] * inv_n - gp[i] * inv_n * W[e];
        }
    }
    __syncthreads();
    const long long c1 = clock64();
    // scale by sqrt(|Gd|_1 |Gd|_inf) >= |Gd|_2
    double rs = 0.0, cs = 0.0;
    for (int i = tid; i < nc; i += kIcaThreads) {
        double r = 0.0, c = 0.0;
        for (int j = 0; j < nc; ++j) {
            r += fabs(X[i * ld + j]);
            c += fabs(X[j * ld + i]);
        }
        rs = fmax(rs, r);
        cs = fmax(cs, c);
    }
    rs = block_reduce_max(rs, red);
    cs = block_reduce_max(cs, red);
    const double scale = sqrt(rs * cs);
    bool ok = (scale > 0.0) && isfinite(scale);
    if (ok) {
        const double inv = 1.0 / scale;
        for (int e = tid; e < nc * nc; e += kIcaThreads) X[(e / nc) * ld + e % nc] *= inv;
    }
    __syncthreads();
    // Newton-Schulz with the optimal per-step scaling for singular values in [l, 1]:
    //   X <- X (1.5 a I - 0.5 a^3 X^T X),  a^2 = 3 / (1 + l + l^2),  l <- 1.5 a l - 0.5 (a l)^3
    // (a = 1 is the plain iteration; a -> sqrt(3) while l is small grows the small singular values 2.6x per
    // step instead of 1.5x).  l is only a guess of the smallest singular value: values below it still grow by
    // >= 1.5x per step, and once l reaches 1 the iteration is the unscaled, quadratically convergent one.
    // l0 = 1e-3: the 1-norm/inf-norm scaling above over-estimates |Gd|_2 by ~5x for the Gd of a FastICA run, whose
    // own conditioning is 1e-2 .. 0.7, so the scaled singular values start at 1e-3 .. 0.1 (12 steps instead of 20).
    bool converged = false;
    int ns_it = 0;
    double lo = 1e-3;
    for (int it = 0; ok && it < 100; ++it) {
        ns_it = it;
        smem_gemm<true>(X, X, Tm, nc, ld);  // T = X^T X
        __syncthreads();
        const double alpha = (lo < 1.0 - 1e-9) ? sqrt(3.0 / (1.0 + lo + lo * lo)) : 1.0;
        const double alpha3 = alpha * alpha * alpha;
        lo = fmin(1.0, 1.5 * alpha * lo - 0.5 * alpha3 * lo * lo * lo);
        double err = 0.0;
        for (int e = tid; e < nc * nc; e += kIcaThreads) {
            const int i = e / nc, j = e % nc;
            const double t = Tm[i * ld + j];
            const double dlt = t - (i == j ? 1.0 : 0.0);
            err += dlt * dlt;
            Tm[i * ld + j] = (i == j ? 1.5 * alpha : 0.0) - 0.5 * alpha3 * t;
        }
        err = block_reduce_sum(err, red);
        if (!(err == err)) {
            ok = false;
            break;
        }
        if (err < 1e-26) {
            converged = true;
            break;
        }
        smem_gemm<false>(X, Tm, Y, nc, ld);  // Y = X (1.5 a I - 0.5 a^3 T)
        __syncthreads();
        double* t = X;
        X = Y;
        Y = t;
    }
    if (!converged) {
        if (tid == 0) {
            out2[0] = 0.0;
            out2[1] = 1.0;
            out2[6] = 2.0;
        }
        return;
    }
    const long long c2 = clock64();
    // lim against the previous W (still in global memory)
    double best = 0.0;
    for (int i = tid; i < nc; i += kIcaThreads) {
        double sdot = 0.0;
        for (int j = 0; j < nc; ++j) sdot += X[i * ld + j] * (lim_variant ? W[(size_t)j * nc + i] : W[(size_t)i * nc + j]);
        best = fmax(best, fabs(fabs(sdot) - 1.0));
    }
    best = block_reduce_max(best, red);
    __syncthreads();
    for (int e = tid; e < nc * nc; e += kIcaThreads) W[e] = X[(e / nc) * ld + e % nc];
    // W~ = W1 K1 for the next pass
    if (K1) {
        __syncthreads();
        for (int e = tid; e < nc * d; e += kIcaThreads) Tm[(e / d) * ld + e % d] = K1[e];  // K1 (nc x d)
        __syncthreads();
        // Y (nc x d) = X (nc x nc) * K1 (nc x d): output columns d <= 64 handled by the same 16 x 16 grid
        {
            const int ty = tid >> 4, tx = tid & 15;
            double acc[4][4] = {};
            for (int k = 0; k < nc; ++k) {
                double a[4], b[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    a[u] = X[min(ty + 16 * u, nc - 1) * ld + k];
                    b[u] = Tm[k * ld + min(tx + 16 * u, d - 1)];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int v = 0; v < 4; ++v) acc[u][v] += a[u] * b[v];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v)
                    if (ty + 16 * u < nc && tx + 16 * v < d) {
                        const int c = ty + 16 * u, k = tx + 16 * v;
                        Wt_out[(size_t)c * d + k] = (T)acc[u][v];
                        if (whi != nullptr) {
                            const float hf = __uint_as_float(__float_as_uint((float)acc[u][v]) & 0xFFFFE000u);
                            whi[c * 64 + k] = hf;
                            wlo[c * 64 + k] = (float)(acc[u][v] - (double)hf);
                        }
                    }
        }
    } else {
        for (int e = tid; e < nc * d; e += kIcaThreads) {
            const double w = X[(e / d) * ld + e % d];
            Wt_out[e] = (T)w;
            if (whi != nullptr) {
                const float hf = __uint_as_float(__float_as_uint((float)w) & 0xFFFFE000u);
                whi[(e / d) * 64 + e % d] = hf;
                wlo[(e / d) * 64 + e % d] = (float)(w - (double)hf);
            }
        }
    }
    if (consumed != nullptr)
        for (int e = tid; e < consumed_n; e += kIcaThreads) consumed[e] = 0.0;
    if (tid == 0) {
        if (Ht != nullptr) {
            out2[7] += 1.0;
            if (best < tol) out2[6] = 1.0;
        }
        out2[0] = best;
        out2[1] = 0.0;
        out2[2] = (double)ns_it;
        out2[3] = (double)(c1 - c0);
        out2[4] = (double)(c2 - c1);
        out2[5] = (double)(clock64() - c2);
    }
}

// One-pass FastICA streaming kernel (ica_kernels.cuh): f32 only, d and nc <= 64, tcgen05 engine selected.
template <typename T>
bool ica_one_pass_supported(petal_ctx*, const T*, int64_t, int64_t, int64_t, int64_t) { return false; }
template <>
bool ica_one_pass_supported<float>(petal_ctx* ctx, const float* X, int64_t ld, int64_t n, int64_t d, int64_t nc) {
    if (const char* e = getenv("PETAL_ICA_ONEPASS")) if (e[0] == '0') return false;
    return ctx->f32_engine == 1 && ica::fused_supported(X, ld, n, d, nc);
}
template <typename T>
struct IcaOnePass {  // f64: never selected
    void init(petal_ctx*, const T*, int64_t, int64_t, int64_t, const T*, int64_t, int, double*, double*, const double*) {
        linalg_error("one-pass FastICA kernel is f32 only");
    }
    void set_w(petal_ctx*, const T*) {}
    void run(petal_ctx*) {}
    float* whi_ptr() { return nullptr; }
    float* wlo_ptr() { return nullptr; }
};
template <>
struct IcaOnePass<float> {
    ica::IcaFused f;
    bool on = false;
    void init(petal_ctx* ctx, const float* X, int64_t ld, int64_t n, int64_t d, const float* mu, int64_t nc, int fun, double* Ht,
              double* gp, const double* state) {
        f.init(ctx, X, ld, n, d, mu, nc, fun, Ht, gp, state);
        on = true;
    }
    void set_w(petal_ctx* ctx, const float* Wt) { f.set_w(ctx, Wt); }
    void run(petal_ctx* ctx) { f.run(ctx); }
    float* whi_ptr() { return on ? f.whi.p : nullptr; }
    float* wlo_ptr() { return on ? f.wlo.p : nullptr; }
};

// ica_par (reference src/ica.rs:319-361) on data X[n x d] with whitening folded in:
// the whitened sample is x1 = K1 (x - mu) with K1 = sqrt(n) K (nc x d); K1 == nullptr means the
// data is already white (d == nc).  Returns W (nc x nc, f64, device) and the iteration count.
template <typename T>
void ica_par(petal_ctx* ctx, RowStream<T>& Xs, int64_t d, int64_t n_total, const T* mu, const double* K1,
             int64_t nc, int fun, double tol, int64_t max_iter, int lim_variant, const double* w_init, double* W,
             int64_t* n_iter_out, double* lim_out, bool one_pass_all_ranks = true) {
    const int64_t n = Xs.n;
    Xs.load();  // a host X that fits becomes resident here (the fit flow has already loaded it with its first pass)
    // out-of-core X (ring): every fixed-point iteration re-streams the rows through the generic three-kernel pass,
    // chunk by chunk (PCIe-bound; the one-pass kernel addresses all of X through one tensor map)
    const T* X = Xs.ring ? nullptr : Xs.resident_ptr();
    if (fun != PETAL_ICA_LOGCOSH && fun != PETAL_ICA_EXP && fun != PETAL_ICA_CUBE) invalid_input("unknown contrast function");
    DBuf<double> Wk(ctx, (size_t)(nc * d)), Hg(ctx, (size_t)(nc * d)), Htg(ctx, (size_t)(nc * d + nc)), HK(ctx, (size_t)(nc * nc)),
        Gd(ctx, (size_t)(nc * nc)), W1(ctx, (size_t)(nc * nc)), limd(ctx, 1);
    // the one-pass kernel changes how often the host polls (and with it the number of collectives per batch):
    // taken only when every rank can run it on its shard
    const bool one_pass = one_pass_all_ranks && !Xs.ring && ica_one_pass_supported<T>(ctx, X, d, n, d, nc);
    if (one_pass_all_ranks && !one_pass && ctx->world > 1 && ica_one_pass_supported<T>(ctx, nullptr, d, n, d, nc))
        linalg_error("inconsistent one-pass decision across ranks");
    const int64_t u_rows = Xs.ring ? Xs.slot_rows : n;
    DBuf<T> Wt(ctx, (size_t)(nc * d)), U(ctx, one_pass ? (size_t)1 : (size_t)(u_rows * nc));
    double* H = Hg.p;
    const int htg_n = (int)(nc * d + nc);
    // Htg = this rank's partial [H^T (d x nc) | sum g' (nc)]; with several ranks the sum over ranks goes to a second
    // buffer, so that re-issuing the all-reduce with an unconsumed partial (launches of a batch after the device-side
    // stop flag was raised) cannot scale the sums
    DBuf<double> Hred(ctx, ctx->world > 1 ? (size_t)htg_n : 0);
    double* Ht_loc = Htg.p;
    double* gp_loc = Htg.p + nc * d;
    double* Ht = ctx->world > 1 ? Hred.p : Htg.p;
    double* gp = Ht + nc * d;
    const double inv_n = 1.0 / (double)n_total;
    int64_t iters = max_iter;
    double lim = 0.0;
    const bool fused = (nc <= kIcaFusedMax) && (d <= kIcaFusedMax);
    DBuf<double> state(ctx, 8);
    state.zero();
    Htg.zero();
    const size_t upd_smem = 3 * (size_t)std::max(nc, d) * (std::max(nc, d) + kIcaLdPad) * sizeof(double);
    if (fused)
        ensure_dynamic_smem(ctx, ica_update_kernel<T>, 3 * kIcaFusedMax * (kIcaFusedMax + kIcaLdPad) * sizeof(double));
    IcaOnePass<T> pass;
    if (one_pass) pass.init(ctx, X, d, n, d, mu, nc, fun, Ht_loc, gp_loc, state.p);
    auto make_wt = [&]() {
        // W~ = W K1 so that W x1 = W~ (x - mu): the whitened copy is never materialised
        const double* Wfull = W;
        if (K1) {
            gemm_xb<double>(ctx, W, nc, nc, nc, K1, d, false, d, nullptr, nullptr, Wk.p, d);
            Wfull = Wk.p;
        }
        launch_cast<double, T>(ctx, Wfull, Wt.p, nc * d);
        if (one_pass) pass.set_w(ctx, Wt.p);
    };
    auto read_state = [&](double* h) {
        PETAL_CUDA(cudaMemcpyAsync(h, state.p, 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        PETAL_CUDA(cudaStreamSynchronize(ctx->stream));
    };
    auto launch_update = [&](bool init) {
        KTimer kt(ctx, "ica_update", 0.0);
        ica_update_kernel<T><<<1, kIcaThreads, upd_smem, ctx->stream>>>(init ? nullptr : Ht, init ? nullptr : gp, W, K1, (int)nc,
                                                                       (int)d, inv_n, lim_variant, Wt.p, state.p, tol,
                                                                       pass.whi_ptr(), pass.wlo_ptr(), init ? nullptr : Htg.p, htg_n);
        launch1(ctx);
    };
    // W = symmetric_decorrelation(w_init) (src/ica.rs:329) and the first W~
    bool init_done = false;
    double hs[8];
    if (fused) {
        PETAL_CUDA(cudaMemcpyAsync(W, w_init, (size_t)(nc * nc) * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        launch_update(true);
        read_state(hs);
        init_done = (hs[6] == 0.0);
        if (!init_done) state.zero();
    }
    if (!init_done) {
        symmetric_decorrelation(ctx, w_init, nc, W);
        make_wt();
    }
    // streaming half of one fixed-point iteration: Ht, gp (summed over ranks)
    auto stream_pass = [&]() {
        if (one_pass) {
            // U, g(U), sum g'(U) and H^T in a single pass over X (tcgen05; U and g(U) never reach HBM)
            pass.run(ctx);
        } else {
            PETAL_CUDA(cudaMemsetAsync(Htg.p, 0, (size_t)htg_n * sizeof(double), ctx->stream));
            Xs.traverse([&](const T* Xc, int64_t, int64_t rows) {
                // U = (X - mu) W~^T  (n x nc)    [w.dot(input), src/ica.rs:332]
                gemm_xb<T>(ctx, Xc, d, rows, d, Wt.p, d, true, nc, mu, nullptr, U.p, nc);
                // g(U) in place and sum of g'(U) per component  [logcosh, src/ica.rs:383-398]
                launch_nonlin<T>(ctx, U.p, rows, nc, nc, fun, gp_loc);
                // H = g(U)^T (X - mu)  (nc x d)   [gwtx.dot(input.t()), src/ica.rs:333, before whitening],
                // computed as H^T = (X - mu)^T g(U) so that the pass runs on the X^T*Y engine (tcgen05 for f32)
                gemm_atb<T>(ctx, Xc, d, d, mu, U.p, nc, nc, nullptr, rows, Ht_loc, /*zero=*/false);
            });
        }
        if (ctx->world > 1) allreduce_sum_to(ctx, Htg.p, Hred.p, (size_t)htg_n);
    };
    // small half on the Jacobi path (any nc; also the fallback when Newton-Schulz hits a singular Gd)
    auto slow_update = [&]() {
        launch_transpose(ctx, Ht, d, nc, H);
        // Gd = (H K1^T) / n - diag(mean g') W   (src/ica.rs:334-342)
        const double* HKp = H;
        if (K1) {
            gemm_xb<double>(ctx, H, d, nc, d, K1, d, true, nc, nullptr, nullptr, HK.p, nc);
            HKp = HK.p;
        }
        ica_gd_kernel<<<(unsigned)ceil_div(nc * nc, 256), 256, 0, ctx->stream>>>(HKp, gp, W, nc, inv_n, Gd.p);
        launch1(ctx);
        symmetric_decorrelation(ctx, Gd.p, nc, W1.p);  // src/ica.rs:343
        ica_lim_kernel<<<1, 256, 0, ctx->stream>>>(W1.p, W, (int)nc, lim_variant, limd.p);
        launch1(ctx);
        PETAL_CUDA(cudaMemcpyAsync(W, W1.p, (size_t)(nc * nc) * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        PETAL_CUDA(cudaMemcpyAsync(&lim, limd.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        PETAL_CUDA(cudaStreamSynchronize(ctx->stream));
        make_wt();
        Htg.zero();
    };
    // The convergence test (src/ica.rs:344-357) runs on the device: the update kernel raises state[6] when
    // lim < tol and every later launch of the batch returns at once, so the host only looks every `batch` iterations.
    const int64_t batch = (fused && one_pass) ? 4 : 1;
    int64_t it = 0;
    while (it < max_iter) {
        if (!fused) {
            stream_pass();
            slow_update();
            ++it;
            if (lim < tol) {  // src/ica.rs:355-357
                iters = it;
                break;
            }
            continue;
        }
        const int64_t b = std::min<int64_t>(batch, max_iter - it);
        for (int64_t j = 0; j < b; ++j) {
            stream_pass();
            launch_update(false);
        }
        read_state(hs);
        if (getenv("PETAL_PHASES"))
            fprintf(stderr, "[ica_update] iters %.0f done %.0f lim %.3e ns_steps %.0f cycles: setup %.0f ns %.0f tail %.0f\n", hs[7],
                    hs[6], hs[0], hs[2], hs[3], hs[4], hs[5]);
        it = (int64_t)hs[7];
        lim = hs[0];
        if (hs[6] == 1.0) {
            iters = it;
            break;
        }
        if (hs[6] == 2.0) {
            // singular Gd: redo this iteration's small half with the Jacobi path (Ht / gp are still intact)
            slow_update();
            ++it;
            const double st[2] = {lim < tol ? 1.0 : 0.0, (double)it};
            PETAL_CUDA(cudaMemcpyAsync(state.p + 6, st, 2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            PETAL_CUDA(cudaStreamSynchronize(ctx->stream));
            if (lim < tol) {
                iters = it;
                break;
            }
        }
    }
    *n_iter_out = iters;
    if (lim_out) *lim_out = lim;
}

// Deflation FastICA (ica_deflation.cuh; sklearn `_ica_def`) on data X[n x d] with the whitening folded in like ica_par:
// x1 = K1 (x - mu).  Returns W (nc x nc, f64, device) and the largest iteration count over the components.
template <typename T>
void ica_defl(petal_ctx* ctx, RowStream<T>& Xs, int64_t d, int64_t n_total, const T* mu, const double* K1, int64_t nc,
              int fun, double tol, int64_t max_iter, const double* w_init, double* W, int64_t* n_iter_out, double* lim_out) {
    if (fun != PETAL_ICA_LOGCOSH && fun != PETAL_ICA_EXP && fun != PETAL_ICA_CUBE) invalid_input("unknown contrast function");
    if (nc > 4096) invalid_input("deflation FastICA supports at most 4096 components");
    const bool fused_pass = defl::pass_supported<T>(d);
    DBuf<double> hacc(ctx, (size_t)d + 1), w(ctx, (size_t)nc), state(ctx, 4);
    DBuf<T> wt(ctx, (size_t)d), U;
    if (!fused_pass) U.alloc(ctx, (size_t)(Xs.ring ? Xs.slot_rows : Xs.n));
    state.zero();
    PETAL_CUDA(cudaMemsetAsync(W, 0, (size_t)(nc * nc) * sizeof(double), ctx->stream));
    const double inv_n = 1.0 / (double)n_total;
    const size_t upd_smem = (size_t)(2 * nc + defl::kThreads) * sizeof(double);
    ensure_dynamic_smem(ctx, defl::defl_update_kernel<T>, upd_smem);
    // iterations queued between two looks at the device-side state (the generic engines do not read the stop flag)
    const int64_t batch = fused_pass ? 4 : 1;
    double hs[4] = {0, 0, 0, 0};
    double worst_lim = 0.0;
    for (int64_t j = 0; j < nc; ++j) {
        defl::defl_init_kernel<T><<<1, defl::kThreads, 0, ctx->stream>>>(w_init, (int)j, (int)nc, (int)d, K1, w.p, wt.p, state.p, hacc.p);
        launch1(ctx);
        int64_t it = 0;
        while (it < max_iter) {
            const int64_t b = std::min<int64_t>(batch, max_iter - it);
            for (int64_t t = 0; t < b; ++t) {
                Xs.traverse([&](const T* Xc, int64_t, int64_t rows) {
                    if (fused_pass) {
                        defl::launch_pass<T>(ctx, Xc, rows, d, d, mu, wt.p, fun, hacc.p, state.p);
                    } else {
                        // wide rows: u = (X - mu) w~ (one column), g(u) in place, h += (X - mu)^T g(u) on the generic engines
                        gemm_xb<T>(ctx, Xc, d, rows, d, wt.p, d, true, 1, mu, nullptr, U.p, 1);
                        launch_nonlin<T>(ctx, U.p, rows, 1, 1, fun, hacc.p + d);
                        gemm_atb<T>(ctx, Xc, d, d, mu, U.p, 1, 1, nullptr, rows, hacc.p, /*zero=*/false);
                    }
                });
                allreduce_sum(ctx, hacc.p, (size_t)d + 1);
                {
                    KTimer kt(ctx, "ica_defl_update", 0.0);
                    defl::defl_update_kernel<T><<<1, defl::kThreads, upd_smem, ctx->stream>>>(hacc.p, K1, W, (int)j, (int)nc, (int)d, inv_n,
                                                                                            tol, w.p, wt.p, state.p);
                    launch1(ctx);
                }
            }
            PETAL_CUDA(cudaMemcpyAsync(hs, state.p, sizeof hs, cudaMemcpyDeviceToHost, ctx->stream));
            PETAL_CUDA(cudaStreamSynchronize(ctx->stream));
            it = (int64_t)hs[1];
            if (hs[2] != 0.0) break;
        }
        worst_lim = std::max(worst_lim, hs[0]);
        defl::defl_commit_kernel<<<1, 256, 0, ctx->stream>>>(w.p, W, (int)j, (int)nc, state.p);
        launch1(ctx);
    }
    PETAL_CUDA(cudaMemcpyAsync(hs, state.p, sizeof hs, cudaMemcpyDeviceToHost, ctx->stream));
    PETAL_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_iter_out = (int64_t)hs[3];
    if (lim_out) *lim_out = worst_lim;
}

template <typename T>
void fastica_fit(petal_ctx* ctx, const T* x_user, int64_t n, int64_t d, int fun, double tol, int64_t max_iter,
                 int lim_variant, const T* w_init_user, T* comps_u, T* mean_u, int64_t* n_iter_u, double* lim_u,
                 T* sources_u, bool deflation = false) {
    if (n < 0 || d < 0 || max_iter < 0) invalid_input("negative dimension");
    RowStream<T> X;
    X.open(ctx, x_user, n, d, (size_t)(2 * n * std::min<int64_t>(d, n)) * sizeof(T) + (size_t)(8 * d * d) * sizeof(double));
    const bool one_pass_local = !deflation && !X.ring && aligned_on_device(x_user) &&
                                ica_one_pass_supported<T>(ctx, nullptr, d, n, d, std::min<int64_t>(d, 64));
    const GlobalInfo ginfo = global_info(ctx, n, one_pass_local);
    const int64_t n_total = ginfo.n_total;
    if (n_total == 0 || d == 0) return;  // src/ica.rs:174-176
    const int64_t nc = std::min<int64_t>(n_total, d);  // src/ica.rs:173
    if (w_init_user == nullptr) invalid_input("w_init (nc x nc) is required");

    DevIn<T> Winit(ctx, w_init_user, (size_t)(nc * nc));
    DevOut<T> comps(ctx, comps_u, (size_t)(nc * d)), mean(ctx, mean_u, (size_t)d), sources(ctx, sources_u, (size_t)(n * nc));

    // whitening (src/ica.rs:189-208): the reference takes U, sigma from gesvd of the d x n centred
    // matrix; the same U, sigma^2 are the eigenpairs of the d x d Gram Xc^T Xc.
    ColMean<T> cm;
    DBuf<double> G(ctx, (size_t)(d * d)), Jt(ctx, (size_t)(d * d)), lam(ctx, (size_t)d);
    mean_and_gram<T>(ctx, X, d, n_total, true, cm, G.p);
    jacobi_rows(ctx, G.p, d, d, nullptr, Jt.p, lam.p, kGramNoise);
    DBuf<double> K(ctx, (size_t)(nc * d)), K1(ctx, (size_t)(nc * d));
    const double wcut = 64.0 * 2.220446049250313e-16;  // relative eigenvalue floor of an f64-accumulated Gram matrix
    whitening_kernel<<<(unsigned)ceil_div(nc * d, 256), 256, 0, ctx->stream>>>(Jt.p, lam.p, nc, d, 1.0, wcut, K.p);
    launch1(ctx);
    whitening_kernel<<<(unsigned)ceil_div(nc * d, 256), 256, 0, ctx->stream>>>(Jt.p, lam.p, nc, d,
                                                                              std::sqrt((double)n_total), wcut, K1.p);
    launch1(ctx);

    DBuf<double> Wd(ctx, (size_t)(nc * nc)), Winit_d(ctx, (size_t)(nc * nc));
    launch_cast<T, double>(ctx, Winit.p, Winit_d.p, nc * nc);
    int64_t iters = 0;
    double lim = 0.0;
    if (deflation)
        ica_defl<T>(ctx, X, d, n_total, cm.mu, K1.p, nc, fun, tol, max_iter, Winit_d.p, Wd.p, &iters, &lim);
    else
        ica_par<T>(ctx, X, d, n_total, cm.mu, K1.p, nc, fun, tol, max_iter, lim_variant, Winit_d.p, Wd.p, &iters, &lim,
                   ginfo.cap[0]);

    // components = W K (src/ica.rs:217)
    DBuf<double> Cd(ctx, (size_t)(nc * d));
    gemm_xb<double>(ctx, Wd.p, nc, nc, nc, K.p, d, false, d, nullptr, nullptr, Cd.p, d);
    DBuf<T> comps_tmp;
    T* comps_dev = comps.p;
    if (!comps_dev) {
        comps_tmp.alloc(ctx, (size_t)(nc * d));
        comps_dev = comps_tmp.p;
    }
    launch_cast<double, T>(ctx, Cd.p, comps_dev, nc * d);
    if (sources)  // fit_transform = (components * xc)^T (src/ica.rs:155-156)
        X.traverse([&](const T* Xc, int64_t r0, int64_t rows) {
            gemm_xb<T>(ctx, Xc, d, rows, d, comps_dev, d, true, nc, cm.mu, nullptr, sources.p + r0 * nc, nc);
        });
    if (mean) launch_cast<T, T>(ctx, cm.mean_t.p, mean.p, d);
    if (n_iter_u) *n_iter_u = iters;
    if (lim_u) *lim_u = lim;
    comps.commit(ctx); mean.commit(ctx); sources.commit(ctx);
    finish_call(ctx, comps.to_host || mean.to_host || sources.to_host);
}

template <typename T>
void ica_nonlin(petal_ctx* ctx, T* u_user, int64_t n, int64_t nc, int fun, int engine, double* gsum_user) {
    if (n <= 0 || nc <= 0) invalid_input("empty input");
    if (fun != PETAL_ICA_LOGCOSH && fun != PETAL_ICA_EXP && fun != PETAL_ICA_CUBE) invalid_input("unknown contrast function");
    DevIn<T> Uin(ctx, u_user, (size_t)(n * nc));
    DevOut<T> U(ctx, u_user, (size_t)(n * nc));
    DevOut<double> gs(ctx, gsum_user, (size_t)nc);
    if (!U || !gs) invalid_input("output buffer is null");
    if (U.p != Uin.p) PETAL_CUDA(cudaMemcpyAsync(U.p, Uin.p, (size_t)(n * nc) * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
    PETAL_CUDA(cudaMemsetAsync(gs.p, 0, (size_t)nc * sizeof(double), ctx->stream));
    if (engine == 0) {
        launch_nonlin<T>(ctx, U.p, n, nc, nc, fun, gs.p);
    } else {
        if constexpr (sizeof(T) == 4) {
            const unsigned blocks = (unsigned)ceil_div(n * nc, 256);
            if (fun == PETAL_ICA_LOGCOSH) ica::ica_g_probe_kernel<PETAL_ICA_LOGCOSH><<<blocks, 256, 0, ctx->stream>>>(U.p, n, nc, gs.p);
            else if (fun == PETAL_ICA_EXP) ica::ica_g_probe_kernel<PETAL_ICA_EXP><<<blocks, 256, 0, ctx->stream>>>(U.p, n, nc, gs.p);
            else ica::ica_g_probe_kernel<PETAL_ICA_CUBE><<<blocks, 256, 0, ctx->stream>>>(U.p, n, nc, gs.p);
            launch1(ctx);
        } else {
            invalid_input("engine 1 (one-pass kernel epilogue) is f32 only");
        }
    }
    U.commit(ctx);
    gs.commit(ctx);
    finish_call(ctx, U.to_host || gs.to_host);
}

}  // namespace

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

int petal_ctx_create(int device, petal_ctx** out) {
    if (!out) return PETAL_INVALID_INPUT;
    *out = nullptr;
    try {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
            throw Error{PETAL_LINALG_ERROR, std::string("no CUDA device available (libpetal_b200 has no CPU fallback): ") +
                                                cudaGetErrorString(e)};
        if (device < 0 || device >= count) throw Error{PETAL_INVALID_INPUT, "invalid device ordinal"};
        PETAL_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        PETAL_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10)
            throw Error{PETAL_LINALG_ERROR, std::string("libpetal_b200 is built for sm_100a only; device is sm_") +
                                                std::to_string(prop.major) + std::to_string(prop.minor)};
        petal_ctx* ctx = new petal_ctx;
        ctx->device = device;
        ctx->sm_count = prop.multiProcessorCount;
        PETAL_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        PETAL_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->dev_status), sizeof(int)));
        PETAL_CUDA(cudaMemset(ctx->dev_status, 0, sizeof(int)));
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            // keep freed workspaces in the stream-ordered pool: a fit re-uses multi-GB buffers (Y, scores)
            // every call and returning them to the driver costs tens of ms (petal_ctx_trim releases them)
            uint64_t thresh = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh);
        }
        *out = ctx;
        return PETAL_OK;
    } catch (const Error& e) {
        std::lock_guard<std::mutex> lk(g_global_mutex);
        g_global_error = e.msg;
        cudaGetLastError();
        return e.code;
    }
}

void petal_ctx_destroy(petal_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    delete ctx->comm;
    if (ctx->dev_status) cudaFree(ctx->dev_status);
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        for (auto& e : ctx->copy_ev) if (e) cudaEventDestroy(e);
        cudaStreamDestroy(ctx->copy_stream);
    }
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* petal_last_error(const petal_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }
const char* petal_last_global_error(void) { return g_global_error.c_str(); }

int petal_ctx_set_stream(petal_ctx* ctx, void* cuda_stream) {
    return guarded(ctx, [&] {
        PETAL_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
        ctx->stream = static_cast<cudaStream_t>(cuda_stream);
        ctx->own_stream = false;
    });
}

int petal_ctx_trim(petal_ctx* ctx) {
    return guarded(ctx, [&] {
        PETAL_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaMemPool_t pool;
        PETAL_CUDA(cudaDeviceGetDefaultMemPool(&pool, ctx->device));
        PETAL_CUDA(cudaMemPoolTrimTo(pool, 0));
    });
}

int petal_ctx_synchronize(petal_ctx* ctx) {
    return guarded(ctx, [&] { PETAL_CUDA(cudaStreamSynchronize(ctx->stream)); });
}

int64_t petal_ctx_launch_count(const petal_ctx* ctx) { return ctx ? ctx->launches : 0; }

int petal_ctx_set_f32_engine(petal_ctx* ctx, int engine) {
    if (!ctx) return -1;
    if (engine >= 0) ctx->f32_engine = engine ? 1 : 0;
    return ctx->f32_engine;
}

int petal_ctx_set_f64_engine(petal_ctx* ctx, int engine) {
    if (!ctx) return -1;
    if (engine >= 0) ctx->f64_engine = engine ? 1 : 0;
    return ctx->f64_engine;
}

int petal_ctx_set_host_staging(petal_ctx* ctx, int mode, int64_t chunk_bytes) {
    if (!ctx) return -1;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (mode >= 0 && mode <= 2) ctx->host_staging = mode;
    if (chunk_bytes > 0) ctx->host_chunk_bytes = chunk_bytes;
    return ctx->host_staging;
}

int petal_ctx_set_host_gram(petal_ctx* ctx, int enable) {
    if (!ctx) return -1;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (enable >= 0) ctx->host_gram = enable ? 1 : 0;
    return ctx->host_gram;
}

int petal_ctx_host_stream_stats(const petal_ctx* ctx, int64_t* h2d_bytes, int64_t* traversals, int* ring) {
    if (!ctx) return PETAL_INVALID_INPUT;
    if (h2d_bytes) *h2d_bytes = ctx->last_h2d_bytes;
    if (traversals) *traversals = ctx->last_traversals;
    if (ring) *ring = ctx->last_ring;
    return PETAL_OK;
}

int petal_ctx_set_profiling(petal_ctx* ctx, int enable) {
    if (!ctx) return PETAL_INVALID_INPUT;
    ctx->profiling = enable != 0;
    return PETAL_OK;
}

// JSON: {"kernel": {"count": c, "total_ms": t, "min_ms": a, "max_ms": b, "work": w}, ...}; clears the log.
int64_t petal_ctx_profile_json(petal_ctx* ctx, char* buf, int64_t cap) {
    if (!ctx) return -1;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    struct Agg { const char* name; int64_t count; double total, mn, mx, work; };
    std::vector<Agg> aggs;
    for (auto& e : ctx->prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e.a, e.b);
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
        Agg* a = nullptr;
        for (auto& x : aggs) if (std::strcmp(x.name, e.name) == 0) a = &x;
        if (!a) { aggs.push_back(Agg{e.name, 0, 0.0, 1e30, 0.0, 0.0}); a = &aggs.back(); }
        a->count++; a->total += ms; a->mn = std::min(a->mn, (double)ms); a->mx = std::max(a->mx, (double)ms);
        a->work += e.work;
    }
    ctx->prof.clear();
    std::string out = "{";
    for (size_t i = 0; i < aggs.size(); ++i) {
        char tmp[512];
        std::snprintf(tmp, sizeof tmp, "%s\"%s\": {\"count\": %lld, \"total_ms\": %.6f, \"min_ms\": %.6f, \"max_ms\": %.6f, \"work\": %.1f}",
                      i ? ", " : "", aggs[i].name, (long long)aggs[i].count, aggs[i].total, aggs[i].mn, aggs[i].mx, aggs[i].work);
        out += tmp;
    }
    out += "}";
    if (buf && cap > 0) {
        std::strncpy(buf, out.c_str(), (size_t)cap - 1);
        buf[cap - 1] = 0;
    }
    return (int64_t)out.size() + 1;
}

int petal_comm_unique_id(void* out_id) {
    try {
        ncclUniqueId id;
        PETAL_NCCL(nccl_api().GetUniqueId(&id));
        std::memcpy(out_id, &id, sizeof id);
        return PETAL_OK;
    } catch (const Error& e) {
        std::lock_guard<std::mutex> lk(g_global_mutex);
        g_global_error = e.msg;
        return e.code;
    }
}

int petal_comm_init(petal_ctx* ctx, const void* id, int rank, int world_size) {
    return guarded(ctx, [&] {
        if (world_size < 1 || rank < 0 || rank >= world_size) invalid_input("invalid rank / world size");
        delete ctx->comm;
        ctx->comm = nullptr;
        ctx->rank = rank;
        ctx->world = world_size;
        if (world_size == 1) return;
        ncclUniqueId uid;
        std::memcpy(&uid, id, sizeof uid);
        Comm* c = new Comm;
        ncclResult_t r = nccl_api().CommInitRank(&c->comm, world_size, uid, rank);
        if (r != ncclSuccess) {
            delete c;
            ctx->world = 1;
            linalg_error(std::string("ncclCommInitRank failed: ") + nccl_api().GetErrorString(r));
        }
        ctx->comm = c;
    });
}

#define PETAL_DEFINE_TYPED(SUFFIX, T)                                                                          \
    int petal_pca_fit_##SUFFIX(petal_ctx* ctx, const T* x, int64_t n, int64_t d, int64_t k, int centering,      \
                               T* components, T* mean, T* singular, T* total_variance, T* scores) {            \
        return guarded(ctx, [&] {                                                                              \
            pca_fit<T>(ctx, x, n, d, k, centering != 0, components, mean, singular, total_variance, scores);   \
        });                                                                                                    \
    }                                                                                                          \
    int petal_rpca_fit_##SUFFIX(petal_ctx* ctx, const T* x, int64_t n, int64_t d, int64_t k, int centering,     \
                                int64_t n_oversamples, int64_t n_power_iter, const T* omega, T* components,    \
                                T* mean, T* singular, T* total_variance, T* scores) {                          \
        return guarded(ctx, [&] {                                                                              \
            rpca_fit<T>(ctx, x, n, d, k, centering != 0, n_oversamples, n_power_iter, omega, components, mean, \
                        singular, total_variance, scores);                                                     \
        });                                                                                                    \
    }                                                                                                          \
    int petal_transform_##SUFFIX(petal_ctx* ctx, const T* x, int64_t n, int64_t d, const T* components,         \
                                 int64_t k, const T* mean, T* out) {                                           \
        return guarded(ctx, [&] { transform<T>(ctx, x, n, d, components, k, mean, out); });                    \
    }                                                                                                          \
    int petal_inverse_transform_##SUFFIX(petal_ctx* ctx, const T* y, int64_t n, int64_t k,                      \
                                         const T* components, int64_t d, const T* mean, T* out) {              \
        return guarded(ctx, [&] { inverse_transform<T>(ctx, y, n, k, components, d, mean, out); });            \
    }                                                                                                          \
    int petal_fastica_fit_##SUFFIX(petal_ctx* ctx, const T* x, int64_t n, int64_t d, int fun, double tol,       \
                                   int64_t max_iter, int lim_variant, const T* w_init, T* components, T* mean, \
                                   int64_t* n_iter, double* final_lim, T* sources) {                           \
        return guarded(ctx, [&] {                                                                              \
            fastica_fit<T>(ctx, x, n, d, fun, tol, max_iter, lim_variant, w_init, components, mean, n_iter,    \
                           final_lim, sources);                                                                \
        });                                                                                                    \
    }                                                                                                          \
    int petal_fastica_deflation_fit_##SUFFIX(petal_ctx* ctx, const T* x, int64_t n, int64_t d, int fun,         \
                                             double tol, int64_t max_iter, const T* w_init, T* components,     \
                                             T* mean, int64_t* n_iter, double* final_lim, T* sources) {        \
        return guarded(ctx, [&] {                                                                              \
            fastica_fit<T>(ctx, x, n, d, fun, tol, max_iter, 0, w_init, components, mean, n_iter, final_lim,   \
                           sources, /*deflation=*/true);                                                       \
        });                                                                                                    \
    }                                                                                                          \
    int petal_colmean_gram_##SUFFIX(petal_ctx* ctx, const T* x, int64_t n, int64_t d, int centering,            \
                                    double* mean, double* gram) {                                              \
        return guarded(ctx, [&] {                                                                              \
            const int64_t n_total = global_rows(ctx, n);                                                       \
            RowStream<T> X;                                                                                    \
            X.open(ctx, x, n, d, (size_t)(4 * d * d) * sizeof(double));                                        \
            DevOut<double> m(ctx, mean, (size_t)d), g(ctx, gram, (size_t)(d * d));                             \
            if (n_total == 0 || d == 0) return;                                                                \
            ColMean<T> cm;                                                                                     \
            DBuf<double> gtmp;                                                                                 \
            if (!g) gtmp.alloc(ctx, (size_t)(d * d));                                                          \
            mean_and_gram<T>(ctx, X, d, n_total, centering != 0, cm, g ? g.p : gtmp.p);                        \
            if (m) launch_cast<double, double>(ctx, cm.mean_d.p, m.p, d);                                      \
            m.commit(ctx);                                                                                     \
            g.commit(ctx);                                                                                     \
            finish_call(ctx, m.to_host || g.to_host);                                                          \
        });                                                                                                    \
    }

#define PETAL_DEFINE_XTY(SUFFIX, T)                                                                            \
    int petal_xty_##SUFFIX(petal_ctx* ctx, const T* x, int64_t n, int64_t d, const T* mean, const T* y, int64_t l, \
                           double* out) {                                                                      \
        return guarded(ctx, [&] {                                                                              \
            if (n <= 0 || d <= 0 || l <= 0) invalid_input("empty input");                                      \
            DevIn<T> X(ctx, x, (size_t)(n * d)), Y(ctx, y, (size_t)(n * l)), mu(ctx, mean, (size_t)d);         \
            DevOut<double> O(ctx, out, (size_t)(d * l));                                                       \
            if (!O) invalid_input("output buffer is null");                                                    \
            gemm_atb<T>(ctx, X.p, d, d, mu.p, Y.p, l, l, nullptr, n, O.p);                                     \
            allreduce_sum(ctx, O.p, (size_t)(d * l));                                                          \
            O.commit(ctx);                                                                                     \
            finish_call(ctx, O.to_host);                                                                       \
        });                                                                                                    \
    }
PETAL_DEFINE_XTY(f32, float)
PETAL_DEFINE_XTY(f64, double)

PETAL_DEFINE_TYPED(f32, float)
PETAL_DEFINE_TYPED(f64, double)

#define PETAL_DEFINE_ICA_PAR(SUFFIX, T)                                                                        \
    int petal_ica_par_##SUFFIX(petal_ctx* ctx, const T* x1t, int64_t n, int64_t nc, int fun, double tol,           \
                               int64_t max_iter, int lim_variant, const double* w_init, double* w_out,            \
                               int64_t* n_iter, double* final_lim) {                                              \
        return guarded(ctx, [&] {                                                                                 \
            if (n <= 0 || nc <= 0) invalid_input("empty input");                                                  \
            RowStream<T> X;                                                                                       \
            X.open(ctx, x1t, n, nc, (size_t)(2 * n * nc) * sizeof(T));                                            \
            const bool one_pass_local = !X.ring && aligned_on_device(x1t) &&                                      \
                                        ica_one_pass_supported<T>(ctx, nullptr, nc, n, nc, nc);                   \
            const GlobalInfo gi = global_info(ctx, n, one_pass_local);                                            \
            DevIn<double> Wi(ctx, w_init, (size_t)(nc * nc));                                                     \
            DevOut<double> W(ctx, w_out, (size_t)(nc * nc));                                                      \
            if (!W) invalid_input("output buffer is null");                                                       \
            int64_t iters = 0;                                                                                    \
            double lim = 0.0;                                                                                     \
            ica_par<T>(ctx, X, nc, gi.n_total, nullptr, nullptr, nc, fun, tol, max_iter, lim_variant, Wi.p,       \
                       W.p, &iters, &lim, gi.cap[0]);                                                             \
            if (n_iter) *n_iter = iters;                                                                          \
            if (final_lim) *final_lim = lim;                                                                      \
            W.commit(ctx);                                                                                        \
            finish_call(ctx, W.to_host);                                                                          \
        });                                                                                                       \
    }
PETAL_DEFINE_ICA_PAR(f32, float)
PETAL_DEFINE_ICA_PAR(f64, double)

#define PETAL_DEFINE_ICA_DEFL(SUFFIX, T)                                                                       \
    int petal_ica_defl_##SUFFIX(petal_ctx* ctx, const T* x1t, int64_t n, int64_t nc, int fun, double tol,         \
                                int64_t max_iter, const double* w_init, double* w_out, int64_t* n_iter,           \
                                double* final_lim) {                                                              \
        return guarded(ctx, [&] {                                                                                 \
            if (n <= 0 || nc <= 0) invalid_input("empty input");                                                  \
            RowStream<T> X;                                                                                       \
            X.open(ctx, x1t, n, nc, (size_t)(2 * n) * sizeof(T));                                                 \
            const int64_t n_total = global_rows(ctx, n);                                                          \
            DevIn<double> Wi(ctx, w_init, (size_t)(nc * nc));                                                     \
            DevOut<double> W(ctx, w_out, (size_t)(nc * nc));                                                      \
            if (!W) invalid_input("output buffer is null");                                                       \
            int64_t iters = 0;                                                                                    \
            double lim = 0.0;                                                                                     \
            ica_defl<T>(ctx, X, nc, n_total, nullptr, nullptr, nc, fun, tol, max_iter, Wi.p, W.p, &iters, &lim);  \
            if (n_iter) *n_iter = iters;                                                                          \
            if (final_lim) *final_lim = lim;                                                                      \
            W.commit(ctx);                                                                                        \
            finish_call(ctx, W.to_host);                                                                          \
        });                                                                                                       \
    }
PETAL_DEFINE_ICA_DEFL(f32, float)
PETAL_DEFINE_ICA_DEFL(f64, double)

int petal_ica_nonlin_f32(petal_ctx* ctx, float* u, int64_t n, int64_t nc, int fun, int engine, double* gprime_sum) {
    return guarded(ctx, [&] { ica_nonlin<float>(ctx, u, n, nc, fun, engine, gprime_sum); });
}
int petal_ica_nonlin_f64(petal_ctx* ctx, double* u, int64_t n, int64_t nc, int fun, int engine, double* gprime_sum) {
    return guarded(ctx, [&] { ica_nonlin<double>(ctx, u, n, nc, fun, engine, gprime_sum); });
}

int petal_symmetric_decorrelation_f64(petal_ctx* ctx, const double* w, int64_t m, double* out) {
    return guarded(ctx, [&] {
        if (m <= 0) invalid_input("empty input");
        DevIn<double> W(ctx, w, (size_t)(m * m));
        DevOut<double> O(ctx, out, (size_t)(m * m));
        if (!O) invalid_input("output buffer is null");
        symmetric_decorrelation(ctx, W.p, m, O.p);
        O.commit(ctx);
        finish_call(ctx, O.to_host);
    });
}

int petal_small_svd_f64(petal_ctx* ctx, const double* a, int64_t m, int64_t len, double* u, double* s, double* vt) {
    return guarded(ctx, [&] {
        if (m <= 0 || len <= 0) invalid_input("empty input");
        DevIn<double> A(ctx, a, (size_t)(m * len));
        DevOut<double> U(ctx, u, (size_t)(m * m)), S(ctx, s, (size_t)m), Vt(ctx, vt, (size_t)(m * len));
        DBuf<double> Aout(ctx, (size_t)(m * len)), Jt(ctx, (size_t)(m * m)), sig(ctx, (size_t)m);
        jacobi_rows(ctx, A.p, m, len, Aout.p, Jt.p, sig.p);
        if (Vt) launch_normalize_rows(ctx, Aout.p, sig.p, m, len, 0.0, Vt.p);
        if (U) launch_transpose(ctx, Jt.p, m, m, U.p);
        if (S) launch_cast<double, double>(ctx, sig.p, S.p, m);
        U.commit(ctx);
        S.commit(ctx);
        Vt.commit(ctx);
        finish_call(ctx, U.to_host || S.to_host || Vt.to_host);
    });
}

int petal_probe_dmma_tflops(petal_ctx* ctx, int ctas_per_sm, double* out) {
    return guarded(ctx, [&] {
        if (!out || ctas_per_sm < 1 || ctas_per_sm > 8) invalid_input("bad probe arguments");
        DBuf<double> sink(ctx, 1);
        const int iters = 1 << 14, grid = ctx->sm_count * ctas_per_sm;
        cudaEvent_t a, b;
        PETAL_CUDA(cudaEventCreate(&a));
        PETAL_CUDA(cudaEventCreate(&b));
        dmma_probe_kernel<<<grid, 256, 0, ctx->stream>>>(sink.p, 256, 1.0);  // warm-up
        launch1(ctx);
        PETAL_CUDA(cudaEventRecord(a, ctx->stream));
        dmma_probe_kernel<<<grid, 256, 0, ctx->stream>>>(sink.p, iters, 1.0);
        launch1(ctx);
        PETAL_CUDA(cudaEventRecord(b, ctx->stream));
        PETAL_CUDA(cudaEventSynchronize(b));
        float ms = 0.f;
        PETAL_CUDA(cudaEventElapsedTime(&ms, a, b));
        cudaEventDestroy(a);
        cudaEventDestroy(b);
        const double flops = (double)grid * 8.0 * iters * 16.0 * 512.0;
        *out = flops / (ms * 1e-3) / 1e12;
    });
}

}  // extern "C"
