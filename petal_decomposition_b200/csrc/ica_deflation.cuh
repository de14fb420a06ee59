// FastICA, deflation scheme: the components are extracted one at a time, each by the one-unit fixed-point iteration
//     w+ = E[x1 g(w^T x1)] - E[g'(w^T x1)] w,   w+ <- w+ - sum_{i<j} (w+ . w_i) w_i,   w+ <- w+ / |w+|
// (Hyvarinen's algorithm; sklearn `_ica_def`, _fastica.py:65-100, restated in oracle/ica.py::ica_def).  The reference
// crate only has the symmetric ("parallel") scheme (src/ica.rs:319-361); SURVEY 8(f) rank 4 lists this one as the
// callers' next need, with the same whitening (src/ica.rs:189-208) and the same contrast functions around it.
//
// Streaming side: ONE pass over X per iteration, HBM-bound at d * sizeof(T) bytes per sample - with the whitening folded
// into the weight vector (w~ = K1^T w, so w^T x1 = w~ . (x - mu)) a row is read once and used twice:
//     u = w~ . (x - mu),   h += g(u) (x - mu),   gp += g'(u)
// A group of G lanes (8, 16 or 32, so that a group's vector loads cover one row) owns a row at a time: 16-byte loads,
// the dot product reduced with xor-shuffles inside the group, the rank-one update h += g x kept in registers and
// folded through shared memory into d + 1 f64 atomics per CTA at the end.  No tensor cores: 4 flop per loaded element.
// Small side: one single-CTA kernel per iteration (h -> K1 h / n, Gram-Schmidt against the finished rows, the
// convergence test |<w+, w>| -> 1, the next w~), with a device-side stop flag so that the host only looks at the
// state every few iterations (as in ica_par).
#pragma once
#include "common.cuh"
#include "stream_kernels.cuh"

namespace petal {
namespace defl {

constexpr int kThreads = 256;
constexpr int kMaxWordsPerLane = 32;  // row elements a lane keeps in registers (x, mu, w~ and h: 4 x this)

__device__ __forceinline__ int64_t ceil_div_dev(int64_t a, int64_t b) { return (a + b - 1) / b; }
// out[i] = sum_e M[i][e] v[e] for rows i < rows of a row-major matrix: one warp per row, coalesced, shuffle-reduced
__device__ __forceinline__ void cta_matvec(const double* __restrict__ M, int rows, int cols, const double* v, double* out) {
    // four rows per trip: their loads are in flight together (the kernel is a chain of L2 latencies, not of flops)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int W = kThreads / 32, R = 4;
    for (int i0 = warp; i0 < rows; i0 += W * R) {
        double a[R];
#pragma unroll
        for (int r = 0; r < R; ++r) a[r] = 0.0;
        for (int e = lane; e < cols; e += 32) {
            const double ve = v[e];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = i0 + r * W;
                if (i < rows) a[r] += M[(int64_t)i * cols + e] * ve;
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a[r] += __shfl_xor_sync(0xffffffffu, a[r], o);
            if (lane == 0 && i0 + r * W < rows) out[i0 + r * W] = a[r];
        }
    }
}
// out[e] = sum_i M[i][e] c[i] for e < cols (columns of a row-major matrix): thread per column, coalesced, eight rows in
// flight per thread
__device__ __forceinline__ double col_dot(const double* __restrict__ M, int rows, int cols, const double* c, int e) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int i = 0;
    for (; i + 8 <= rows; i += 8) {
        double m[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) m[u] = M[(int64_t)(i + u) * cols + e];
        a0 += m[0] * c[i] + m[4] * c[i + 4];
        a1 += m[1] * c[i + 1] + m[5] * c[i + 5];
        a2 += m[2] * c[i + 2] + m[6] * c[i + 6];
        a3 += m[3] * c[i + 3] + m[7] * c[i + 7];
    }
    for (; i < rows; ++i) a0 += M[(int64_t)i * cols + e] * c[i];
    return (a0 + a1) + (a2 + a3);
}

// state[0] lim, state[1] iterations done for the current component, state[2] 1 = converged (kernels of the batch that
// follow return at once), state[3] max iterations over the finished components
template <typename T, int NV, bool VEC>
__global__ void __launch_bounds__(kThreads)
defl_pass_kernel(const T* __restrict__ X, int64_t n, int64_t d, int64_t ld, const T* __restrict__ mu,
                 const T* __restrict__ wt, int fun, int G, double* __restrict__ hacc, int64_t rows_per_cta,
                 const double* __restrict__ state) {
    constexpr int V = VEC ? Pack<T>::N : 1;
    using Acc = T;  // f32 data: a group sums a few thousand rows in fp32 before the f64 reduction
    extern __shared__ double sh[];  // d + 1
    if (state[2] != 0.0) return;
    const int tid = threadIdx.x, gl = tid % G, grp = tid / G, ngrp = kThreads / G;
    for (int e = tid; e < d + 1; e += kThreads) sh[e] = 0.0;
    __syncthreads();
    T w[NV][V], m[NV][V];
    Acc h[NV][V];
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int c = 0; c < V; ++c) {
            const int64_t e = ((int64_t)gl + (int64_t)G * j) * V + c;
            w[j][c] = e < d ? wt[e] : T(0);
            m[j][c] = (e < d && mu != nullptr) ? mu[e] : T(0);
            h[j][c] = Acc(0);
        }
    double gpsum = 0.0;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta, r1 = min(n, r0 + rows_per_cta);
    // the trip count is uniform over the CTA (the group reductions are full-warp shuffles); rows past the end are zeros
    const int64_t trips = ceil_div_dev(r1 - r0, (int64_t)(2 * ngrp));
    for (int64_t it = 0; it < trips; ++it) {
        const int64_t r = r0 + grp + it * 2 * ngrp;
        // two rows per trip: both rows' loads are issued before the first reduction
        T x[2][NV][V];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int64_t rr = r + (int64_t)q * ngrp;
            const T* row = X + rr * ld;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int64_t e0 = ((int64_t)gl + (int64_t)G * j) * V;
                if (rr < r1 && e0 < d) {
                    if constexpr (VEC) {
                        const Pack<T> p = *reinterpret_cast<const Pack<T>*>(row + e0);
#pragma unroll
                        for (int c = 0; c < V; ++c) x[q][j][c] = p.v[c] - m[j][c];
                    } else {
                        x[q][j][0] = row[e0] - m[j][0];
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < V; ++c) x[q][j][c] = T(0);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            T u = T(0);
#pragma unroll
            for (int j = 0; j < NV; ++j)
#pragma unroll
                for (int c = 0; c < V; ++c) u += x[q][j][c] * w[j][c];
            for (int o = G >> 1; o > 0; o >>= 1) u += __shfl_xor_sync(0xffffffffu, u, o);
            T g, gp;
            ica_g<T>(fun, u, g, gp);
            if (r + (int64_t)q * ngrp < r1) {
                gpsum += (double)gp;
#pragma unroll
                for (int j = 0; j < NV; ++j)
#pragma unroll
                    for (int c = 0; c < V; ++c) h[j][c] += g * x[q][j][c];
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int c = 0; c < V; ++c) {
            const int64_t e = ((int64_t)gl + (int64_t)G * j) * V + c;
            if (e < d) atomicAdd(&sh[e], (double)h[j][c]);
        }
    if (gl == 0) atomicAdd(&sh[d], gpsum);
    __syncthreads();
    for (int e = tid; e < d + 1; e += kThreads) atomicAdd(&hacc[e], sh[e]);
}

// w <- w_init[j] / |w_init[j]|, w~ <- K1^T w, per-component state reset
template <typename T>
__global__ void __launch_bounds__(kThreads)
defl_init_kernel(const double* __restrict__ w_init, int j, int nc, int d, const double* __restrict__ K1,
                 double* __restrict__ w, T* __restrict__ wt, double* __restrict__ state, double* __restrict__ hacc) {
    __shared__ double red[kThreads];
    double s = 0.0;
    for (int i = threadIdx.x; i < nc; i += kThreads) s += w_init[(int64_t)j * nc + i] * w_init[(int64_t)j * nc + i];
    red[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < kThreads; ++i) t += red[i];
        red[0] = t;
    }
    __syncthreads();
    const double inv = rsqrt(red[0]);
    __syncthreads();
    for (int i = threadIdx.x; i < nc; i += kThreads) w[i] = w_init[(int64_t)j * nc + i] * inv;
    __syncthreads();
    for (int e = threadIdx.x; e < d; e += kThreads) {
        double a = 0.0;
        if (K1) {
            for (int i = 0; i < nc; ++i) a += K1[(int64_t)i * d + e] * w[i];
        } else {
            a = w[e];
        }
        wt[e] = (T)a;
    }
    for (int e = threadIdx.x; e < d + 1; e += kThreads) hacc[e] = 0.0;
    if (threadIdx.x == 0) {
        state[0] = 0.0;
        state[1] = 0.0;
        state[2] = 0.0;
    }
}

__device__ __forceinline__ double cta_sum(double v, double* red) {
    __syncthreads();
    red[threadIdx.x] = v;
    __syncthreads();
    for (int o = kThreads / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    return red[0];
}

// one small-side step (sklearn _fastica.py:84-97): w1 = K1 h / n - mean(g') w; Gram-Schmidt against rows 0..j-1 of W
// (all coefficients from the same w1, `_gs_decorrelation`); normalise; lim = ||<w1, w>| - 1|; w <- w1; w~ <- K1^T w
template <typename T>
__global__ void __launch_bounds__(kThreads)
defl_update_kernel(double* __restrict__ hacc, const double* __restrict__ K1, const double* __restrict__ W, int j, int nc,
                   int d, double inv_n, double tol, double* __restrict__ w, T* __restrict__ wt, double* __restrict__ state) {
    extern __shared__ double sm[];  // w1[nc] | coef[j] | red[kThreads]
    double* w1 = sm;
    double* coef = sm + nc;
    double* red = coef + max(j, 1);
    if (state[2] != 0.0) return;
    const int tid = threadIdx.x;
    const double gpm = hacc[d] * inv_n;
    if (K1) {
        cta_matvec(K1, nc, d, hacc, w1);
    } else {
        for (int i = tid; i < nc; i += kThreads) w1[i] = hacc[i];
    }
    __syncthreads();
    for (int i = tid; i < nc; i += kThreads) w1[i] = w1[i] * inv_n - gpm * w[i];
    __syncthreads();
    cta_matvec(W, j, nc, w1, coef);
    __syncthreads();
    double ss = 0.0;
    for (int e = tid; e < nc; e += kThreads) {
        const double v = w1[e] - col_dot(W, j, nc, coef, e);
        w1[e] = v;
        ss += v * v;
    }
    const double nrm = sqrt(cta_sum(ss, red));
    double dot = 0.0;
    for (int e = tid; e < nc; e += kThreads) {
        const double v = w1[e] / nrm;
        w1[e] = v;
        dot += v * w[e];
    }
    const double dt = cta_sum(dot, red);
    const double lim = fabs(fabs(dt) - 1.0);
    for (int e = tid; e < nc; e += kThreads) w[e] = w1[e];
    __syncthreads();
    for (int e = tid; e < d; e += kThreads) wt[e] = (T)(K1 ? col_dot(K1, nc, d, w1, e) : w1[e]);
    for (int e = tid; e < d + 1; e += kThreads) hacc[e] = 0.0;
    if (tid == 0) {
        state[0] = lim;
        state[1] += 1.0;
        if (lim < tol || !(lim == lim)) state[2] = 1.0;  // NaN: stop (reported through lim)
    }
}

// end of a component: W[j] <- w, state[3] <- max(state[3], iterations)
__global__ void defl_commit_kernel(const double* __restrict__ w, double* __restrict__ W, int j, int nc, double* __restrict__ state) {
    for (int e = threadIdx.x; e < nc; e += blockDim.x) W[(int64_t)j * nc + e] = w[e];
    if (threadIdx.x == 0) state[3] = fmax(state[3], state[1]);
}

template <typename T>
inline bool pass_supported(int64_t d) {
    return d >= 1 && d * (int64_t)(sizeof(T) / 4) <= 32 * kMaxWordsPerLane;
}

// hacc[d + 1] += [sum g(u) (x - mu) | sum g'(u)] over the n rows of X, u = (x - mu) . wt
template <typename T>
void launch_pass(petal_ctx* ctx, const T* X, int64_t n, int64_t d, int64_t ld, const T* mu, const T* wt, int fun,
                 double* hacc, const double* state) {
    if (n == 0) return;
    constexpr int V = Pack<T>::N;
    const bool vec = (d % V == 0) && (ld % V == 0) && is_aligned16(X) && (mu == nullptr || is_aligned16(mu)) && is_aligned16(wt);
    const int64_t per_row = vec ? d / V : d;  // loads per row
    int G = 32;
    while (G > 8 && per_row <= G / 2) G >>= 1;
    const int nv = (int)ceil_div(per_row, G);
    const int64_t target = (int64_t)ctx->sm_count * 4;
    const int ngrp = kThreads / G;
    int64_t grid = std::max<int64_t>(1, std::min<int64_t>(target, ceil_div(n, (int64_t)ngrp * 8)));
    const int64_t rows_per_cta = ceil_div(n, grid);
    grid = ceil_div(n, rows_per_cta);
    const size_t smem = (size_t)(d + 1) * sizeof(double);
    KTimer kt(ctx, kname<T>("ica_defl_pass_f32", "ica_defl_pass_f64"), (double)n * d * sizeof(T));
#define PETAL_DEFL_LAUNCH(NVV, VECV)                                                                                           \
    defl_pass_kernel<T, NVV, VECV><<<(unsigned)grid, kThreads, smem, ctx->stream>>>(X, n, d, ld, mu, wt, fun, G, hacc, rows_per_cta, \
                                                                                    state)
    constexpr int kMaxNvVec = kMaxWordsPerLane / (V * (int)(sizeof(T) / 4));  // 8 for both types
    if (vec) {
        if (nv <= 1) PETAL_DEFL_LAUNCH(1, true);
        else if (nv <= 2) PETAL_DEFL_LAUNCH(2, true);
        else if (nv <= 4) PETAL_DEFL_LAUNCH(4, true);
        else if (nv <= kMaxNvVec) PETAL_DEFL_LAUNCH(kMaxNvVec, true);
        else linalg_error("deflation pass: row too wide");
    } else {
        constexpr int kMaxNvScalar = kMaxWordsPerLane / (int)(sizeof(T) / 4);  // 32 (f32), 16 (f64)
        if (nv <= 2) PETAL_DEFL_LAUNCH(2, false);
        else if (nv <= 8) PETAL_DEFL_LAUNCH(8, false);
        else if (nv <= kMaxNvScalar) PETAL_DEFL_LAUNCH(kMaxNvScalar, false);
        else linalg_error("deflation pass: row too wide");
    }
#undef PETAL_DEFL_LAUNCH
    check_launch(ctx);
}

}  // namespace defl
}  // namespace petal
