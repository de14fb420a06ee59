// Small replicated dense factorizations on device, all f64: a one-sided (Hestenes) Jacobi SVD.
//
// It replaces every LAPACK call the reference makes (src/linalg/lapack.rs: gesvd :103-132,
// gesdd :70-101, syev/heev :134-184, gelqf/orglq :49-68,186-202) on the small matrices that
// are replicated on every GPU: the d x d Gram of exact PCA / whitening, the l x l Grams of the
// range finder, the l x d projected matrix B and the nc x nc FastICA update.
//
// Formulation ("rows"): given m vectors of length len (row-major A[m][len]) find an orthogonal
// Jt[m][m] such that Jt * A has mutually orthogonal rows:  Jt * A = diag(s) * N,  N N^T = I.
//   -> SVD            A = U diag(s) Vt   with U = Jt^T, Vt = N
//   -> eigh (A sym. PSD): eigenvalues s, eigenvectors = rows of Jt
//   -> polar factor   (A A^T)^-1/2 A = Jt^T N                      (symmetric decorrelation)
// Rows are returned sorted by s descending.
//
// Two engines: one CTA with everything in shared memory (m*(len+m)*8 B <= 200 KB; latency-bound
// problems such as the 64 x 64 FastICA update), and a multi-CTA engine working in global memory
// (L2-resident for moderate sizes) with one launch per round-robin step.
#pragma once
#include "common.cuh"

namespace petal {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// round-robin (circle method) pairing: step s in [0, me-1), pair index i in [0, me/2)
__device__ __forceinline__ void rr_pair(int me, int s, int i, int& p, int& q) {
    const int r = me - 1;
    if (i == 0) {
        p = r;
        q = s;
    } else {
        p = (s + i) % r;
        q = (s - i + r) % r;
    }
    if (p > q) {
        int t = p;
        p = q;
        q = t;
    }
}

// Rotation (c, s) that annihilates gamma between two rows of squared norms alpha, beta (exact IEEE f64 math:
// an approximate tangent - MUFU reciprocal / rsqrt + Newton-refined cosine - was tried in r02 and lost: every
// rotation then leaves a ~1e-6 residual, which costs the inner solves two extra sweeps, more than the cheaper
// parameter computation saves; the steps of the shared-memory engines are bound by barriers and dependent
// shared-memory accesses, not by this arithmetic).
__device__ __forceinline__ void rotation_from(double alpha, double beta, double gamma, double& c, double& s, double* t_out) {
    const double zeta = (beta - alpha) / (2.0 * gamma);
    const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    c = rsqrt(1.0 + t * t);
    s = c * t;
    if (t_out) *t_out = t;
}

// Computes the rotation for a row pair; returns false when already orthogonal to tolerance.
// `noise` = (relative entry noise of the input) * (largest row norm): rows of a Gram matrix carry
// absolute noise eps * lambda_max from the start, so two small rows cannot be made orthogonal
// beyond that floor - without the floor the sweep loop never reports convergence.
__device__ __forceinline__ bool jacobi_rotation(double alpha, double beta, double gamma, double tol, double noise,
                                                double& c, double& s, double* t_out = nullptr) {
    const double na = sqrt(alpha), nb = sqrt(beta);
    if (!(fabs(gamma) > tol * na * nb + noise * (na + nb))) return false;  // also false for NaN/zero rows
    rotation_from(alpha, beta, gamma, c, s, t_out);
    return true;
}

// Rows whose norm is below kJacobiZeroRow * (largest row norm) are numerically zero (rank-deficient inputs: m > len,
// repeated samples): their direction is rounding noise, rotating them never terminates, and what they are rotated
// against does not change beyond eps.  They take part in no rotation.
constexpr double kJacobiZeroRow = 4.0 * 2.220446049250313e-16;

// ------------------------------------------------------------------------------------------
// single-CTA engine
// ------------------------------------------------------------------------------------------
constexpr int kJacobiSmemThreads = 1024;

__device__ __forceinline__ double group16_sum(double v, unsigned mask) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}

__global__ void __launch_bounds__(kJacobiSmemThreads)
jacobi_smem_kernel(const double* __restrict__ A, int m, int len, double* __restrict__ Aout,
                   double* __restrict__ Jt, double* __restrict__ sig, int max_sweeps, double tol,
                   double noise_rel, const int* __restrict__ run_flag, int* __restrict__ info,
                   int* __restrict__ status) {
    if (run_flag != nullptr && *run_flag == 0) return;  // fast path succeeded: nothing to do
    extern __shared__ double sm[];
    double* M = sm;                       // m x len
    double* J = sm + (size_t)m * len;     // m x m
    double* nrm = J + (size_t)m * m;      // m
    __shared__ int rotated;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = kJacobiSmemThreads / 32;

    for (int i = tid; i < m * len; i += kJacobiSmemThreads) M[i] = A[i];
    for (int i = tid; i < m * m; i += kJacobiSmemThreads) J[i] = ((i / m) == (i % m)) ? 1.0 : 0.0;
    if (tid == 0) rotated = 0;
    __syncthreads();
    // largest row norm (for the input-noise floor)
    for (int j = warp; j < m; j += nwarps) {
        double a = 0.0;
        for (int e = lane; e < len; e += 32) a += M[(size_t)j * len + e] * M[(size_t)j * len + e];
        a = warp_sum(a);
        if (lane == 0) nrm[j] = a;
    }
    __syncthreads();
    double amax = 0.0;
    for (int j = 0; j < m; ++j) amax = fmax(amax, nrm[j]);
    const double noise = fmax(noise_rel, kJacobiZeroRow) * sqrt(amax);
    __syncthreads();

    const int me = (m + 1) & ~1;
    int sweeps = 0;
    bool converged = (m <= 1);
    // one 16-lane group per row pair: all pairs of a round-robin step run concurrently
    const int gid = tid >> 4, gl = tid & 15, ngroups = kJacobiSmemThreads / 16;
    const unsigned gmask = 0xFFFFu << (lane & 16);
    for (int sweep = 0; sweep < max_sweeps && m > 1; ++sweep) {
        // squared row norms: recomputed exactly once per sweep, then carried through the rotations
        // (|a'|^2 = c^2 |a|^2 - 2 c s g + s^2 |b|^2, ...), so a step needs one inner product instead of three
        for (int j = warp; j < m; j += nwarps) {
            double a = 0.0;
            for (int e = lane; e < len; e += 32) a += M[(size_t)j * len + e] * M[(size_t)j * len + e];
            a = warp_sum(a);
            if (lane == 0) nrm[j] = a;
        }
        __syncthreads();
        for (int step = 0; step < me - 1; ++step) {
            for (int pi = gid; pi < me / 2; pi += ngroups) {
                int p, q;
                rr_pair(me, step, pi, p, q);
                if (q >= m) continue;
                double* mp = M + (size_t)p * len;
                double* mq = M + (size_t)q * len;
                double ga = 0.0;
                for (int e = gl; e < len; e += 16) ga += mp[e] * mq[e];
                ga = group16_sum(ga, gmask);
                const double al = nrm[p], be = nrm[q];
                double c, s, t;
                if (!jacobi_rotation(al, be, ga, tol, noise, c, s, &t)) continue;
                __syncwarp(gmask);  // every lane of the group has read nrm[p], nrm[q]
                if (gl == 0) {  // exact for any rotation (c, s)
                    nrm[p] = fmax(c * c * al - 2.0 * c * s * ga + s * s * be, 0.0);
                    nrm[q] = fmax(s * s * al + 2.0 * c * s * ga + c * c * be, 0.0);
                }
                for (int e = gl; e < len; e += 16) {
                    double x = mp[e], y = mq[e];
                    mp[e] = c * x - s * y;
                    mq[e] = s * x + c * y;
                }
                double* jp = J + (size_t)p * m;
                double* jq = J + (size_t)q * m;
                for (int e = gl; e < m; e += 16) {
                    double x = jp[e], y = jq[e];
                    jp[e] = c * x - s * y;
                    jq[e] = s * x + c * y;
                }
                if (gl == 0) rotated = 1;
            }
            __syncthreads();
        }
        sweeps = sweep + 1;
        int r = rotated;
        __syncthreads();
        if (tid == 0) rotated = 0;
        __syncthreads();
        if (!r) {
            converged = true;
            break;
        }
    }
    // sweeps exhausted while rotations were still being applied: the reference's LAPACK drivers return info > 0
    // here (src/linalg.rs:84,115 -> "did not converge"); the host turns this bit into PETAL_LINALG_ERROR
    if (!converged && tid == 0 && status != nullptr) atomicOr(status, kStatusJacobiNotConverged);

    // norms
    for (int j = warp; j < m; j += nwarps) {
        double a = 0.0;
        for (int e = lane; e < len; e += 32) a += M[(size_t)j * len + e] * M[(size_t)j * len + e];
        a = warp_sum(a);
        if (lane == 0) nrm[j] = sqrt(a);
    }
    __syncthreads();
    // rank sort (descending, stable) and scatter
    for (int j = warp; j < m; j += nwarps) {
        double sj = nrm[j];
        int rank = 0;
        for (int i = lane; i < m; i += 32) {
            double si = nrm[i];
            rank += (si > sj || (si == sj && i < j)) ? 1 : 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
        if (lane == 0) sig[rank] = sj;
        if (Aout)
            for (int e = lane; e < len; e += 32) Aout[(size_t)rank * len + e] = M[(size_t)j * len + e];
        if (Jt)
            for (int e = lane; e < m; e += 32) Jt[(size_t)rank * m + e] = J[(size_t)j * m + e];
    }
    if (tid == 0 && info) *info = sweeps;
}

// ------------------------------------------------------------------------------------------
// multi-CTA engine (global memory)
// ------------------------------------------------------------------------------------------
__global__ void set_identity_kernel(double* J, int64_t m) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m * m) J[i] = ((i / m) == (i % m)) ? 1.0 : 0.0;
}

__global__ void __launch_bounds__(256)
jacobi_step_kernel(double* __restrict__ M, int m, int len, double* __restrict__ J, int me, int step,
                   double tol, const double* __restrict__ noise_ptr, int* __restrict__ rotated) {
    int p, q;
    rr_pair(me, step, blockIdx.x, p, q);
    if (q >= m) return;
    double* mp = M + (size_t)p * len;
    double* mq = M + (size_t)q * len;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double al = 0.0, be = 0.0, ga = 0.0;
    for (int e = tid; e < len; e += 256) {
        double x = mp[e], y = mq[e];
        al += x * x;
        be += y * y;
        ga += x * y;
    }
    __shared__ double red[3][8];
    al = warp_sum(al);
    be = warp_sum(be);
    ga = warp_sum(ga);
    if (lane == 0) {
        red[0][warp] = al;
        red[1][warp] = be;
        red[2][warp] = ga;
    }
    __syncthreads();
    al = be = ga = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        al += red[0][w];
        be += red[1][w];
        ga += red[2][w];
    }
    double c, s;
    if (!jacobi_rotation(al, be, ga, tol, *noise_ptr, c, s)) return;
    for (int e = tid; e < len; e += 256) {
        double x = mp[e], y = mq[e];
        mp[e] = c * x - s * y;
        mq[e] = s * x + c * y;
    }
    double* jp = J + (size_t)p * m;
    double* jq = J + (size_t)q * m;
    for (int e = tid; e < m; e += 256) {
        double x = jp[e], y = jq[e];
        jp[e] = c * x - s * y;
        jq[e] = s * x + c * y;
    }
    if (tid == 0) *rotated = 1;
}

// Persistent multi-CTA engine: all sweeps in one cooperative launch.  Every CTA takes the pairs blockIdx.x,
// blockIdx.x + gridDim.x, ... of a round-robin step; steps are separated by a grid-wide barrier (one atomic
// counter; cooperative launch guarantees co-residency).  Squared row norms are recomputed once per sweep and
// carried through the rotations, so a step costs one inner product per pair.  rot[sweep] is raised by any
// rotation; a sweep without rotations ends the factorization.  info[0] = sweeps used.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        } while ((int)(v - target) < 0);
        __threadfence();  // acquire side: later loads of this CTA (ordered behind the barrier below) see the other CTAs' rows
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256)
jacobi_coop_kernel(double* __restrict__ M, int m, int len, double* __restrict__ J, double* __restrict__ nrm2, int me,
                   int max_sweeps, double tol, double noise_rel, unsigned* __restrict__ bar, int* __restrict__ rot,
                   int* __restrict__ info, int* __restrict__ status) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ double red[8];
    unsigned target = 0;
    auto row_norm2 = [&](int j) {  // block-wide, result in bc[0] for every thread after the sync
        double a = 0.0;
        for (int e = tid; e < len; e += 256) {
            const double x = M[(size_t)j * len + e];
            a += x * x;
        }
        a = warp_sum(a);
        __syncthreads();
        if (lane == 0) red[warp] = a;
        __syncthreads();
        double sacc = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) sacc += red[w];
        return sacc;
    };
    double noise = 0.0;
    int sweeps = 0;
    bool converged = (m <= 1);
    for (int sweep = 0; sweep < max_sweeps && m > 1; ++sweep) {
        for (int j = blockIdx.x; j < m; j += gridDim.x) {
            const double a = row_norm2(j);
            if (tid == 0) nrm2[j] = a;
        }
        grid_barrier(bar, target);
        if (sweep == 0) {  // input-noise floor from the largest row norm
            double amax = 0.0;
            for (int j = tid; j < m; j += 256) amax = fmax(amax, nrm2[j]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
            __syncthreads();
            if (lane == 0) red[warp] = amax;
            __syncthreads();
            amax = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) amax = fmax(amax, red[w]);
            noise = fmax(noise_rel, kJacobiZeroRow) * sqrt(amax);
        }
        for (int step = 0; step < me - 1; ++step) {
            for (int pi = blockIdx.x; pi < me / 2; pi += gridDim.x) {
                int p, q;
                rr_pair(me, step, pi, p, q);
                if (q >= m) continue;
                double* mp = M + (size_t)p * len;
                double* mq = M + (size_t)q * len;
                double ga = 0.0;
                for (int e = tid; e < len; e += 256) ga += mp[e] * mq[e];
                ga = warp_sum(ga);
                __syncthreads();
                if (lane == 0) red[warp] = ga;
                __syncthreads();
                ga = 0.0;
#pragma unroll
                for (int w = 0; w < 8; ++w) ga += red[w];
                const double al = nrm2[p], be = nrm2[q];
                double c, s, t;
                if (!jacobi_rotation(al, be, ga, tol, noise, c, s, &t)) continue;  // block-uniform
                __syncthreads();
                if (tid == 0) {
                    nrm2[p] = fmax(c * c * al - 2.0 * c * s * ga + s * s * be, 0.0);  // exact for any rotation (c, s)
                    nrm2[q] = fmax(s * s * al + 2.0 * c * s * ga + c * c * be, 0.0);
                    rot[sweep] = 1;
                }
                for (int e = tid; e < len; e += 256) {
                    const double x = mp[e], y = mq[e];
                    mp[e] = c * x - s * y;
                    mq[e] = s * x + c * y;
                }
                double* jp = J + (size_t)p * m;
                double* jq = J + (size_t)q * m;
                for (int e = tid; e < m; e += 256) {
                    const double x = jp[e], y = jq[e];
                    jp[e] = c * x - s * y;
                    jq[e] = s * x + c * y;
                }
            }
            grid_barrier(bar, target);
        }
        sweeps = sweep + 1;
        if (reinterpret_cast<volatile int*>(rot)[sweep] == 0) {  // written before the last barrier of the sweep: same value in every CTA
            converged = true;
            break;
        }
    }
    if (blockIdx.x == 0 && tid == 0 && !converged && status != nullptr) atomicOr(status, kStatusJacobiNotConverged);
    if (blockIdx.x == 0 && tid == 0 && info) info[0] = sweeps;
}

__global__ void __launch_bounds__(256)
row_norm_kernel(const double* __restrict__ M, int m, int len, double* __restrict__ nrm) {
    const int j = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double a = 0.0;
    for (int e = tid; e < len; e += 256) {
        double x = M[(size_t)j * len + e];
        a += x * x;
    }
    __shared__ double red[8];
    a = warp_sum(a);
    if (lane == 0) red[warp] = a;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w];
        nrm[j] = sqrt(s);
    }
}

// noise[0] = noise_rel * max_j nrm[j]
__global__ void noise_floor_kernel(const double* __restrict__ nrm, int m, double noise_rel, double* __restrict__ noise) {
    double a = 0.0;
    for (int j = 0; j < m; ++j) a = fmax(a, nrm[j]);
    noise[0] = fmax(noise_rel, kJacobiZeroRow) * a;
}

__global__ void __launch_bounds__(256)
sort_scatter_kernel(const double* __restrict__ M, const double* __restrict__ J,
                    const double* __restrict__ nrm, int m, int len, double* __restrict__ Aout,
                    double* __restrict__ Jt, double* __restrict__ sig) {
    const int j = blockIdx.x;
    const int tid = threadIdx.x;
    __shared__ int rank_s;
    if (tid == 0) rank_s = 0;
    __syncthreads();
    double sj = nrm[j];
    int rank = 0;
    for (int i = tid; i < m; i += 256) {
        double si = nrm[i];
        rank += (si > sj || (si == sj && i < j)) ? 1 : 0;
    }
    if (rank) atomicAdd(&rank_s, rank);
    __syncthreads();
    rank = rank_s;
    if (tid == 0) sig[rank] = sj;
    if (Aout)
        for (int e = tid; e < len; e += 256) Aout[(size_t)rank * len + e] = M[(size_t)j * len + e];
    if (Jt)
        for (int e = tid; e < m; e += 256) Jt[(size_t)rank * m + e] = J[(size_t)j * m + e];
}

// A[m][len] (device, f64, not modified) -> Aout[m][len] (= diag(sig) N, may be null),
// Jt[m][m] (may be null), sig[m].  Returns the number of sweeps used (-1 if unknown).
// input_noise_rel: relative entry noise of A (0 for an exactly given matrix; ~eps for a Gram matrix
// accumulated in f64) - sets the orthogonality floor, see jacobi_rotation.
// block engine for large m (dense_f64.cuh)
inline void block_jacobi_rows(petal_ctx* ctx, const double* A, int64_t m, int64_t len, double* Aout, double* Jt, double* sig,
                              double input_noise_rel);
inline int64_t jacobi_block_min() {
    if (const char* e = getenv("PETAL_JACOBI_BLOCK_MIN")) return std::max<int64_t>(2, atoll(e));
    return 2048;
}

inline bool jacobi_fits_smem(int64_t m, int64_t len) {
    return ((size_t)m * len + (size_t)m * m + (size_t)m) * sizeof(double) <= 200 * 1024;
}

// run_flag (device, optional): when given and *run_flag == 0 the factorization is skipped (used as the
// fallback of the Cholesky fast path; only honoured by the shared-memory engine).
inline int jacobi_rows(petal_ctx* ctx, const double* A, int64_t m, int64_t len, double* Aout, double* Jt,
                       double* sig, double input_noise_rel = 0.0, bool force_global = false,
                       const int* run_flag = nullptr) {
    if (m == 0) return 0;
    if (m > (int64_t)1 << 20 || len > (int64_t)1 << 30) invalid_input("matrix too large for the Jacobi solver");
    int max_sweeps = 60;
    if (const char* e = getenv("PETAL_JACOBI_MAX_SWEEPS")) max_sweeps = std::max(1, atoi(e));  // testing: forces "did not converge"
    // |<a_p, a_q>| <= tol * |a_p| |a_q| counts as orthogonal. The computed inner product carries
    // ~sqrt(len) * eps of rounding noise, so the threshold sits a small factor above that - a
    // tighter one never reports a rotation-free sweep and runs to max_sweeps.
    const double tol = 8.0 * 2.220446049250313e-16 * std::sqrt((double)std::max<int64_t>(len, 1));
    size_t smem = ((size_t)m * len + (size_t)m * m + (size_t)m) * sizeof(double);
    const bool use_smem = !force_global && smem <= 200 * 1024;
    if (!use_smem && run_flag == nullptr && m >= jacobi_block_min()) {
        // rotations as DMMA GEMMs on pairs of 32-row blocks: the scalar engines below are DFMA-issue bound
        block_jacobi_rows(ctx, A, m, len, Aout, Jt, sig, input_noise_rel);
        return -1;
    }
    KTimer kt(ctx, use_smem ? "jacobi_smem" : "jacobi_global", 0.0);
    ctx->status_armed = true;
    if (use_smem) {
        ensure_dynamic_smem(ctx, jacobi_smem_kernel, 208 * 1024);
        DBuf<int> info;
        const bool want_info = getenv("PETAL_JACOBI_INFO") != nullptr;
        if (want_info) {
            info.alloc(ctx, 1);
            info.zero();
        }
        jacobi_smem_kernel<<<1, kJacobiSmemThreads, smem, ctx->stream>>>(A, (int)m, (int)len, Aout, Jt, sig,
                                                                         max_sweeps, tol, input_noise_rel, run_flag, info.p,
                                                                         ctx->dev_status);
        check_launch(ctx);
        if (want_info) {
            int h = 0;
            PETAL_CUDA(cudaMemcpyAsync(&h, info.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            PETAL_CUDA(cudaStreamSynchronize(ctx->stream));
            fprintf(stderr, "[jacobi_smem] m %d len %d sweeps %d\n", (int)m, (int)len, h);
        }
        return -1;
    }
    DBuf<double> M(ctx, (size_t)(m * len));
    DBuf<double> J(ctx, (size_t)(m * m));
    DBuf<double> nrm(ctx, (size_t)m);
    DBuf<int> rotated(ctx, 1);
    PETAL_CUDA(cudaMemcpyAsync(M.p, A, (size_t)(m * len) * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    set_identity_kernel<<<(unsigned)ceil_div(m * m, 256), 256, 0, ctx->stream>>>(J.p, m);
    check_launch(ctx);
    DBuf<double> noise(ctx, 1);
    row_norm_kernel<<<(unsigned)m, 256, 0, ctx->stream>>>(M.p, (int)m, (int)len, nrm.p);
    check_launch(ctx);
    noise_floor_kernel<<<1, 1, 0, ctx->stream>>>(nrm.p, (int)m, input_noise_rel, noise.p);
    check_launch(ctx);
    const int me = (int)((m + 1) & ~(int64_t)1);
    int sweeps = 0;
    // persistent cooperative kernel (one launch for all sweeps) when the device supports it
    if (ctx->coop_ok < 0) {
        int attr = 0, per_sm = 0;
        cudaDeviceGetAttribute(&attr, cudaDevAttrCooperativeLaunch, ctx->device);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jacobi_coop_kernel, 256, 0);
        ctx->coop_max_ctas = per_sm * ctx->sm_count;
        ctx->coop_ok = (attr != 0 && ctx->coop_max_ctas > 0 && getenv("PETAL_JACOBI_COOP_OFF") == nullptr) ? 1 : 0;
    }
    if (ctx->coop_ok == 1 && m > 1) {
        DBuf<unsigned> bar(ctx, 1);
        DBuf<int> rot(ctx, (size_t)max_sweeps + 1);  // [sweep flags | sweeps used]
        bar.zero();
        rot.zero();
        int grid = std::min<int>(me / 2, ctx->coop_max_ctas);
        grid = std::max(grid, 1);
        double* Mp = M.p;
        double* Jp = J.p;
        double* np = nrm.p;
        int mi = (int)m, li = (int)len, ms = max_sweeps;
        double tl = tol, nr = input_noise_rel;
        unsigned* bp = bar.p;
        int* rp = rot.p;
        int* ip = rot.p + max_sweeps;
        int mee = me;
        int* sp = ctx->dev_status;
        void* args[] = {&Mp, &mi, &li, &Jp, &np, &mee, &ms, &tl, &nr, &bp, &rp, &ip, &sp};
        PETAL_CUDA(cudaLaunchCooperativeKernel((void*)jacobi_coop_kernel, dim3((unsigned)grid), dim3(256), args, 0, ctx->stream));
        check_launch(ctx);
        row_norm_kernel<<<(unsigned)m, 256, 0, ctx->stream>>>(M.p, (int)m, (int)len, nrm.p);
        check_launch(ctx);
        sort_scatter_kernel<<<(unsigned)m, 256, 0, ctx->stream>>>(M.p, J.p, nrm.p, (int)m, (int)len, Aout, Jt, sig);
        check_launch(ctx);
        return -1;
    }
    for (int sweep = 0; sweep < max_sweeps && m > 1; ++sweep) {
        rotated.zero();
        for (int step = 0; step < me - 1; ++step) {
            jacobi_step_kernel<<<me / 2, 256, 0, ctx->stream>>>(M.p, (int)m, (int)len, J.p, me, step, tol,
                                                                noise.p, rotated.p);
            check_launch(ctx);
        }
        int h = 0;
        PETAL_CUDA(cudaMemcpyAsync(&h, rotated.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        PETAL_CUDA(cudaStreamSynchronize(ctx->stream));
        sweeps = sweep + 1;
        if (!h) break;
        if (sweep == max_sweeps - 1) linalg_error("did not converge");  // src/linalg.rs:84,115
    }
    row_norm_kernel<<<(unsigned)m, 256, 0, ctx->stream>>>(M.p, (int)m, (int)len, nrm.p);
    check_launch(ctx);
    sort_scatter_kernel<<<(unsigned)m, 256, 0, ctx->stream>>>(M.p, J.p, nrm.p, (int)m, (int)len, Aout, Jt, sig);
    check_launch(ctx);
    return sweeps;
}

// ------------------------------------------------------------------------------------------
// helpers built on the factorization
// ------------------------------------------------------------------------------------------
// N[j][:] = Aout[j][:] / sig[j]  (zero row when sig[j] <= cutoff * sig[0])
__global__ void normalize_rows_kernel(const double* __restrict__ Aout, const double* __restrict__ sig,
                                      int64_t m, int64_t len, double cutoff, double* __restrict__ N) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * len) return;
    int64_t j = idx / len;
    double s = sig[j];
    double s0 = sig[0];
    N[idx] = (s > cutoff * s0 && s > 0.0) ? Aout[idx] / s : 0.0;
}

inline void launch_normalize_rows(petal_ctx* ctx, const double* Aout, const double* sig, int64_t m,
                                  int64_t len, double cutoff, double* N) {
    if (m * len == 0) return;
    normalize_rows_kernel<<<(unsigned)ceil_div(m * len, 256), 256, 0, ctx->stream>>>(Aout, sig, m, len, cutoff, N);
    check_launch(ctx);
}

// P[a][j] = Jt[j][a] * f(sig[j]),   mode 0: f = 1/sqrt(s)  (s = eigenvalue of a Gram matrix)
//                                    mode 1: f = 1/s
//                                    mode 2: f = 1
// f = 0 when s <= cutoff * s[0]  (rank-deficient directions are dropped, not amplified)
__global__ void scaled_transpose_kernel(const double* __restrict__ Jt, const double* __restrict__ sig,
                                        int64_t m, int mode, double cutoff, double* __restrict__ P,
                                        const int* __restrict__ run_flag) {
    if (run_flag != nullptr && *run_flag == 0) return;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * m) return;
    int64_t a = idx / m, j = idx % m;
    double s = sig[j], s0 = sig[0];
    double f = 0.0;
    if (mode == 2) f = 1.0;
    else if (s > cutoff * s0 && s > 0.0) f = (mode == 0) ? rsqrt(s) : 1.0 / s;
    P[idx] = Jt[j * m + a] * f;
}

inline void launch_scaled_transpose(petal_ctx* ctx, const double* Jt, const double* sig, int64_t m, int mode,
                                    double cutoff, double* P, const int* run_flag = nullptr) {
    if (m == 0) return;
    scaled_transpose_kernel<<<(unsigned)ceil_div(m * m, 256), 256, 0, ctx->stream>>>(Jt, sig, m, mode, cutoff, P,
                                                                                     run_flag);
    check_launch(ctx);
}


// ------------------------------------------------------------------------------------------
// Cholesky fast path: G (m x m, SPD) = R^T R, P = R^-1 (upper triangular), so that Z P has orthonormal
// columns when G = Z^T Z.  Single CTA, everything in shared memory.  Sets *fail = 1 (and leaves P
// untouched) when a pivot drops below cutoff * max diag - the caller then falls back to the Jacobi
// eigensolver, which handles rank-deficient G.
// ------------------------------------------------------------------------------------------
constexpr int kCholMax = 104;

__global__ void __launch_bounds__(256)
chol_inverse_kernel(const double* __restrict__ G, int m, double cutoff, double* __restrict__ P,
                    int* __restrict__ fail) {
    extern __shared__ double sm[];
    const int ld = m + 1;
    double* A = sm;               // m x ld : R in the upper triangle
    double* X = sm + (size_t)m * ld;  // m x ld : R^-1
    __shared__ double maxdiag;
    __shared__ int bad;
    const int tid = threadIdx.x;
    for (int i = tid; i < m * m; i += 256) {
        A[(i / m) * ld + (i % m)] = G[i];
        X[(i / m) * ld + (i % m)] = 0.0;
    }
    if (tid == 0) {
        double md = 0.0;
        for (int i = 0; i < m; ++i) md = fmax(md, G[(size_t)i * m + i]);
        maxdiag = md;
        bad = !(md > 0.0) || !isfinite(md);
    }
    __syncthreads();
    for (int j = 0; j < m && !bad; ++j) {
        const double d = A[j * ld + j];
        if (!(d > cutoff * maxdiag)) {
            __syncthreads();
            if (tid == 0) bad = 1;
            __syncthreads();
            break;
        }
        const double r = sqrt(d);
        __syncthreads();
        for (int c = j + tid; c < m; c += 256) A[j * ld + c] = (c == j) ? r : A[j * ld + c] / r;
        __syncthreads();
        // trailing update: A[i][c] -= R[j][i] * R[j][c] for j < i <= c
        const int t = m - j - 1;
        for (int e = tid; e < t * t; e += 256) {
            const int i = j + 1 + e / t, c = j + 1 + e % t;
            if (c >= i) A[i * ld + c] -= A[j * ld + i] * A[j * ld + c];
        }
        __syncthreads();
    }
    if (bad) {
        if (tid == 0) *fail = 1;
        return;
    }
    // back substitution, one column of R^-1 per thread
    for (int c = tid; c < m; c += 256) {
        X[c * ld + c] = 1.0 / A[c * ld + c];
        for (int i = c - 1; i >= 0; --i) {
            double acc = 0.0;
            for (int k = i + 1; k <= c; ++k) acc += A[i * ld + k] * X[k * ld + c];
            X[i * ld + c] = -acc / A[i * ld + i];
        }
    }
    __syncthreads();
    for (int i = tid; i < m * m; i += 256) P[i] = X[(i / m) * ld + (i % m)];
}

inline bool chol_supported(int64_t m) { return m >= 1 && m <= kCholMax; }

inline void launch_chol_inverse(petal_ctx* ctx, const double* G, int64_t m, double cutoff, double* P, int* fail) {
    size_t smem = 2 * (size_t)m * (m + 1) * sizeof(double);
    ensure_dynamic_smem(ctx, chol_inverse_kernel, 200 * 1024);
    KTimer kt(ctx, "cholesky", 0.0);
    chol_inverse_kernel<<<1, 256, smem, ctx->stream>>>(G, (int)m, cutoff, P, fail);
    check_launch(ctx);
}

// P (m x m) with (Z P)^T (Z P) = I on the numerical range of Z, given G = Z^T Z:
// Cholesky when G is safely positive definite, Jacobi eigensolver (W Lambda^-1/2, null directions
// dropped) otherwise - decided on the device, no host round trip.
inline void gram_to_orthonormalizer(petal_ctx* ctx, const double* G, int64_t m, double cutoff, double noise_rel,
                                    double* P) {
    DBuf<double> Jt(ctx, (size_t)(m * m)), sig(ctx, (size_t)m);
    if (chol_supported(m) && jacobi_fits_smem(m, m)) {
        DBuf<int> fail(ctx, 1);
        fail.zero();
        // a pivot below sqrt(cutoff)-ish of the largest diagonal means the columns are numerically
        // dependent at the level where one Cholesky round can no longer fix it
        launch_chol_inverse(ctx, G, m, std::max(cutoff, 1e-10), P, fail.p);
        jacobi_rows(ctx, G, m, m, nullptr, Jt.p, sig.p, noise_rel, false, fail.p);
        launch_scaled_transpose(ctx, Jt.p, sig.p, m, 0, cutoff, P, fail.p);
    } else {
        jacobi_rows(ctx, G, m, m, nullptr, Jt.p, sig.p, noise_rel);
        launch_scaled_transpose(ctx, Jt.p, sig.p, m, 0, cutoff, P);
    }
}

}  // namespace petal
