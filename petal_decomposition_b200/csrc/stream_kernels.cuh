// Streaming (row-sharded, HBM-bound) kernels of the fit/transform hot path - SIMT engine.
//
//   colsum   : column sums                      (reference: mean_axis, src/pca.rs:207,521; src/ica.rs:174)
//   xb       : Y = (A - mu) * B + bias          (X*Omega, X*P, transform, inverse_transform;
//                                                src/pca.rs:707,714,745,806; src/ica.rs:130,332)
//   atb      : C += (A - mua)^T (B - mub)       (Gram / X^T*Q / Q^T*X; src/pca.rs:681,711; src/ica.rs:333)
//   nonlin   : g(u), sum g'(u)                  (logcosh, src/ica.rs:383-398; exp / cube extensions)
//   colabsmax: per column (max |.|, first row, sign)   (svd_flip, src/pca.rs:815-850)
//
// The centred copy of X that the reference materialises (src/pca.rs:217,531; src/ica.rs:178-188)
// never exists here: `mu` is subtracted while a tile is staged into shared memory.
// All kernels take row-major inputs with explicit leading dimensions, are fully bounds-guarded
// (ragged / tiny shapes), use 128-bit loads when the layout allows (ALIGNED), and reduce across
// CTAs with f64 atomics.
#pragma once
#include "common.cuh"

namespace petal {

// ------------------------------------------------------------------------------------------
// colsum
// ------------------------------------------------------------------------------------------
template <typename T, bool VECLOAD>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ X, int64_t n, int64_t d, int64_t ld, double* __restrict__ sum,
              int tx, int64_t rows_per_cta) {
    constexpr int V = VECLOAD ? Pack<T>::N : 1;
    const int ty = 256 / tx;
    const int cx = threadIdx.x % tx;
    const int ry = threadIdx.x / tx;
    const int64_t col0 = ((int64_t)blockIdx.x * tx + cx) * V;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
    const int64_t r1 = min(n, r0 + rows_per_cta);
    double acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = 0.0;
    if (col0 < d) {
        const T* base = X + col0;
        int64_t r = r0 + ry;
        if constexpr (VECLOAD) {
            for (; r + 3 * (int64_t)ty < r1; r += 4 * (int64_t)ty) {
                Pack<T> a0 = *reinterpret_cast<const Pack<T>*>(base + r * ld);
                Pack<T> a1 = *reinterpret_cast<const Pack<T>*>(base + (r + ty) * ld);
                Pack<T> a2 = *reinterpret_cast<const Pack<T>*>(base + (r + 2 * ty) * ld);
                Pack<T> a3 = *reinterpret_cast<const Pack<T>*>(base + (r + 3 * ty) * ld);
#pragma unroll
                for (int v = 0; v < V; ++v)
                    acc[v] += ((double)a0.v[v] + (double)a1.v[v]) + ((double)a2.v[v] + (double)a3.v[v]);
            }
            for (; r < r1; r += ty) {
                Pack<T> a0 = *reinterpret_cast<const Pack<T>*>(base + r * ld);
#pragma unroll
                for (int v = 0; v < V; ++v) acc[v] += (double)a0.v[v];
            }
        } else {
            for (; r + 3 * (int64_t)ty < r1; r += 4 * (int64_t)ty) {
                T a0 = base[r * ld], a1 = base[(r + ty) * ld], a2 = base[(r + 2 * ty) * ld],
                  a3 = base[(r + 3 * ty) * ld];
                acc[0] += ((double)a0 + (double)a1) + ((double)a2 + (double)a3);
            }
            for (; r < r1; r += ty) acc[0] += (double)base[r * ld];
        }
    }
    __shared__ double red[256 * 4];
#pragma unroll
    for (int v = 0; v < V; ++v) red[(ry * tx + cx) * V + v] = acc[v];
    __syncthreads();
    if (ry == 0 && col0 < d) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
            double s = 0.0;
            for (int y = 0; y < ty; ++y) s += red[(y * tx + cx) * V + v];
            atomicAdd(&sum[col0 + v], s);
        }
    }
}

inline int pow2_ceil(int64_t x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

// sum[d] (f64) += column sums of X. `sum` must be zeroed by the caller.
template <typename T>
void launch_colsum(petal_ctx* ctx, const T* X, int64_t n, int64_t d, int64_t ld, double* sum) {
    if (n == 0 || d == 0) return;
    constexpr int V = Pack<T>::N;
    bool vec = (d % V == 0) && (ld % V == 0) && is_aligned16(X);
    int64_t cols = vec ? d / V : d;
    int tx = (int)std::min<int64_t>(256, pow2_ceil(cols));
    int ty = 256 / tx;
    int64_t gx = ceil_div(cols, tx);
    int64_t target_ctas = (int64_t)ctx->sm_count * 8;
    int64_t gy = std::max<int64_t>(1, std::min<int64_t>(target_ctas / gx, ceil_div(n, (int64_t)ty * 8)));
    gy = std::min<int64_t>(gy, 65535);
    int64_t rows_per_cta = ceil_div(n, gy);
    gy = ceil_div(n, rows_per_cta);
    dim3 grid((unsigned)gx, (unsigned)gy);
    KTimer kt(ctx, kname<T>("colsum_f32", "colsum_f64"), (double)n * d * sizeof(T));
    if (vec)
        colsum_kernel<T, true><<<grid, 256, 0, ctx->stream>>>(X, n, d, ld, sum, tx, rows_per_cta);
    else
        colsum_kernel<T, false><<<grid, 256, 0, ctx->stream>>>(X, n, d, ld, sum, tx, rows_per_cta);
    check_launch(ctx);
}

// ------------------------------------------------------------------------------------------
// xb : Y[n x L] = (A[n x K] - mu) * B + bias        (B is K x L row-major, or L x K if b_trans)
// ------------------------------------------------------------------------------------------
template <typename T>
struct XbParams {
    const T* A;
    int64_t lda;
    int64_t n;
    int64_t K;
    const T* B;
    int64_t ldb;
    int b_trans;
    int64_t L;
    const T* mu;    // nullable [K]
    const T* bias;  // nullable [L]
    T* Y;
    int64_t ldy;
    double* sumsq;  // nullable: += sum((A - mu)^2)
};

template <typename T, int TN, bool ALIGNED>
__global__ void __launch_bounds__(256) xb_kernel(XbParams<T> p) {
    constexpr int V = Pack<T>::N;
    constexpr int BM = 128, TM = 8, BN = 16 * TN, BK = 8 * V;
    constexpr int LDS_A = BK + V;
    constexpr int A_VECS = BM * BK / V / 256;  // 4
    constexpr int B_ELEMS = (BK * BN + 255) / 256;
    __shared__ __align__(16) T As[BM * LDS_A];
    __shared__ __align__(16) T Bs[BK * BN];
    __shared__ double red[8];

    const int tid = threadIdx.x;
    const int tr = tid >> 4, tc = tid & 15;
    const int64_t row0 = (int64_t)blockIdx.x * BM;
    const int64_t col0 = (int64_t)blockIdx.y * BN;
    const bool want_ss = (p.sumsq != nullptr) && (blockIdx.y == 0);

    T acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = T(0);

    Pack<T> a_stage[A_VECS];
    T b_stage[B_ELEMS];
    double ss = 0.0;

    auto load_tiles = [&](int64_t k0) {
#pragma unroll
        for (int i = 0; i < A_VECS; ++i) {
            int idx = tid + i * 256;
            int m = idx >> 3;  // BK / V == 8 vectors per row
            int kv = idx & 7;
            int64_t r = row0 + m;
            int64_t kk = k0 + kv * V;
            Pack<T> z;
#pragma unroll
            for (int v = 0; v < V; ++v) z.v[v] = T(0);
            if (r < p.n && kk < p.K) {
                if constexpr (ALIGNED) {
                    z = *reinterpret_cast<const Pack<T>*>(p.A + r * p.lda + kk);
                    if (p.mu) {
                        Pack<T> m4 = *reinterpret_cast<const Pack<T>*>(p.mu + kk);
#pragma unroll
                        for (int v = 0; v < V; ++v) z.v[v] -= m4.v[v];
                    }
                } else {
#pragma unroll
                    for (int v = 0; v < V; ++v)
                        if (kk + v < p.K) {
                            T x = p.A[r * p.lda + kk + v];
                            if (p.mu) x -= p.mu[kk + v];
                            z.v[v] = x;
                        }
                }
                if (want_ss) {
#pragma unroll
                    for (int v = 0; v < V; ++v) ss += (double)z.v[v] * (double)z.v[v];
                }
            }
            a_stage[i] = z;
        }
#pragma unroll
        for (int i = 0; i < B_ELEMS; ++i) {
            int idx = tid + i * 256;
            T val = T(0);
            if (idx < BK * BN) {
                int kk = idx / BN, c = idx % BN;
                int64_t gk = k0 + kk, gc = col0 + c;
                if (gk < p.K && gc < p.L) val = p.b_trans ? p.B[gc * p.ldb + gk] : p.B[gk * p.ldb + gc];
            }
            b_stage[i] = val;
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int i = 0; i < A_VECS; ++i) {
            int idx = tid + i * 256;
            int m = idx >> 3, kv = idx & 7;
            *reinterpret_cast<Pack<T>*>(&As[m * LDS_A + kv * V]) = a_stage[i];
        }
#pragma unroll
        for (int i = 0; i < B_ELEMS; ++i) {
            int idx = tid + i * 256;
            if (idx < BK * BN) Bs[idx] = b_stage[i];
        }
    };

    load_tiles(0);
    store_tiles();
    __syncthreads();
    for (int64_t k0 = 0; k0 < p.K; k0 += BK) {
        const bool has_next = (k0 + BK) < p.K;
        if (has_next) load_tiles(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; kk += V) {
            Pack<T> a[TM];
#pragma unroll
            for (int i = 0; i < TM; ++i)
                a[i] = *reinterpret_cast<const Pack<T>*>(&As[(tr * TM + i) * LDS_A + kk]);
            T b[V][TN];
#pragma unroll
            for (int v = 0; v < V; ++v)
#pragma unroll
                for (int j = 0; j < TN; ++j) b[v][j] = Bs[(kk + v) * BN + tc + 16 * j];
#pragma unroll
            for (int v = 0; v < V; ++v)
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] += a[i].v[v] * b[v][j];
        }
        __syncthreads();
        if (has_next) {
            store_tiles();
            __syncthreads();
        }
    }

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int64_t r = row0 + tr * TM + i;
        if (r >= p.n) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int64_t c = col0 + tc + 16 * j;
            if (c < p.L) {
                T v = acc[i][j];
                if (p.bias) v += p.bias[c];
                p.Y[r * p.ldy + c] = v;
            }
        }
    }
    if (want_ss) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if ((tid & 31) == 0) red[tid >> 5] = ss;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < 8; ++w) s += red[w];
            atomicAdd(p.sumsq, s);
        }
    }
}

template <typename T, int TN>
void launch_xb_tn(petal_ctx* ctx, const XbParams<T>& p, bool aligned) {
    dim3 grid((unsigned)ceil_div(p.n, 128), (unsigned)ceil_div(p.L, 16 * TN));
    KTimer kt(ctx, kname<T>("xb_f32", "xb_f64"), (double)p.n * (p.K + p.L) * sizeof(T));
    if (aligned)
        xb_kernel<T, TN, true><<<grid, 256, 0, ctx->stream>>>(p);
    else
        xb_kernel<T, TN, false><<<grid, 256, 0, ctx->stream>>>(p);
    check_launch(ctx);
}

inline bool launch_xb_dmma(petal_ctx* ctx, const XbParams<double>& p);  // FP64 tensor path, defined with the DMMA kernels

template <typename T>
void launch_xb(petal_ctx* ctx, const XbParams<T>& p) {
    if (p.n == 0 || p.L == 0) return;
    if constexpr (sizeof(T) == 8) {
        if (launch_xb_dmma(ctx, p)) return;
    }
    constexpr int V = Pack<T>::N;
    bool aligned = (p.K % V == 0) && (p.lda % V == 0) && is_aligned16(p.A) &&
                   (p.mu == nullptr || is_aligned16(p.mu));
    // smallest column tile that covers L in one pass over A (A is then read exactly once)
    if (p.L <= 16) launch_xb_tn<T, 1>(ctx, p, aligned);
    else if (p.L <= 32) launch_xb_tn<T, 2>(ctx, p, aligned);
    else if (p.L <= 48) launch_xb_tn<T, 3>(ctx, p, aligned);
    else if (p.L <= 64) launch_xb_tn<T, 4>(ctx, p, aligned);
    else if (p.L <= 80) launch_xb_tn<T, 5>(ctx, p, aligned);
    else if (p.L <= 96) launch_xb_tn<T, 6>(ctx, p, aligned);
    else launch_xb_tn<T, 8>(ctx, p, aligned);
}

// ------------------------------------------------------------------------------------------
// atb : C[da x db] (f64) += (A - mua)^T (B - mub), reduction over the n rows, split across CTAs
// ------------------------------------------------------------------------------------------
template <typename T>
struct AtbParams {
    const T* A;
    int64_t lda;
    int64_t da;
    const T* mua;  // nullable
    const T* B;
    int64_t ldb;
    int64_t db;
    const T* mub;  // nullable
    int64_t n;
    int64_t rows_per_cta;
    double* C;
    int64_t ldc;
    int symmetric;  // A == B: compute tiles with tile_j >= tile_i only (mirror afterwards)
    int tiles_j;
    int negate;     // C -= A^T B (trailing update of the blocked Cholesky)
};

template <typename T, int TI, int TJ, bool ALIGNED>
__global__ void __launch_bounds__(256) atb_kernel(AtbParams<T> p) {
    constexpr int V = Pack<T>::N;
    constexpr int BI = 16 * TI, BJ = 16 * TJ, BR = 16;
    constexpr bool VEC_I = (TI % V == 0);
    constexpr bool VEC_J = (TJ % V == 0);
    constexpr int64_t FLUSH_ROWS = (sizeof(T) == 4) ? 8192 : (int64_t(1) << 62);
    static_assert(VEC_I, "TI must be a multiple of the vector width");
    constexpr int A_VECS = BR * BI / V / 256;
    constexpr int B_ELEMS = BR * BJ / 256;  // == TJ
    __shared__ __align__(16) T As[BR * BI];
    __shared__ __align__(16) T Bs[BR * BJ];

    const int tile_i = blockIdx.x / p.tiles_j, tile_j = blockIdx.x % p.tiles_j;
    if (p.symmetric && tile_j < tile_i) return;
    const int64_t i0 = (int64_t)tile_i * BI, j0 = (int64_t)tile_j * BJ;
    const int tid = threadIdx.x;
    const int ti = tid >> 4, tj = tid & 15;
    const int64_t r_begin = (int64_t)blockIdx.y * p.rows_per_cta;
    const int64_t r_end = min(p.n, r_begin + p.rows_per_cta);
    if (r_begin >= r_end) return;

    T acc[TI][TJ];
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TJ; ++j) acc[i][j] = T(0);

    Pack<T> a_stage[A_VECS];
    T b_stage[B_ELEMS];

    auto load_tiles = [&](int64_t r0) {
#pragma unroll
        for (int q = 0; q < A_VECS; ++q) {
            int idx = tid + q * 256;
            int rr = idx / (BI / V), cv = idx % (BI / V);
            int64_t r = r0 + rr, c = i0 + cv * V;
            Pack<T> z;
#pragma unroll
            for (int v = 0; v < V; ++v) z.v[v] = T(0);
            if (r < r_end && c < p.da) {
                if constexpr (ALIGNED) {
                    z = *reinterpret_cast<const Pack<T>*>(p.A + r * p.lda + c);
                    if (p.mua) {
                        Pack<T> m4 = *reinterpret_cast<const Pack<T>*>(p.mua + c);
#pragma unroll
                        for (int v = 0; v < V; ++v) z.v[v] -= m4.v[v];
                    }
                } else {
#pragma unroll
                    for (int v = 0; v < V; ++v)
                        if (c + v < p.da) {
                            T x = p.A[r * p.lda + c + v];
                            if (p.mua) x -= p.mua[c + v];
                            z.v[v] = x;
                        }
                }
            }
            a_stage[q] = z;
        }
#pragma unroll
        for (int q = 0; q < B_ELEMS; ++q) {
            int idx = tid + q * 256;
            int rr = idx / BJ, cc = idx % BJ;
            int64_t r = r0 + rr, c = j0 + cc;
            T x = T(0);
            if (r < r_end && c < p.db) {
                x = p.B[r * p.ldb + c];
                if (p.mub) x -= p.mub[c];
            }
            b_stage[q] = x;
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int q = 0; q < A_VECS; ++q) {
            int idx = tid + q * 256;
            *reinterpret_cast<Pack<T>*>(&As[idx * V]) = a_stage[q];
        }
#pragma unroll
        for (int q = 0; q < B_ELEMS; ++q) Bs[tid + q * 256] = b_stage[q];
    };
    auto flush = [&]() {
#pragma unroll
        for (int i = 0; i < TI; ++i) {
            int64_t gi = i0 + ti * TI + i;
#pragma unroll
            for (int j = 0; j < TJ; ++j) {
                int64_t gj;
                if constexpr (VEC_J) gj = j0 + (j / V) * (16 * V) + tj * V + (j % V);
                else gj = j0 + tj + 16 * j;
                if (gi < p.da && gj < p.db) atomicAdd(&p.C[gi * p.ldc + gj], p.negate ? -(double)acc[i][j] : (double)acc[i][j]);
                acc[i][j] = T(0);
            }
        }
    };

    load_tiles(r_begin);
    store_tiles();
    __syncthreads();
    int64_t since_flush = 0;
    for (int64_t r0 = r_begin; r0 < r_end; r0 += BR) {
        const bool has_next = (r0 + BR) < r_end;
        if (has_next) load_tiles(r0 + BR);
#pragma unroll
        for (int rr = 0; rr < BR; ++rr) {
            T a[TI], b[TJ];
#pragma unroll
            for (int i = 0; i < TI; i += V) {
                Pack<T> t = *reinterpret_cast<const Pack<T>*>(&As[rr * BI + ti * TI + i]);
#pragma unroll
                for (int v = 0; v < V; ++v) a[i + v] = t.v[v];
            }
            if constexpr (VEC_J) {
#pragma unroll
                for (int j = 0; j < TJ; j += V) {
                    Pack<T> t = *reinterpret_cast<const Pack<T>*>(&Bs[rr * BJ + (j / V) * (16 * V) + tj * V]);
#pragma unroll
                    for (int v = 0; v < V; ++v) b[j + v] = t.v[v];
                }
            } else {
#pragma unroll
                for (int j = 0; j < TJ; ++j) b[j] = Bs[rr * BJ + tj + 16 * j];
            }
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int j = 0; j < TJ; ++j) acc[i][j] += a[i] * b[j];
        }
        since_flush += BR;
        if (since_flush >= FLUSH_ROWS) {
            flush();
            since_flush = 0;
        }
        __syncthreads();
        if (has_next) {
            store_tiles();
            __syncthreads();
        }
    }
    flush();
}

template <typename T, int TI, int TJ>
void launch_atb_t(petal_ctx* ctx, AtbParams<T> p, bool aligned) {
    constexpr int BI = 16 * TI, BJ = 16 * TJ;
    int64_t tiles_i = ceil_div(p.da, BI), tiles_j = ceil_div(p.db, BJ);
    p.tiles_j = (int)tiles_j;
    int64_t tiles = tiles_i * tiles_j;
    int64_t eff_tiles = p.symmetric ? (tiles_i * (tiles_i + 1)) / 2 : tiles;
    int64_t target = (int64_t)ctx->sm_count * 4;
    int64_t gy = std::max<int64_t>(1, target / std::max<int64_t>(1, eff_tiles));
    gy = std::min<int64_t>(gy, ceil_div(p.n, 64));
    gy = std::min<int64_t>(std::max<int64_t>(gy, 1), 65535);
    int64_t rows = ceil_div(ceil_div(p.n, gy), 16) * 16;
    p.rows_per_cta = rows;
    gy = ceil_div(p.n, rows);
    dim3 grid((unsigned)tiles, (unsigned)gy);
    KTimer kt(ctx, kname<T>("atb_f32", "atb_f64"),
              (double)p.n * ((p.symmetric ? 0 : p.da) + p.db) * sizeof(T));
    if (aligned)
        atb_kernel<T, TI, TJ, true><<<grid, 256, 0, ctx->stream>>>(p);
    else
        atb_kernel<T, TI, TJ, false><<<grid, 256, 0, ctx->stream>>>(p);
    check_launch(ctx);
}


// ------------------------------------------------------------------------------------------
// atb on the FP64 tensor path (DMMA, mma.sync.m8n8k4.f64): C[da x db] (f64) += (A - mua)^T (B - mub)
// 128 x 128 output tile per CTA over a slice of rows; 8 warps as 4 (i) x 2 (j), each 32 x 64 = 4 x 8 DMMA tiles.
// Shared tiles keep the natural [row][feature] layout with a 64 B row skew so that the four K rows a fragment
// load touches fall into both halves of the bank space (2 wavefronts = the minimum for 256 B).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

template <bool ALIGNED>
__global__ void __launch_bounds__(256) atb_dmma_kernel(AtbParams<double> p) {
    constexpr int BI = 128, BJ = 128, BR = 16, LDS_ = BI + 8;
    constexpr int A_VECS = BR * BI / 2 / 256;  // 4 double2 per thread per operand
    __shared__ __align__(16) double As[BR * LDS_];
    __shared__ __align__(16) double Bs[BR * LDS_];

    const int tile_i = blockIdx.x / p.tiles_j, tile_j = blockIdx.x % p.tiles_j;
    if (p.symmetric && tile_j < tile_i) return;
    const int64_t i0 = (int64_t)tile_i * BI, j0 = (int64_t)tile_j * BJ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wi = warp >> 1, wj = warp & 1;
    const int64_t r_begin = (int64_t)blockIdx.y * p.rows_per_cta;
    const int64_t r_end = min(p.n, r_begin + p.rows_per_cta);
    if (r_begin >= r_end) return;
    const bool same = p.symmetric && (tile_i == tile_j);  // diagonal tile: B tile == A tile

    double acc[4][8][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    // The global loads of the next row chunk are issued before the MMA loop and consumed after it: nothing may
    // touch the staged registers in between (an earlier version subtracted the column means right after the load,
    // which made every warp wait for HBM before its DMMAs: 32 % of the stall samples, r02 ncu).  The means are
    // subtracted when the chunk goes to shared memory; a thread's columns do not change, so its means live in
    // registers.
    Pack<double> a_stage[A_VECS], b_stage[A_VECS];
    Pack<double> a_mu[A_VECS], b_mu[A_VECS];
#pragma unroll
    for (int q = 0; q < A_VECS; ++q) {
        const int cv = (tid + q * 256) % (BI / 2);
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            const int64_t ca = i0 + cv * 2 + v, cb = j0 + cv * 2 + v;
            a_mu[q].v[v] = (p.mua && ca < p.da) ? p.mua[ca] : 0.0;
            b_mu[q].v[v] = (p.mub && cb < p.db) ? p.mub[cb] : 0.0;
        }
    }
    auto load_one = [&](const double* M, int64_t ld, int64_t ncol, int64_t c0, int64_t r0, Pack<double>* stage) {
#pragma unroll
        for (int q = 0; q < A_VECS; ++q) {
            int idx = tid + q * 256;
            int rr = idx / (BI / 2), cv = idx % (BI / 2);
            int64_t r = r0 + rr, c = c0 + cv * 2;
            Pack<double> z;
            z.v[0] = z.v[1] = 0.0;
            if (r < r_end && c < ncol) {
                if constexpr (ALIGNED) {
                    z = *reinterpret_cast<const Pack<double>*>(M + r * ld + c);
                } else {
#pragma unroll
                    for (int v = 0; v < 2; ++v)
                        if (c + v < ncol) z.v[v] = M[r * ld + c + v];
                }
            }
            stage[q] = z;
        }
    };
    auto load_tiles = [&](int64_t r0) {
        load_one(p.A, p.lda, p.da, i0, r0, a_stage);
        if (!same) load_one(p.B, p.ldb, p.db, j0, r0, b_stage);
    };
    auto store_tiles = [&](int64_t r0) {
#pragma unroll
        for (int q = 0; q < A_VECS; ++q) {
            int idx = tid + q * 256;
            int rr = idx / (BI / 2), cv = idx % (BI / 2);
            const bool row_ok = (r0 + rr) < r_end;  // rows past the slice stay exactly zero
            Pack<double> a = a_stage[q];
#pragma unroll
            for (int v = 0; v < 2; ++v)
                if (row_ok && i0 + cv * 2 + v < p.da) a.v[v] -= a_mu[q].v[v];
            *reinterpret_cast<Pack<double>*>(&As[rr * LDS_ + cv * 2]) = a;
            if (!same) {
                Pack<double> b = b_stage[q];
#pragma unroll
                for (int v = 0; v < 2; ++v)
                    if (row_ok && j0 + cv * 2 + v < p.db) b.v[v] -= b_mu[q].v[v];
                *reinterpret_cast<Pack<double>*>(&Bs[rr * LDS_ + cv * 2]) = b;
            }
        }
    };

    load_tiles(r_begin);
    store_tiles(r_begin);
    __syncthreads();
    const double* Bt = same ? As : Bs;
    const int kq = lane & 3, rq = lane >> 2;
    for (int64_t r0 = r_begin; r0 < r_end; r0 += BR) {
        const bool has_next = (r0 + BR) < r_end;
        if (has_next) load_tiles(r0 + BR);
#pragma unroll
        for (int k4 = 0; k4 < BR / 4; ++k4) {
            double af[4], bf[8];
#pragma unroll
            for (int a = 0; a < 4; ++a) af[a] = As[(k4 * 4 + kq) * LDS_ + wi * 32 + a * 8 + rq];
#pragma unroll
            for (int b = 0; b < 8; ++b) bf[b] = Bt[(k4 * 4 + kq) * LDS_ + wj * 64 + b * 8 + rq];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) dmma_m8n8k4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
        __syncthreads();
        if (has_next) {
            store_tiles(r0 + BR);
            __syncthreads();
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int64_t gi = i0 + wi * 32 + a * 8 + rq;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int64_t gj = j0 + wj * 64 + b * 8 + 2 * kq;
            if (gi < p.da) {
                if (gj < p.db) atomicAdd(&p.C[gi * p.ldc + gj], p.negate ? -acc[a][b][0] : acc[a][b][0]);
                if (gj + 1 < p.db) atomicAdd(&p.C[gi * p.ldc + gj + 1], p.negate ? -acc[a][b][1] : acc[a][b][1]);
            }
        }
    }
}

inline void launch_atb_dmma(petal_ctx* ctx, AtbParams<double> p, bool aligned) {
    int64_t tiles_i = ceil_div(p.da, 128), tiles_j = ceil_div(p.db, 128);
    p.tiles_j = (int)tiles_j;
    int64_t tiles = tiles_i * tiles_j;
    int64_t eff_tiles = p.symmetric ? (tiles_i * (tiles_i + 1)) / 2 : tiles;
    int64_t target = (int64_t)ctx->sm_count * 2;
    int64_t gy = std::max<int64_t>(1, target / std::max<int64_t>(1, eff_tiles));
    gy = std::min<int64_t>(gy, ceil_div(p.n, 256));
    gy = std::min<int64_t>(std::max<int64_t>(gy, 1), 65535);
    int64_t rows = ceil_div(ceil_div(p.n, gy), 16) * 16;
    p.rows_per_cta = rows;
    gy = ceil_div(p.n, rows);
    dim3 grid((unsigned)tiles, (unsigned)gy);
    KTimer kt(ctx, "atb_dmma_f64", (double)p.n * ((p.symmetric ? 0 : p.da) + p.db) * sizeof(double));
    if (aligned)
        atb_dmma_kernel<true><<<grid, 256, 0, ctx->stream>>>(p);
    else
        atb_dmma_kernel<false><<<grid, 256, 0, ctx->stream>>>(p);
    check_launch(ctx);
}

// C must be zeroed by the caller (the kernel accumulates).
template <typename T>
void launch_atb(petal_ctx* ctx, AtbParams<T> p) {
    if (p.n == 0 || p.da == 0 || p.db == 0) return;
    constexpr int V = Pack<T>::N;
    bool aligned = (p.da % V == 0) && (p.lda % V == 0) && is_aligned16(p.A) &&
                   (p.mua == nullptr || is_aligned16(p.mua));
    if constexpr (sizeof(T) == 8) {
        // FP64 tensor path for Gram-shaped work (both output dimensions fill a good part of a 128 x 128 tile)
        if (ctx->f64_engine == 1 && p.da >= 96 && p.db >= 96 && p.n >= 64) {
            bool al2 = aligned && (p.db % V == 0) && (p.ldb % V == 0) && is_aligned16(p.B);
            launch_atb_dmma(ctx, p, al2);
            return;
        }
    }
    if (p.symmetric) {
        launch_atb_t<T, 8, 8>(ctx, p, aligned);
        return;
    }
    if (p.db <= 16) launch_atb_t<T, 8, 1>(ctx, p, aligned);
    else if (p.db <= 32) launch_atb_t<T, 8, 2>(ctx, p, aligned);
    else if (p.db <= 48) launch_atb_t<T, 8, 3>(ctx, p, aligned);
    else if (p.db <= 64) launch_atb_t<T, 8, 4>(ctx, p, aligned);
    else if (p.db <= 80) launch_atb_t<T, 8, 5>(ctx, p, aligned);
    else if (p.db <= 96) launch_atb_t<T, 8, 6>(ctx, p, aligned);
    else launch_atb_t<T, 8, 8>(ctx, p, aligned);
}


// ------------------------------------------------------------------------------------------
// Panel-major Y ([row block of 32][np][32] floats, np = columns rounded up to 16; see tc_kernels.cuh):
//   panel_gram : G[np x np] (f64) += Y^T Y
//   panel_xb   : out[n x k] (row-major) = Y * S,  S[l x k] row-major
// Skinny, Y-sized passes of the randomized-PCA epilogue (G1 = Y^T Y, scores = Q U_B Sigma).
// ------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------
// xb on the FP64 tensor path (DMMA): Y[n x L] = (A - mu) * B (+ optional sum of squares of the centred A).
// CTA = 128 rows x 64 output columns (grid.y walks the column blocks), 8 warps of 16 rows x 64 columns = 2 x 8 DMMA
// tiles; K in chunks of 16 staged through registers (the next chunk's global loads overlap the current chunk's
// MMAs).  Row pitches 20 / 68 doubles keep every fragment load of a half-warp on 16 distinct bank pairs.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) xb_dmma_kernel(XbParams<double> p) {
    constexpr int BM = 128, BN = 64, KC = 16, LDA = KC + 4, LDB = BN + 4;
    __shared__ __align__(16) double As[BM * LDA];
    __shared__ __align__(16) double Bs[KC * LDB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kq = lane & 3, rq = lane >> 2;
    const int64_t r0 = (int64_t)blockIdx.x * BM;
    const int64_t c0 = (int64_t)blockIdx.y * BN;
    double acc[2][8][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    // staging registers: A tile 128 x 16 = 2048 doubles -> 8 per thread (row = tid / 2, 8 consecutive k);
    // B chunk 16 x 64 = 1024 doubles -> 4 per thread (k = tid / 16, 4 consecutive columns)
    double a_st[8], b_st[4];
    const int ar = tid >> 1, ak = (tid & 1) * 8;
    const int bk = tid >> 4, bc = (tid & 15) * 4;
    double ss = 0.0;
    auto load_chunk = [&](int64_t k0) {
        const int64_t r = r0 + ar;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t k = k0 + ak + j;
            double v = 0.0;
            if (r < p.n && k < p.K) v = p.A[r * p.lda + k];  // centred when stored (keeps the prefetch asynchronous)
            a_st[j] = v;
        }
        const int64_t k = k0 + bk;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t c = c0 + bc + j;
            double v = 0.0;
            if (k < p.K && c < p.L) v = p.b_trans ? p.B[c * p.ldb + k] : p.B[k * p.ldb + c];
            b_st[j] = v;
        }
    };
    auto store_chunk = [&](int64_t k0) {
        const bool row_ok = (r0 + ar) < p.n;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t k = k0 + ak + j;
            const double v = (p.mu && row_ok && k < p.K) ? a_st[j] - p.mu[k] : a_st[j];
            As[ar * LDA + ak + j] = v;
            ss += v * v;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) Bs[bk * LDB + bc + j] = b_st[j];
    };
    load_chunk(0);
    store_chunk(0);
    __syncthreads();
    for (int64_t k0 = 0; k0 < p.K; k0 += KC) {
        const bool has_next = (k0 + KC) < p.K;
        if (has_next) load_chunk(k0 + KC);
#pragma unroll
        for (int k4 = 0; k4 < KC / 4; ++k4) {
            double af[2], bf[8];
#pragma unroll
            for (int a = 0; a < 2; ++a) af[a] = As[(warp * 16 + a * 8 + rq) * LDA + k4 * 4 + kq];
#pragma unroll
            for (int b = 0; b < 8; ++b) bf[b] = Bs[(k4 * 4 + kq) * LDB + b * 8 + rq];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) dmma_m8n8k4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
        __syncthreads();
        if (has_next) {
            store_chunk(k0 + KC);
            __syncthreads();
        }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int64_t r = r0 + warp * 16 + a * 8 + rq;
        if (r < p.n) {
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const int64_t c = c0 + b * 8 + 2 * kq;
                if (c < p.L) p.Y[r * p.ldy + c] = acc[a][b][0];
                if (c + 1 < p.L) p.Y[r * p.ldy + c + 1] = acc[a][b][1];
            }
        }
    }
    if (p.sumsq != nullptr && blockIdx.y == 0) {
        ss = warp_sum(ss);
        __shared__ double red[8];
        if (lane == 0) red[warp] = ss;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += red[w];
            atomicAdd(p.sumsq, t);
        }
    }
}

inline bool launch_xb_dmma(petal_ctx* ctx, const XbParams<double>& p) {
    if (ctx->f64_engine != 1 || p.bias != nullptr || p.n < 256 || p.K < 16 || p.L < 8) return false;
    dim3 grid((unsigned)ceil_div(p.n, 128), (unsigned)ceil_div(p.L, 64));
    KTimer kt(ctx, "xb_dmma_f64", (double)p.n * (p.K + p.L) * sizeof(double));
    xb_dmma_kernel<<<grid, 256, 0, ctx->stream>>>(p);
    check_launch(ctx);
    return true;
}

struct AbsMax {
    double a;     // |value|
    double sgn;   // +1 / -1 (f64::signum semantics: -0.0 -> -1)
    int64_t idx;  // row index (local)
};

__device__ __forceinline__ bool absmax_better(const AbsMax& x, const AbsMax& y) {
    // first maximum wins (reference src/pca.rs:830 `abs <= absmax -> continue`)
    return x.a > y.a || (x.a == y.a && x.idx < y.idx);
}

__global__ void colabsmax_final_kernel(const AbsMax* __restrict__ partial, int64_t chunks, int64_t k,
                                       double* __restrict__ out3);

template <int NP>
__global__ void __launch_bounds__(256)
panel_gram_kernel(const float* __restrict__ Yp, int64_t nblocks, int64_t blocks_per_cta, double* __restrict__ G) {
    constexpr int TI = NP / 16;  // outputs per thread per dimension (16 x 16 thread grid, interleaved)
    __shared__ float Ts[32][NP + 1];  // [row in block][column]
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int64_t b0 = (int64_t)blockIdx.x * blocks_per_cta;
    const int64_t b1 = min(nblocks, b0 + blocks_per_cta);
    float acc[TI][TI];
    double dacc[TI][TI];
#pragma unroll
    for (int u = 0; u < TI; ++u)
#pragma unroll
        for (int v = 0; v < TI; ++v) {
            acc[u][v] = 0.f;
            dacc[u][v] = 0.0;
        }
    int since = 0;
    auto flush = [&]() {  // fp32 partial sums (<= 2048 rows) -> f64 registers
#pragma unroll
        for (int u = 0; u < TI; ++u)
#pragma unroll
            for (int v = 0; v < TI; ++v) {
                dacc[u][v] += (double)acc[u][v];
                acc[u][v] = 0.f;
            }
    };
    for (int64_t b = b0; b < b1; ++b) {
        const float* src = Yp + b * NP * 32;
        __syncthreads();
        for (int e = tid; e < NP * 32; e += 256) Ts[e & 31][e >> 5] = src[e];  // coalesced read, transposed store
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < 32; ++i) {
            float a[TI], bb[TI];
#pragma unroll
            for (int u = 0; u < TI; ++u) {
                a[u] = Ts[i][ty + 16 * u];
                bb[u] = Ts[i][tx + 16 * u];
            }
#pragma unroll
            for (int u = 0; u < TI; ++u)
#pragma unroll
                for (int v = 0; v < TI; ++v) acc[u][v] += a[u] * bb[v];
        }
        if (++since == 64) {
            flush();
            since = 0;
        }
    }
    flush();
#pragma unroll
    for (int u = 0; u < TI; ++u)
#pragma unroll
        for (int v = 0; v < TI; ++v) atomicAdd(&G[(ty + 16 * u) * NP + tx + 16 * v], dacc[u][v]);
}

// G must be zeroed by the caller; np in {16, 32, ..., 128}
inline void launch_panel_gram(petal_ctx* ctx, const float* Yp, int64_t n, int np, double* G) {
    const int64_t nblocks = ceil_div(n, 32);
    int64_t ctas = std::min<int64_t>(nblocks, (int64_t)ctx->sm_count * 4);
    const int64_t per = ceil_div(nblocks, ctas);
    ctas = ceil_div(nblocks, per);
    KTimer kt(ctx, "panel_gram_f32", (double)n * np * sizeof(float));
    switch (np) {
        case 16: panel_gram_kernel<16><<<(unsigned)ctas, 256, 0, ctx->stream>>>(Yp, nblocks, per, G); break;
        case 32: panel_gram_kernel<32><<<(unsigned)ctas, 256, 0, ctx->stream>>>(Yp, nblocks, per, G); break;
        case 48: panel_gram_kernel<48><<<(unsigned)ctas, 256, 0, ctx->stream>>>(Yp, nblocks, per, G); break;
        case 64: panel_gram_kernel<64><<<(unsigned)ctas, 256, 0, ctx->stream>>>(Yp, nblocks, per, G); break;
        case 80: panel_gram_kernel<80><<<(unsigned)ctas, 256, 0, ctx->stream>>>(Yp, nblocks, per, G); break;
        case 96: panel_gram_kernel<96><<<(unsigned)ctas, 256, 0, ctx->stream>>>(Yp, nblocks, per, G); break;
        case 112: panel_gram_kernel<112><<<(unsigned)ctas, 256, 0, ctx->stream>>>(Yp, nblocks, per, G); break;
        default: panel_gram_kernel<128><<<(unsigned)ctas, 256, 0, ctx->stream>>>(Yp, nblocks, per, G); break;
    }
    check_launch(ctx);
}

// out[r][j] = sum_c Y[r][c] * S[c][j]; CTAs loop over pairs of row blocks (64 rows); S staged once in shared
// memory (l x kp floats, kp = k rounded up to 32*KV).  256 threads = 8 (row groups) x 32 (column groups): each thread
// owns the 8 consecutive rows 8 i0 .. 8 i0 + 7 (two broadcast LDS.128 per c) and the KV contiguous columns KV*j0 ...
// (one vector load of S per c).
template <int KV>
__global__ void __launch_bounds__(256)
panel_xb_kernel(const float* __restrict__ Yp, int64_t n, int np, int l, const float* __restrict__ S, int k,
                float* __restrict__ out /* nullable */, int64_t ldo, AbsMax* __restrict__ partial /* nullable */) {
    extern __shared__ float psm[];
    constexpr int KP = 32 * KV;
    float* Ss = psm;                   // [l][KP]
    constexpr int TS = 68;             // 64 rows + pad, 16 B aligned rows
    float* Ts = psm + (size_t)l * KP;  // [np][TS]: column c of the 64 rows at c*TS + i
    const int tid = threadIdx.x;
    for (int e = tid; e < l * KP; e += 256) {
        const int c = e / KP, j = e % KP;
        Ss[e] = (j < k) ? S[c * k + j] : 0.f;
    }
    const int64_t nblocks = (n + 31) / 32;
    const int64_t npairs = (nblocks + 1) / 2;
    const int i0 = tid >> 5, j0 = tid & 31;
    // per column (max |score|, first row, sign) over the rows this thread produces: the u-based sign flip
    // (reference src/pca.rs:815-850) needs it, and taking it here saves re-reading the scores
    // (kept in fp32 - the scores are fp32 - and widened at the end: f64 compares per element would cost as much as
    // the product itself)
    float best_a[KV], best_s[KV];
    int64_t best_i[KV];
#pragma unroll
    for (int v = 0; v < KV; ++v) {
        best_a[v] = -1.f;
        best_s[v] = 1.f;
        best_i[v] = INT64_MAX;
    }
    for (int64_t pb = blockIdx.x; pb < npairs; pb += gridDim.x) {
        __syncthreads();
        for (int h = 0; h < 2; ++h) {
            const int64_t b = pb * 2 + h;
            if (b < nblocks) {
                const float* src = Yp + b * np * 32;
                for (int e = tid; e < np * 32; e += 256) Ts[(e >> 5) * TS + h * 32 + (e & 31)] = src[e];
            }
        }
        __syncthreads();
        float acc[8][KV];
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int v = 0; v < KV; ++v) acc[u][v] = 0.f;
        for (int c = 0; c < l; ++c) {
            float y[8], sv[KV];
            {
                const float4 y0 = *reinterpret_cast<const float4*>(Ts + c * TS + 8 * i0);  // warp broadcast
                const float4 y1 = *reinterpret_cast<const float4*>(Ts + c * TS + 8 * i0 + 4);
                y[0] = y0.x; y[1] = y0.y; y[2] = y0.z; y[3] = y0.w;
                y[4] = y1.x; y[5] = y1.y; y[6] = y1.z; y[7] = y1.w;
            }
#pragma unroll
            for (int v = 0; v < KV; ++v) sv[v] = Ss[c * KP + KV * j0 + v];
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int v = 0; v < KV; ++v) acc[u][v] += y[u] * sv[v];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int64_t r = pb * 64 + 8 * i0 + u;
            if (r < n) {
#pragma unroll
                for (int v = 0; v < KV; ++v)
                    if (KV * j0 + v < k) {
                        if (out != nullptr) out[r * ldo + KV * j0 + v] = acc[u][v];
                        // rows are visited in increasing order by this thread: a strict > keeps the first maximum
                        const float a = fabsf(acc[u][v]);
                        if (a > best_a[v]) {
                            best_a[v] = a;
                            best_s[v] = acc[u][v];
                            best_i[v] = r;
                        }
                    }
            }
        }
    }
    if (partial != nullptr) {
        // combine the 8 row groups of the CTA (the operand staging area is free now)
        __syncthreads();
        AbsMax* red = reinterpret_cast<AbsMax*>(psm);  // [8][32 * KV]
#pragma unroll
        for (int v = 0; v < KV; ++v) {
            AbsMax b;
            b.a = (double)best_a[v];
            b.sgn = signbit(best_s[v]) ? -1.0 : 1.0;
            b.idx = best_i[v];
            red[i0 * (32 * KV) + KV * j0 + v] = b;
        }
        __syncthreads();
        if (i0 == 0) {
#pragma unroll
            for (int v = 0; v < KV; ++v) {
                const int col = KV * j0 + v;
                if (col < k) {
                    AbsMax b = red[col];
                    for (int y = 1; y < 8; ++y) {
                        const AbsMax c = red[y * (32 * KV) + col];
                        if (absmax_better(c, b)) b = c;
                    }
                    partial[(int64_t)blockIdx.x * k + col] = b;
                }
            }
        }
    }
}

// Register-tiled variant for k <= 4 * CG (CG = 8, 16): each thread owns 8 consecutive rows x 4 consecutive columns
// (per panel column c: two broadcast LDS.128 of Y, one LDS.128 of S, 32 FFMA - the 8 x KV tiling above spends as many
// cycles on shared-memory loads as on FFMAs).  256 threads = CG column groups x (256 / CG) row groups: tiles of
// 128 (CG = 16) or 256 (CG = 8) rows.
template <int CG>
__global__ void __launch_bounds__(256)
panel_xb4_kernel(const float* __restrict__ Yp, int64_t n, int np, int l, const float* __restrict__ S, int k,
                 float* __restrict__ out /* nullable */, int64_t ldo, AbsMax* __restrict__ partial /* nullable */) {
    extern __shared__ float psm[];
    constexpr int RG = 256 / CG, ROWS = RG * 8, NB = ROWS / 32, TS = ROWS + 4, KP = 4 * CG;
    float* Ss = psm;                   // [l][KP]
    float* Ts = psm + (size_t)l * KP;  // [np][TS]: column c of the tile's rows at c*TS + i
    const int tid = threadIdx.x;
    for (int e = tid; e < l * KP; e += 256) {
        const int c = e / KP, j = e % KP;
        Ss[e] = (j < k) ? S[c * k + j] : 0.f;
    }
    const int64_t nblocks = (n + 31) / 32;
    const int64_t ntiles = (nblocks + NB - 1) / NB;
    const int cg = tid % CG, rg = tid / CG;
    float best_a[4], best_s[4];
    int64_t best_i[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        best_a[v] = -1.f;
        best_s[v] = 1.f;
        best_i[v] = INT64_MAX;
    }
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        __syncthreads();
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int64_t blk = t * NB + b;
            const float* src = Yp + blk * np * 32;
            if (blk < nblocks) {
                for (int e = tid; e < np * 32; e += 256) Ts[(e >> 5) * TS + b * 32 + (e & 31)] = src[e];
            } else {
                for (int e = tid; e < np * 32; e += 256) Ts[(e >> 5) * TS + b * 32 + (e & 31)] = 0.f;
            }
        }
        __syncthreads();
        float acc[8][4];
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[u][v] = 0.f;
#pragma unroll 2
        for (int c = 0; c < l; ++c) {
            const float4 y0 = *reinterpret_cast<const float4*>(Ts + c * TS + 8 * rg);
            const float4 y1 = *reinterpret_cast<const float4*>(Ts + c * TS + 8 * rg + 4);
            const float4 s4 = *reinterpret_cast<const float4*>(Ss + c * KP + 4 * cg);
            const float y[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
            const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(y[u], sv[v], acc[u][v]);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int64_t r = t * ROWS + 8 * rg + u;
            if (r < n) {
#pragma unroll
                for (int v = 0; v < 4; ++v)
                    if (4 * cg + v < k) {
                        if (out != nullptr) out[r * ldo + 4 * cg + v] = acc[u][v];
                        const float a = fabsf(acc[u][v]);  // rows increase: a strict > keeps the first maximum
                        if (a > best_a[v]) {
                            best_a[v] = a;
                            best_s[v] = acc[u][v];
                            best_i[v] = r;
                        }
                    }
            }
        }
    }
    if (partial != nullptr) {
        __syncthreads();
        AbsMax* red = reinterpret_cast<AbsMax*>(psm);  // [RG][KP]
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            AbsMax b;
            b.a = (double)best_a[v];
            b.sgn = signbit(best_s[v]) ? -1.0 : 1.0;
            b.idx = best_i[v];
            red[rg * KP + 4 * cg + v] = b;
        }
        __syncthreads();
        if (rg == 0) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const int col = 4 * cg + v;
                if (col < k) {
                    AbsMax b = red[col];
                    for (int y = 1; y < RG; ++y) {
                        const AbsMax c2 = red[y * KP + col];
                        if (absmax_better(c2, b)) b = c2;
                    }
                    partial[(int64_t)blockIdx.x * k + col] = b;
                }
            }
        }
    }
}

// out (n x k, row-major) may be null when only absmax3 (k x 3: |max|, first local row, sign) is wanted
inline void launch_panel_xb(petal_ctx* ctx, const float* Yp, int64_t n, int np, int l, const float* S, int k,
                            float* out, int64_t ldo, double* absmax3 = nullptr) {
    if (n == 0 || k == 0) return;
    const int64_t nblocks = ceil_div(n, 32);
    const char* tile_env = getenv("PETAL_PANEL_XB4");
    const bool tiled = k <= 64 && !(tile_env && atoi(tile_env) == 0);
    DBuf<AbsMax> partial;
    int grid = 0;
    if (tiled) {
        const int cgn = k <= 32 ? 8 : 16, rows = (256 / cgn) * 8, kp = 4 * cgn;
        size_t smem = ((size_t)l * kp + (size_t)np * (rows + 4)) * sizeof(float);
        smem = std::max(smem, (size_t)(256 / cgn) * kp * sizeof(AbsMax));
        ensure_dynamic_smem(ctx, panel_xb4_kernel<8>, 160 * 1024);
        ensure_dynamic_smem(ctx, panel_xb4_kernel<16>, 160 * 1024);
        grid = (int)std::min<int64_t>(ceil_div(nblocks, rows / 32), (int64_t)ctx->sm_count * 3);
        if (absmax3 != nullptr) partial.alloc(ctx, (size_t)grid * k);
        KTimer kt(ctx, "panel_xb_f32", (double)n * (np + (out ? k : 0)) * sizeof(float));
        if (cgn == 8) panel_xb4_kernel<8><<<grid, 256, smem, ctx->stream>>>(Yp, n, np, l, S, k, out, ldo, partial.p);
        else panel_xb4_kernel<16><<<grid, 256, smem, ctx->stream>>>(Yp, n, np, l, S, k, out, ldo, partial.p);
        check_launch(ctx);
    } else {
    const int kv = k <= 32 ? 1 : (k <= 64 ? 2 : 4);
    size_t smem = ((size_t)l * 32 * kv + 68 * (size_t)np) * sizeof(float);
    ensure_dynamic_smem(ctx, panel_xb_kernel<1>, 100 * 1024);
    ensure_dynamic_smem(ctx, panel_xb_kernel<2>, 100 * 1024);
    ensure_dynamic_smem(ctx, panel_xb_kernel<4>, 100 * 1024);
    grid = (int)std::min<int64_t>(nblocks, (int64_t)ctx->sm_count * 4);
    smem = std::max(smem, (size_t)8 * 32 * kv * sizeof(AbsMax));
    if (absmax3 != nullptr) partial.alloc(ctx, (size_t)grid * k);
    {
        KTimer kt(ctx, "panel_xb_f32", (double)n * (np + (out ? k : 0)) * sizeof(float));
        if (k <= 32) panel_xb_kernel<1><<<grid, 256, smem, ctx->stream>>>(Yp, n, np, l, S, k, out, ldo, partial.p);
        else if (k <= 64) panel_xb_kernel<2><<<grid, 256, smem, ctx->stream>>>(Yp, n, np, l, S, k, out, ldo, partial.p);
        else panel_xb_kernel<4><<<grid, 256, smem, ctx->stream>>>(Yp, n, np, l, S, k, out, ldo, partial.p);
        check_launch(ctx);
    }
    }
    if (absmax3 != nullptr) {
        colabsmax_final_kernel<<<(unsigned)ceil_div(k, 128), 128, 0, ctx->stream>>>(partial.p, grid, k, absmax3);
        check_launch(ctx);
    }
}

__global__ void symmetrize_kernel(double* C, int64_t d, int64_t ldc) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= d * d) return;
    int64_t i = idx / d, j = idx % d;
    if (i > j) C[i * ldc + j] = C[j * ldc + i];
}

inline void launch_symmetrize(petal_ctx* ctx, double* C, int64_t d) {
    if (d == 0) return;
    symmetrize_kernel<<<(unsigned)ceil_div(d * d, 256), 256, 0, ctx->stream>>>(C, d, d);
    check_launch(ctx);
}

// ------------------------------------------------------------------------------------------
// nonlin : in place U <- g(U), gsum[c] += sum_r g'(U[r][c])      (FastICA contrast functions)
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void ica_g(int fun, T u, T& g, T& gp) {
    if (fun == PETAL_ICA_LOGCOSH) {
        g = tanh(u);
        gp = T(1) - g * g;
    } else if (fun == PETAL_ICA_EXP) {
        T e = exp(-u * u * T(0.5));
        g = u * e;
        gp = (T(1) - u * u) * e;
    } else {
        g = u * u * u;
        gp = T(3) * u * u;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
nonlin_kernel(T* __restrict__ U, int64_t n, int64_t nc, int64_t ld, int fun, double* __restrict__ gsum,
              int tx, int64_t rows_per_cta) {
    const int ty = 256 / tx;
    const int cx = threadIdx.x % tx, ry = threadIdx.x / tx;
    const int64_t col = (int64_t)blockIdx.x * tx + cx;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
    const int64_t r1 = min(n, r0 + rows_per_cta);
    double acc = 0.0;
    if (col < nc) {
        for (int64_t r = r0 + ry; r < r1; r += ty) {
            T g, gp;
            ica_g<T>(fun, U[r * ld + col], g, gp);
            U[r * ld + col] = g;
            acc += (double)gp;
        }
    }
    __shared__ double red[256];
    red[ry * tx + cx] = acc;
    __syncthreads();
    if (ry == 0 && col < nc) {
        double s = 0.0;
        for (int y = 0; y < ty; ++y) s += red[y * tx + cx];
        atomicAdd(&gsum[col], s);
    }
}

template <typename T>
void launch_nonlin(petal_ctx* ctx, T* U, int64_t n, int64_t nc, int64_t ld, int fun, double* gsum) {
    if (n == 0 || nc == 0) return;
    int tx = (int)std::min<int64_t>(256, pow2_ceil(nc));
    int ty = 256 / tx;
    int64_t gx = ceil_div(nc, tx);
    int64_t gy = std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->sm_count * 8 / gx, ceil_div(n, (int64_t)ty * 4)));
    gy = std::min<int64_t>(gy, 65535);
    int64_t rows = ceil_div(n, gy);
    gy = ceil_div(n, rows);
    KTimer kt(ctx, kname<T>("nonlin_f32", "nonlin_f64"), 2.0 * n * nc * sizeof(T));
    nonlin_kernel<T><<<dim3((unsigned)gx, (unsigned)gy), 256, 0, ctx->stream>>>(U, n, nc, ld, fun, gsum, tx, rows);
    check_launch(ctx);
}

// ------------------------------------------------------------------------------------------
// colabsmax : per column of S[n x k]: (max |s|, first row attaining it, sign of that entry)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
colabsmax_kernel(const T* __restrict__ S, int64_t n, int64_t k, int64_t ld, AbsMax* __restrict__ partial,
                 int tx, int64_t rows_per_cta) {
    const int ty = 256 / tx;
    const int cx = threadIdx.x % tx, ry = threadIdx.x / tx;
    const int64_t col = (int64_t)blockIdx.x * tx + cx;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
    const int64_t r1 = min(n, r0 + rows_per_cta);
    AbsMax best;
    best.a = -1.0;
    best.sgn = 1.0;
    best.idx = INT64_MAX;
    if (col < k) {
        for (int64_t r = r0 + ry; r < r1; r += ty) {
            double v = (double)S[r * ld + col];
            AbsMax c;
            c.a = fabs(v);
            c.sgn = signbit(v) ? -1.0 : 1.0;
            c.idx = r;
            if (absmax_better(c, best)) best = c;
        }
    }
    __shared__ AbsMax red[256];
    red[ry * tx + cx] = best;
    __syncthreads();
    if (ry == 0 && col < k) {
        AbsMax b = red[cx];
        for (int y = 1; y < ty; ++y) {
            AbsMax c = red[y * tx + cx];
            if (absmax_better(c, b)) b = c;
        }
        partial[(int64_t)blockIdx.y * k + col] = b;
    }
}

// out3[k*3] = (absmax, first local row index, sign) per column
__global__ void colabsmax_final_kernel(const AbsMax* __restrict__ partial, int64_t chunks, int64_t k,
                                       double* __restrict__ out3) {
    int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= k) return;
    AbsMax b = partial[col];
    for (int64_t c = 1; c < chunks; ++c) {
        AbsMax x = partial[c * k + col];
        if (absmax_better(x, b)) b = x;
    }
    if (b.a < 0.0) {  // no rows at all
        b.a = -1.0;
        b.sgn = 1.0;
        b.idx = 0;
    }
    out3[col * 3 + 0] = b.a;
    out3[col * 3 + 1] = (double)b.idx;
    out3[col * 3 + 2] = b.sgn;
}

template <typename T>
void launch_colabsmax(petal_ctx* ctx, const T* S, int64_t n, int64_t k, int64_t ld, double* out3) {
    if (k == 0) return;
    int tx = (int)std::min<int64_t>(256, pow2_ceil(k));
    int ty = 256 / tx;
    int64_t gx = ceil_div(k, tx);
    int64_t gy = std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->sm_count * 4 / gx,
                                                        ceil_div(std::max<int64_t>(n, 1), (int64_t)ty * 4)));
    gy = std::min<int64_t>(gy, 65535);
    int64_t rows = std::max<int64_t>(1, ceil_div(std::max<int64_t>(n, 1), gy));
    gy = std::max<int64_t>(1, ceil_div(n, rows));
    DBuf<AbsMax> partial(ctx, (size_t)(gy * k));
    KTimer kt(ctx, kname<T>("colabsmax_f32", "colabsmax_f64"), (double)n * k * sizeof(T));
    colabsmax_kernel<T><<<dim3((unsigned)gx, (unsigned)gy), 256, 0, ctx->stream>>>(S, n, k, ld, partial.p, tx, rows);
    check_launch(ctx);
    colabsmax_final_kernel<<<(unsigned)ceil_div(k, 128), 128, 0, ctx->stream>>>(partial.p, gy, k, out3);
    check_launch(ctx);
}

// flip[j] in {+1,-1}; S[:, j] *= flip[j] (n x k) and C[j, :] *= flip[j] (k x d)
template <typename T>
__global__ void apply_flip_kernel(T* S, int64_t n, int64_t k, int64_t lds, T* C, int64_t d,
                                  const double* __restrict__ flip3 /* [k][3], sign at +2 */) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t ns = S ? n * k : 0;
    if (idx < ns) {
        int64_t r = idx / k, j = idx % k;
        if (flip3[j * 3 + 2] < 0.0) S[r * lds + j] = -S[r * lds + j];
    } else if (idx < ns + k * d) {
        int64_t e = idx - ns;
        int64_t j = e / d;
        if (flip3[j * 3 + 2] < 0.0) C[e] = -C[e];
    }
}

template <typename T>
void launch_apply_flip(petal_ctx* ctx, T* S, int64_t n, int64_t k, int64_t lds, T* C, int64_t d,
                       const double* flip3) {
    int64_t total = (S ? n * k : 0) + k * d;
    if (total == 0) return;
    apply_flip_kernel<T><<<(unsigned)ceil_div(total, 256), 256, 0, ctx->stream>>>(S, n, k, lds, C, d, flip3);
    check_launch(ctx);
}

// ------------------------------------------------------------------------------------------
// small elementwise helpers
// ------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ in, TO* __restrict__ out, int64_t count, double scale) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = (TO)((double)in[i] * scale);
}

template <typename TI, typename TO>
void launch_cast(petal_ctx* ctx, const TI* in, TO* out, int64_t count, double scale = 1.0) {
    if (count == 0) return;
    cast_kernel<TI, TO><<<(unsigned)ceil_div(count, 256), 256, 0, ctx->stream>>>(in, out, count, scale);
    check_launch(ctx);
}

// out[c][r] = in[r][c]
__global__ void transpose_kernel(const double* __restrict__ in, int64_t rows, int64_t cols,
                                 double* __restrict__ out) {
    __shared__ double tile[32][33];
    int64_t c = (int64_t)blockIdx.x * 32 + threadIdx.x;
    int64_t r0 = (int64_t)blockIdx.y * 32;
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        int64_t r = r0 + y;
        tile[y][threadIdx.x] = (r < rows && c < cols) ? in[r * cols + c] : 0.0;
    }
    __syncthreads();
    int64_t orow0 = (int64_t)blockIdx.x * 32;
    int64_t ocol = r0 + threadIdx.x;
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        int64_t orow = orow0 + y;
        if (orow < cols && ocol < rows) out[orow * rows + ocol] = tile[threadIdx.x][y];
    }
}

inline void launch_transpose(petal_ctx* ctx, const double* in, int64_t rows, int64_t cols, double* out) {
    if (rows == 0 || cols == 0) return;
    dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32));
    transpose_kernel<<<grid, dim3(32, 8), 0, ctx->stream>>>(in, rows, cols, out);
    check_launch(ctx);
}

}  // namespace petal
