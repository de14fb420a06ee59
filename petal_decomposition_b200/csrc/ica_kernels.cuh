// Fused FastICA fixed-point pass on tcgen05 / TMEM (f32, d <= 64 features, nc <= 64 components):
//
//     U = (X - mu) W~^T          (reference: w.dot(input),            src/ica.rs:332)
//     G = g(U),  gp = sum g'(U)  (reference: logcosh,                 src/ica.rs:383-398)
//     H^T = (X - mu)^T G         (reference: gwtx.dot(&input.t()),    src/ica.rs:333)
//
// in ONE pass over X: a 128-row tile is pulled in by TMA once, used as the A operand of the first MMA
// (lanes = rows) and, transposed through the transform warps, as the A operand of the second MMA
// (lanes = features); G never leaves the SM (TMEM -> registers -> shared-memory operand tiles).
// Both contractions are 3xTF32 (hi/lo split, fp32 accumulate in TMEM); the H accumulator chain is cut
// every tile (128 rows) into fp32 registers, and only the CTA's totals go to global memory (f64 atomics).
//
// Warp roles (22 warps):
//   0-15  row warps     : lane quarter q = w & 3, 16-column group cq = w >> 2
//                         transform-A (x - mu, hi/lo -> TMEM A1), epilogue-1 (U -> g -> smem G tiles, g' sums);
//                         the warps of lane quarters 0, 1 also drain the H accumulator (lanes = features)
//   16,17,20,21 feature warps: q = w & 3 (features 0-63), quarter tiles (w - 16) >> 2 and + 2
//                         transform-B (transposed x - mu, hi/lo -> the TMEM A2 ring)
//   18    TMA producer  (X tile ring; W~ operand tiles once)
//   19    MMA issuer + TMEM allocation
#pragma once
#include "tc_kernels.cuh"

namespace petal {
namespace ica {

using namespace tc;

constexpr int kRows = 128;                 // rows per tile
constexpr int kD = 64;                     // padded features
constexpr int kNC = 64;                    // padded components
constexpr int kThreadsIca = 704;           // 22 warps: 88 registers per thread
constexpr int kRowWarps = 16;              // warps 0-15
constexpr int kFeatWarps = 4;              // warps 16, 17, 20, 21 (TMEM lane quarters 0, 1 = features 0-63)
constexpr int kTmaWarp = 18, kMmaWarp = 19;
constexpr int kXStage = kRows * kD * 4;    // 32 KB: two [128 rows][32 floats] SWIZZLE_128B sub-tiles
constexpr int kXStages = 2;
constexpr int kWTile = kNC * 128;          // 8 KB: [64 comps][32 k] K-major SWIZZLE_128B
constexpr int kGTile = kNC * 128;          // 8 KB: [64 comps][32 rows]
constexpr int kGBuf = 8 * kGTile;          // one G buffer: hi tiles 0-3, lo tiles 4-7 (64 KB); two buffers
// TMEM columns (all 512)
constexpr int kA1 = 0;                     // [hi 64 | lo 64]                          lanes = rows
constexpr int kA2 = 128;                   // ring of 3 quarter-tile slots [hi 32 | lo 32]  lanes = features
constexpr int kA2Slots = 3;
constexpr int kAccU = 320;                 // two U accumulators of 64 columns
constexpr int kAccH = 448;                 // 64 columns

struct IcaParams {
    CUtensorMap map_x;    // X as {features inner, rows}, box {32, 128}, SWIZZLE_128B
    CUtensorMap map_whi;  // W~ hi as {k inner (64), comps (64)}, box {32, 64}, SWIZZLE_128B
    CUtensorMap map_wlo;
    const float* mu_pad;  // [64]
    int64_t n;
    int d, nc, fun;
    double* Ht;           // [d x nc] f64, accumulated
    double* gp;           // [nc] f64, accumulated
    const double* state;  // optional: state[6] != 0 -> the fixed point has converged (or failed), do nothing
    long long* trace;     // optional clock64 timeline of CTA 0 (PETAL_ICA_TRACE)
};

// smem carve-up (offsets from a 1024 B aligned base): 64 + 32 + 128 KB
constexpr uint32_t kOffX = 0;
constexpr uint32_t kOffW = kOffX + kXStages * kXStage;       // W hi: 2 tiles, W lo: 2 tiles
constexpr uint32_t kOffG = kOffW + 4 * kWTile;               // two G buffers
constexpr uint32_t kOffMu = kOffG + 2 * kGBuf;
constexpr uint32_t kOffBars = kOffMu + 256;
constexpr uint32_t kOffSlot = kOffBars + 32 * 8;
constexpr uint32_t kSmemIca = kOffSlot + 16 + 1024;

__device__ __forceinline__ uint32_t ib(uint32_t bars, int i) { return bars + 8u * (uint32_t)i; }
// barrier ids
constexpr int B_XFULL = 0, B_XEMPTY = 2, B_WFULL = 4, B_A1 = 5, B_UFULL = 6, B_UEMPTY = 8, B_GREADY = 10, B_GFREE = 12,
              B_A2READY = 14, B_A2FREE = 17, B_HFULL = 20, B_HEMPTY = 21;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// g(u) and the quantity summed per component.  For logcosh the kernel sums g^2 (sum g' = rows - sum g^2, one FFMA
// instead of two); rows past n have u == 0 exactly, which gives g == 0 for every contrast function.
template <int FUN>
__device__ __forceinline__ float ica_g_acc(float u, float& acc) {
    if (FUN == PETAL_ICA_LOGCOSH) {
        // tanh(u) = sign(u) (1 - 2 / (exp(2|u|) + 1)): two MUFU ops, absolute error ~1e-7, saturates cleanly
        const float e = ex2_approx(fabsf(u) * 2.8853900817779268f);
        const float t = fmaf(-2.0f, rcp_approx(e + 1.0f), 1.0f);
        const float g = copysignf(t, u);
        acc = fmaf(g, g, acc);
        return g;
    } else if (FUN == PETAL_ICA_EXP) {
        const float u2 = u * u;
        const float e = ex2_approx(u2 * -0.72134752044448170f);
        acc = fmaf(1.0f - u2, e, acc);
        return u * e;
    } else {
        const float u2 = u * u;
        acc = fmaf(3.0f, u2, acc);
        return u * u2;
    }
}

// test hook: the epilogue's device function applied elementwise to U[n x nc] (see petal_ica_nonlin_f32)
template <int FUN>
__global__ void ica_g_probe_kernel(float* __restrict__ U, int64_t n, int64_t nc, double* __restrict__ gsum) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * nc) return;
    float acc = 0.f;
    const float g = ica_g_acc<FUN>(U[i], acc);
    U[i] = g;
    // the kernel carries sum g^2 for logcosh and turns it into sum (1 - g^2) at the end
    atomicAdd(&gsum[i % nc], (double)(FUN == PETAL_ICA_LOGCOSH ? 1.0f - acc : acc));
}

constexpr int kTraceTiles = 64;
constexpr int kTraceEv = 16;
__device__ __forceinline__ void ica_trace(const IcaParams& p, int ev, uint32_t it) {
#ifdef PETAL_TC_TRACE_BUILD
    if (p.trace != nullptr && blockIdx.x == 0 && it < (uint32_t)kTraceTiles) p.trace[ev * kTraceTiles + it] = clock64();
#else
    (void)p; (void)ev; (void)it;
#endif
}

// Software pipeline (tile index t per CTA; tensor pipe order M1(0), M1(1), M2(0), M1(2), M2(1), ...):
//   M1(t): U[t & 1] = A1 * W~^T                    A1 (TMEM) written by the row warps from X(t)
//   epi(t): U -> g -> G[t & 1] (smem operand tiles)  row warps, while M2(t-1) runs
//   M2(t): H^T = A2 * G[t & 1], four quarter tiles   A2 ring slots written by the feature warps from X(t)
//   H is drained into registers every tile (row warps of lane quarters 0, 1) while M1(t+2) runs.
template <int FUN>
__global__ void __launch_bounds__(kThreadsIca, 1) ica_fused_kernel(const __grid_constant__ IcaParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + kOffBars;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ntiles = (p.n + kRows - 1) / kRows;
    if (p.state != nullptr && p.state[6] != 0.0) return;  // on-device convergence flag (uniform over the grid)
    // tiles of this CTA: blockIdx.x, + gridDim.x, ...
    const uint32_t T = (uint32_t)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kXStages; ++s) {
            mbar_init(ib(bars, B_XFULL + s), 1);
            mbar_init(ib(bars, B_XEMPTY + s), kRowWarps + kFeatWarps);
        }
        mbar_init(ib(bars, B_WFULL), 1);
        mbar_init(ib(bars, B_A1), kRowWarps);
        for (int b = 0; b < 2; ++b) {
            mbar_init(ib(bars, B_UFULL + b), 1);
            mbar_init(ib(bars, B_UEMPTY + b), kRowWarps);
            mbar_init(ib(bars, B_GREADY + b), kRowWarps);
            mbar_init(ib(bars, B_GFREE + b), 1);
        }
        for (int s = 0; s < kA2Slots; ++s) {
            mbar_init(ib(bars, B_A2READY + s), 2);   // the two feature warps (features 0-31, 32-63) of a quarter tile
            mbar_init(ib(bars, B_A2FREE + s), 1);
        }
        mbar_init(ib(bars, B_HFULL), 1);
        mbar_init(ib(bars, B_HEMPTY), kRowWarps / 2);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < kD; i += kThreadsIca) reinterpret_cast<float*>(bp + kOffMu)[i] = p.mu_pad[i];
    if (warp == kMmaWarp) tmem_alloc(base + kOffSlot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(bp + kOffSlot);
    const float* mus = reinterpret_cast<const float*>(bp + kOffMu);

    if (warp == kTmaWarp) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            mbar_expect_tx(ib(bars, B_WFULL), 4 * kWTile);
            for (int kb = 0; kb < 2; ++kb) {
                tma_load_2d(base + kOffW + (uint32_t)kb * kWTile, &p.map_whi, kb * 32, 0, ib(bars, B_WFULL));
                tma_load_2d(base + kOffW + (uint32_t)(2 + kb) * kWTile, &p.map_wlo, kb * 32, 0, ib(bars, B_WFULL));
            }
            for (uint32_t it = 0; it < T; ++it) {
                const int64_t t = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
                const int s = (int)(it & 1u);
                mbar_wait(ib(bars, B_XEMPTY + s), ((it >> 1) & 1u) ^ 1u);
                mbar_expect_tx(ib(bars, B_XFULL + s), kXStage);
                tma_load_2d(base + kOffX + (uint32_t)s * kXStage, &p.map_x, 0, (int)(t * kRows), ib(bars, B_XFULL + s));
                tma_load_2d(base + kOffX + (uint32_t)s * kXStage + 16384u, &p.map_x, 32, (int)(t * kRows), ib(bars, B_XFULL + s));
            }
        }
    } else if (warp == kMmaWarp) {
        // ================================ MMA issuer ================================
        const uint32_t idesc = make_idesc_tf32(kNC, 0);
        mbar_wait(ib(bars, B_WFULL), 0);
        auto issue_m1 = [&](uint32_t tau) {  // U[tau & 1] = A1 * W~^T
            const int ub = (int)(tau & 1u);
            mbar_wait(ib(bars, B_A1), tau & 1u);
            if (tau >= 2) mbar_wait(ib(bars, B_UEMPTY + ub), ((tau >> 1) - 1u) & 1u);
            tc_fence_after();
            if (lane == 0) ica_trace(p, 4, tau);
            if (elect_one()) {
                const uint32_t acc = tmem_base + (uint32_t)(kAccU + 64 * ub);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint32_t woff = (uint32_t)(ks >> 2) * kWTile + (uint32_t)(ks & 3) * 32u;
                    const uint64_t dhi = make_desc_sw128(base + kOffW + woff, 16u, 1024u);
                    const uint64_t dlo = make_desc_sw128(base + kOffW + 2u * kWTile + woff, 16u, 1024u);
                    const uint32_t a_hi = tmem_base + (uint32_t)(kA1 + ks * 8);
                    const uint32_t a_lo = a_hi + 64u;
                    mma_tf32_ts(acc, a_lo, dhi, idesc, ks > 0 ? 1u : 0u);
                    mma_tf32_ts(acc, a_hi, dlo, idesc, 1u);
                    mma_tf32_ts(acc, a_hi, dhi, idesc, 1u);
                }
                tc_commit(ib(bars, B_UFULL + ub));
            }
            __syncwarp();
        };
        issue_m1(0);
        for (uint32_t it = 0; it < T; ++it) {
            if (it + 1 < T) issue_m1(it + 1);
            // ---- M2(it): H^T = A2 * G[it & 1], quarter tile by quarter tile
            const int gb = (int)(it & 1u);
            mbar_wait(ib(bars, B_GREADY + gb), (it >> 1) & 1u);
            if (it >= 1) mbar_wait(ib(bars, B_HEMPTY), (it - 1u) & 1u);
            if (lane == 0) ica_trace(p, 5, it);
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
                const uint32_t gran = 4u * it + (uint32_t)q;
                const int slot = (int)(gran % 3u);
                mbar_wait(ib(bars, B_A2READY + slot), (gran / 3u) & 1u);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t goff = kOffG + (uint32_t)gb * kGBuf + (uint32_t)q * kGTile;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t dhi = make_desc_sw128(base + goff + (uint32_t)ks * 32u, 16u, 1024u);
                        const uint64_t dlo = make_desc_sw128(base + goff + 4u * kGTile + (uint32_t)ks * 32u, 16u, 1024u);
                        const uint32_t a_hi = tmem_base + (uint32_t)(kA2 + 64 * slot + ks * 8);
                        const uint32_t a_lo = a_hi + 32u;
                        mma_tf32_ts(tmem_base + kAccH, a_lo, dhi, idesc, (q == 0 && ks == 0) ? 0u : 1u);
                        mma_tf32_ts(tmem_base + kAccH, a_hi, dlo, idesc, 1u);
                        mma_tf32_ts(tmem_base + kAccH, a_hi, dhi, idesc, 1u);
                    }
                    tc_commit(ib(bars, B_A2FREE + slot));
                    if (q == 3) {
                        tc_commit(ib(bars, B_GFREE + gb));
                        tc_commit(ib(bars, B_HFULL));
                    }
                }
                __syncwarp();
            }
            if (lane == 0) ica_trace(p, 6, it);
        }
    } else if (warp < kRowWarps) {
        // ================================ row warps ================================
        const int q = warp & 3, cq = warp >> 2;           // lane quarter, 16-column group
        const int r = q * 32 + lane;                      // row inside the tile = TMEM lane
        const uint32_t lane_field = (uint32_t)(q * 32) << 16;
        const bool tr = (warp == 0 && lane == 0);
        const bool h_owner = (q < 2);                     // lanes 0-63 of H = features: these warps drain H
        float acc[16], hacc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            acc[j] = 0.f;
            hacc[j] = 0.f;
        }
        int nvalid = 0;
        // x - mu of this thread's row (features 16 cq ..) -> A1 hi / lo, for the tile in X stage `it`
        auto transform_a = [&](uint32_t it) {
            const int64_t t = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
            const int s = (int)(it & 1u);
            const bool valid = t * kRows + r < p.n;
            nvalid += valid ? 1 : 0;
            if (tr) ica_trace(p, 7, it);
            mbar_wait(ib(bars, B_XFULL + s), (it >> 1) & 1u);
            if (tr) ica_trace(p, 8, it);
            uint32_t v[16];
            const uint8_t* xr = bp + kOffX + (uint32_t)s * kXStage + (uint32_t)(cq >> 1) * 16384u + r * 128;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const int c = (cq & 1) * 4 + c4;
                const float4 x4 = *reinterpret_cast<const float4*>(xr + ((c ^ (r & 7)) << 4));
                const float4 m4 = *reinterpret_cast<const float4*>(mus + cq * 16 + c4 * 4);
                v[c4 * 4 + 0] = __float_as_uint(valid ? x4.x - m4.x : 0.f);
                v[c4 * 4 + 1] = __float_as_uint(valid ? x4.y - m4.y : 0.f);
                v[c4 * 4 + 2] = __float_as_uint(valid ? x4.z - m4.z : 0.f);
                v[c4 * 4 + 3] = __float_as_uint(valid ? x4.w - m4.w : 0.f);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(ib(bars, B_XEMPTY + s));
            // A1 is free: M1 of the previous tile has completed (this warp has seen its u_full)
            const uint32_t a = tmem_base + lane_field + (uint32_t)(kA1 + cq * 16);
            tmem_st16(a, v);
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = __float_as_uint(__uint_as_float(v[k]) - __uint_as_float(v[k] & 0xFFFFE000u));
            tmem_st16(a + 64u, v);
            if (tr) ica_trace(p, 9, it);
            tmem_st_wait();
            if (tr) ica_trace(p, 10, it);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(ib(bars, B_A1));
        };
        auto drain_h = [&](uint32_t tau) {  // H of tile tau -> registers
            mbar_wait(ib(bars, B_HFULL), tau & 1u);
            tc_fence_after();
            uint32_t w[16];
            tmem_ld16(tmem_base + lane_field + (uint32_t)(kAccH + cq * 16), w);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) hacc[j] += __uint_as_float(w[j]);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(ib(bars, B_HEMPTY));
        };
        transform_a(0);
        for (uint32_t it = 0; it < T; ++it) {
            const int ub = (int)(it & 1u);
            if (h_owner && it >= 2) drain_h(it - 2);  // M2(it-2) precedes M1(it) on the tensor pipe
            // ---- epilogue-1: U[r][16*cq ..] -> g
            mbar_wait(ib(bars, B_UFULL + ub), (it >> 1) & 1u);
            tc_fence_after();
            if (tr) ica_trace(p, 0, it);
            uint32_t u[16];
            tmem_ld16(tmem_base + lane_field + (uint32_t)(kAccU + 64 * ub + cq * 16), u);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(ib(bars, B_UEMPTY + ub));
#pragma unroll
            for (int j = 0; j < 16; ++j) u[j] = __float_as_uint(ica_g_acc<FUN>(__uint_as_float(u[j]), acc[j]));
            if (tr) ica_trace(p, 1, it);
            // ---- A operand of the next tile's M1 (issued before M2(it), so it overlaps the G store below)
            if (it + 1 < T) transform_a(it + 1);
            if (tr) ica_trace(p, 2, it);
            // ---- G[it & 1]: free once M2(it-2) has completed
            if (it >= 2) mbar_wait(ib(bars, B_GFREE + ub), ((it >> 1) - 1u) & 1u);
            {
                // K-major operand tiles [comp][32 rows] per 32-row block (= q), SWIZZLE_128B
                uint8_t* gh = bp + kOffG + (uint32_t)ub * kGBuf + (uint32_t)q * kGTile + (lane & 3) * 4;
                const int rc = lane >> 2;  // 16 B chunk of this row inside the 128 B line
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int c = cq * 16 + j;
                    const int off = c * 128 + ((rc ^ (j & 7)) << 4);   // (c & 7) == (j & 7)
                    const float g = __uint_as_float(u[j]);
                    *reinterpret_cast<float*>(gh + off) = g;
                    *reinterpret_cast<float*>(gh + 4u * kGTile + off) = g - __uint_as_float(u[j] & 0xFFFFE000u);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(ib(bars, B_GREADY + ub));
            }
            if (tr) ica_trace(p, 3, it);
        }
        if (h_owner) {
            if (T >= 2) drain_h(T - 2);
            drain_h(T - 1);
            const int f = q * 32 + lane;
            if (f < p.d) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int c = cq * 16 + j;
                    if (c < p.nc) atomicAdd(&p.Ht[(size_t)f * p.nc + c], (double)hacc[j]);
                }
            }
        }
        // per-column totals of this warp's rows -> sum of g' -> one atomic per column per warp
        const float fvalid = (float)nvalid, finvalid = (float)((int)T - nvalid);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float sgp = acc[j];
            if (FUN == PETAL_ICA_LOGCOSH) sgp = fvalid - sgp;        // sum (1 - g^2) over valid rows
            else if (FUN == PETAL_ICA_EXP) sgp = sgp - finvalid;     // u == 0 rows contributed exactly 1 each
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sgp += __shfl_xor_sync(0xffffffffu, sgp, o);
            const int c = cq * 16 + j;
            if (lane == 0 && c < p.nc) atomicAdd(&p.gp[c], (double)sgp);
        }
    } else if ((warp & 3) < 2 && warp < 22) {
        // ================================ feature warps ================================
        const int q = warp & 3;                      // 0, 1: features 0-31, 32-63
        const int rh = (warp - kRowWarps) >> 2;      // quarter tiles rh and rh + 2
        const int f = q * 32 + lane;
        const uint32_t lane_field = (uint32_t)(q * 32) << 16;
        const float mu_f = mus[f];
        for (uint32_t it = 0; it < T; ++it) {
            const int s = (int)(it & 1u);
            mbar_wait(ib(bars, B_XFULL + s), (it >> 1) & 1u);
            const uint8_t* xs = bp + kOffX + (uint32_t)s * kXStage + (uint32_t)(f >> 5) * 16384u + (f & 3) * 4;
            const int fc = (f & 31) >> 2;
            // both quarter tiles into registers, then the X stage can be recycled
            uint32_t va[32], vb[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const int ra = rh * 32 + k, rb = (rh + 2) * 32 + k;
                va[k] = __float_as_uint(*reinterpret_cast<const float*>(xs + ra * 128 + ((fc ^ (ra & 7)) << 4)) - mu_f);
                vb[k] = __float_as_uint(*reinterpret_cast<const float*>(xs + rb * 128 + ((fc ^ (rb & 7)) << 4)) - mu_f);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(ib(bars, B_XEMPTY + s));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t* v = h ? vb : va;
                const uint32_t gran = 4u * it + (uint32_t)(rh + 2 * h);
                const int slot = (int)(gran % 3u);
                mbar_wait(ib(bars, B_A2FREE + slot), ((gran / 3u) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t a = tmem_base + lane_field + (uint32_t)(kA2 + 64 * slot);
                tmem_st16(a, v);
                tmem_st16(a + 16u, v + 16);
#pragma unroll
                for (int k = 0; k < 32; ++k) v[k] = __float_as_uint(__uint_as_float(v[k]) - __uint_as_float(v[k] & 0xFFFFE000u));
                tmem_st16(a + 32u, v);
                tmem_st16(a + 48u, v + 16);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(ib(bars, B_A2READY + slot));
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512);
}

inline bool fused_supported(const void* X, int64_t ld, int64_t n, int64_t d, int64_t nc) {
    return n >= 1024 && d >= 4 && d <= kD && nc >= 1 && nc <= kNC && (ld % 4 == 0) && is_aligned16(X) && n < ((int64_t)1 << 31);
}

// Plan for the repeated fused pass of one FastICA fit: operand buffers and tensor maps are built once,
// run() is a single kernel launch.  Ht[d x nc] += (X - mu)^T g((X - mu) W~^T), gp[nc] += sum g'(.); both must be
// zero on entry (the update kernel re-zeroes them after consuming them).
struct IcaFused {
    DBuf<float> whi, wlo, mup;   // W~ hi / lo operand arrays [64][64] (zero padded), mu [64]
    IcaParams p;
    int grid = 0;
    int fun = 0;
    int64_t ntiles = 0;

    void init(petal_ctx* ctx, const float* X, int64_t ld, int64_t n, int64_t d, const float* mu, int64_t nc, int fun_,
              double* Ht, double* gp, const double* state) {
        whi.alloc(ctx, (size_t)(kNC * kD));
        wlo.alloc(ctx, (size_t)(kNC * kD));
        mup.alloc(ctx, (size_t)kD);
        whi.zero();
        wlo.zero();
        prep_mu_kernel<<<1, 256, 0, ctx->stream>>>(mu, d, kD, mup.p);
        check_launch(ctx);
        std::memset(&p, 0, sizeof p);
        p.map_x = make_map_2d(X, (uint64_t)d, (uint64_t)n, (uint64_t)ld, 32, kRows, true);
        p.map_whi = make_map_2d(whi.p, kD, kNC, kD, 32, kNC, true);
        p.map_wlo = make_map_2d(wlo.p, kD, kNC, kD, 32, kNC, true);
        p.mu_pad = mup.p;
        p.n = n;
        p.d = (int)d;
        p.nc = (int)nc;
        p.fun = fun_;
        p.Ht = Ht;
        p.gp = gp;
        p.state = state;
        fun = fun_;
        ntiles = ceil_div(n, kRows);
        grid = (int)std::min<int64_t>(ntiles, ctx->sm_count);
    }
    // W~ (nc x d, f32 row-major) -> operand arrays (used when the update did not write them itself)
    void set_w(petal_ctx* ctx, const float* Wt) {
        prep_b_kernel<float><<<(unsigned)ceil_div((int64_t)kNC * kD, 256), 256, 0, ctx->stream>>>(Wt, p.d, 1, p.d, kD, p.nc, kNC,
                                                                                                  whi.p, wlo.p);
        check_launch(ctx);
    }
    void run(petal_ctx* ctx) {
        DBuf<long long> trbuf;
        IcaParams q = p;
        if (getenv("PETAL_ICA_TRACE")) {
            trbuf.alloc(ctx, (size_t)kTraceEv * kTraceTiles);
            trbuf.zero();
            q.trace = trbuf.p;
        }
        auto launch = [&](auto kernel) {
            ensure_dynamic_smem(ctx, kernel, kSmemIca);
            KTimer kt(ctx, "ica_fused_f32", (double)p.n * p.d * sizeof(float));
            kernel<<<grid, kThreadsIca, kSmemIca, ctx->stream>>>(q);
            check_launch(ctx);
        };
        if (fun == PETAL_ICA_LOGCOSH) launch(ica_fused_kernel<PETAL_ICA_LOGCOSH>);
        else if (fun == PETAL_ICA_EXP) launch(ica_fused_kernel<PETAL_ICA_EXP>);
        else launch(ica_fused_kernel<PETAL_ICA_CUBE>);
        if (q.trace) {
            std::vector<long long> h((size_t)kTraceEv * kTraceTiles);
            PETAL_CUDA(cudaMemcpyAsync(h.data(), q.trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
            PETAL_CUDA(cudaStreamSynchronize(ctx->stream));
            auto ev = [&](int e, int i) { return h[(size_t)e * kTraceTiles + i]; };
            const int a = 8, b = (int)std::min<int64_t>(40, ceil_div(ntiles, grid) - 1);
            if (b > a) {
                auto mean = [&](int e1, int e0, int sh) {
                    double sacc = 0;
                    for (int i = a; i < b; ++i) sacc += (double)(ev(e1, i + sh) - ev(e0, i));
                    return sacc / (b - a);
                };
                fprintf(stderr,
                        "[ica trace] period %.0f | row: ufull->tanh done %.0f  ->trA(next) done %.0f  ->G stored %.0f  ->next ufull %.0f | "
                        "mma: M1 start->M2 start %.0f  M2 start->issued %.0f  M2 issued->next M1 start %.0f\n",
                        mean(0, 0, 1), mean(1, 0, 0), mean(2, 1, 0), mean(3, 2, 0), mean(0, 3, 1), mean(5, 4, 0), mean(6, 5, 0),
                        mean(4, 6, 1));
                fprintf(stderr, "[ica trace]   trA: wait X %.0f  read+st issue %.0f  st wait %.0f\n", mean(8, 7, 0), mean(9, 8, 0),
                        mean(10, 9, 0));
            }
        }
    }
};

}  // namespace ica
}  // namespace petal
