// tcgen05 / TMEM / TMA engine for the f32 streaming contractions (sm_100a only).
//
//   tc_xb  : Y[n x L]  = (X - mu) * B            X*Omega, X*P, transform    (src/pca.rs:707,714,745)
//   tc_atb : Z[da x L] += (X - mu)^T * Y          X^T*Q, Q^T*X               (src/pca.rs:681,711)
//
// fp32 accuracy is kept with the 3xTF32 split: every operand v is used as hi = tf32(v) (the tensor
// core truncates the low 13 mantissa bits itself) and lo = v - hi (exact in fp32), and each product is
// accumulated as lo*hi + hi*lo + hi*hi in the fp32 TMEM accumulator.
//
// Data flow per CTA (persistent, 1 CTA / SM, 640 threads = 20 warps):
//   warp 16    TMA producer : X tiles -> shared memory ring (mbarrier complete_tx); stages are released as soon as the
//                             transform warps have read them.  tc_atb with in-kernel B_lo: also feeds the B ring.
//   warp 19    B side       : tc_xb: TMA producer of the B^T hi / lo tiles.  tc_atb: splitter (Y panel block = B_hi
//                             operand as loaded; B_lo = y - tf32(y) written next to it).  TMEM alloc / free.
//   warps 0-15 transform    : shared X tile -> registers: subtract mu, split hi/lo -> tcgen05.st into a TMEM operand
//                             ring of half-K-block slots (the A operand of the MMA lives in TMEM: lane = output row);
//                             for tc_atb this is also where the X tile is transposed (lane = feature)
//   warps 17,18 MMA issuers : one per M tile; tcgen05.mma.kind::tf32 (A from TMEM, B from shared memory descriptors),
//                             3 MMAs per K-step, accumulators stay in TMEM
//   warps 0-15 epilogue     : tcgen05.ld accumulators -> Y (tc_xb: panel-major or row-major) / fp32 registers ->
//                             f64 atomics once per CTA (tc_atb)
// The centred copy of X never exists; mu is subtracted in the transform stage.  Kernel modes (fast / precise = cut
// accumulation chains): see ModeTraits.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace petal {
namespace tc {

constexpr int kTransformWarps = 16;          // two warps per 32 TMEM lanes: each owns one half (16 values) of a K block
constexpr int kEpilogueWarps = 8;            // tc_atb: only the first half keeps the register accumulators
constexpr int kThreads = (kTransformWarps + 4) * 32;  // + X TMA, 2 x MMA, B TMA (+ TMEM alloc); a 21st warp would round
                                                       // the register file split up to 24 warps (80 regs per thread)
constexpr int kMT = 2;            // M tiles (128 TMEM lanes each) per CTA
constexpr int kKB = 32;           // K block: 32 fp32 = 128 B
constexpr int kXStageBytes = kMT * 128 * kKB * 4;  // 32 KB
constexpr int kTmemCols = 512;
constexpr int kYBufs = 3;         // tc_atb, row-major Y: ring of transposed Y tiles (decoupled from the TMEM operand slots)
constexpr int kASlotsMax = 5;              // TMEM operand ring: one slot per half K block (16 k-values); 3, 4 or 5 slots
constexpr int kASlotCols = kMT * 32;       // per slot: MT x (16 hi + 16 lo) columns
// accumulators start after the operand ring: 4 slots -> column 256, one accumulator set (MT x n_pad columns);
// 3 slots -> column 192, two sets (tc_xb with n_pad <= 80): the next super-tile's MMAs run while the previous
// accumulators are drained

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a pipeline bug must trap, not hang the GPU.
// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes (or the hint elapses)
// instead of letting it spin.  A software spin loop (try_wait without hint + iteration bookkeeping) was measured
// (ncu source page, r02) at 25-40 % of all warp instructions of the tc kernels - issue slots and power taken from the
// transform warps that share the scheduler.  kWaitHintNs * kWaitTries bounds the wait at ~4 s.
#ifndef PETAL_WAIT_HINT_NS
#define PETAL_WAIT_HINT_NS 1000000
#endif
constexpr uint32_t kWaitHintNs = PETAL_WAIT_HINT_NS;   // 0: no hint (the hardware's default, short, suspend time)
constexpr uint32_t kWaitTries = kWaitHintNs >= 1000u ? 4000000000u / kWaitHintNs : 40000000u;
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    if (kWaitHintNs != 0u)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(kWaitHintNs)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    return done != 0;
}
__device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity) {
    printf("petal tc kernel: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
    __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
#pragma unroll 1
    for (uint32_t it = 0; it < kWaitTries; ++it)
        if (mbar_try_wait(bar, parity)) return;
    mbar_timeout(bar, parity);
}
// Wait with back-off for roles that run ahead of their consumer (TMA producers waiting for a free stage, transform
// warps waiting for a TMEM operand slot): between probes the thread sleeps ~ns instead of re-issuing try_wait - the
// hardware's suspended try_wait wakes on every barrier event of the CTA, which made these loops 40 % of tc_atb's
// warp instructions (r02 ncu source page).  Wake-up is late by at most ~2 ns_ : only for waits with slack.
#ifndef PETAL_WAIT_BACKOFF
#define PETAL_WAIT_BACKOFF 0
#endif
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, uint32_t ns) {
#if PETAL_WAIT_BACKOFF
    if (mbar_try_wait(bar, parity)) return;
#pragma unroll 1
    for (uint32_t it = 0; it < 40000000u; ++it) {
        __nanosleep(ns);
        if (mbar_try_wait(bar, parity)) return;
    }
    mbar_timeout(bar, parity);
#else
    (void)ns;
    mbar_wait(bar, parity);
#endif
}
// non-blocking probe of a barrier phase
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
// true in exactly one (elected) lane of a fully converged warp; keeps the surrounding code warp-uniform so
// that descriptors / addresses live in uniform registers (the tensor-core issue path reads those)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xFFFFFFFF;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc], kind::tf32
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc], kind::f16 (bf16 operands, K = 16 per instruction, fp32 accumulate)
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// two fp32 -> one 32-bit word of two bf16 (round to nearest): `lo` in bits 0-15 (the lower K index), `hi` in 16-31
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

// 2-MMA split ("cross" mode).  x y = xh yh + (xl yh + xh yl) + xl yl with xh = tf32(x), xl = x - xh.  The leading term
// needs tf32 operands; the two cross terms are 2^-11 of it, so their factors only need ~8 significant bits: they are
// fed as bf16 and, concatenated along K, cost ONE kind::f16 MMA with K = 16 ([xl | xh] . [yh ; yl] over 8 k values)
// instead of two kind::tf32 MMAs with K = 8.  Tensor work per K step: 2 MMA slots instead of 3; the error of the
// bf16 rounding enters at 2^-11 * 2^-9 = 2^-20 relative to the product, unbiased (round to nearest).
#ifndef PETAL_TC_CROSS
#define PETAL_TC_CROSS 1
#endif
constexpr bool kCrossEnabled = PETAL_TC_CROSS != 0;
#ifndef PETAL_TC_CROSS_ATB
#define PETAL_TC_CROSS_ATB 0
#endif
// tc_atb with panel-major Y (build option, OFF): the B-side cross tile ([bf16 y | bf16 (y - tf32 y)] per K step) is
// derived from the fp32 Y panel block by the sixteen transform warps, one 16 B chunk per thread, before they turn to
// their own operand.  Correct (parity tests pass) but slower: tc_atb is bound by the instruction throughput of its
// transform warps (transposing scalar shared-memory reads), not by the tensor pipe, and the cross operand adds to
// exactly that - measured (r02, same box): X^T Y pass 10.9 -> 12.5 ms with this variant, 13.1 ms with a single
// splitter warp, 18.9 ms with the X-producer warp as a second splitter.
constexpr bool kCrossAtb = kCrossEnabled && (PETAL_TC_CROSS_ATB != 0);
// Used by tc_xb (X B pass at c2: 11.3 -> 10.4 ms under the power cap, 1.93 -> 1.68 ms = 80 % of the HBM roofline in a
// 2M-row burst; same accuracy in every parity test).  tc_atb keeps the 3 x tf32 form, see kCrossAtb.

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]),
        "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]),
        "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// 16 lanes x 16 columns: thread t holds (lane t/4, cols 2(t%4)+{0,1}) in v[0..1], (lane t/4+8, same cols) in
// v[2..3], and the same two rows for cols 8+2(t%4)+{0,1} in v[4..5] / v[6..7]  -> 4 lanes cover one 32 B sector
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor (sm_100 UMMA): SWIZZLE_128B, version 1
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;  // SWIZZLE_128B
    return d;
}
// instruction descriptor kind::tf32, fp32 accumulate, M = 128, N = n, A K-major (TMEM), B major as given
__host__ __device__ inline uint32_t make_idesc_tf32(int n, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(b_mn_major ? 1 : 0) << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(128 >> 4) << 24);
}
// instruction descriptor kind::f16 with bf16 A / B (format 1), fp32 accumulate, M = 128, N = n, both K-major
__host__ __device__ inline uint32_t make_idesc_bf16(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// ------------------------------------------------------------------------------------------
// kernel parameters
// ------------------------------------------------------------------------------------------
struct TcParams {
    CUtensorMap map_x;    // tc_xb : X as {K inner, rows}, box {32, 256}, SWIZZLE_128B
                          // tc_atb: X as {features inner, rows}, 8 boxes {32, 32}, SWIZZLE_128B
    CUtensorMap map_bhi;  // tc_xb : B^T hi as {K inner, n_pad}, box {32, n_pad}, SWIZZLE_128B
                          // tc_atb: Y as {cols inner, rows}, box {n_pad, 32}, no swizzle
    CUtensorMap map_blo;  // tc_xb : B^T lo (same shape as hi); unused by tc_atb
    const float* mu_pad;  // tc_xb: [K rounded up to 32] zero padded; tc_atb: [features rounded up to 256]; never null
    const float* mub_pad; // tc_atb, row-major Y only: column means of Y [n_pad] subtracted on load (nullable)
    int64_t n;            // rows
    int64_t K;            // tc_xb: reduction length (features); tc_atb: da (features)
    int n_pad;            // MMA N (multiple of 16, <= 128)
    int L;                // valid output columns
    int stages;           // X ring depth (X tile + mu slice; released as soon as the transform has read it)
    int stages_b;         // B ring depth (B / Y tiles; released by the MMAs, or by the transform for row-major Y)
    // tc_xb outputs
    float* Y;
    float* Ylo;           // tc_xb panel mode: second panel with y - tf32(y) (the B_lo operand of a later tc_atb)
    int64_t ldy;
    int y_vec;            // Y rows may be written with 16 B stores
    int y_panel;          // tc_xb: write Y panel-major [row block of 32][n_pad][32]; tc_atb: B operand is panel-major
    int b_split;          // tc_atb, panel-major Y: only the Y panel is loaded; a splitter warp derives the y - tf32(y)
                          // operand tile in shared memory (no Y_lo panel in HBM)
    double* sumsq;        // nullable
    const float* bias;    // tc_xb, row-major Y: per output column, added in the epilogue (nullable)
    int64_t y_cols;       // tc_xb, row-major Y: columns of a Y row that may be written (the row pitch, or the width of
                          // a column block when Y is a window of a wider matrix)
    int nwin;             // tc_xb, row-major Y wider than one MMA N: number of n_pad-column windows of Y produced per
                          // 256-row super-tile (B^T is [nwin * n_pad][K]); 1 = the plain kernel.  A CTA walks the windows
                          // of a super-tile back to back, so the narrow A tile is re-read from L2, not from HBM
    int ones_col_p1;      // tc_xb, panel-major Y: 1 + index of a padding column that is written as 1.0 (valid rows) instead of
                          // 0 - the X^T Y pass that follows then delivers the column sums of X - mu for free; 0 = none
    // tc_atb outputs / decomposition
    double* Z;            // [da x ldz] f64, atomically accumulated
    int64_t ldz;
    int64_t chunk_rows;   // rows per TMEM accumulation (multiple of 32)
    int64_t slice_rows;   // rows per CTA slice (multiple of 32)
    int fgroups;          // feature groups of 256
    int dbg;
    long long* trace;     // optional timeline trace (PETAL_TC_TRACE), [role][event][kblock]
};

struct SmemLayout {
    uint32_t x, bhi, blo, mu, ylo, bars, tmem_slot, total;
    uint32_t stage_x, stage_b, stage_mu, ylo_bytes;
};

// One carve-up shared by host (size) and device (offsets). Offsets are relative to a 1024 B aligned base.
__host__ __device__ inline SmemLayout make_layout(bool atb, int n_pad, int stages, int stages_b, bool panel = false) {
    SmemLayout l;
    l.stage_x = kXStageBytes;
    l.stage_b = (uint32_t)n_pad * 128u;   // tc_xb: B^T tile [n_pad][32]; tc_atb: Y tile (raw [32][n_pad] or panel [n_pad][32])
    l.stage_mu = 128;
    // tc_atb, row-major Y: ring of kYBufs K-major operand tiles [n_pad][32], Y_hi | Y_lo, produced by the transform
    // warps (transpose + split).  Panel-major Y needs none of that: Y_hi / Y_lo panels are TMA-loaded like tc_xb's B.
    l.ylo_bytes = (atb && !panel) ? 2u * (uint32_t)n_pad * 128u : 0u;
    uint32_t off = 0;
    l.x = off;
    off += (uint32_t)stages * l.stage_x;
    l.bhi = off;
    off += (uint32_t)stages_b * l.stage_b;
    l.blo = off;
    off += (atb && !panel) ? 0u : (uint32_t)stages_b * l.stage_b;
    l.ylo = off;
    off += (uint32_t)kYBufs * l.ylo_bytes;
    l.mu = off;
    off += atb ? 0u : (uint32_t)stages * l.stage_mu;
    l.bars = off;
    off += 64 * 8;
    l.tmem_slot = off;
    off += 16;
    l.total = off + 1024;  // slack for the 1024 B alignment of the base
    return l;
}

// The X ring is what hides HBM latency (a stage is recycled as soon as the transform warps have pulled it into
// registers); the B ring only has to cover the MMA latency.  Give B three stages and X the rest.
inline bool pick_stages(bool atb, int n_pad, bool panel, int& sx, int& sb) {
    for (sb = 3; sb >= 2; --sb)
        for (sx = 7; sx >= 2; --sx)
            if (make_layout(atb, n_pad, sx, sb, panel).total <= 222 * 1024) return true;
    return false;
}

// barrier indices inside the `bars` block (64 slots)
__device__ __forceinline__ uint32_t bar_full(uint32_t base, int s) { return base + 8u * (uint32_t)s; }            // X ring
__device__ __forceinline__ uint32_t bar_empty_x(uint32_t base, int s) { return base + 8u * (8 + (uint32_t)s); }
__device__ __forceinline__ uint32_t bar_full_b(uint32_t base, int s) { return base + 8u * (16 + (uint32_t)s); }   // B ring
__device__ __forceinline__ uint32_t bar_empty_b(uint32_t base, int s) { return base + 8u * (24 + (uint32_t)s); }
__device__ __forceinline__ uint32_t bar_a_ready(uint32_t base, int t) { return base + 8u * (t < 4 ? 32u + (uint32_t)t : 56u); }  // 5 slots
__device__ __forceinline__ uint32_t bar_a_free(uint32_t base, int t) { return base + 8u * (t < 4 ? 36u + (uint32_t)t : 57u); }
__device__ __forceinline__ uint32_t bar_acc_full(uint32_t base, int b) { return base + 8u * (40 + (uint32_t)b); }   // 40 .. 42
__device__ __forceinline__ uint32_t bar_acc_empty(uint32_t base, int b) { return base + 8u * (47 + (uint32_t)b); }  // 47 .. 49
__device__ __forceinline__ uint32_t bar_y_free(uint32_t base, int t) { return base + 8u * (44 + (uint32_t)t); }
__device__ __forceinline__ uint32_t bar_full_b2(uint32_t base, int s) { return base + 8u * (50 + (uint32_t)s); }  // B ring, after the splitter

// ------------------------------------------------------------------------------------------
// the kernel (ATB = false: tc_xb, ATB = true: tc_atb; NP = compile-time n_pad for tc_atb)
//
// Work is cut into "groups": one TMEM accumulation each.
//   tc_xb : group = super-tile of 256 rows, K loop over ceil(K / 32) feature blocks; the epilogue
//           stores the 256 x n_pad tile of Y.  Groups are dealt round-robin to the persistent CTAs.
//   tc_atb: the CTA owns one feature group (256 features) and one contiguous slice of rows.  Fast mode: a group
//           is a chunk of <= 1024 rows of that slice (K loop over its 32-row blocks) = one accumulation chain;
//           precise mode: one group per CTA, chains of 4 K blocks staggered between the M tiles (ModeTraits).  The
//           tensor core adds into its fp32 accumulator with truncation, which biases long chains: finished chains
//           are added into fp32 registers (round-to-nearest) and only the CTA's final sums go to global memory
//           (f64 atomics).
// ------------------------------------------------------------------------------------------
constexpr int kTraceKB = 512;   // K blocks traced (CTA 0 only)
constexpr int kTraceEvents = 12;
// timeline tracing is a build option (-DPETAL_TC_TRACE_BUILD): even a not-taken trace check costs a parameter load, a
// special-register read and a compare per event in loops that are issue-bound
__device__ __forceinline__ void trace_ev(const TcParams& p, int ev, uint32_t it) {
#ifdef PETAL_TC_TRACE_BUILD
    if (p.trace != nullptr && blockIdx.x == 0 && it < (uint32_t)kTraceKB) p.trace[ev * kTraceKB + it] = clock64();
#else
    (void)p; (void)ev; (void)it;
#endif
}

struct Group {
    int64_t row0;
    int64_t kblocks;
    int f0;
    int win;  // tc_xb: output window (see TcParams::nwin)
};

// Position in a ring of `n` stages, advanced incrementally: a runtime `it % n`, `it / n` pair costs an
// I2F / MUFU.RCP / IMAD.HI sequence per use in loops that are issue-bound.
struct RingPos {
    int s = 0;
    uint32_t ph = 0;
    int n;
    __device__ __forceinline__ explicit RingPos(int n_) : n(n_) {}
    __device__ __forceinline__ void advance() {
        if (++s == n) {
            s = 0;
            ph ^= 1u;
        }
    }
};

template <bool ATB>
__device__ __forceinline__ bool get_group(const TcParams& p, int64_t g, Group& out) {
    if (ATB) {
        const int fg = (int)(blockIdx.x % (unsigned)p.fgroups);
        const int64_t sl = blockIdx.x / (unsigned)p.fgroups;
        const int64_t s0 = sl * p.slice_rows;
        const int64_t s1 = min(p.n, s0 + p.slice_rows);
        const int64_t r0 = s0 + g * p.chunk_rows;
        if (r0 >= s1) return false;
        out.row0 = r0;
        out.kblocks = (min(p.chunk_rows, s1 - r0) + kKB - 1) / kKB;
        out.f0 = fg * 256;
        return true;
    } else {
        const uint32_t tile_round = (uint32_t)g / (uint32_t)p.nwin;
        const int64_t item = (int64_t)blockIdx.x + (int64_t)tile_round * (int64_t)gridDim.x;
        if (item * 256 >= p.n) return false;
        out.row0 = item * 256;
        out.kblocks = (p.K + kKB - 1) / kKB;
        out.f0 = 0;
        out.win = (int)((uint32_t)g - tile_round * (uint32_t)p.nwin);
        return true;
    }
}

template <int N>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

// Kernel modes (template parameter MODE); everything derived from it is a compile-time constant - the transform and
// MMA issue loops are instruction-bound enough that a runtime modulo in them costs 10-20 %.
//   0  fast     : 4-slot operand ring, one accumulator set, one accumulation chain per group
//   1  fast2    : tc_xb only, n_pad <= 80: 3-slot ring, two accumulator sets alternating per super-tile, the epilogue of
//                 tile g drained inside the K loop of tile g+1.  Not instantiated: measured 6-8 % slower than mode 0 (the
//                 3-slot ring costs more than the hidden epilogue gains); with a 4-slot ring (n_pad <= 64) it was a wash.
//   2  precise  : accumulation chains cut (the tensor core adds into its accumulator with truncation):
//                 tc_xb : 3-slot ring, two sets, chains of 4 K blocks
//                 tc_atb: 4-slot ring, three one-M-tile buffers, chains of 4 K blocks staggered between the M tiles
template <bool ATB, int MODE>
struct ModeTraits {
    static constexpr bool kCut = (MODE == 2);
    // MODE 3 (tc_xb, n_pad <= 96): the fast mode with a 5-slot operand ring (2.5 K blocks between the transform warps
    // and the MMA issuers instead of 2): both sides spend about half of their time waiting for each other (r02 ncu
    // source page: 37 % of all warp samples of tc_xb are transform warps waiting for a free operand slot while the
    // tensor pipe is 57 % busy), so the deeper ring buys overlap
    static constexpr int kSlots = (!ATB && MODE == 3) ? 5 : ((!ATB && MODE != 0) ? 3 : 4);
    static constexpr int kAccBase = kSlots * kASlotCols;
    static constexpr int kAccBufs = ATB ? (MODE == 2 ? 3 : 1) : ((MODE == 0 || MODE == 3) ? 1 : 2);
    static constexpr bool kStagger = ATB && MODE == 2;
    static constexpr int kChainKB = 4;  // tc_xb precise: K blocks per chain (48 accumulating MMAs)
};

// transform + epilogue role (warps 0 .. 15).  warp w: M tile (w >> 2) & 1, TMEM lane quarter w & 3,
// K-block half w >> 3.  EPI: this warp also runs the epilogue (tc_xb: all; tc_atb: first half only -
// they hold the register accumulators, hence the separate instantiation and register budget).
template <bool ATB, int NP, bool PANEL, int MODE>
__device__ __forceinline__ void transform_role(const TcParams& p, uint8_t* base_ptr, const SmemLayout& L,
                                               uint32_t bars, uint32_t tmem_base, int warp, int lane, int n_pad,
                                               int S) {
    {
        const int half = warp >> 3;
        const int mt = (warp >> 2) & 1;
        const int q = warp & 3;
        const int lrow = q * 32 + lane;           // lane (= row / feature) inside the M tile
        const uint32_t lane_field = (uint32_t)(q * 32) << 16;
        const int ttid = threadIdx.x;             // 0 .. 511
        constexpr bool panel = PANEL;
        uint32_t it = 0;
        double ss = 0.0;
        // tc_atb: the two warps of a lane-quarter pair split the accumulator columns in alternate 16-column chunks,
        // so each thread keeps at most ceil(NP/32) * 16 running sums in registers
        // tc_xb with chain cutting (n_pad <= 80): 3 chunks of 16 (panel) or 5 chunks of 8 (row-major) -> 48
        using MT_ = ModeTraits<ATB, MODE>;
        constexpr bool CUT = MT_::kCut;
        constexpr int kSlots = MT_::kSlots;
        constexpr int kAccBaseC = MT_::kAccBase;
        constexpr int acc_bufs = MT_::kAccBufs;
        constexpr int kAccChunks = ATB ? (NP / 16 + 1) / 2 : 3;  // tc_xb: only touched when CUT
        float racc[kAccChunks * 16];
#pragma unroll
        for (int j = 0; j < kAccChunks * 16; ++j) racc[j] = 0.f;
        // tc_atb: this thread's share of the Y-tile transposition (loop invariant): chunk i covers
        // column nn = i % n_pad, K rows 4*cc .. 4*cc+3 with cc = i / n_pad
        constexpr int kYIter = (ATB && !PANEL) ? (NP * 8 + kTransformWarps * 32 - 1) / (kTransformWarps * 32) : 1;
        int y_src[kYIter], y_dst[kYIter], y_row[kYIter];
        float y_mu[kYIter];
        if (ATB && !PANEL) {
#pragma unroll
            for (int u = 0; u < kYIter; ++u) {
                const int i = ttid + u * kTransformWarps * 32;
                const int nn = i % n_pad, cc = i / n_pad;
                y_src[u] = (i < n_pad * 8) ? (cc * 4 * n_pad + nn) : -1;
                y_dst[u] = nn * 128 + ((cc ^ (nn & 7)) << 4);
                y_row[u] = cc * 4;
                y_mu[u] = (p.mub_pad != nullptr && i < n_pad * 8) ? p.mub_pad[nn] : 0.f;
            }
        }
        bool pend_on = false;
        int64_t pend_row0 = 0;
        // ======== fast mode (CUT == false): one chain per group ========
        // ---- tc_xb epilogue, drained one 16-column chunk at a time.  With two accumulator sets the drain of
        // super-tile g is deferred into the K loop of super-tile g+1 (one chunk after each of its first K blocks),
        // so the tensor pipe keeps running while Y is written.
        const int xb_chunks = ATB ? 0 : (panel ? (n_pad - half * 16 + 31) / 32 : n_pad / 16);
        uint32_t pend_gi = 0;
        int pend_next = 0;
        int pend_wcol = 0;  // first column of the pending super-tile's output window
        auto store_rowmajor_chunk = [&](int64_t row0, int c0w, const float* w) {
            const int c0 = pend_wcol + c0w;
            // 16x256b pattern: four lanes hold 8 consecutive columns of one row, so every store instruction writes
            // whole 32 B sectors (8 rows x 32 B).  This warp covers lanes 16*half .. 16*half+15 of its quarter.
            const int t0 = lane & 3, t1 = lane >> 2;
            const int64_t ra = row0 + mt * 128 + q * 32 + half * 16 + t1;
            const int64_t rb = ra + 8;
            const int ca = c0 + 2 * t0, cb = ca + 8;
            // optional per-column bias (inverse_transform: + mean)
            float ba0 = 0.f, ba1 = 0.f, bb0 = 0.f, bb1 = 0.f;
            if (p.bias != nullptr) {
                if (ca < p.L) ba0 = p.bias[ca];
                if (ca + 1 < p.L) ba1 = p.bias[ca + 1];
                if (cb < p.L) bb0 = p.bias[cb];
                if (cb + 1 < p.L) bb1 = p.bias[cb + 1];
            }
            if (p.y_vec) {
                if (ra < p.n) {
                    if (ca + 1 < p.y_cols) *reinterpret_cast<float2*>(p.Y + ra * p.ldy + ca) = make_float2(w[0] + ba0, w[1] + ba1);
                    if (cb + 1 < p.y_cols) *reinterpret_cast<float2*>(p.Y + ra * p.ldy + cb) = make_float2(w[4] + bb0, w[5] + bb1);
                }
                if (rb < p.n) {
                    if (ca + 1 < p.y_cols) *reinterpret_cast<float2*>(p.Y + rb * p.ldy + ca) = make_float2(w[2] + ba0, w[3] + ba1);
                    if (cb + 1 < p.y_cols) *reinterpret_cast<float2*>(p.Y + rb * p.ldy + cb) = make_float2(w[6] + bb0, w[7] + bb1);
                }
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int64_t r = (e & 2) ? rb : ra;
                    const int c = ((e & 4) ? cb : ca) + (e & 1);
                    const float bv = (e & 4) ? ((e & 1) ? bb1 : bb0) : ((e & 1) ? ba1 : ba0);
                    if (r < p.n && c < p.y_cols) p.Y[r * p.ldy + c] = w[e] + bv;
                }
            }
        };
        auto xb_drain_chunk = [&](int64_t row0, int buf, int j) {
            const uint32_t acc_col = (uint32_t)(kAccBaseC + buf * kMT * n_pad + mt * n_pad);
            if (panel) {
                // panel-major Y [row block of 32][n_pad][32]: lane = row inside the block, so for each
                // column the warp writes one full 128 B line.  The two warps of a lane-quarter pair
                // take alternate 16-column chunks.  Rows past n are written as zeros.
                const int c0 = half * 16 + 32 * j;
                const int64_t r = row0 + mt * 128 + lrow;
                const int64_t rblk = (row0 + mt * 128 + q * 32) >> 5;
                float* yb = p.Y + (rblk * n_pad) * 32 + lane;
                float* yl = p.Ylo + (rblk * n_pad) * 32 + lane;
                const bool valid = r < p.n;
                const bool blk_valid = rblk * 32 < p.n;  // the panel buffer ends at the last partial row block
                uint32_t w[16];
                tmem_ld16(tmem_base + lane_field + acc_col + (uint32_t)c0, w);
                tmem_ld_wait();
                if (blk_valid) {
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) {
                        float y = valid ? __uint_as_float(w[jj]) : 0.f;
                        if (c0 + jj + 1 == p.ones_col_p1) y = valid ? 1.f : 0.f;
                        yb[(c0 + jj) * 32] = y;
                        if (p.Ylo != nullptr) yl[(c0 + jj) * 32] = y - __uint_as_float(__float_as_uint(y) & 0xFFFFE000u);
                    }
                }
            } else {
                // Each warp of a lane-quarter pair drains 16 of the 32 lanes with the 16x256b pattern:
                // four lanes hold 8 consecutive columns of one row, so every store instruction writes
                // whole 32 B sectors (8 rows x 32 B).
                const int c0 = 16 * j;
                const uint32_t acc16 = tmem_base + ((uint32_t)(q * 32 + half * 16) << 16) + acc_col;
                uint32_t w[8];
                tmem_ld_16x256b_x2(acc16 + (uint32_t)c0, w);
                tmem_ld_wait();
                store_rowmajor_chunk(row0, c0, reinterpret_cast<const float*>(w));
            }
        };
        auto xb_drain_step = [&]() {
            const int buf = (acc_bufs == 2) ? (int)(pend_gi & 1u) : 0;
            if (pend_next == 0) {
                mbar_wait(bar_acc_full(bars, buf), (acc_bufs == 2) ? ((pend_gi >> 1) & 1u) : (pend_gi & 1u));
                tc_fence_after();
            }
            if (pend_next < xb_chunks) xb_drain_chunk(pend_row0, buf, pend_next);
            if (++pend_next >= xb_chunks) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_acc_empty(bars, buf));
                pend_on = false;
            }
        };
        // ======== precise mode (CUT == true) ========
        // ---- accumulator hand-off.  The tensor core adds into its fp32 accumulator with truncation (bias ~ -2e-8
        // per add), so accumulation chains are kept short: with two accumulator sets in TMEM (`cut`) a chain is
        // p.chain_kb K blocks; each finished chain is added into fp32 registers (round to nearest) one K block
        // later, while the MMAs already fill the other set.  tc_xb writes Y from the registers at the end of a
        // super-tile, tc_atb sends the registers to global memory once per CTA.  With a single accumulator set
        // (n_pad > 80) a chain is a whole group and tc_xb drains TMEM straight to global memory.
        constexpr bool cut = true;
        constexpr int chain_kb = MT_::kChainKB;
        uint32_t chains = 0;  // chains handed to the MMA warps so far (both sides count alike)
        bool pend_flush = false;
        uint32_t pend_ch = 0;
        // TMEM accumulator set -> += registers
        auto drain_acc = [&](uint32_t acc_col) {
            if constexpr (ATB) {
#pragma unroll
                for (int ch = 0; ch < kAccChunks; ++ch) {
                    const int c0 = (2 * ch + half) * 16;
                    if (c0 < NP) {
                        uint32_t w[16];
                        tmem_ld16(tmem_base + lane_field + acc_col + (uint32_t)c0, w);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) racc[ch * 16 + j] += __uint_as_float(w[j]);
                    }
                }
            } else if (panel) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const int c0 = half * 16 + 32 * ch;
                    if (c0 < n_pad) {
                        uint32_t w[16];
                        tmem_ld16(tmem_base + lane_field + acc_col + (uint32_t)c0, w);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) racc[ch * 16 + j] += __uint_as_float(w[j]);
                    }
                }
            } else {
                const uint32_t acc16 = tmem_base + ((uint32_t)(q * 32 + half * 16) << 16) + acc_col;
#pragma unroll
                for (int ch = 0; ch < 5; ++ch) {
                    if (16 * ch < n_pad) {
                        uint32_t w[8];
                        tmem_ld_16x256b_x2(acc16 + (uint32_t)(16 * ch), w);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 8; ++e) racc[ch * 8 + e] += __uint_as_float(w[e]);
                    }
                }
            }
        };
        // one 16-column chunk of a tc_xb super-tile -> global memory; w = the chunk's values in this thread's layout
        auto store_panel_chunk = [&](int64_t row0, int c0, const float* w) {
            // panel-major Y [row block of 32][n_pad][32]: lane = row inside the block, so for each column the warp
            // writes one full 128 B line.  Rows past n are written as zeros.
            const int64_t r = row0 + mt * 128 + lrow;
            const int64_t rblk = (row0 + mt * 128 + q * 32) >> 5;
            if (rblk * 32 >= p.n) return;  // the panel buffer ends at the last partial row block
            float* yb = p.Y + (rblk * n_pad) * 32 + lane;
            float* yl = p.Ylo + (rblk * n_pad) * 32 + lane;
            const bool valid = r < p.n;
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                float y = valid ? w[jj] : 0.f;
                if (c0 + jj + 1 == p.ones_col_p1) y = valid ? 1.f : 0.f;
                yb[(c0 + jj) * 32] = y;
                if (p.Ylo != nullptr) yl[(c0 + jj) * 32] = y - __uint_as_float(__float_as_uint(y) & 0xFFFFE000u);
            }
        };
        // tc_xb, registers -> Y for the super-tile starting at row0, registers cleared
        auto flush_y = [&](int64_t row0) {
            if (panel) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch)
                    if (half * 16 + 32 * ch < n_pad) store_panel_chunk(row0, half * 16 + 32 * ch, racc + ch * 16);
            } else {
#pragma unroll
                for (int ch = 0; ch < 5; ++ch)
                    if (16 * ch < n_pad) store_rowmajor_chunk(row0, 16 * ch, racc + ch * 8);
            }
#pragma unroll
            for (int j = 0; j < kAccChunks * 16; ++j) racc[j] = 0.f;
        };
        // tc_xb with a single accumulator set: TMEM -> global, chunk by chunk
        auto xb_direct = [&](int64_t row0) {
            const uint32_t acc_col = (uint32_t)(kAccBaseC + mt * n_pad);
            if (panel) {
                for (int c0 = half * 16; c0 < n_pad; c0 += 32) {
                    uint32_t w[16];
                    tmem_ld16(tmem_base + lane_field + acc_col + (uint32_t)c0, w);
                    tmem_ld_wait();
                    store_panel_chunk(row0, c0, reinterpret_cast<const float*>(w));
                }
            } else {
                const uint32_t acc16 = tmem_base + ((uint32_t)(q * 32 + half * 16) << 16) + acc_col;
                for (int c0 = 0; c0 < n_pad; c0 += 16) {
                    uint32_t w[8];
                    tmem_ld_16x256b_x2(acc16 + (uint32_t)c0, w);
                    tmem_ld_wait();
                    store_rowmajor_chunk(row0, c0, reinterpret_cast<const float*>(w));
                }
            }
        };
        constexpr bool stagger = MT_::kStagger;
        int pend_kb = 0;
        auto do_pending = [&]() {
            // two full sets: chain c uses set c & 1.  staggered: chain j (of one M tile) uses buffer j % 3.
            const int buf = stagger ? (int)(pend_ch % 3u) : (int)(pend_ch & 1u);
            const uint32_t par = stagger ? ((pend_ch / 3u) & 1u) : ((pend_ch >> 1) & 1u);
            const uint32_t acc_col = stagger ? (uint32_t)(kAccBaseC + buf * n_pad)
                                             : (uint32_t)(kAccBaseC + buf * kMT * n_pad + mt * n_pad);
            mbar_wait(bar_acc_full(bars, buf), par);
            tc_fence_after();
            if (ATB || cut) drain_acc(acc_col);
            else xb_direct(pend_row0);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty(bars, buf));
            if (!ATB && cut && pend_flush) flush_y(pend_row0);
            pend_on = false;
        };
        // cross mode, tc_atb panel: the (up to two) 16 B chunks of the Y block this thread converts, loop invariant.
        // Chunk e = (row n = e / 8, logical chunk c = e % 8) holds y[k = 4c .. 4c+3]; K step g = c / 2; its bf16 words go
        // to bytes 8 (c % 2) .. of output chunk 2g (values) and 2g + 1 (low parts), all with the row's 128 B swizzle.
        uint32_t yx_in[2] = {0xFFFFFFFFu, 0xFFFFFFFFu}, yx_hi[2] = {0, 0}, yx_lo[2] = {0, 0};
        if (kCrossAtb && ATB && PANEL) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int e = ttid + u * kTransformWarps * 32;
                if (e < NP * 8) {
                    const int nrow = e >> 3, c = e & 7, sw = nrow & 7, g2 = c >> 1, hf = c & 1;
                    yx_in[u] = (uint32_t)(nrow * 128 + ((c ^ sw) << 4));
                    yx_hi[u] = (uint32_t)(nrow * 128 + (((2 * g2) ^ sw) << 4) + hf * 8);
                    yx_lo[u] = (uint32_t)(nrow * 128 + (((2 * g2 + 1) ^ sw) << 4) + hf * 8);
                }
            }
        }
        Group g;
        g.f0 = 0;
        RingPos rx(S), rb(p.stages_b), ry(kYBufs);
        for (int64_t gi = 0; get_group<ATB>(p, gi, g); ++gi) {
            float mu_f = 0.f;
            if (ATB) mu_f = p.mu_pad[g.f0 + mt * 128 + lrow];
            const bool row_valid = ATB ? true : (g.row0 + mt * 128 + lrow < p.n);
            float ss0 = 0.f, ss1 = 0.f, ss2 = 0.f, ss3 = 0.f;
            for (int64_t kb = 0; kb < g.kblocks; ++kb, ++it, rx.advance()) {
                const int s = rx.s;
                const uint32_t ph = rx.ph;
                const uint32_t gran = 2u * it + (uint32_t)half;  // this warp's half K block
                const int ta = (int)(gran % (uint32_t)kSlots);
                const uint32_t pa = (gran / (uint32_t)kSlots) & 1u;
                mbar_wait(bar_full(bars, s), ph);
                if (warp == 0 && lane == 0) trace_ev(p, 3, it);
                uint32_t v[16];
                const uint8_t* xs = base_ptr + L.x + (uint32_t)s * L.stage_x;
                if (ATB) {
                    // 8 sub-tiles [32 rows][32 features] (128 B rows, SWIZZLE_128B): this thread owns one feature
                    // (transpose) and the K rows 16*half .. 16*half+15
                    const int f = mt * 128 + lrow;
                    const uint8_t* xsub = xs + (f >> 5) * 4096 + (f & 3) * 4;
                    const int fc = (f & 31) >> 2;  // 16 B chunk of the feature inside its 128 B row
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const int kr = half * 16 + k;
                        v[k] = __float_as_uint(*reinterpret_cast<const float*>(xsub + kr * 128 + ((fc ^ (kr & 7)) << 4)) - mu_f);
                    }
                    __syncwarp();  // smem X stage consumed (values are in registers)
                    if (lane == 0) mbar_arrive(bar_empty_x(bars, s));
                    if (warp == 0 && lane == 0) trace_ev(p, 7, it);
                    if constexpr (kCrossAtb && PANEL) {
                        // this thread's share of the B-side cross tile of the K block (see kCrossAtb)
                        const int sbq = rb.s;
                        mbar_wait(bar_full_b(bars, sbq), rb.ph);
                        const uint8_t* bh8 = base_ptr + L.bhi + (uint32_t)sbq * L.stage_b;
                        uint8_t* bl8 = base_ptr + L.blo + (uint32_t)sbq * L.stage_b;
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            if (yx_in[u] != 0xFFFFFFFFu) {
                                const float4 y = *reinterpret_cast<const float4*>(bh8 + yx_in[u]);
                                const float l0 = y.x - __uint_as_float(__float_as_uint(y.x) & 0xFFFFE000u);
                                const float l1 = y.y - __uint_as_float(__float_as_uint(y.y) & 0xFFFFE000u);
                                const float l2 = y.z - __uint_as_float(__float_as_uint(y.z) & 0xFFFFE000u);
                                const float l3 = y.w - __uint_as_float(__float_as_uint(y.w) & 0xFFFFE000u);
                                *reinterpret_cast<uint2*>(bl8 + yx_hi[u]) = make_uint2(pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w));
                                *reinterpret_cast<uint2*>(bl8 + yx_lo[u]) = make_uint2(pack_bf16x2(l0, l1), pack_bf16x2(l2, l3));
                            }
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_full_b2(bars, sbq));
                        rb.advance();
                    }
                } else {
                    // tile [256 rows][32 floats], 128 B rows, SWIZZLE_128B: this thread owns one row and
                    // the 16 B chunks 4*half .. 4*half+3 of it
                    const int r = mt * 128 + lrow;
                    const uint8_t* xr = xs + r * 128;
                    const float4* mus = reinterpret_cast<const float4*>(base_ptr + L.mu + (uint32_t)s * L.stage_mu);
#pragma unroll
                    for (int cq = 0; cq < 4; ++cq) {
                        const int c = half * 4 + cq;
                        float4 x4 = *reinterpret_cast<const float4*>(xr + ((c ^ (r & 7)) << 4));
                        float4 m = mus[c];
                        float e0 = x4.x - m.x, e1 = x4.y - m.y, e2 = x4.z - m.z, e3 = x4.w - m.w;
                        v[cq * 4 + 0] = __float_as_uint(e0);
                        v[cq * 4 + 1] = __float_as_uint(e1);
                        v[cq * 4 + 2] = __float_as_uint(e2);
                        v[cq * 4 + 3] = __float_as_uint(e3);
                        if (p.sumsq) {
                            ss0 += e0 * e0;
                            ss1 += e1 * e1;
                            ss2 += e2 * e2;
                            ss3 += e3 * e3;
                        }
                    }
                    __syncwarp();  // smem X stage consumed (values are in registers)
                    if (lane == 0) mbar_arrive(bar_empty_x(bars, s));
                }
                if (ATB && !panel) {
                    const int sb = rb.s;
                    mbar_wait(bar_full_b(bars, sb), rb.ph);  // Y tile landed
                    if (warp == 0 && lane == 0) trace_ev(p, 8, it);
                    const int yb = ry.s;
                    mbar_wait(bar_y_free(bars, yb), ry.ph ^ 1u);  // MMAs of K block it-3 done
                    rb.advance();
                    ry.advance();
                    if (warp == 0 && lane == 0) trace_ev(p, 9, it);
                    {
                    // B operand of this K block: raw Y tile [32 rows][n_pad] (row-major) -> transposed
                    // K-major tiles Y_hi / Y_lo [n_pad][32 rows] with the 128 B swizzle the MMA expects.
                    const float* yr = reinterpret_cast<const float*>(base_ptr + L.bhi + (uint32_t)sb * L.stage_b);
                    uint8_t* bh = base_ptr + L.ylo + (uint32_t)yb * L.ylo_bytes;
                    uint8_t* bl = bh + n_pad * 128;
#pragma unroll
                    for (int u = 0; u < kYIter; ++u) {
                        if (y_src[u] < 0) continue;
                        const float* ys = yr + y_src[u];
                        float4 h, l4;
                        h.x = ys[0];
                        h.y = ys[n_pad];
                        h.z = ys[2 * n_pad];
                        h.w = ys[3 * n_pad];
                        if (p.mub_pad != nullptr) {
                            // centred Y: rows past n are TMA zero fill and must stay zero
                            const int64_t r0 = g.row0 + kb * kKB + y_row[u];
                            h.x = (r0 + 0 < p.n) ? h.x - y_mu[u] : 0.f;
                            h.y = (r0 + 1 < p.n) ? h.y - y_mu[u] : 0.f;
                            h.z = (r0 + 2 < p.n) ? h.z - y_mu[u] : 0.f;
                            h.w = (r0 + 3 < p.n) ? h.w - y_mu[u] : 0.f;
                        }
                        l4.x = h.x - __uint_as_float(__float_as_uint(h.x) & 0xFFFFE000u);
                        l4.y = h.y - __uint_as_float(__float_as_uint(h.y) & 0xFFFFE000u);
                        l4.z = h.z - __uint_as_float(__float_as_uint(h.z) & 0xFFFFE000u);
                        l4.w = h.w - __uint_as_float(__float_as_uint(h.w) & 0xFFFFE000u);
                        *reinterpret_cast<float4*>(bh + y_dst[u]) = h;
                        *reinterpret_cast<float4*>(bl + y_dst[u]) = l4;
                    }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_empty_b(bars, sb));  // TMA-landed Y tile consumed
                }
                // TMEM operand slot `ta` must have been drained by the MMAs of two K blocks ago
                if (warp == 0 && lane == 0) trace_ev(p, 4, it);
                mbar_wait_relaxed(bar_a_free(bars, ta), pa ^ 1u, 48);
                tc_fence_after();
                if (warp == 0 && lane == 0) trace_ev(p, 5, it);
                const uint32_t a_addr = tmem_base + lane_field + (uint32_t)(ta * kASlotCols + mt * 32);
                tmem_st16(a_addr, v);  // hi: the tensor core ignores the low 13 mantissa bits
                if constexpr ((kCrossEnabled && !ATB) || (kCrossAtb && ATB && PANEL)) {
                    // cross operand, per K step of 8 values: 4 words of bf16 pairs of lo = v - tf32(v), then 4 words of bf16
                    // pairs of v (K order [lo 0..7 | hi 0..7], matching the [hi ; lo] order of the B-side cross tile)
                    uint32_t cw[16];
#pragma unroll
                    for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float f0 = __uint_as_float(v[8 * s2 + 2 * j]), f1 = __uint_as_float(v[8 * s2 + 2 * j + 1]);
                            const float l0 = f0 - __uint_as_float(v[8 * s2 + 2 * j] & 0xFFFFE000u);
                            const float l1 = f1 - __uint_as_float(v[8 * s2 + 2 * j + 1] & 0xFFFFE000u);
                            cw[8 * s2 + j] = pack_bf16x2(l0, l1);
                            cw[8 * s2 + 4 + j] = pack_bf16x2(f0, f1);
                        }
                    tmem_st16(a_addr + 16u, cw);
                } else {
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const float f = __uint_as_float(v[k]);
                        v[k] = __float_as_uint(f - __uint_as_float(v[k] & 0xFFFFE000u));
                    }
                    tmem_st16(a_addr + 16u, v);  // lo = v - tf32(v), exact
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_a_ready(bars, ta));
                if (warp == 0 && lane == 0) trace_ev(p, 6, it);
                if constexpr (CUT) {
                  if (stagger) {
                    // tc_atb: chains of 4 K blocks per M tile, M tile 1 shifted by 2 K blocks, three accumulator buffers
                    // rotating over the chains in start order (j = 2c + 1 for M tile 0, 2c for M tile 1).  A chain is
                    // added into the registers two K blocks after it ended (its MMAs have completed by then), which
                    // is before the buffer's next owner (chain j + 3) starts.  One group per CTA (host guarantees it).
                    const int kbi = (int)kb;
                    if (pend_on && kbi == pend_kb + 2) do_pending();
                    const bool chain_end = (kb == g.kblocks - 1) || (((kbi + 2 * mt) & 3) == 3);
                    if (chain_end) {
                        if (pend_on) do_pending();
                        pend_on = true;
                        pend_ch = (mt == 0) ? (uint32_t)(2 * (kbi >> 2) + 1) : (uint32_t)(2 * ((kbi + 2) >> 2));
                        pend_kb = kbi;
                    }
                  } else {
                    const bool chain_end = (kb == g.kblocks - 1) || (((int)kb & (chain_kb - 1)) == chain_kb - 1);
                    if (chain_end) {
                        // the previous chain ended chain_kb K blocks ago: its MMAs have long completed, so this never
                        // blocks, and the MMA warps need its accumulator set only for the chain after this one
                        if (pend_on) do_pending();
                        pend_on = true;
                        pend_ch = chains++;
                        pend_flush = (kb == g.kblocks - 1);
                        pend_row0 = g.row0;
                        pend_wcol = g.win * n_pad;
                    }
                  }
                } else {
                    if (!ATB && pend_on && kb >= 1) xb_drain_step();  // previous super-tile, one chunk per K block
                }
            }
            if (!ATB && row_valid) ss += (double)((ss0 + ss1) + (ss2 + ss3));

            if constexpr (!CUT) {
            if constexpr (ATB) {
                mbar_wait(bar_acc_full(bars, 0), (uint32_t)gi & 1u);
                tc_fence_after();
                const uint32_t acc = tmem_base + lane_field + (uint32_t)(kAccBaseC + mt * n_pad);
#pragma unroll
                for (int ch = 0; ch < kAccChunks; ++ch) {
                    const int c0 = (2 * ch + half) * 16;
                    if (c0 < NP) {
                        uint32_t w[16];
                        tmem_ld16(acc + (uint32_t)c0, w);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) racc[ch * 16 + j] += __uint_as_float(w[j]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_acc_empty(bars, 0));
            } else {
                while (pend_on) xb_drain_step();  // whatever is left of the previous super-tile
                pend_on = true;
                pend_row0 = g.row0;
                pend_wcol = g.win * n_pad;
                pend_gi = (uint32_t)gi;
                pend_next = 0;
                if (acc_bufs != 2)
                    while (pend_on) xb_drain_step();  // single accumulator set: drain now
            }
            }
        }
        if constexpr (CUT) {
            if (pend_on) do_pending();
        } else {
            if (!ATB)
                while (pend_on) xb_drain_step();
        }
        if constexpr (ATB) {
            // the CTA's partial (256 features x L) -> global f64 accumulator
            const int64_t f = (int64_t)g.f0 + mt * 128 + lrow;
            if (f < p.K) {
#pragma unroll
                for (int ch = 0; ch < kAccChunks; ++ch)
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int c = (2 * ch + half) * 16 + j;
                        if (c < p.L) atomicAdd(&p.Z[f * p.ldz + c], (double)racc[ch * 16 + j]);
                    }
            }
        } else if (p.sumsq) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            if (lane == 0) atomicAdd(p.sumsq, ss);
        }
    }
}

template <bool ATB, int NP, bool PANEL, int MODE>
__global__ void __launch_bounds__(kThreads, 1) tc_gemm_kernel(const __grid_constant__ TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const int n_pad = ATB ? NP : p.n_pad;
    constexpr bool panel = PANEL;
    const SmemLayout L = make_layout(ATB, n_pad, p.stages, p.stages_b, ATB && panel);
    const uint32_t bars = base + L.bars;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.stages, SB = p.stages_b;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(bar_full(bars, s), 1);
            mbar_init(bar_empty_x(bars, s), kTransformWarps);
        }
        for (int s = 0; s < SB; ++s) {
            mbar_init(bar_full_b(bars, s), 1);
            mbar_init(bar_full_b2(bars, s), (kCrossAtb && ATB && panel) ? kTransformWarps : 1);
            // released by the two MMA warps, or by the transform warps when they consume the raw row-major Y tile
            mbar_init(bar_empty_b(bars, s), (ATB && !panel) ? kTransformWarps : kMT);
        }
        for (int t = 0; t < kASlotsMax; ++t) {
            mbar_init(bar_a_ready(bars, t), kTransformWarps / 2);  // the 8 warps that own this half of a K block
            mbar_init(bar_a_free(bars, t), kMT);
        }
        // staggered tc_atb chains (acc_bufs == 3): a buffer belongs to one M tile at a time
        for (int b = 0; b < 3; ++b) {
            mbar_init(bar_acc_full(bars, b), ModeTraits<ATB, MODE>::kStagger ? 1 : kMT);
            mbar_init(bar_acc_empty(bars, b), ModeTraits<ATB, MODE>::kStagger ? kTransformWarps / 2 : kTransformWarps);
        }
        for (int t = 0; t < kYBufs; ++t) mbar_init(bar_y_free(bars, t), kMT);
        fence_barrier_init();
    }
    if (warp == kTransformWarps + 3) tmem_alloc(base + L.tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(base_ptr + L.tmem_slot);

    if (warp >= kTransformWarps) {
        if (ATB && panel && p.b_split && warp == kTransformWarps + 3) {
            // ================ tc_atb with in-kernel B_lo: B-ring loader + splitter warp ================
            // Only the Y panel is in HBM.  The rows of a CTA's slice are contiguous, so K block i of the CTA starts at
            // row s0 + 32 i whatever the group structure.  This warp loads the panel blocks itself (lane 0, blocking
            // waits, refilling the stage of block i-1 after block i has been split: the MMAs of block i-1 are finishing
            // then, and SB-1 blocks stay in flight) - the earlier single producer lane polling both rings with
            // non-blocking probes took every other issue slot of its scheduler (r02 ncu: 38 % of the kernel's warp
            // instructions).  The X ring keeps its own producer (warp 16, below).
            const int64_t s0 = (int64_t)(blockIdx.x / (unsigned)p.fgroups) * p.slice_rows;
            const int64_t s1 = min(p.n, s0 + p.slice_rows);
            const uint32_t total = (s1 > s0) ? (uint32_t)((s1 - s0 + kKB - 1) / kKB) : 0u;
            auto load_block = [&](uint32_t blk, int stage) {
                const uint32_t fullb = bar_full_b(bars, stage);
                const int64_t r = s0 + (int64_t)blk * kKB;
                mbar_expect_tx(fullb, L.stage_b);
                tma_load_2d(base + L.bhi + (uint32_t)stage * L.stage_b, &p.map_bhi, 0, (int)(r / 32) * n_pad, fullb);
            };
            if constexpr (kCrossAtb) {
                // cross mode: the transform warps derive the second operand tile; this warp only keeps the B ring full
                if (lane == 0) {
                    RingPos rp(SB);
                    for (uint32_t blk = 0; blk < total; ++blk, rp.advance()) {
                        mbar_wait_relaxed(bar_empty_b(bars, rp.s), rp.ph ^ 1u, 64);
                        load_block(blk, rp.s);
                    }
                }
            } else {
            if (lane == 0)
                for (uint32_t i = 0; i < total && i < (uint32_t)SB; ++i) load_block(i, (int)i);
            RingPos rb(SB), rr(SB);  // block being split / stage being refilled (one block behind)
            for (uint32_t it = 0; it < total; ++it, rb.advance()) {
                const int sb = rb.s;
                mbar_wait(bar_full_b(bars, sb), rb.ph);
                // the block TMA dropped into the B ring is the B_hi operand as it is (the tensor core ignores the low
                // mantissa bits); B_lo = y - tf32(y) is derived here, element-wise in the same swizzled layout
                const float4* bh = reinterpret_cast<const float4*>(base_ptr + L.bhi + (uint32_t)sb * L.stage_b);
                float4* bl = reinterpret_cast<float4*>(base_ptr + L.blo + (uint32_t)sb * L.stage_b);
                constexpr int kPer = ATB ? (NP * 8 + 31) / 32 : 1;  // 16 B chunks per lane
                float4 h[kPer];
#pragma unroll
                for (int u = 0; u < kPer; ++u) {
                    const int e = lane + 32 * u;
                    if (e < NP * 8) h[u] = bh[e];
                }
#pragma unroll
                for (int u = 0; u < kPer; ++u) {
                    const int e = lane + 32 * u;
                    if (e < NP * 8) {
                        float4 l4;
                        l4.x = h[u].x - __uint_as_float(__float_as_uint(h[u].x) & 0xFFFFE000u);
                        l4.y = h[u].y - __uint_as_float(__float_as_uint(h[u].y) & 0xFFFFE000u);
                        l4.z = h[u].z - __uint_as_float(__float_as_uint(h[u].z) & 0xFFFFE000u);
                        l4.w = h[u].w - __uint_as_float(__float_as_uint(h[u].w) & 0xFFFFE000u);
                        bl[e] = l4;
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_full_b2(bars, sb));
                if (it >= 1) {
                    const uint32_t nxt = it - 1u + (uint32_t)SB;
                    if (nxt < total && lane == 0) {
                        mbar_wait_relaxed(bar_empty_b(bars, rr.s), rr.ph, 64);  // MMAs of block it-1 have read the stage
                        load_block(nxt, rr.s);
                    }
                    rr.advance();
                    __syncwarp();
                }
            }
            }
        } else if (warp == kTransformWarps) {
            // ================================ TMA producer: X ring ================================
            if (lane == 0) {
                uint32_t it = 0;
                Group g;
                RingPos rx(S);
                for (int64_t gi = 0; get_group<ATB>(p, gi, g); ++gi) {
                    for (int64_t kb = 0; kb < g.kblocks; ++kb, ++it, rx.advance()) {
                        const int s = rx.s;
                        const uint32_t ph = rx.ph;
                        mbar_wait_relaxed(bar_empty_x(bars, s), ph ^ 1u, 128);
                        trace_ev(p, 0, it);
                        const uint32_t full = bar_full(bars, s);
                        if (ATB) {
                            // 8 sub-tiles of 32 features x 32 rows (128 B rows, SWIZZLE_128B): the wide un-swizzled box
                            // {256 features, 32 rows} was served at only ~15 B/clk by the TMA unit
                            mbar_expect_tx(full, L.stage_x);
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                tma_load_2d(base + L.x + (uint32_t)s * L.stage_x + (uint32_t)j * 4096u, &p.map_x, g.f0 + 32 * j,
                                            (int)(g.row0 + kb * kKB), full);
                        } else {
                            const int k0 = (int)(kb * kKB);
                            mbar_expect_tx(full, L.stage_x + L.stage_mu);
                            tma_load_2d(base + L.x + (uint32_t)s * L.stage_x, &p.map_x, k0, (int)g.row0, full);
                            bulk_load_1d(base + L.mu + (uint32_t)s * L.stage_mu, p.mu_pad + k0, 128, full);
                        }
                    }
                }
            }
        } else if (warp == kTransformWarps + 3) {
            // ================================ TMA producer: B ring ================================
            if (lane == 0) {
                uint32_t it = 0;
                Group g;
                RingPos rb(SB);
                for (int64_t gi = 0; get_group<ATB>(p, gi, g); ++gi) {
                    for (int64_t kb = 0; kb < g.kblocks; ++kb, ++it, rb.advance()) {
                        const int sb = rb.s;
                        const uint32_t phb = rb.ph;
                        mbar_wait_relaxed(bar_empty_b(bars, sb), phb ^ 1u, 64);
                        const uint32_t fullb = bar_full_b(bars, sb);
                        if (ATB && panel) {
                            // panel-major Y: rows (r/32)*n_pad .. +n_pad of the [blocks*n_pad][32] views are exactly
                            // the K-major Y_hi / Y_lo operand tiles
                            const int r = (int)(g.row0 + kb * kKB);
                            mbar_expect_tx(fullb, p.b_split ? L.stage_b : 2u * L.stage_b);
                            tma_load_2d(base + L.bhi + (uint32_t)sb * L.stage_b, &p.map_bhi, 0, (r / 32) * n_pad, fullb);
                            if (!p.b_split)
                                tma_load_2d(base + L.blo + (uint32_t)sb * L.stage_b, &p.map_blo, 0, (r / 32) * n_pad, fullb);
                        } else if (ATB) {
                            // row-major Y: box {n_pad, 32 rows}, transposed + split by the transform warps
                            mbar_expect_tx(fullb, L.stage_b);
                            tma_load_2d(base + L.bhi + (uint32_t)sb * L.stage_b, &p.map_bhi, 0, (int)(g.row0 + kb * kKB), fullb);
                        } else {
                            const int k0 = (int)(kb * kKB);
                            mbar_expect_tx(fullb, 2u * L.stage_b);
                            tma_load_2d(base + L.bhi + (uint32_t)sb * L.stage_b, &p.map_bhi, k0, g.win * n_pad, fullb);
                            tma_load_2d(base + L.blo + (uint32_t)sb * L.stage_b, &p.map_blo, k0, g.win * n_pad, fullb);
                        }
                    }
                }
            }
        } else if (warp == kTransformWarps + 1 || warp == kTransformWarps + 2) {
            // ================================ MMA issuers ================================
            // One warp per M tile (two independent issue streams).  The whole warp runs the loop so that
            // every address / descriptor is warp-uniform; only the tcgen05 instructions are elected.
            const int mt = warp - (kTransformWarps + 1);
            const uint32_t idesc = make_idesc_tf32(n_pad, 0);
            const uint32_t idesc_x = make_idesc_bf16(n_pad);
            using MT_ = ModeTraits<ATB, MODE>;
            constexpr uint32_t aslots = (uint32_t)MT_::kSlots;
            constexpr int kAccBaseC = MT_::kAccBase;
            constexpr bool stagger = MT_::kStagger;
            constexpr bool two_sets = (MT_::kAccBufs == 2);
            constexpr bool cut2 = two_sets && MT_::kCut;  // tc_xb precise: chains of kChainKB K blocks
            constexpr int chain_kb = MT_::kChainKB;
            uint32_t it = 0;
            uint32_t chains = 0;   // accumulation chains started so far (same count as in the transform warps)
            int abuf = 0;
            uint32_t acc = 0;
            Group g;
            RingPos rb(SB), ry(kYBufs);
            for (int64_t gi = 0; get_group<ATB>(p, gi, g); ++gi) {
                const int nkb = (int)g.kblocks;
                for (int kb = 0; kb < nkb; ++kb, ++it, rb.advance(), ry.advance()) {
                    bool chain_start, chain_end;
                    if constexpr (stagger) {
                        // see transform_role: this M tile's chains, buffers rotate over the chains of both M tiles
                        chain_start = (kb == 0) || (((kb + 2 * mt) & 3) == 0);
                        chain_end = (kb == nkb - 1) || (((kb + 2 * mt) & 3) == 3);
                        if (chain_start) {
                            const uint32_t j = (mt == 0) ? (uint32_t)(2 * (kb >> 2) + 1) : (uint32_t)(2 * ((kb + 2) >> 2));
                            abuf = (int)(j % 3u);
                            acc = tmem_base + (uint32_t)(kAccBaseC + abuf * n_pad);
                            mbar_wait(bar_acc_empty(bars, abuf), ((j / 3u) & 1u) ^ 1u);
                            tc_fence_after();
                        }
                    } else {
                        chain_start = (kb == 0) || (cut2 && (kb & (chain_kb - 1)) == 0);
                        chain_end = (kb == nkb - 1) || (cut2 && (kb & (chain_kb - 1)) == chain_kb - 1);
                        if (chain_start) {
                            abuf = two_sets ? (int)(chains & 1u) : 0;
                            acc = tmem_base + (uint32_t)(kAccBaseC + abuf * kMT * n_pad + mt * n_pad);
                            mbar_wait(bar_acc_empty(bars, abuf), (two_sets ? ((chains >> 1) & 1u) : (chains & 1u)) ^ 1u);
                            tc_fence_after();
                            ++chains;
                        }
                    }
                    const int sb = rb.s;
                    const uint32_t phb = rb.ph;
                    // tc_xb reads its B tiles straight from the TMA ring; tc_atb's B tiles are TMA-loaded panels, or
                    // (row-major Y) produced by ALL transform warps - then both halves must have checked in first
                    if (!ATB || panel) mbar_wait((ATB && p.b_split) ? bar_full_b2(bars, sb) : bar_full_b(bars, sb), phb);
                    // (tc_xb also checks both halves in first: measured faster than issuing per half there)
                    constexpr bool kSplitIssue = ATB && panel;
                    if (!kSplitIssue) {
                        mbar_wait(bar_a_ready(bars, (int)((2u * it) % aslots)), ((2u * it) / aslots) & 1u);
                        mbar_wait(bar_a_ready(bars, (int)((2u * it + 1u) % aslots)), ((2u * it + 1u) / aslots) & 1u);
                    }
                    // B operand tiles, K-major [n_pad][32 fp32] SWIZZLE_128B
                    const int yb = ry.s;
                    const uint32_t ring = base + L.ylo + (uint32_t)yb * L.ylo_bytes;
                    const bool from_ring = ATB && !panel;
                    const uint32_t bhi_addr = from_ring ? ring : (base + L.bhi + (uint32_t)sb * L.stage_b);
                    const uint32_t blo_addr = from_ring ? (ring + (uint32_t)n_pad * 128u) : (base + L.blo + (uint32_t)sb * L.stage_b);
                    const uint64_t dhi0 = make_desc_sw128(bhi_addr, 16u, 1024u);
                    const uint64_t dlo0 = make_desc_sw128(blo_addr, 16u, 1024u);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t gran = 2u * it + (uint32_t)h;
                        const int ta = (int)(gran % aslots);
                        if (kSplitIssue) mbar_wait(bar_a_ready(bars, ta), (gran / aslots) & 1u);
                        tc_fence_after();
                        if (h == 0 && mt == 0 && lane == 0) trace_ev(p, 1, it);
                        const uint32_t a_hi0 = tmem_base + (uint32_t)(ta * kASlotCols + mt * 32);
                        if (elect_one()) {
#pragma unroll
                            for (int k2 = 0; k2 < ((p.dbg & 2) ? 0 : 2); ++k2) {
                                const int ks = 2 * h + k2;
                                // 32 B (= 2 descriptor address units) per K step inside the 128 B swizzle row
                                const uint64_t dhi = dhi0 + (uint64_t)(ks * 2);
                                const uint64_t dlo = dlo0 + (uint64_t)(ks * 2);
                                const uint32_t a_hi = a_hi0 + (uint32_t)k2 * 8u;
                                const uint32_t a_lo = a_hi + 16u;
                                if constexpr ((kCrossEnabled && !ATB) || (kCrossAtb && ATB && PANEL)) {
                                    // a_lo: this K step's 8 columns of bf16 [lo | hi]; dlo: the B-side cross tile [hi ; lo]
                                    mma_tf32_ts(acc, a_hi, dhi, idesc, (!chain_start || ks > 0) ? 1u : 0u);
                                    mma_f16_ts(acc, a_lo, dlo, idesc_x, 1u);
                                } else {
                                    mma_tf32_ts(acc, a_lo, dhi, idesc, (!chain_start || ks > 0) ? 1u : 0u);
                                    mma_tf32_ts(acc, a_hi, dlo, idesc, 1u);
                                    mma_tf32_ts(acc, a_hi, dhi, idesc, 1u);
                                }
                            }
                            tc_commit(bar_a_free(bars, ta));
                            if (h == 1) {
                                if (!ATB || panel) tc_commit(bar_empty_b(bars, sb));
                                if (ATB && !panel) tc_commit(bar_y_free(bars, yb));
                            }
                        }
                        __syncwarp();
                    }
                    __syncwarp();
                    if (mt == 0 && lane == 0) trace_ev(p, 2, it);
                    if (chain_end) {
                        if (elect_one()) tc_commit(bar_acc_full(bars, abuf));
                        __syncwarp();
                    }
                }
            }
        }
    } else {
        // ================================ transform + epilogue ================================
        transform_role<ATB, NP, PANEL, MODE>(p, base_ptr, L, bars, tmem_base, warp, lane, n_pad, S);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kTransformWarps + 3) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------
// operand preparation (tiny kernels)
// ------------------------------------------------------------------------------------------
// Bt_hi / Bt_lo [n_pad][Kp] from B (K x L row-major with ldb, or L x K if b_trans); TS = float or double
template <typename TS>
__global__ void prep_b_kernel(const TS* __restrict__ B, int64_t ldb, int b_trans, int64_t K, int64_t Kp, int L,
                              int n_pad, float* __restrict__ hi, float* __restrict__ lo) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n_pad * Kp) return;
    int64_t c = idx / Kp, k = idx % Kp;
    double b = 0.0;
    if (c < L && k < K) b = (double)(b_trans ? B[c * ldb + k] : B[k * ldb + c]);
    float bf = (float)b;
    float h = __uint_as_float(__float_as_uint(bf) & 0xFFFFE000u);
    hi[idx] = h;
    lo[idx] = (float)(b - (double)h);
}

// cross mode: hi[n_pad][Kp] as above; xw[n_pad][Kp] 32-bit words - per group of 8 k values: 4 words of bf16 pairs of
// hi, then 4 words of bf16 pairs of lo (the [hi ; lo] K order the A-side [lo | hi] operand is paired with)
template <typename TS>
__global__ void prep_b_cross_kernel(const TS* __restrict__ B, int64_t ldb, int b_trans, int64_t K, int64_t Kp, int L,
                                    int n_pad, float* __restrict__ hi, uint32_t* __restrict__ xw) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per pair of k values
    if (idx >= (int64_t)n_pad * (Kp / 2)) return;
    const int64_t c = idx / (Kp / 2), kp = idx % (Kp / 2), k0 = 2 * kp;
    float h[2], l[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int64_t k = k0 + e;
        double b = 0.0;
        if (c < L && k < K) b = (double)(b_trans ? B[c * ldb + k] : B[k * ldb + c]);
        const float bf = (float)b;
        h[e] = __uint_as_float(__float_as_uint(bf) & 0xFFFFE000u);
        l[e] = (float)(b - (double)h[e]);
        hi[c * Kp + k] = h[e];
    }
    const int64_t g = k0 >> 3, j = (k0 & 7) >> 1;
    xw[c * Kp + 8 * g + j] = pack_bf16x2(h[0], h[1]);
    xw[c * Kp + 8 * g + 4 + j] = pack_bf16x2(l[0], l[1]);
}

__global__ void prep_mu_kernel(const float* __restrict__ mu, int64_t K, int64_t Kp, float* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Kp) out[i] = (mu != nullptr && i < K) ? mu[i] : 0.f;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    PETAL_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) linalg_error("cuTensorMapEncodeTiled is unavailable");
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

// 2-D fp32 tensor map: dims {inner, outer}, row pitch in elements
inline CUtensorMap make_map_2d(const float* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_elems, uint32_t box_inner,
                               uint32_t box_outer, bool swizzle128) {
    CUtensorMap m;
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {pitch_elems * sizeof(float)};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE,
                             swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) linalg_error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return m;
}

inline int round_up(int64_t v, int m) { return (int)(((v + m - 1) / m) * m); }

// precise = -1: the context default (petal_ctx::tc_precise), PETAL_TC_PRECISE overrides everything (testing)
inline bool want_precise(petal_ctx* ctx, int precise, const char* kind_env = nullptr) {
    if (kind_env)
        if (const char* e = getenv(kind_env)) return atoi(e) != 0;
    if (const char* e = getenv("PETAL_TC_PRECISE")) return atoi(e) != 0;
    return precise < 0 ? ctx->tc_precise : (precise != 0);
}
inline bool xb_supported(const void* A, int64_t lda, int64_t n, int64_t K, int64_t L) {
    return n >= 512 && K >= 32 && L >= 1 && L <= 128 && (lda % 4 == 0) && is_aligned16(A) && n < ((int64_t)1 << 31) &&
           K < ((int64_t)1 << 31);
}

template <bool ATB, int NP, bool PANEL, int MODE>
inline void launch_kernel(petal_ctx* ctx, const TcParams& p, int grid, size_t smem) {
    ensure_dynamic_smem(ctx, tc_gemm_kernel<ATB, NP, PANEL, MODE>, smem);
    const char* trace_path = getenv("PETAL_TC_TRACE");
    if (trace_path == nullptr) {
        tc_gemm_kernel<ATB, NP, PANEL, MODE><<<grid, kThreads, smem, ctx->stream>>>(p);
        check_launch(ctx);
        return;
    }
    // debug: record a clock64 timeline of CTA 0 and append it to the file
    TcParams q = p;
    const size_t cnt = (size_t)kTraceEvents * kTraceKB;
    DBuf<long long> tr(ctx, cnt);
    tr.zero();
    q.trace = tr.p;
    tc_gemm_kernel<ATB, NP, PANEL, MODE><<<grid, kThreads, smem, ctx->stream>>>(q);
    check_launch(ctx);
    std::vector<long long> h(cnt);
    PETAL_CUDA(cudaMemcpyAsync(h.data(), tr.p, cnt * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    PETAL_CUDA(cudaStreamSynchronize(ctx->stream));
    FILE* f = fopen(trace_path, "a");
    if (f) {
        fprintf(f, "kernel atb=%d np=%d n=%lld K=%lld n_pad=%d grid=%d\n", (int)ATB, NP, (long long)p.n, (long long)p.K, p.n_pad, grid);
        for (int e = 0; e < kTraceEvents; ++e) {
            for (int i = 0; i < kTraceKB; ++i) fprintf(f, "%lld ", h[(size_t)e * kTraceKB + i]);
            fprintf(f, "\n");
        }
        fclose(f);
    }
}

// Y[n x ldy] = (A - mu) * B.  B: TS in {float, double}, K x L row-major (ldb) or L x K when b_trans.
// Columns [L, min(ldy, n_pad)) of Y are written as zeros (padding for later TMA reads).
// y_panel: Y is written panel-major, [ceil(n/32)][n_pad][32] floats (n_pad = L rounded up to 16), rows >= n zero.
template <typename TS>
void launch_tc_xb(petal_ctx* ctx, const float* A, int64_t lda, int64_t n, int64_t K, const TS* B, int64_t ldb,
                  bool b_trans, int64_t L, const float* mu, float* Y, int64_t ldy, double* sumsq,
                  bool y_panel = false, float* Y_lo_panel = nullptr, int precise = -1, const float* bias = nullptr,
                  int64_t y_cols = -1, int ones_col = -1) {
    // L > 128 (row-major Y only): windows of 128 output columns walked inside the kernel (TcParams::nwin)
    const int nwin = (L > 128) ? (int)ceil_div(L, 128) : 1;
    if (nwin > 1 && (y_panel || sumsq != nullptr)) linalg_error("tc_xb: wide outputs are row-major, without a sum of squares");
    const int n_pad = nwin > 1 ? 128 : round_up(L, 16);
    const int b_rows = nwin * n_pad;  // rows of the B^T operand arrays
    int stages = 0, stages_b = 0;
    if (!pick_stages(false, n_pad, false, stages, stages_b)) linalg_error("tc_xb: no pipeline configuration fits in shared memory");
    const int64_t Kp = round_up(K, 32);
    DBuf<float> bhi(ctx, (size_t)(b_rows * Kp)), blo(ctx, (size_t)(b_rows * Kp)), mup(ctx, (size_t)Kp);
    if (kCrossEnabled)
        prep_b_cross_kernel<TS><<<(unsigned)ceil_div((int64_t)b_rows * (Kp / 2), 256), 256, 0, ctx->stream>>>(
            B, ldb, b_trans ? 1 : 0, K, Kp, (int)L, b_rows, bhi.p, reinterpret_cast<uint32_t*>(blo.p));
    else
        prep_b_kernel<TS><<<(unsigned)ceil_div((int64_t)b_rows * Kp, 256), 256, 0, ctx->stream>>>(B, ldb, b_trans ? 1 : 0, K, Kp,
                                                                                                 (int)L, b_rows, bhi.p, blo.p);
    check_launch(ctx);
    prep_mu_kernel<<<(unsigned)ceil_div(Kp, 256), 256, 0, ctx->stream>>>(mu, K, Kp, mup.p);
    check_launch(ctx);
    TcParams p;
    std::memset(&p, 0, sizeof p);
    p.map_x = make_map_2d(A, (uint64_t)K, (uint64_t)n, (uint64_t)lda, 32, 256, true);
    p.map_bhi = make_map_2d(bhi.p, (uint64_t)Kp, (uint64_t)b_rows, (uint64_t)Kp, 32, (uint32_t)n_pad, true);
    p.map_blo = make_map_2d(blo.p, (uint64_t)Kp, (uint64_t)b_rows, (uint64_t)Kp, 32, (uint32_t)n_pad, true);
    p.mu_pad = mup.p;
    p.n = n;
    p.K = K;
    p.n_pad = n_pad;
    p.nwin = nwin;
    p.L = (int)L;
    p.stages = stages;
    p.stages_b = stages_b;
    p.Y = Y;
    p.ldy = ldy;
    p.y_cols = (y_cols < 0) ? ldy : y_cols;
    p.bias = bias;
    p.L = (int)L;
    p.y_vec = (is_aligned16(Y) && (ldy % 4 == 0) && (p.y_cols % 2 == 0)) ? 1 : 0;
    p.y_panel = y_panel ? 1 : 0;
    p.ones_col_p1 = (y_panel && ones_col >= 0 && ones_col < n_pad) ? ones_col + 1 : 0;
    p.Ylo = Y_lo_panel;
    p.sumsq = sumsq;
    // mode (see ModeTraits): two accumulator sets need 192 + 4 n_pad <= 512
    const bool two_sets_fit = (n_pad <= 80);
    int mode = 0;
    if (two_sets_fit && want_precise(ctx, precise, "PETAL_XB_PRECISE")) mode = 2;
    if (mode == 0 && n_pad <= 96) {  // 5 * 64 + 2 * n_pad <= 512
        const char* e = getenv("PETAL_XB_SLOTS5");
        if (e && atoi(e) != 0) mode = 3;  // opt-in: measured 5 % slower than the 4-slot ring on c2 (r02), kept for experiments
    }
    const SmemLayout lay = make_layout(false, n_pad, stages, stages_b);
    const int64_t items = ceil_div(n, 256);
    const int grid = (int)std::min<int64_t>(items, ctx->sm_count);
    KTimer kt(ctx, nwin > 1 ? "tc_xb_f32_wide" : (K >= 256 ? "tc_xb_f32" : "tc_xb_f32_skinny"), (double)n * (K + L) * sizeof(float));
    switch (mode * 2 + (y_panel ? 1 : 0)) {
        case 6: launch_kernel<false, 0, false, 3>(ctx, p, grid, lay.total); break;
        case 7: launch_kernel<false, 0, true, 3>(ctx, p, grid, lay.total); break;
        case 0: launch_kernel<false, 0, false, 0>(ctx, p, grid, lay.total); break;
        case 1: launch_kernel<false, 0, true, 0>(ctx, p, grid, lay.total); break;
        case 4: launch_kernel<false, 0, false, 2>(ctx, p, grid, lay.total); break;
        default: launch_kernel<false, 0, true, 2>(ctx, p, grid, lay.total); break;
    }
}

inline bool atb_supported(const void* A, int64_t lda, int64_t da, const void* B, int64_t ldb, int64_t db, int64_t n) {
    return n >= 1024 && da >= 32 && db >= 1 && db <= 128 && (lda % 4 == 0) && (ldb % 4 == 0) && is_aligned16(A) &&
           is_aligned16(B) && n < ((int64_t)1 << 31) && da < ((int64_t)1 << 31);
}

// Z[da x ldz] (f64, accumulated; caller zeroes) += (A - mua)^T * B, B is n x db (ldb), not centred.
// b_panel: B is panel-major [ceil(n/32)][n_pad][32] (as written by launch_tc_xb with y_panel).
inline void launch_tc_atb(petal_ctx* ctx, const float* A, int64_t lda, int64_t da, const float* mua, const float* B,
                          int64_t ldb, int64_t db, int64_t n, double* Z, int64_t ldz, bool b_panel = false,
                          const float* B_lo_panel = nullptr, const float* mub = nullptr, int precise = -1) {
    const int n_pad = round_up(db, 16);
    int stages = 0, stages_b = 0;
    if (!pick_stages(true, n_pad, b_panel, stages, stages_b)) linalg_error("tc_atb: no pipeline configuration fits in shared memory");
    const int fgroups = (int)ceil_div(da, 256);
    const int64_t Fp = (int64_t)fgroups * 256;
    DBuf<float> mup(ctx, (size_t)Fp);
    prep_mu_kernel<<<(unsigned)ceil_div(Fp, 256), 256, 0, ctx->stream>>>(mua, da, Fp, mup.p);
    check_launch(ctx);
    TcParams p;
    std::memset(&p, 0, sizeof p);
    p.map_x = make_map_2d(A, (uint64_t)da, (uint64_t)n, (uint64_t)lda, 32, 32, true);
    if (b_panel) {
        p.map_bhi = make_map_2d(B, 32, (uint64_t)ceil_div(n, 32) * (uint64_t)n_pad, 32, 32, (uint32_t)n_pad, true);
        p.b_split = (B_lo_panel == nullptr || kCrossAtb) ? 1 : 0;  // no Y_lo panel: the operand tile is derived inside the kernel
        p.map_blo = p.b_split ? p.map_bhi
                              : make_map_2d(B_lo_panel, 32, (uint64_t)ceil_div(n, 32) * (uint64_t)n_pad, 32, 32, (uint32_t)n_pad, true);
    } else {
        p.map_bhi = make_map_2d(B, (uint64_t)db, (uint64_t)n, (uint64_t)ldb, (uint32_t)n_pad, 32, false);
        p.map_blo = p.map_bhi;
    }
    DBuf<float> mubp;
    if (mub != nullptr) {
        if (b_panel) linalg_error("tc_atb: panel-major Y cannot be centred on load");
        mubp.alloc(ctx, (size_t)n_pad);
        prep_mu_kernel<<<1, 256, 0, ctx->stream>>>(mub, db, n_pad, mubp.p);
        check_launch(ctx);
        p.mub_pad = mubp.p;
    }
    p.y_panel = b_panel ? 1 : 0;
    p.mu_pad = mup.p;
    p.n = n;
    p.K = da;
    p.n_pad = n_pad;
    p.L = (int)db;
    p.stages = stages;
    p.stages_b = stages_b;
    // fast: one accumulator set, chain = 1024 rows.  precise (n_pad <= 80): chains of 4 K blocks (48 accumulating MMAs),
    // staggered between the two M tiles over three one-tile accumulator buffers, so the 4-slot operand ring stays
    // (the truncating accumulation of the tensor core biases long chains, see transform_role)
    const bool cut = (n_pad <= 80) && want_precise(ctx, precise, "PETAL_ATB_PRECISE");  // 256 + 3 n_pad <= 512
    p.Z = Z;
    p.ldz = ldz;
    // One CTA = one feature group x one contiguous slice of rows; the TMEM accumulation chain is cut
    // every 1024 rows (see the kernel comment).
    const int64_t max_slices = std::max<int64_t>(1, ctx->sm_count / fgroups);
    const int64_t slices = std::max<int64_t>(1, std::min<int64_t>(max_slices, ceil_div(n, 1024)));
    p.slice_rows = ceil_div(ceil_div(n, slices), 32) * 32;
    p.chunk_rows = cut ? p.slice_rows : 1024;  // precise: chains are cut inside the kernel, one group per CTA
    p.fgroups = fgroups;
    p.dbg = getenv("PETAL_TC_DBG") ? atoi(getenv("PETAL_TC_DBG")) : 0;
    const SmemLayout lay = make_layout(true, n_pad, stages, stages_b, b_panel);
    const int grid = (int)(ceil_div(n, p.slice_rows) * fgroups);
    KTimer kt(ctx, da >= 256 ? "tc_atb_f32" : "tc_atb_f32_skinny", (double)n * (da + db) * sizeof(float));
#define PETAL_ATB_CASE(NPV)                                                                  \
    case NPV:                                                                                \
        if (cut) {                                                                           \
            if (b_panel) launch_kernel<true, (NPV <= 80 ? NPV : 80), true, 2>(ctx, p, grid, lay.total);   \
            else launch_kernel<true, (NPV <= 80 ? NPV : 80), false, 2>(ctx, p, grid, lay.total);          \
        } else {                                                                             \
            if (b_panel) launch_kernel<true, NPV, true, 0>(ctx, p, grid, lay.total);     \
            else launch_kernel<true, NPV, false, 0>(ctx, p, grid, lay.total);            \
        }                                                                                    \
        break;
    switch (n_pad) {
        PETAL_ATB_CASE(16)
        PETAL_ATB_CASE(32)
        PETAL_ATB_CASE(48)
        PETAL_ATB_CASE(64)
        PETAL_ATB_CASE(80)
        PETAL_ATB_CASE(96)
        PETAL_ATB_CASE(112)
        default:
            if (b_panel) launch_kernel<true, 128, true, 0>(ctx, p, grid, lay.total);
            else launch_kernel<true, 128, false, 0>(ctx, p, grid, lay.total);
            break;
    }
#undef PETAL_ATB_CASE
}

}  // namespace tc
}  // namespace petal
