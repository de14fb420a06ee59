/*
 * petal_b200.h - C ABI of libpetal_b200.so
 *
 * B200-native (sm_100a) replacement for the fit/transform hot path of the Rust crate
 * petal-decomposition 0.9.0.  The reference exposes no FFI of its own; its boundary is the
 * public Rust API (src/lib.rs:17-18).  Each entry point below replaces the *private* hot-path
 * function(s) cited next to it, so that a thin Rust (or C++ / Python) host can keep the public
 * types `Pca`, `RandomizedPca`, `FastIca` verbatim and forward the arithmetic here
 * (binding stubs: INTEGRATION.md; host mirrors: include/petal_decomposition.hpp,
 * petal_decomposition_b200/api.py).
 *
 * Conventions
 *  - Matrices are row-major, C-contiguous, samples x features - the layout the reference
 *    asserts (src/linalg.rs:75,106,132).  Dimensions are int64_t (lifts the i32 limit of
 *    src/linalg.rs:76-79 and the i32 `m*m` of src/linalg/lapack.rs:111).
 *  - Every data pointer may be a HOST pointer or a DEVICE pointer of the context's GPU
 *    (detected with cudaPointerGetAttributes).  Host inputs are staged to HBM inside the
 *    call; host outputs are copied back before the call returns.  The caller owns all
 *    buffers; the library owns only context-internal workspaces.
 *  - Generic `A` of the reference maps to the suffix: _f32 / _f64.
 *  - Return value: PETAL_OK, PETAL_INVALID_INPUT (DecompositionError::InvalidInput,
 *    src/lib.rs:24-25), PETAL_LINALG_ERROR (DecompositionError::LinalgError,
 *    src/lib.rs:26-27; also CUDA / NCCL failures).  The message (same wording as the
 *    reference's) is available from petal_last_error().
 *  - Multi-GPU: one process per GPU.  After petal_comm_init() every fit call is
 *    COLLECTIVE: each rank passes its own row shard (n = local rows); model outputs
 *    (components, mean, singular values, ...) are identical on every rank, per-sample
 *    outputs (scores / sources) are the local shard's rows.
 *  - There is no CPU fallback: without a usable sm_100 device petal_ctx_create fails.
 */
#ifndef PETAL_B200_H
#define PETAL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PETAL_OK 0
#define PETAL_INVALID_INPUT 1
#define PETAL_LINALG_ERROR 2

typedef struct petal_ctx petal_ctx;

/* ---- context ------------------------------------------------------------------------- */
/* Creates a context bound to CUDA device `device` (ordinal). Owns one stream + workspaces. */
int petal_ctx_create(int device, petal_ctx** out);
void petal_ctx_destroy(petal_ctx* ctx);
/* Last error message of this context (never NULL). */
const char* petal_last_error(const petal_ctx* ctx);
/* Library-level message for failures that happen before a context exists. */
const char* petal_last_global_error(void);
/* Run on a caller-owned cudaStream_t (e.g. torch's current stream) instead of the internal one. */
int petal_ctx_set_stream(petal_ctx* ctx, void* cuda_stream);
/* Workspaces are kept in a stream-ordered pool between calls; this returns them to the driver. */
int petal_ctx_trim(petal_ctx* ctx);
/* Blocks until all work queued by this context has finished. */
int petal_ctx_synchronize(petal_ctx* ctx);
/* Number of kernels this context has launched since creation (bench.py's `gpu_launches`). */
int64_t petal_ctx_launch_count(const petal_ctx* ctx);
/* Per-kernel timing with CUDA events on the launch stream (bench.py's roofline numbers).
 * petal_ctx_profile_json synchronizes, writes {"kernel": {"count","total_ms","min_ms","max_ms","work"}}
 * (work = algorithmic bytes the launcher accounted for) into buf, clears the log and returns the
 * length needed (including the terminating NUL). */
int petal_ctx_set_profiling(petal_ctx* ctx, int enable);
int64_t petal_ctx_profile_json(petal_ctx* ctx, char* buf, int64_t cap);
/* Select the f32 streaming-GEMM engine: 0 = FFMA SIMT kernels, 1 = tcgen05 3xTF32 (default when
 * the shape is supported). Returns the engine now in force. Negative `engine` only queries. */
int petal_ctx_set_f32_engine(petal_ctx* ctx, int engine);
/* Same for f64 Gram-shaped contractions: 0 = SIMT DFMA kernels, 1 = DMMA (mma.sync.m8n8k4.f64, default). */
int petal_ctx_set_f64_engine(petal_ctx* ctx, int engine);

/* Host-resident X (out-of-core / streaming ingest).  A host `x` passed to a fit / transform entry is
 * copied to HBM in row chunks of about `chunk_bytes` (default 1 GiB; <= 0 keeps the current value) on a second stream,
 * double-buffered, and each chunk is consumed as soon as it has landed (pinned host memory makes the copies
 * asynchronous; pageable memory works, without overlap).
 *   mode 0 (default): keep a resident copy when X fits into the free HBM next to the call's workspaces - the first
 *          streaming pass of the fit then runs underneath the transfer - otherwise
 *   mode 2: out-of-core: X is re-streamed through a two-slot ring by every traversal; the flows pair the two
 *          contractions of a range-finder iteration on the resident chunk (q + 1 trips over PCIe for randomized PCA,
 *          3 for exact PCA with scores, 2 + iterations for FastICA).  mode 1 forces the resident copy.
 * A host `out` of inverse_transform larger than a chunk is produced and drained chunk by chunk in the same way.
 * Returns the mode now in force (negative `mode` only queries). */
int petal_ctx_set_host_staging(petal_ctx* ctx, int mode, int64_t chunk_bytes);
/* Randomized PCA on a host-fed X (every rank's shard a host buffer, d <= 2048, at least one power iteration): the ingest traversal also
 * accumulates the Gram matrix G = Xc^T Xc while the GPU would otherwise wait for PCIe, the power iterations
 * Z <- Xc^T (Xc B) = G B (src/pca.rs:708-715) run on the small side, and only the last pair of products (the ones
 * that define the result) is taken from X: 2 traversals instead of q + 1.  1 = on (default), 0 = off (the plain pass
 * sequence, as for X in HBM).  Returns the setting now in force (negative `enable` only queries). */
int petal_ctx_set_host_gram(petal_ctx* ctx, int enable);
/* What the last call that was given a host X did: bytes copied host->device, traversals of X, ring (1) or resident (0). */
int petal_ctx_host_stream_stats(const petal_ctx* ctx, int64_t* h2d_bytes, int64_t* traversals, int* ring);

/* ---- multi-GPU (row sharding; replaces nothing in the reference, which is single-host) - */
#define PETAL_COMM_ID_BYTES 128
/* Rank 0 creates an NCCL unique id and ships it to the other ranks (torch.distributed / MPI). */
int petal_comm_unique_id(void* out_id /* PETAL_COMM_ID_BYTES */);
int petal_comm_init(petal_ctx* ctx, const void* id, int rank, int world_size);

/* ---- host RNG: the reference's seeded stream ------------------------------------------- */
/* rand_pcg::Mcg128Xsl64 + rand_distr::StandardNormal, as used by src/pca.rs:356-358,701-705
 * and src/ica.rs:75-77,210-214 (restated; see oracle/rng.py for provenance).               */
typedef struct petal_rng petal_rng;
/* `Pcg::from_seed(seed.to_be_bytes())` - seed given as two 64-bit halves of the u128. */
petal_rng* petal_rng_from_seed(uint64_t seed_hi, uint64_t seed_lo);
/* `Pcg64Mcg::new(state)` (src/pca.rs:991). */
petal_rng* petal_rng_from_state(uint64_t state_hi, uint64_t state_lo);
void petal_rng_free(petal_rng* rng);
uint64_t petal_rng_next_u64(petal_rng* rng);
void petal_rng_get_state(const petal_rng* rng, uint64_t* state_hi, uint64_t* state_lo);
/* rows*cols draws in row-major order, one f64 StandardNormal per element, cast to the
 * output type (`A::Real::from_f64`), HOST output. */
void petal_rng_normal_f64(petal_rng* rng, double* out, int64_t count);
void petal_rng_normal_f32(petal_rng* rng, float* out, int64_t count);

/* ---- exact PCA -------------------------------------------------------------------------
 * Replaces Pca::inner_fit (src/pca.rs:195-231) = mean_axis + centred copy + linalg::svd
 * (src/linalg.rs:70-91 -> gesvd, src/linalg/lapack.rs:103-132) + svd_flip (src/pca.rs:815-850),
 * and transform_with_u (src/pca.rs:758-779) when `scores` != NULL.
 *   x[n*d] in;  components[k*d], mean[d] (zeros when !centering), singular[k],
 *   total_variance[1] (= sum of ALL squared singular values, src/pca.rs:224),
 *   scores[n*k] or NULL (= U[:, :k] * sigma, the fit_transform output).
 * n == 0 with centering: returns PETAL_OK and leaves every output untouched (src/pca.rs:207-211).
 * Any dimension < k -> PETAL_INVALID_INPUT "every dimension should be at least {k}" (:199-204). */
int petal_pca_fit_f32(petal_ctx* ctx, const float* x, int64_t n, int64_t d, int64_t k, int centering,
                      float* components, float* mean, float* singular, float* total_variance,
                      float* scores);
int petal_pca_fit_f64(petal_ctx* ctx, const double* x, int64_t n, int64_t d, int64_t k, int centering,
                      double* components, double* mean, double* singular, double* total_variance,
                      double* scores);

/* ---- randomized PCA ---------------------------------------------------------------------
 * Replaces RandomizedPca::inner_fit (src/pca.rs:509-550), randomized_svd (:668-686) and
 * randomized_range_finder (:689-718).  `omega` is the d x (k + n_oversamples) Gaussian test
 * matrix, row-major, drawn by the caller from its RNG exactly as src/pca.rs:701-705 does
 * (petal_rng_normal_*), so the model's RNG state advances like the reference's.
 * Reference constants: n_oversamples = 10 (src/pca.rs:679), n_power_iter = 7 (:680).
 * total_variance = ||Xc||_F^2 (src/pca.rs:533).  Outputs as petal_pca_fit_*. */
int petal_rpca_fit_f32(petal_ctx* ctx, const float* x, int64_t n, int64_t d, int64_t k, int centering,
                       int64_t n_oversamples, int64_t n_power_iter, const float* omega,
                       float* components, float* mean, float* singular, float* total_variance,
                       float* scores);
int petal_rpca_fit_f64(petal_ctx* ctx, const double* x, int64_t n, int64_t d, int64_t k, int centering,
                       int64_t n_oversamples, int64_t n_power_iter, const double* omega,
                       double* components, double* mean, double* singular, double* total_variance,
                       double* scores);

/* ---- transform / inverse_transform -------------------------------------------------------
 * transform (src/pca.rs:726-750; src/ica.rs:120-131): out[n*k] = (x - mean) * components^T.
 * mean == NULL <=> centering(false).  The column-count checks of the reference
 * ("# of columns should be {}", "too many columns") are done by the host mirror, which owns
 * the model shapes; here d is the common column count by construction. */
int petal_transform_f32(petal_ctx* ctx, const float* x, int64_t n, int64_t d, const float* components,
                        int64_t k, const float* mean, float* out);
int petal_transform_f64(petal_ctx* ctx, const double* x, int64_t n, int64_t d, const double* components,
                        int64_t k, const double* mean, double* out);
/* inverse_transform (src/pca.rs:788-811): out[n*d] = y * components + mean. */
int petal_inverse_transform_f32(petal_ctx* ctx, const float* y, int64_t n, int64_t k,
                                const float* components, int64_t d, const float* mean, float* out);
int petal_inverse_transform_f64(petal_ctx* ctx, const double* y, int64_t n, int64_t k,
                                const double* components, int64_t d, const double* mean, double* out);

/* ---- FastICA ---------------------------------------------------------------------------
 * Replaces FastIca::inner_fit (src/ica.rs:167-222), ica_par (:319-361),
 * symmetric_decorrelation (:363-381), logcosh (:383-398) and the fit_transform product (:155-156).
 *   nc = min(n_total, d) components (src/ica.rs:173); w_init[nc*nc] row-major from the caller's
 *   RNG (src/ica.rs:210-214).  Reference constants: tol = 1e-4, max_iter = 200 (src/ica.rs:216).
 *   fun: PETAL_ICA_LOGCOSH is the only nonlinearity the reference has (src/ica.rs:383-398);
 *        EXP / CUBE are extensions named by the north star.
 *   lim_variant: 0 = row.row convergence test (textbook / sklearn), 1 = the reference's literal
 *        row.column test (src/ica.rs:345-349).
 * Outputs: components[nc*d], mean[d], n_iter[1], final_lim[1] (may be NULL),
 *          sources[n*nc] or NULL (= fit_transform output). */
#define PETAL_ICA_LOGCOSH 0
#define PETAL_ICA_EXP 1
#define PETAL_ICA_CUBE 2
int petal_fastica_fit_f32(petal_ctx* ctx, const float* x, int64_t n, int64_t d, int fun, double tol,
                          int64_t max_iter, int lim_variant, const float* w_init, float* components,
                          float* mean, int64_t* n_iter, double* final_lim, float* sources);
int petal_fastica_fit_f64(petal_ctx* ctx, const double* x, int64_t n, int64_t d, int fun, double tol,
                          int64_t max_iter, int lim_variant, const double* w_init, double* components,
                          double* mean, int64_t* n_iter, double* final_lim, double* sources);

/* Deflation scheme (not in the reference, which only has the symmetric scheme above): the
 * components are extracted one at a time by the one-unit fixed-point iteration with Gram-Schmidt against the finished
 * ones - sklearn `_ica_def` (_fastica.py:65-100), restated in oracle/ica.py.  Same whitening (src/ica.rs:189-208),
 * contrast functions and outputs as petal_fastica_fit_*; n_iter = the largest iteration count over the components,
 * final_lim = the largest final | |<w+, w>| - 1 |.  One pass over X per iteration (d * sizeof(T) bytes per sample). */
int petal_fastica_deflation_fit_f32(petal_ctx* ctx, const float* x, int64_t n, int64_t d, int fun, double tol,
                                    int64_t max_iter, const float* w_init, float* components, float* mean,
                                    int64_t* n_iter, double* final_lim, float* sources);
int petal_fastica_deflation_fit_f64(petal_ctx* ctx, const double* x, int64_t n, int64_t d, int fun, double tol,
                                    int64_t max_iter, const double* w_init, double* components, double* mean,
                                    int64_t* n_iter, double* final_lim, double* sources);

/* ---- building blocks exposed for unit tests (tests/ call these through the same ABI) -----
 * ica_par (src/ica.rs:319-361) on an already-whitened nc x n matrix given as its transpose
 * x1t[n*nc] (samples x components, row-major); w_init / w_out[nc*nc] are f64 for both data types (the small
 * side is f64 on the device).  The f32 entry runs the one-pass tcgen05 kernel when the shape allows. */
int petal_ica_par_f32(petal_ctx* ctx, const float* x1t, int64_t n, int64_t nc, int fun, double tol,
                      int64_t max_iter, int lim_variant, const double* w_init, double* w_out,
                      int64_t* n_iter, double* final_lim);
int petal_ica_par_f64(petal_ctx* ctx, const double* x1t, int64_t n, int64_t nc, int fun, double tol,
                      int64_t max_iter, int lim_variant, const double* w_init, double* w_out,
                      int64_t* n_iter, double* final_lim);
/* The deflation scheme on an already-whitened matrix (same conventions as petal_ica_par_*): replays sklearn's
 * `_ica_def` golden vectors (tests/golden/ica_deflation.json). */
int petal_ica_defl_f32(petal_ctx* ctx, const float* x1t, int64_t n, int64_t nc, int fun, double tol, int64_t max_iter,
                       const double* w_init, double* w_out, int64_t* n_iter, double* final_lim);
int petal_ica_defl_f64(petal_ctx* ctx, const double* x1t, int64_t n, int64_t nc, int fun, double tol, int64_t max_iter,
                       const double* w_init, double* w_out, int64_t* n_iter, double* final_lim);
/* logcosh (src/ica.rs:383-398) and the exp / cube extensions, elementwise: u[n*nc] (samples x components) is
 * replaced by g(u), gprime_sum[nc] = sum over the n samples of g'(u) (the reference divides by n).
 * engine 0: the kernel of the generic path (libm-accurate tanh / exp); engine 1 (f32 only): the device function
 * the one-pass tcgen05 kernel applies in its epilogue (ex2.approx / rcp.approx tanh). */
int petal_ica_nonlin_f32(petal_ctx* ctx, float* u, int64_t n, int64_t nc, int fun, int engine, double* gprime_sum);
int petal_ica_nonlin_f64(petal_ctx* ctx, double* u, int64_t n, int64_t nc, int fun, int engine, double* gprime_sum);
/* symmetric_decorrelation (src/ica.rs:363-381), textbook (W W^T)^-1/2 W; w[m*m] -> out[m*m]. */
int petal_symmetric_decorrelation_f64(petal_ctx* ctx, const double* w, int64_t m, double* out);
/* One-sided Jacobi SVD of a[m*len] (row-major, m <= len or not): a = U diag(s) Vt with
 * u[m*m], s[m] descending, vt[m*len] (rows with s == 0 are zero). Replaces the LAPACK calls of
 * src/linalg/lapack.rs (gesvd/gesdd/syev) on the small replicated matrices. */
int petal_small_svd_f64(petal_ctx* ctx, const double* a, int64_t m, int64_t len, double* u, double* s,
                        double* vt);
/* Column means and centred Gram matrix (the streaming passes of exact PCA / whitening):
 * mean[d], gram[d*d] = (x - mean)^T (x - mean) in f64. */
int petal_colmean_gram_f32(petal_ctx* ctx, const float* x, int64_t n, int64_t d, int centering,
                           double* mean, double* gram);
int petal_colmean_gram_f64(petal_ctx* ctx, const double* x, int64_t n, int64_t d, int centering,
                           double* mean, double* gram);

/* out[d*l] (f64) = (x - mean)^T y  with x[n*d], y[n*l] (mean may be NULL): the X^T*Q / Q^T*X pass of
 * the range finder (src/pca.rs:681,711), all-reduced over ranks. */
int petal_xty_f32(petal_ctx* ctx, const float* x, int64_t n, int64_t d, const float* mean, const float* y,
                  int64_t l, double* out);
int petal_xty_f64(petal_ctx* ctx, const double* x, int64_t n, int64_t d, const double* mean, const double* y,
                  int64_t l, double* out);

/* ---- measurement utility ------------------------------------------------------------------
 * FP64 tensor-pipe probe: back-to-back mma.sync.m8n8k4.f64 from registers on every SM (no memory traffic), timed
 * with CUDA events.  out[0] = TFLOP/s with `ctas_per_sm` CTAs of 8 warps per SM; the denominator of the
 * "fraction of FP64 tensor peak" figures bench.py reports for exact PCA (profiles/r02_fp64_dmma_peak.json). */
int petal_probe_dmma_tflops(petal_ctx* ctx, int ctas_per_sm, double* out);

#ifdef __cplusplus
}
#endif
#endif /* PETAL_B200_H */
