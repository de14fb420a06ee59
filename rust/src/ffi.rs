//! Raw bindings of include/petal_b200.h (the subset the shim uses).
use std::os::raw::{c_char, c_double, c_int, c_void};

#[repr(C)]
pub struct PetalCtx {
    _private: [u8; 0],
}

pub const PETAL_OK: c_int = 0;
pub const PETAL_INVALID_INPUT: c_int = 1;

extern "C" {
    pub fn petal_ctx_create(device: c_int, out: *mut *mut PetalCtx) -> c_int;
    pub fn petal_ctx_destroy(ctx: *mut PetalCtx);
    pub fn petal_last_error(ctx: *const PetalCtx) -> *const c_char;
    pub fn petal_last_global_error() -> *const c_char;

    pub fn petal_pca_fit_f32(ctx: *mut PetalCtx, x: *const f32, n: i64, d: i64, k: i64, centering: c_int,
        components: *mut f32, mean: *mut f32, singular: *mut f32, total_variance: *mut f32, scores: *mut f32) -> c_int;
    pub fn petal_pca_fit_f64(ctx: *mut PetalCtx, x: *const f64, n: i64, d: i64, k: i64, centering: c_int,
        components: *mut f64, mean: *mut f64, singular: *mut f64, total_variance: *mut f64, scores: *mut f64) -> c_int;
    pub fn petal_rpca_fit_f32(ctx: *mut PetalCtx, x: *const f32, n: i64, d: i64, k: i64, centering: c_int,
        n_oversamples: i64, n_power_iter: i64, omega: *const f32, components: *mut f32, mean: *mut f32,
        singular: *mut f32, total_variance: *mut f32, scores: *mut f32) -> c_int;
    pub fn petal_rpca_fit_f64(ctx: *mut PetalCtx, x: *const f64, n: i64, d: i64, k: i64, centering: c_int,
        n_oversamples: i64, n_power_iter: i64, omega: *const f64, components: *mut f64, mean: *mut f64,
        singular: *mut f64, total_variance: *mut f64, scores: *mut f64) -> c_int;
    pub fn petal_transform_f32(ctx: *mut PetalCtx, x: *const f32, n: i64, d: i64, components: *const f32, k: i64,
        mean: *const f32, out: *mut f32) -> c_int;
    pub fn petal_transform_f64(ctx: *mut PetalCtx, x: *const f64, n: i64, d: i64, components: *const f64, k: i64,
        mean: *const f64, out: *mut f64) -> c_int;
    pub fn petal_inverse_transform_f32(ctx: *mut PetalCtx, y: *const f32, n: i64, k: i64, components: *const f32,
        d: i64, mean: *const f32, out: *mut f32) -> c_int;
    pub fn petal_inverse_transform_f64(ctx: *mut PetalCtx, y: *const f64, n: i64, k: i64, components: *const f64,
        d: i64, mean: *const f64, out: *mut f64) -> c_int;
    pub fn petal_fastica_fit_f32(ctx: *mut PetalCtx, x: *const f32, n: i64, d: i64, fun: c_int, tol: c_double,
        max_iter: i64, lim_variant: c_int, w_init: *const f32, components: *mut f32, mean: *mut f32,
        n_iter: *mut i64, final_lim: *mut c_double, sources: *mut f32) -> c_int;
    pub fn petal_fastica_fit_f64(ctx: *mut PetalCtx, x: *const f64, n: i64, d: i64, fun: c_int, tol: c_double,
        max_iter: i64, lim_variant: c_int, w_init: *const f64, components: *mut f64, mean: *mut f64,
        n_iter: *mut i64, final_lim: *mut c_double, sources: *mut f64) -> c_int;
    // extensions: deflation scheme; how a host `x` (the caller's ArrayBase) reaches HBM
    pub fn petal_fastica_deflation_fit_f32(ctx: *mut PetalCtx, x: *const f32, n: i64, d: i64, fun: c_int, tol: c_double,
        max_iter: i64, w_init: *const f32, components: *mut f32, mean: *mut f32, n_iter: *mut i64,
        final_lim: *mut c_double, sources: *mut f32) -> c_int;
    pub fn petal_fastica_deflation_fit_f64(ctx: *mut PetalCtx, x: *const f64, n: i64, d: i64, fun: c_int, tol: c_double,
        max_iter: i64, w_init: *const f64, components: *mut f64, mean: *mut f64, n_iter: *mut i64,
        final_lim: *mut c_double, sources: *mut f64) -> c_int;
    pub fn petal_ctx_set_host_staging(ctx: *mut PetalCtx, mode: c_int, chunk_bytes: i64) -> c_int;
    pub fn petal_ctx_set_host_gram(ctx: *mut PetalCtx, enable: c_int) -> c_int;
    pub fn petal_ctx_host_stream_stats(ctx: *const PetalCtx, h2d_bytes: *mut i64, traversals: *mut i64, ring: *mut c_int) -> c_int;
}

#[allow(dead_code)]
pub type Opaque = c_void;
