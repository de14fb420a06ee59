//! petal-decomposition on B200: the crate's public API (reference src/lib.rs:17-18) with the
//! arithmetic forwarded to libpetal_b200.so.  Type names, builders, method signatures and error
//! behaviour follow the reference (src/pca.rs:41-663, src/ica.rs:41-317); the generic `A` is
//! restricted to f32 / f64 through the sealed `Scalar` trait below, which picks the `_f32` / `_f64`
//! entry points.  Omega / w_init are drawn here with the *real* rand_pcg / rand_distr crates in
//! the reference's order (src/pca.rs:701-705, src/ica.rs:210-214) and handed to the library, so
//! seeded results follow the reference's stream exactly and the model's RNG advances per fit.
//!
//! SOURCE ONLY: this image has no Rust toolchain; see INTEGRATION.md for how a maintainer builds it.
mod ffi;

use ndarray::{Array1, Array2, ArrayBase, Data, Ix2};
use rand::{Rng, RngCore, SeedableRng};
use rand_distr::StandardNormal;
use rand_pcg::Mcg128Xsl64 as Pcg;
use std::ffi::CStr;
use std::sync::OnceLock;
use thiserror::Error;

/// reference src/lib.rs:22-28
#[derive(Debug, Error)]
pub enum DecompositionError {
    #[error("invalid matrix: {0}")]
    InvalidInput(String),
    #[error("linear algerba operation failed: {0}")]
    LinalgError(String),
}

struct Ctx(*mut ffi::PetalCtx);
// libpetal_b200 serialises calls on one context with an internal mutex (one call at a time per context, like the
// reference's `&mut self` on fit), so sharing the handle between threads is sound.
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}

fn ctx() -> Result<*mut ffi::PetalCtx, DecompositionError> {
    static CTX: OnceLock<Result<Ctx, String>> = OnceLock::new();
    match CTX.get_or_init(|| unsafe {
        let mut p = std::ptr::null_mut();
        if ffi::petal_ctx_create(0, &mut p) == ffi::PETAL_OK {
            Ok(Ctx(p))
        } else {
            Err(CStr::from_ptr(ffi::petal_last_global_error()).to_string_lossy().into_owned())
        }
    }) {
        Ok(c) => Ok(c.0),
        Err(e) => Err(DecompositionError::LinalgError(e.clone())),
    }
}

fn check(ctx: *mut ffi::PetalCtx, status: i32) -> Result<(), DecompositionError> {
    if status == ffi::PETAL_OK {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(ffi::petal_last_error(ctx)).to_string_lossy().into_owned() };
    if status == ffi::PETAL_INVALID_INPUT {
        Err(DecompositionError::InvalidInput(msg))
    } else {
        Err(DecompositionError::LinalgError(msg))
    }
}

mod sealed {
    pub trait Sealed {}
    impl Sealed for f32 {}
    impl Sealed for f64 {}
}

/// f32 / f64: selects the `_f32` / `_f64` entry points of the C ABI.
pub trait Scalar: sealed::Sealed + Copy + Default + num_like::Float {
    #[allow(clippy::too_many_arguments)]
    unsafe fn pca_fit(c: *mut ffi::PetalCtx, x: *const Self, n: i64, d: i64, k: i64, centering: i32, comps: *mut Self,
        mean: *mut Self, sing: *mut Self, tv: *mut Self, scores: *mut Self) -> i32;
    #[allow(clippy::too_many_arguments)]
    unsafe fn rpca_fit(c: *mut ffi::PetalCtx, x: *const Self, n: i64, d: i64, k: i64, centering: i32, over: i64,
        iters: i64, omega: *const Self, comps: *mut Self, mean: *mut Self, sing: *mut Self, tv: *mut Self,
        scores: *mut Self) -> i32;
    unsafe fn transform(c: *mut ffi::PetalCtx, x: *const Self, n: i64, d: i64, comps: *const Self, k: i64,
        mean: *const Self, out: *mut Self) -> i32;
    unsafe fn inverse_transform(c: *mut ffi::PetalCtx, y: *const Self, n: i64, k: i64, comps: *const Self, d: i64,
        mean: *const Self, out: *mut Self) -> i32;
    #[allow(clippy::too_many_arguments)]
    unsafe fn fastica_fit(c: *mut ffi::PetalCtx, x: *const Self, n: i64, d: i64, w_init: *const Self,
        comps: *mut Self, mean: *mut Self, n_iter: *mut i64, sources: *mut Self) -> i32;
    fn from_f64(v: f64) -> Self;
}

/// minimal float surface the shim needs (kept local so the shim has no extra dependencies)
pub mod num_like {
    pub trait Float: std::ops::Mul<Output = Self> + std::ops::Div<Output = Self> + Sized {}
    impl Float for f32 {}
    impl Float for f64 {}
}

macro_rules! impl_scalar {
    ($t:ty, $pca:ident, $rpca:ident, $tr:ident, $inv:ident, $ica:ident) => {
        impl Scalar for $t {
            unsafe fn pca_fit(c: *mut ffi::PetalCtx, x: *const Self, n: i64, d: i64, k: i64, centering: i32,
                comps: *mut Self, mean: *mut Self, sing: *mut Self, tv: *mut Self, scores: *mut Self) -> i32 {
                ffi::$pca(c, x, n, d, k, centering, comps, mean, sing, tv, scores)
            }
            unsafe fn rpca_fit(c: *mut ffi::PetalCtx, x: *const Self, n: i64, d: i64, k: i64, centering: i32, over: i64,
                iters: i64, omega: *const Self, comps: *mut Self, mean: *mut Self, sing: *mut Self, tv: *mut Self,
                scores: *mut Self) -> i32 {
                ffi::$rpca(c, x, n, d, k, centering, over, iters, omega, comps, mean, sing, tv, scores)
            }
            unsafe fn transform(c: *mut ffi::PetalCtx, x: *const Self, n: i64, d: i64, comps: *const Self, k: i64,
                mean: *const Self, out: *mut Self) -> i32 {
                ffi::$tr(c, x, n, d, comps, k, mean, out)
            }
            unsafe fn inverse_transform(c: *mut ffi::PetalCtx, y: *const Self, n: i64, k: i64, comps: *const Self,
                d: i64, mean: *const Self, out: *mut Self) -> i32 {
                ffi::$inv(c, y, n, k, comps, d, mean, out)
            }
            unsafe fn fastica_fit(c: *mut ffi::PetalCtx, x: *const Self, n: i64, d: i64, w_init: *const Self,
                comps: *mut Self, mean: *mut Self, n_iter: *mut i64, sources: *mut Self) -> i32 {
                // reference constants: logcosh, tol 1e-4, max_iter 200 (src/ica.rs:216); row.row test
                ffi::$ica(c, x, n, d, 0, 1e-4, 200, 0, w_init, comps, mean, n_iter, std::ptr::null_mut(), sources)
            }
            fn from_f64(v: f64) -> Self { v as $t }
        }
    };
}
impl_scalar!(f32, petal_pca_fit_f32, petal_rpca_fit_f32, petal_transform_f32, petal_inverse_transform_f32, petal_fastica_fit_f32);
impl_scalar!(f64, petal_pca_fit_f64, petal_rpca_fit_f64, petal_transform_f64, petal_inverse_transform_f64, petal_fastica_fit_f64);

/// The reference asserts the standard layout (src/linalg.rs:75); a non-contiguous view is copied here.
fn standard<A: Scalar, S: Data<Elem = A>>(x: &ArrayBase<S, Ix2>) -> Array2<A> {
    x.as_standard_layout().into_owned()
}

/// reference `Pca<A>` (src/pca.rs:41-231)
pub struct Pca<A: Scalar> {
    components: Array2<A>,
    n_samples: usize,
    means: Array1<A>,
    total_variance: A,
    singular: Array1<A>,
    centering: bool,
}

impl<A: Scalar> Pca<A> {
    #[must_use]
    pub fn new(n_components: usize) -> Self {
        Self { components: Array2::default((n_components, 0)), n_samples: 0, means: Array1::default(0),
               total_variance: A::default(), singular: Array1::default(0), centering: true }
    }
    pub fn components(&self) -> &Array2<A> { &self.components }
    pub fn mean(&self) -> &Array1<A> { &self.means }
    pub fn n_components(&self) -> usize { self.components.nrows() }
    pub fn singular_values(&self) -> &Array1<A> { &self.singular }
    pub fn explained_variance_ratio(&self) -> Array1<A> {
        self.singular.mapv(|s| s * s / self.total_variance) // src/pca.rs:101-105
    }
    pub fn fit<S: Data<Elem = A>>(&mut self, input: &ArrayBase<S, Ix2>) -> Result<(), DecompositionError> {
        self.inner_fit(input, false).map(|_| ())
    }
    pub fn fit_transform<S: Data<Elem = A>>(&mut self, input: &ArrayBase<S, Ix2>) -> Result<Array2<A>, DecompositionError> {
        self.inner_fit(input, true)
    }
    pub fn transform<S: Data<Elem = A>>(&self, input: &ArrayBase<S, Ix2>) -> Result<Array2<A>, DecompositionError> {
        transform(input, &self.components, &self.means, self.centering)
    }
    pub fn inverse_transform<S: Data<Elem = A>>(&self, input: &ArrayBase<S, Ix2>) -> Result<Array2<A>, DecompositionError> {
        inverse_transform(input, &self.components, &self.means, self.centering)
    }
    fn inner_fit<S: Data<Elem = A>>(&mut self, input: &ArrayBase<S, Ix2>, scores: bool) -> Result<Array2<A>, DecompositionError> {
        let k = self.n_components();
        if input.shape().iter().any(|v| *v < k) {
            return Err(DecompositionError::InvalidInput(format!("every dimension should be at least {k}")));
        }
        let (n, d) = input.dim();
        if n == 0 { return Ok(Array2::default((0, k))); } // src/pca.rs:207-211
        let x = standard(input);
        let c = ctx()?;
        let mut comps = Array2::<A>::default((k, d));
        let mut means = Array1::<A>::default(d);
        let mut sing = Array1::<A>::default(k);
        let mut tv = A::default();
        let mut y = Array2::<A>::default((if scores { n } else { 0 }, k));
        let yp = if scores { y.as_mut_ptr() } else { std::ptr::null_mut() };
        check(c, unsafe { A::pca_fit(c, x.as_ptr(), n as i64, d as i64, k as i64, self.centering as i32,
            comps.as_mut_ptr(), means.as_mut_ptr(), sing.as_mut_ptr(), &mut tv, yp) })?;
        self.components = comps; self.means = means; self.singular = sing; self.total_variance = tv; self.n_samples = n;
        Ok(y)
    }
}

/// reference `PcaBuilder` (src/pca.rs:246-283)
pub struct PcaBuilder { n_components: usize, centering: bool }
impl PcaBuilder {
    #[must_use] pub fn new(n_components: usize) -> Self { Self { n_components, centering: true } }
    #[must_use] pub fn centering(mut self, centering: bool) -> Self { self.centering = centering; self }
    #[must_use] pub fn build<A: Scalar>(self) -> Pca<A> { let mut p = Pca::new(self.n_components); p.centering = self.centering; p }
}

/// reference `RandomizedPca<A, R>` (src/pca.rs:317-550)
pub struct RandomizedPca<A: Scalar, R: RngCore = Pcg> { inner: Pca<A>, rng: R }
impl<A: Scalar> RandomizedPca<A, Pcg> {
    #[must_use] pub fn new(n_components: usize) -> Self { Self::with_rng(n_components, Pcg::from_rng(&mut rand::rng())) }
    #[must_use] pub fn with_seed(n_components: usize, seed: u128) -> Self {
        Self::with_rng(n_components, Pcg::from_seed(seed.to_be_bytes())) // src/pca.rs:356-358
    }
}
impl<A: Scalar, R: RngCore> RandomizedPca<A, R> {
    #[must_use] pub fn with_rng(n_components: usize, rng: R) -> Self { Self { inner: Pca::new(n_components), rng } }
    pub fn components(&self) -> &Array2<A> { self.inner.components() }
    pub fn mean(&self) -> &Array1<A> { self.inner.mean() }
    pub fn n_components(&self) -> usize { self.inner.n_components() }
    pub fn singular_values(&self) -> &Array1<A> { self.inner.singular_values() }
    pub fn explained_variance_ratio(&self) -> Array1<A> { self.inner.explained_variance_ratio() }
    pub fn fit<S: Data<Elem = A>>(&mut self, input: &ArrayBase<S, Ix2>) -> Result<(), DecompositionError> { self.inner_fit(input, false).map(|_| ()) }
    pub fn fit_transform<S: Data<Elem = A>>(&mut self, input: &ArrayBase<S, Ix2>) -> Result<Array2<A>, DecompositionError> { self.inner_fit(input, true) }
    pub fn transform<S: Data<Elem = A>>(&self, input: &ArrayBase<S, Ix2>) -> Result<Array2<A>, DecompositionError> { self.inner.transform(input) }
    pub fn inverse_transform<S: Data<Elem = A>>(&self, input: &ArrayBase<S, Ix2>) -> Result<Array2<A>, DecompositionError> { self.inner.inverse_transform(input) }
    fn inner_fit<S: Data<Elem = A>>(&mut self, input: &ArrayBase<S, Ix2>, scores: bool) -> Result<Array2<A>, DecompositionError> {
        let k = self.n_components();
        if input.shape().iter().any(|v| *v < k) {
            return Err(DecompositionError::InvalidInput(format!("every dimension should be at least {k}")));
        }
        let (n, d) = input.dim();
        if n == 0 { return Ok(Array2::default((0, k))); }
        let l = k + 10; // src/pca.rs:679
        // src/pca.rs:701-705: row-major visiting order, one f64 StandardNormal per element, cast to A
        let rng = &mut self.rng;
        let omega = Array2::<A>::from_shape_fn((d, l), |_| A::from_f64(rng.sample::<f64, _>(StandardNormal)));
        let x = standard(input);
        let c = ctx()?;
        let mut comps = Array2::<A>::default((k, d));
        let mut means = Array1::<A>::default(d);
        let mut sing = Array1::<A>::default(k);
        let mut tv = A::default();
        let mut y = Array2::<A>::default((if scores { n } else { 0 }, k));
        let yp = if scores { y.as_mut_ptr() } else { std::ptr::null_mut() };
        check(c, unsafe { A::rpca_fit(c, x.as_ptr(), n as i64, d as i64, k as i64, self.inner.centering as i32, 10, 7,
            omega.as_ptr(), comps.as_mut_ptr(), means.as_mut_ptr(), sing.as_mut_ptr(), &mut tv, yp) })?;
        self.inner.components = comps; self.inner.means = means; self.inner.singular = sing;
        self.inner.total_variance = tv; self.inner.n_samples = n;
        Ok(y)
    }
}

/// reference `RandomizedPcaBuilder<R>` (src/pca.rs:564-663)
pub struct RandomizedPcaBuilder<R: RngCore = Pcg> { rng: R, n_components: usize, centering: bool }
impl RandomizedPcaBuilder<Pcg> {
    #[must_use] pub fn new(n_components: usize) -> Self { Self { rng: Pcg::from_rng(&mut rand::rng()), n_components, centering: true } }
    #[must_use] pub fn seed(mut self, seed: u128) -> Self { self.rng = Pcg::from_seed(seed.to_be_bytes()); self }
}
impl<R: RngCore> RandomizedPcaBuilder<R> {
    #[must_use] pub fn with_rng(rng: R, n_components: usize) -> Self { Self { rng, n_components, centering: true } } // src/pca.rs:643
    #[must_use] pub fn centering(mut self, centering: bool) -> Self { self.centering = centering; self }
    #[must_use] pub fn build<A: Scalar>(self) -> RandomizedPca<A, R> {
        let mut p = RandomizedPca::with_rng(self.n_components, self.rng); p.inner.centering = self.centering; p
    }
}

/// reference `FastIca<A, R>` (src/ica.rs:41-222): fit / transform / fit_transform only
pub struct FastIca<A: Scalar, R: RngCore = Pcg> { rng: R, components: Array2<A>, means: Array1<A>, n_iter: usize }
impl<A: Scalar> FastIca<A, Pcg> {
    #[must_use] pub fn new() -> Self { Self::with_rng(Pcg::from_rng(&mut rand::rng())) }
    #[must_use] pub fn with_seed(seed: u128) -> Self { Self::with_rng(Pcg::from_seed(seed.to_be_bytes())) }
}
impl<A: Scalar> Default for FastIca<A, Pcg> { fn default() -> Self { Self::new() } }
impl<A: Scalar, R: RngCore> FastIca<A, R> {
    #[must_use] pub fn with_rng(rng: R) -> Self { Self { rng, components: Array2::default((0, 0)), means: Array1::default(0), n_iter: 0 } }
    pub fn fit<S: Data<Elem = A>>(&mut self, input: &ArrayBase<S, Ix2>) -> Result<(), DecompositionError> { self.inner_fit(input, false).map(|_| ()) }
    pub fn fit_transform<S: Data<Elem = A>>(&mut self, input: &ArrayBase<S, Ix2>) -> Result<Array2<A>, DecompositionError> { self.inner_fit(input, true) }
    pub fn transform<S: Data<Elem = A>>(&self, input: &ArrayBase<S, Ix2>) -> Result<Array2<A>, DecompositionError> {
        if input.ncols() != self.means.len() {
            return Err(DecompositionError::InvalidInput("too many columns".to_string())); // src/ica.rs:124-128
        }
        transform(input, &self.components, &self.means, true)
    }
    fn inner_fit<S: Data<Elem = A>>(&mut self, input: &ArrayBase<S, Ix2>, sources: bool) -> Result<Array2<A>, DecompositionError> {
        let (n, d) = input.dim();
        let nc = n.min(d); // src/ica.rs:173
        if n == 0 { return Ok(Array2::default((0, d))); }
        let rng = &mut self.rng;
        let w_init = Array2::<A>::from_shape_fn((nc, nc), |_| A::from_f64(rng.sample::<f64, _>(StandardNormal))); // :210-214
        let x = standard(input);
        let c = ctx()?;
        let mut comps = Array2::<A>::default((nc, d));
        let mut means = Array1::<A>::default(d);
        let mut n_iter = 0i64;
        let mut s = Array2::<A>::default((if sources { n } else { 0 }, nc));
        let sp = if sources { s.as_mut_ptr() } else { std::ptr::null_mut() };
        check(c, unsafe { A::fastica_fit(c, x.as_ptr(), n as i64, d as i64, w_init.as_ptr(), comps.as_mut_ptr(),
            means.as_mut_ptr(), &mut n_iter, sp) })?;
        self.components = comps; self.means = means; self.n_iter = n_iter as usize;
        Ok(s)
    }
}

/// reference `FastIcaBuilder<R>` (src/ica.rs:244-317)
pub struct FastIcaBuilder<R: RngCore = Pcg> { rng: R }
impl FastIcaBuilder<Pcg> {
    #[must_use] pub fn new() -> Self { Self { rng: Pcg::from_rng(&mut rand::rng()) } }
    #[must_use] pub fn seed(mut self, seed: u128) -> Self { self.rng = Pcg::from_seed(seed.to_be_bytes()); self }
}
impl Default for FastIcaBuilder<Pcg> { fn default() -> Self { Self::new() } }
impl<R: RngCore> FastIcaBuilder<R> {
    #[must_use] pub fn with_rng(rng: R) -> Self { Self { rng } }
    #[must_use] pub fn build<A: Scalar>(self) -> FastIca<A, R> { FastIca::with_rng(self.rng) }
}

/// reference `transform` (src/pca.rs:726-750)
fn transform<A: Scalar, S: Data<Elem = A>>(input: &ArrayBase<S, Ix2>, components: &Array2<A>, means: &Array1<A>,
    centering: bool) -> Result<Array2<A>, DecompositionError> {
    if input.ncols() != means.len() {
        return Err(DecompositionError::InvalidInput(format!("# of columns should be {}", means.len())));
    }
    let x = standard(input);
    let (n, d) = x.dim();
    let k = components.nrows();
    let c = ctx()?;
    let mut out = Array2::<A>::default((n, k));
    let mp = if centering { means.as_ptr() } else { std::ptr::null() };
    check(c, unsafe { A::transform(c, x.as_ptr(), n as i64, d as i64, components.as_ptr(), k as i64, mp, out.as_mut_ptr()) })?;
    Ok(out)
}

/// reference `inverse_transform` (src/pca.rs:788-811)
fn inverse_transform<A: Scalar, S: Data<Elem = A>>(input: &ArrayBase<S, Ix2>, components: &Array2<A>,
    means: &Array1<A>, centering: bool) -> Result<Array2<A>, DecompositionError> {
    if input.ncols() != components.nrows() {
        return Err(DecompositionError::InvalidInput(format!("# of columns should be {}", components.nrows())));
    }
    let y = standard(input);
    let (n, k) = y.dim();
    let d = components.ncols();
    let c = ctx()?;
    let mut out = Array2::<A>::default((n, d));
    let mp = if centering { means.as_ptr() } else { std::ptr::null() };
    check(c, unsafe { A::inverse_transform(c, y.as_ptr(), n as i64, k as i64, components.as_ptr(), d as i64, mp, out.as_mut_ptr()) })?;
    Ok(out)
}
