"""ORACLE (test infrastructure only) - CPU restatement of the reference's PCA path.

Restates, on numpy/scipy (OpenBLAS LAPACK - the same LAPACK family as the reference's
`openblas-static` backend), the algorithms of /root/reference/src/pca.rs with the
LAPACK conventions of src/linalg.rs and src/linalg/lapack.rs.  Each function cites
the reference lines it follows.  The reference crate itself cannot be compiled here
(no cargo/rustc), so this restatement is pinned against the reference's own unit
tests / doctests / README example (tests/test_oracle_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may
import this package - never the product path.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla


class InvalidInput(ValueError):
    """DecompositionError::InvalidInput (reference src/lib.rs:24-25)."""


def svd_flip(u: np.ndarray, v: np.ndarray) -> None:
    """reference src/pca.rs:815-850 - u-based sign fix, first max-|.| entry wins,
    pairs u columns with v rows (zip -> min(u.ncols, v.nrows) pairs). In place."""
    npairs = min(u.shape[1], v.shape[0])
    if u.shape[0] == 0 or npairs == 0:
        return
    idx = np.argmax(np.abs(u[:, :npairs]), axis=0)  # first max wins (pca.rs:830 `abs <= absmax`)
    sign = np.copysign(1.0, u[idx, np.arange(npairs)])  # f64::signum: +0 -> 1, -0 -> -1
    neg = sign < 0
    u[:, :npairs][:, neg] *= -1
    v[:npairs][neg, :] *= -1


def _svd_full(a: np.ndarray, calc_vt: bool = True):
    """reference src/linalg.rs:70-91 -> src/linalg/lapack.rs:103-132:
    gesvd(jobu='A'|'N', jobvt='A') - FULL n x n U."""
    u, s, vt = sla.svd(a, full_matrices=True, lapack_driver="gesvd", check_finite=False)
    return u, s, (vt if calc_vt else None)


class Pca:
    """reference src/pca.rs:41-231 (`Pca<A>`, `PcaBuilder`)."""

    def __init__(self, n_components: int, centering: bool = True, economy: bool = False):
        self.k = n_components
        self.centering = centering
        self.economy = economy  # economy=True: same values, skips the n x n U (baseline timing only)
        self.components = None
        self.means = None
        self.singular = None
        self.total_variance = None
        self.n_samples = 0

    def _inner_fit(self, x: np.ndarray) -> np.ndarray:
        # pca.rs:199-204
        if any(v < self.k for v in x.shape):
            raise InvalidInput(f"every dimension should be at least {self.k}")
        # pca.rs:206-214
        if self.centering:
            if x.shape[0] == 0:
                return np.zeros((0, x.shape[1]), dtype=x.dtype)
            means = x.mean(axis=0, dtype=x.dtype)
        else:
            means = np.zeros(x.shape[1], dtype=x.dtype)
        xc = (x - means) if self.centering else x.copy()  # pca.rs:216-220
        if self.economy:
            u, s, vt = sla.svd(xc, full_matrices=False, lapack_driver="gesvd", check_finite=False)
        else:
            u, s, vt = _svd_full(xc, True)
        u = np.array(u)
        vt = np.array(vt)
        svd_flip(u, vt)  # pca.rs:223
        self.total_variance = float(np.dot(s, s))  # pca.rs:224
        self.components = vt[: self.k].copy()  # pca.rs:225
        self.n_samples = x.shape[0]
        self.means = means
        self.singular = s[: self.k].copy()  # pca.rs:228
        return u

    def fit(self, x):
        self._inner_fit(np.ascontiguousarray(x))

    def fit_transform(self, x):
        x = np.ascontiguousarray(x)
        u = self._inner_fit(x)
        if u.shape[0] == 0 and self.components is None:
            return np.zeros((0, self.k), dtype=x.dtype)
        # transform_with_u, pca.rs:758-779
        return (u[:, : self.k] * self.singular[None, : self.k]).astype(x.dtype)

    def transform(self, x):
        return transform(np.asarray(x), self.components, self.means, self.centering)

    def inverse_transform(self, y):
        return inverse_transform(np.asarray(y), self.components, self.means, self.centering)

    def explained_variance_ratio(self):
        # pca.rs:101-105 - sigma_k^2 / total_variance (no 1/(n-1))
        return self.singular * self.singular / self.total_variance

    def singular_values(self):
        return self.singular


def transform(x, components, means, centering):
    """reference src/pca.rs:726-750."""
    if x.shape[1] != means.shape[0]:
        raise InvalidInput(f"# of columns should be {means.shape[0]}")
    xc = x - means if centering else x
    return xc @ components.T


def inverse_transform(y, components, means, centering):
    """reference src/pca.rs:788-811."""
    if y.shape[1] != components.shape[0]:
        raise InvalidInput(f"# of columns should be {components.shape[0]}")
    out = y @ components
    return out + means if centering else out


def _pl(q: np.ndarray) -> np.ndarray:
    """`lu::Factorized::from(q).into_pl()` then `[:, 0..min(rows, cols)]`
    (reference src/pca.rs:709-710,712-713; lair 0.8 partial-pivot LU, not vendored)."""
    pl, _ = sla.lu(q, permute_l=True, check_finite=False)
    return pl[:, : min(q.shape)]


def _qr_thin(q: np.ndarray) -> np.ndarray:
    """reference src/linalg.rs:127-147 (gelqf + orglq on the transposed view == thin QR)."""
    qq, _ = sla.qr(q, mode="economic", check_finite=False)
    return qq[:, : min(q.shape)]


def randomized_range_finder(x, size, n_iter, omega, normalizer="LU"):
    """reference src/pca.rs:689-718. `omega` is the d x size Gaussian test matrix
    (drawn by the caller from the reference stream, oracle/rng.py)."""
    q = x @ omega  # pca.rs:707
    norm = _pl if normalizer == "LU" else _qr_thin
    for _ in range(n_iter):  # pca.rs:708-715
        pl = norm(q)
        q = x.T @ pl
        pl = norm(q)
        q = x @ pl
    return _qr_thin(q)  # pca.rs:716


def randomized_svd(x, k, omega, n_iter=7, normalizer="LU"):
    """reference src/pca.rs:668-686 (n_random = k + 10 is decided by the shape of omega)."""
    q = randomized_range_finder(x, omega.shape[1], n_iter, omega, normalizer)
    b = q.T @ x  # pca.rs:681
    ub, s, vt = sla.svd(b, full_matrices=False, lapack_driver="gesdd", check_finite=False)  # :682
    u = q @ ub  # pca.rs:683
    vt = np.array(vt)
    svd_flip(u, vt)  # pca.rs:684
    return u, s, vt


class RandomizedPca:
    """reference src/pca.rs:317-550. The RNG lives in the caller: pass `rng` (an
    oracle.rng.Mcg128Xsl64, advanced by every fit like pca.rs:532,536) or `omega`."""

    N_OVERSAMPLES = 10  # pca.rs:679
    N_ITER = 7  # pca.rs:680

    def __init__(self, n_components, rng=None, centering=True, n_iter=None, normalizer="LU"):
        self.k = n_components
        self.rng = rng
        self.centering = centering
        self.n_iter = self.N_ITER if n_iter is None else n_iter
        self.normalizer = normalizer
        self.components = None
        self.means = None
        self.singular = None
        self.total_variance = None
        self.n_samples = 0

    def draw_omega(self, d, dtype):
        return self.rng.normal_matrix(d, self.k + self.N_OVERSAMPLES, dtype)  # pca.rs:701-705

    def _inner_fit(self, x, omega=None):
        if any(v < self.k for v in x.shape):  # pca.rs:513-518
            raise InvalidInput(f"every dimension should be at least {self.k}")
        if self.centering:  # pca.rs:520-528
            if x.shape[0] == 0:
                return np.zeros((0, x.shape[1]), dtype=x.dtype)
            means = x.mean(axis=0, dtype=x.dtype)
        else:
            means = np.zeros(x.shape[1], dtype=x.dtype)
        xc = (x - means) if self.centering else x
        if omega is None:
            omega = self.draw_omega(x.shape[1], x.dtype)
        u, s, vt = randomized_svd(xc, self.k, omega.astype(x.dtype), self.n_iter, self.normalizer)
        # pca.rs:533 - total variance = ||Xc||_F^2 accumulated in A
        self.total_variance = float(np.sum(np.square(xc), dtype=x.dtype))
        self.components = vt[: self.k].copy()
        self.n_samples = x.shape[0]
        self.means = means
        self.singular = s[: self.k].copy()
        return u

    def fit(self, x, omega=None):
        self._inner_fit(np.ascontiguousarray(x), omega)

    def fit_transform(self, x, omega=None):
        x = np.ascontiguousarray(x)
        u = self._inner_fit(x, omega)
        if u.shape[0] == 0 and self.components is None:
            return np.zeros((0, self.k), dtype=x.dtype)
        return (u[:, : self.k] * self.singular[None, : self.k]).astype(x.dtype)

    def transform(self, x):
        return transform(np.asarray(x), self.components, self.means, self.centering)

    def inverse_transform(self, y):
        return inverse_transform(np.asarray(y), self.components, self.means, self.centering)

    def explained_variance_ratio(self):
        return self.singular * self.singular / self.total_variance

    def singular_values(self):
        return self.singular


# ----------------------------------------------------------------------------------
# comparison helpers shared by the parity tests
# ----------------------------------------------------------------------------------
def sign_normalize_rows(c: np.ndarray) -> np.ndarray:
    """Flip each row so its max-|.| entry is positive (sign/permutation normalisation
    named by BASELINE.json north_star)."""
    c = np.array(c, dtype=np.float64)
    idx = np.argmax(np.abs(c), axis=1)
    sg = np.sign(c[np.arange(c.shape[0]), idx])
    sg[sg == 0] = 1
    return c * sg[:, None]


def principal_angles(a_rows: np.ndarray, b_rows: np.ndarray) -> np.ndarray:
    """Principal angles (radians) between the row spaces of two k x d matrices."""
    qa, _ = np.linalg.qr(np.asarray(a_rows, dtype=np.float64).T)
    qb, _ = np.linalg.qr(np.asarray(b_rows, dtype=np.float64).T)
    s = np.clip(np.linalg.svd(qa.T @ qb, compute_uv=False), -1.0, 1.0)
    return np.arccos(s)
