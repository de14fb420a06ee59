"""ORACLE (test infrastructure only) - CPU restatement of the reference's FastICA path.

Follows /root/reference/src/ica.rs:167-222 (inner_fit), :319-361 (ica_par),
:363-381 (symmetric_decorrelation), :383-398 (logcosh) with the LAPACK conventions of
src/linalg.rs:39-60 (eigh -> syev 'V','L') and :70-91 (svd -> gesvd).  Pinned against
the reference's own unit tests (tests/test_oracle_golden.py).

Two places where the reference's literal code differs from the textbook algorithm are
switchable:

* symdec = "textbook": (W W^T)^-1/2 W = V D V^T W        (what the product implements)
  symdec = "literal" : the row-major/col-major mix-up of src/linalg.rs:57-59 +
                       src/ica.rs:370-380 computes V^T D V W (identical on the 2x2 goldens).
* lim    = "rowrow"  : max_i | |w1_i . w_i| - 1 |  (sklearn; textbook)
  lim    = "rowcol"  : rows of w1 zipped with COLUMNS of w, src/ica.rs:345-349 (literal).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may
import this package - never the product path.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
from scipy.linalg import lapack


class InvalidInput(ValueError):
    pass


def _syev(a: np.ndarray):
    fn = lapack.ssyev if a.dtype == np.float32 else lapack.dsyev
    w, v, info = fn(a, compute_v=1, lower=1)
    if info != 0:
        raise np.linalg.LinAlgError("cannot compute eigenvalues")
    return w, v  # ascending eigenvalues, eigenvectors as COLUMNS


def symmetric_decorrelation(w: np.ndarray, symdec: str = "textbook") -> np.ndarray:
    """reference src/ica.rs:363-381."""
    e, v = _syev(w @ w.T)  # ica.rs:369
    d = (1.0 / np.sqrt(e)).astype(w.dtype)  # ica.rs:371-374
    if symdec == "textbook":
        return (v * d[None, :]) @ v.T @ w
    # literal: the buffer LAPACK filled column-major is read row-major, so the
    # reference's `v` is V^T; it scales v[i][j] *= d[j] and multiplies by v.t() = V.
    vt = v.T
    return (vt * d[None, :]) @ v @ w


def logcosh(wx: np.ndarray):
    """reference src/ica.rs:383-398 - (tanh(wx), row means of 1 - tanh^2)."""
    g = np.tanh(wx)
    gp = np.sum(1.0 - g * g, axis=1, dtype=wx.dtype) / wx.dtype.type(wx.shape[1])
    return g, gp


def exp_fun(wx: np.ndarray):
    """`exp` contrast function. Not in the reference (src/ica.rs:383-398 has logcosh only); BASELINE.json's
    north_star names it, so it follows the algorithm the crate mirrors: sklearn
    decomposition/_fastica.py:160-164 (`_exp`): g = x exp(-x^2/2), g' = (1 - x^2) exp(-x^2/2)."""
    e = np.exp(-(wx * wx) / 2)
    return wx * e, ((1 - wx * wx) * e).mean(axis=-1)


def cube_fun(wx: np.ndarray):
    """`cube` contrast function, sklearn decomposition/_fastica.py:167-168 (`_cube`): g = x^3, g' = 3 x^2."""
    return wx ** 3, (3 * wx * wx).mean(axis=-1)


G_FUNS = {"logcosh": logcosh, "exp": exp_fun, "cube": cube_fun}


def ica_par(x1: np.ndarray, tol: float, max_iter: int, w_init: np.ndarray,
            symdec: str = "textbook", lim: str = "rowrow", fun: str = "logcosh"):
    """reference src/ica.rs:319-361. x1 is nc x n (whitened, features x samples)."""
    g = G_FUNS[fun]
    w = symmetric_decorrelation(w_init, symdec)  # ica.rs:329
    p_inv = x1.dtype.type(1.0) / x1.dtype.type(x1.shape[1])
    for i in range(max_iter):
        gwtx, g_wtx = g(w @ x1)  # ica.rs:332
        gd = gwtx @ x1.T  # ica.rs:333
        gd = gd * p_inv - g_wtx[:, None] * w  # ica.rs:334-342
        w1 = symmetric_decorrelation(gd, symdec)  # ica.rs:343
        if lim == "rowrow":
            dots = np.einsum("ij,ij->i", w1, w)
        else:  # ica.rs:345-349: rows of w1 zipped with columns of w
            dots = np.einsum("ij,ji->i", w1, w)
        limv = np.max(np.abs(np.abs(dots) - 1.0))
        if limv < tol:  # ica.rs:355-357
            return w1, i + 1
        w = w1
    return w, max_iter


def ica_def(x1: np.ndarray, tol: float, max_iter: int, w_init: np.ndarray, fun: str = "logcosh"):
    """Deflation FastICA: sklearn `_ica_def` (sklearn/decomposition/_fastica.py:65-100; `_gs_decorrelation` :38-62).
    Not in the reference (src/ica.rs has the symmetric scheme only) - SURVEY 8(f) rank 4.  x1 is nc x n (whitened).
    Pinned against sklearn's own function by tests/golden/ica_deflation.json (tests/golden/make_ica_deflation_fixture.py).
    Returns (W, largest iteration count over the components)."""
    g = G_FUNS[fun]
    nc = w_init.shape[0]
    w_all = np.zeros((nc, nc), dtype=x1.dtype)
    n_iter = []
    for j in range(nc):
        w = w_init[j, :].copy()
        w /= np.sqrt((w ** 2).sum())  # _fastica.py:79-80
        i = 0
        for i in range(max_iter):
            gwtx, g_wtx = g((w @ x1)[None, :])  # :83
            w1 = (x1 * gwtx[0]).mean(axis=1) - g_wtx[0] * w  # :85
            w1 -= (w1 @ w_all[:j].T) @ w_all[:j]  # _gs_decorrelation, :60-61
            w1 /= np.sqrt((w1 ** 2).sum())  # :89
            lim = np.abs(np.abs((w1 * w).sum()) - 1)  # :91
            w = w1
            if lim < tol:  # :93-94
                break
        n_iter.append(i + 1)
        w_all[j, :] = w
    return w_all, max(n_iter)


class FastIca:
    """reference src/ica.rs:41-222. `rng` is an oracle.rng.Mcg128Xsl64 or pass w_init.
    algorithm="deflation" swaps ica_par for ica_def (extension, see ica_def)."""

    TOL = 1e-4  # ica.rs:216
    MAX_ITER = 200

    def __init__(self, rng=None, symdec="textbook", lim="rowrow", max_iter=None, tol=None, algorithm="parallel",
                 fun="logcosh"):
        self.rng = rng
        self.algorithm = algorithm
        self.fun = fun
        self.symdec = symdec
        self.lim = lim
        self.max_iter = self.MAX_ITER if max_iter is None else max_iter
        self.tol = self.TOL if tol is None else tol
        self.components = None
        self.means = None
        self.n_iter = 0
        self.whitening = None

    def _inner_fit(self, x: np.ndarray, w_init=None):
        n, d = x.shape
        nc = min(n, d)  # ica.rs:173
        if n == 0:
            return np.zeros((0, d), dtype=x.dtype)
        means = x.mean(axis=0, dtype=x.dtype)  # ica.rs:174
        xt = np.ascontiguousarray((x - means).T)  # ica.rs:178-188, d x n
        # ica.rs:189: svd(x.clone(), calc_vt=false) -> sigma and the d x d U
        u, s, _ = sla.svd(xt, full_matrices=False, lapack_driver="gesvd", check_finite=False)
        # ica.rs:190-203: K[i][j] = U[j][i] / sigma[i]
        k = (u[:, :nc] / s[None, :nc]).T.astype(x.dtype)
        x1 = (k @ xt) * x.dtype.type(np.sqrt(x.dtype.type(n)))  # ica.rs:204-208
        if w_init is None:
            w_init = self.rng.normal_matrix(nc, nc, x.dtype)  # ica.rs:210-214
        if self.algorithm == "deflation":
            w, n_iter = ica_def(x1, x.dtype.type(self.tol), self.max_iter, w_init.astype(x.dtype), self.fun)
        else:
            w, n_iter = ica_par(x1, x.dtype.type(self.tol), self.max_iter, w_init.astype(x.dtype),
                                self.symdec, self.lim, self.fun)
        self.components = w @ k  # ica.rs:217
        self.means = means
        self.n_iter = n_iter
        self.whitening = k
        return xt

    def fit(self, x, w_init=None):
        self._inner_fit(np.ascontiguousarray(x), w_init)

    def fit_transform(self, x, w_init=None):
        xt = self._inner_fit(np.ascontiguousarray(x), w_init)
        return np.ascontiguousarray((self.components @ xt).T)  # ica.rs:155-156

    def transform(self, x):
        x = np.asarray(x)
        if x.shape[1] != self.means.shape[0]:  # ica.rs:124-128
            raise InvalidInput("too many columns")
        return (x - self.means) @ self.components.T  # ica.rs:129-130


# ----------------------------------------------------------------------------------
# comparison helpers
# ----------------------------------------------------------------------------------
def match_rows(a: np.ndarray, b: np.ndarray):
    """Greedy sign/permutation matching of the rows of `a` to the rows of `b`.
    Returns (a_matched, max_abs_row_cosine_defect)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    an = a / np.linalg.norm(a, axis=1, keepdims=True)
    bn = b / np.linalg.norm(b, axis=1, keepdims=True)
    c = an @ bn.T
    out = np.zeros_like(b)
    used = set()
    defect = 0.0
    order = np.argsort(-np.max(np.abs(c), axis=0))
    for j in order:
        cand = [(abs(c[i, j]), i) for i in range(a.shape[0]) if i not in used]
        val, i = max(cand)
        used.add(i)
        out[j] = a[i] * np.sign(c[i, j])
        defect = max(defect, 1.0 - val)
    return out, defect


def amari_index(w: np.ndarray, a: np.ndarray) -> float:
    """Amari distance of P = W A from a scaled permutation (0 = perfect unmixing)."""
    p = np.abs(np.asarray(w, np.float64) @ np.asarray(a, np.float64))
    d = p.shape[0]
    r = (p / p.max(axis=1, keepdims=True)).sum(axis=1) - 1.0
    c = (p / p.max(axis=0, keepdims=True)).sum(axis=0) - 1.0
    return float((r.sum() + c.sum()) / (2.0 * d * (d - 1)))
