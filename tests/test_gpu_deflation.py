"""Deflation FastICA (sklearn `_ica_def`) on the GPU through the C ABI: sklearn's golden vectors
replayed on whitened inputs, full fits against the oracle restatement, the fused one-pass kernel at every group width
/ vector count it is instantiated for, and the wide-row fallback."""
import json
import os

import numpy as np
import pytest

from oracle import ica as oica
from oracle.rng import Mcg128Xsl64
from tests import synth

pytestmark = pytest.mark.gpu

RNG_SEED = 1_234_567_891_011_121_314
FUNS = {"logcosh": 0, "exp": 1, "cube": 2}


@pytest.fixture(scope="module")
def pd():
    import petal_decomposition_b200 as m
    return m


def _rows_close(a, b, atol):
    s = np.sign(np.sum(a * b, axis=1, keepdims=True))
    return np.max(np.abs(a - s * b)) < atol, float(np.max(np.abs(a - s * b)))


def test_ica_def_replays_sklearn_golden_vectors(pd):
    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ica_deflation.json")))
    for c in fx["cases"]:
        x1 = np.array(c["x1"])
        w, it = pd.ica_def(np.ascontiguousarray(x1.T), c["tol"], c["max_iter"], np.array(c["w_init"]), FUNS[c["fun"]])
        assert it == c["n_iter"], (c["fun"], it, c["n_iter"])
        ok, err = _rows_close(w, np.array(c["w"]), 1e-9)
        assert ok, (c["fun"], err)


@pytest.mark.parametrize("fun", ["logcosh", "exp", "cube"])
@pytest.mark.parametrize("dtype,n,d", [(np.float64, 20_000, 4), (np.float64, 30_000, 12), (np.float32, 50_000, 8)])
def test_fastica_deflation_fit_vs_oracle(pd, fun, dtype, n, d):
    """Full fit (mean, whitening, deflation, components = W K, sources) against the oracle, iterate for iterate.
    The whitened coordinates are only defined up to the signs of the singular vectors (LAPACK's choice in the oracle,
    the Jacobi solver's here), and with deflation the same w_init is then a different starting point: extraction order,
    iteration counts and - at the level of the sampling error - even the constrained optima differ.  So the product's
    own whitening matrix is read back (a fit with max_iter = 0 and w_init = I returns components = K), checked against
    the oracle's row by row up to sign, and the oracle's deflation runs in those coordinates."""
    x, _ = synth.mixed_sources(n, d, seed=11, dtype=dtype)
    x64 = x.astype(np.float64)
    w0 = Mcg128Xsl64.from_seed_u128(RNG_SEED).normal_matrix(d, d, dtype)
    f64 = dtype == np.float64
    # whitening parity (src/ica.rs:189-208)
    mk = pd.FastIca(pd.Pcg.from_seed(RNG_SEED), fun=FUNS[fun], max_iter=0, algorithm=pd.DEFLATION)
    mk.fit(x, np.eye(d, dtype=dtype))
    assert mk.n_iter == 0
    kp = mk.components.astype(np.float64)
    ref = oica.FastIca(algorithm="deflation", fun=fun)
    ref.fit(x64, w0.astype(np.float64))
    ok, err = _rows_close(kp / np.linalg.norm(ref.whitening, axis=1, keepdims=True),
                          ref.whitening / np.linalg.norm(ref.whitening, axis=1, keepdims=True), 1e-8 if f64 else 1e-4)
    assert ok, err
    assert np.allclose(mk.means, ref.means, atol=1e-13 if f64 else 1e-6)
    # the deflation itself, in the product's whitened coordinates
    xc = (x64 - mk.means.astype(np.float64)).T
    x1 = (kp @ xc) * np.sqrt(n)
    w_ref, it_ref = oica.ica_def(x1, 1e-4, 200, w0.astype(np.float64), fun)
    comps_ref = w_ref @ kp
    m = pd.FastIca(pd.Pcg.from_seed(RNG_SEED), fun=FUNS[fun], algorithm=pd.DEFLATION)
    s = m.fit_transform(x, w0)
    assert abs(m.n_iter - it_ref) <= (0 if f64 else 2), (m.n_iter, it_ref)
    scale = np.linalg.norm(comps_ref, axis=1, keepdims=True)
    ok, err = _rows_close(m.components / scale, comps_ref / scale, 1e-7 if f64 else 2e-3)
    assert ok, err
    assert np.allclose(n * (m.components.astype(np.float64) @ np.cov(x64.T, bias=True) @ m.components.astype(np.float64).T),
                       np.eye(d), atol=1e-6 if f64 else 1e-3)  # unmixed signals are white
    sr = (comps_ref @ xc).T
    sg = np.sign(np.sum(np.asarray(s, np.float64) * sr, axis=0, keepdims=True))
    assert np.allclose(s, sg * sr, atol=(1e-7 if f64 else 5e-3) * np.abs(sr).max())


@pytest.mark.parametrize("dtype,d", [(np.float32, 20), (np.float32, 64), (np.float32, 250), (np.float32, 1024), (np.float32, 130),
                                     (np.float64, 64), (np.float64, 33), (np.float64, 512)])
def test_deflation_pass_kernel_instantiations(pd, dtype, d):
    """Every group width / vectors-per-lane instantiation of the fused pass (aligned and scalar-load variants) on the
    whitened-data entry (K1 = null, so the pass runs at width d), against the oracle's iterates at a fixed iteration
    count (so that f32 and f64 walk the same path)."""
    n = 6000
    rng = np.random.default_rng(d)
    # d unit-variance Laplace sources under an orthonormal map: white up to sampling error, which is all ica_def asks
    q, _ = np.linalg.qr(rng.standard_normal((d, d)))
    xfull = (rng.laplace(size=(n, d)) / np.sqrt(2.0)) @ q.T
    x = np.ascontiguousarray(xfull.astype(dtype))
    w0 = rng.standard_normal((d, d))
    # whitened-data entry (K1 = null, d = nc): exercises the pass kernel at width d directly
    xw = x.astype(np.float64)
    w_ref, it_ref = oica.ica_def(np.ascontiguousarray(xw.T), 0.0, 2, w0.copy(), "logcosh")
    w, it = pd.ica_def(x, 0.0, 2, w0, 0) if d <= 64 else (None, None)
    if w is not None:
        assert it == it_ref == 2
        ok, err = _rows_close(w[:8], w_ref[:8], 1e-9 if dtype == np.float64 else 2e-4)
        assert ok, err
    else:
        # wide rows: one iteration per component (lim < 2 always), d passes in all
        w, it = pd.ica_def(x, 2.0, 1, w0, 0)
        w_ref, _ = oica.ica_def(np.ascontiguousarray(xw.T), 2.0, 1, w0.copy(), "logcosh")
        ok, err = _rows_close(w[:4], w_ref[:4], 1e-9 if dtype == np.float64 else 5e-4)
        assert ok, err


def test_deflation_wide_rows_use_generic_engines(pd):
    """d beyond the fused pass (f64: 512): the xb / nonlin / atb fallback gives the same iterates."""
    n, d = 3000, 520
    rng = np.random.default_rng(3)
    x = rng.laplace(size=(n, d)) / np.sqrt(2.0)
    w0 = rng.standard_normal((d, d))
    w, it = pd.ica_def(x, 2.0, 1, w0, 0)
    w_ref, _ = oica.ica_def(np.ascontiguousarray(x.T), 2.0, 1, w0.copy(), "logcosh")
    ok, err = _rows_close(w[:3], w_ref[:3], 1e-9)
    assert ok, err
