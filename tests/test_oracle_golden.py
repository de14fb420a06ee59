"""Pins the oracle (oracle/) against every golden vector the reference's own tests hold
for the path. Reference line numbers are cited per test."""
import numpy as np
import pytest

from oracle import ica as oica
from oracle import pca as opca
from oracle.rng import Mcg128Xsl64, ZIG_NORM_X

RNG_SEED = 1_234_567_891_011_121_314  # reference src/pca.rs:860, src/ica.rs:405
X3 = np.array([[0.0, 0.0], [3.0, 4.0], [6.0, 8.0]])
X6 = np.array([[-1.0, -1], [-2, -1], [-3, -2], [1, 1], [2, 1], [3, 2]])


# ---------------------------------------------------------------- exact PCA
def test_pca_zero_component():  # src/pca.rs:863-875
    p = opca.Pca(0)
    y = p.fit_transform(np.zeros((0, 5), dtype=np.float32))
    assert y.shape == (0, 0)
    y = p.fit_transform(X3.astype(np.float32))
    assert y.shape == (3, 0)


def test_pca_single_sample():  # src/pca.rs:878-883
    y = opca.Pca(1).fit_transform(np.array([[1.0, 1.0]], dtype=np.float32))
    assert y.shape == (1, 1) and y[0, 0] == 0.0


def test_pca():  # src/pca.rs:886-906
    p = opca.Pca(1)
    y = p.fit_transform(X3)
    assert abs(abs(y[0, 0]) - 5) < 1e-10 and abs(y[1, 0]) < 1e-10 and abs(abs(y[2, 0]) - 5) < 1e-10
    z = p.inverse_transform(y)
    assert np.allclose(z, X3, atol=1e-10, rtol=0)
    p = opca.Pca(1)
    p.fit(X3)
    # the sign golden [[-0.6,-0.8]] (pca.rs:901) is a 2-ulp tie in svd_flip:
    # pin magnitudes, accept either sign.
    assert np.allclose(np.abs(p.components), [[0.6, 0.8]], atol=1e-10, rtol=0)
    y = p.transform(X3)
    assert abs(abs(y[0, 0]) - 5) < 1e-10 and abs(y[1, 0]) < 1e-10 and abs(abs(y[2, 0]) - 5) < 1e-10


def test_pca_without_centering():  # src/pca.rs:909-916
    y = opca.Pca(1, centering=False).fit_transform(X3)
    assert abs(y[0, 0]) < 1e-10 and abs(y[1, 0] - 5) < 1e-10 and abs(abs(y[2, 0]) - 10) < 1e-10


def test_pca_explained_variance_ratio():  # src/pca.rs:919-933
    p = opca.Pca(2)
    p.fit(X6)
    r = p.explained_variance_ratio()
    assert r[0] > 0.99244 and r[1] < 0.00756


def test_readme_example():  # README.md:37-48
    x = np.array([[0.0, 0], [1, 1], [2, 2]])
    p = opca.Pca(2)
    p.fit(x)
    assert np.allclose(p.singular_values(), [2.0, 0.0], atol=1e-12)
    assert np.allclose(p.explained_variance_ratio(), [1.0, 0.0], atol=1e-12)
    y = p.transform(x)
    assert np.allclose(np.abs(y[:, 0]), [np.sqrt(2), 0, np.sqrt(2)], atol=1e-12)


def test_pca_doctest():  # src/pca.rs:27-35
    x = np.array([[0.0, 0], [1, 1], [2, 2]])
    y = opca.Pca(1).fit_transform(x)
    assert abs(abs(y[0, 0]) - np.sqrt(2)) < 1e-8 and abs(y[1, 0]) < 1e-8


def test_svd_flip():  # src/pca.rs:1044-1050
    u = np.array([[2.0, -1, 3], [-1, -3, 2]])
    v = np.array([[1.0, 1], [-2, 2], [3, -3]])
    opca.svd_flip(u, v)
    assert np.array_equal(u, [[2, 1, 3], [-1, 3, 2]])
    assert np.array_equal(v, [[1, 1], [2, -2], [3, -3]])


# ---------------------------------------------------------------- randomized PCA
def test_randomized_pca():  # src/pca.rs:950-970 (rank-1 data: independent of Omega)
    p = opca.RandomizedPca(1, rng=Mcg128Xsl64.from_seed_u128(RNG_SEED))
    p.fit(X3)
    y = p.transform(X3)
    assert abs(abs(y[0, 0]) - 5) < 1e-10 and abs(y[1, 0]) < 1e-10 and abs(abs(y[2, 0]) - 5) < 1e-10
    z = p.inverse_transform(y)
    assert np.allclose(z, X3, atol=1e-10, rtol=0)
    p = opca.RandomizedPca(1, rng=Mcg128Xsl64(987654321))
    y = p.fit_transform(X3)
    assert abs(abs(y[0, 0]) - 5) < 1e-10 and abs(y[1, 0]) < 1e-10 and abs(abs(y[2, 0]) - 5) < 1e-10


def test_randomized_pca_doctest():  # src/pca.rs:293-302
    x = np.array([[0.0, 0], [1, 1], [2, 2]])
    y = opca.RandomizedPca(1, rng=Mcg128Xsl64(5)).fit_transform(x)
    assert abs(abs(y[0, 0]) - np.sqrt(2)) < 1e-8 and abs(y[1, 0]) < 1e-8


def test_randomized_pca_explained_variance_ratio():  # src/pca.rs:973-987
    p = opca.RandomizedPca(2, rng=Mcg128Xsl64(42))
    p.fit(X6)
    r = p.explained_variance_ratio()
    assert r[0] > 0.99244 and r[1] < 0.00756


def _x_100x80():
    rng = Mcg128Xsl64(RNG_SEED)  # Pcg64Mcg::new(seed), src/pca.rs:991
    return rng, rng.normal_matrix(100, 80)


def test_randomized_vs_exact_equivalence():  # src/pca.rs:990-1027 (5 % relative)
    rng, x = _x_100x80()
    p = opca.Pca(2)
    pr = opca.RandomizedPca(2, rng=rng)
    p.fit(x)
    pr.fit(x)
    assert np.allclose(p.explained_variance_ratio(), pr.explained_variance_ratio(), rtol=0.05)
    assert np.allclose(p.singular_values(), pr.singular_values(), rtol=0.05)


# ---------------------------------------------------------------- FastICA
XI = np.array([[0.0, 0.0], [1.0, 1.0], [1.0, -1.0]])


@pytest.mark.parametrize("symdec,lim", [("textbook", "rowrow"), ("textbook", "rowcol"),
                                        ("literal", "rowcol")])
def test_fast_ica_fit_transform(symdec, lim):  # src/ica.rs:408-420
    ica = oica.FastIca(Mcg128Xsl64.from_seed_u128(RNG_SEED), symdec=symdec, lim=lim)
    ica.fit(XI)
    n_fit = ica.n_iter
    a = ica.transform(XI)
    ica2 = oica.FastIca(Mcg128Xsl64.from_seed_u128(RNG_SEED), symdec=symdec, lim=lim)
    b = ica2.fit_transform(XI)
    assert ica2.n_iter == n_fit
    assert np.allclose(a, b, atol=1e-12)
    if lim == "rowcol":
        # the reference asserts n_iter == 1 (ica.rs:412): holds for the literal
        # row-column convergence test with the restated RNG stream (1-bit pin)
        assert n_fit == 1


def test_w_init_stream_pin():  # SURVEY App. A probe values
    w = Mcg128Xsl64.from_seed_u128(RNG_SEED).normal_matrix(2, 2)
    assert np.allclose(w.ravel(), [1.00209109, 0.61455719, 0.91880942, -0.56393262], atol=5e-9)
    assert abs(ZIG_NORM_X[0] - 3.910757959537090045) < 1e-14
    assert abs(ZIG_NORM_X[2] - 3.449278298560964462) < 1e-14


@pytest.mark.parametrize("symdec,lim", [("textbook", "rowrow"), ("literal", "rowcol")])
def test_ica_par_single_iter(symdec, lim):  # src/ica.rs:435-444
    x = np.array([[-0.5, 0.5], [-0.3, 0.3]])
    w = np.array([[1.0, 2], [3, 4]])
    y, n = oica.ica_par(x, 0.5, 1, w, symdec, lim)
    assert np.allclose(y, [[0.51449576, -0.85749293], [-0.85749293, -0.51449576]], atol=1e-8, rtol=0)
    assert n == 1


@pytest.mark.parametrize("symdec,lim", [("textbook", "rowrow"), ("literal", "rowcol")])
def test_ica_par_multi_iter(symdec, lim):  # src/ica.rs:447-456
    x = np.array([[1.0, -1], [0, 0]])
    w = np.array([[1.0, 2], [3, 4]])
    y, n = oica.ica_par(x, 1e-4, 200, w, symdec, lim)
    assert np.allclose(y, [[-0.00172682, 0.99999851], [0.99999851, 0.00172682]], atol=1e-8, rtol=0)
    assert n == 6


def test_logcosh():  # src/ica.rs:459-468
    g, gp = oica.logcosh(np.array([[1.0, 2], [3, 4]]))
    assert np.allclose(g, [[0.76159416, 0.96402758], [0.99505475, 0.99932930]], rtol=1e-8)
    assert np.allclose(gp, [0.24531258, 0.00560349], rtol=1e-6)


@pytest.mark.parametrize("symdec", ["textbook", "literal"])
def test_symmetric_decorrelation(symdec):  # src/ica.rs:471-478
    w = oica.symmetric_decorrelation(np.array([[33.0, 24], [48, 57]]), symdec)
    assert np.allclose(w, [[0.96623494, -0.25766265], [0.25766265, 0.96623494]], rtol=1e-8)


def test_literal_symdec_breaks_for_d3():  # SURVEY F5 (documented, not a reference test)
    rng = np.random.default_rng(0)
    w = rng.standard_normal((3, 3))
    wt = oica.symmetric_decorrelation(w, "textbook")
    wl = oica.symmetric_decorrelation(w, "literal")
    assert np.allclose(wt @ wt.T, np.eye(3), atol=1e-10)
    assert not np.allclose(wl @ wl.T, np.eye(3), atol=1e-3)


# ------------------------------------------------------------------ deflation FastICA (extension)
def test_ica_def_matches_sklearn_golden_vectors():
    """oracle.ica.ica_def against the committed outputs of sklearn's own `_ica_def` (tests/golden/ica_deflation.json,
    generated by tests/golden/make_ica_deflation_fixture.py): same iterates, same iteration counts."""
    import json
    import os
    from oracle import ica as oica_
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ica_deflation.json")
    fx = json.load(open(path))
    assert fx["generator"] == "sklearn.decomposition._fastica._ica_def" and len(fx["cases"]) == 9
    for c in fx["cases"]:
        w, it = oica_.ica_def(np.array(c["x1"]), c["tol"], c["max_iter"], np.array(c["w_init"]), c["fun"])
        assert it == c["n_iter"], (c["fun"], it, c["n_iter"])
        assert np.allclose(w, np.array(c["w"]), atol=1e-13, rtol=0)
        assert np.allclose(w @ w.T, np.eye(w.shape[0]), atol=1e-10)  # Gram-Schmidt keeps the rows orthonormal
