"""GPU parity tests: the CUDA path (through the C ABI, via the Python host mirror) against the
oracle on the same seeded inputs, plus the reference's own unit tests replayed on the product
API.  Tolerances follow BASELINE.json north_star: 1e-10 relative (f64) / 1e-4 (f32) on singular
values and explained variance; components after sign normalisation; principal angles for the
randomized / ICA subspaces."""
import numpy as np
import pytest

from oracle import ica as oica
from oracle import pca as opca
from oracle.rng import Mcg128Xsl64
from tests import synth

pytestmark = pytest.mark.gpu

RNG_SEED = 1_234_567_891_011_121_314
X3 = np.array([[0.0, 0.0], [3.0, 4.0], [6.0, 8.0]])
X6 = np.array([[-1.0, -1], [-2, -1], [-3, -2], [1, 1], [2, 1], [3, 2]])


@pytest.fixture(scope="module")
def pd():
    import petal_decomposition_b200 as m
    return m


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


# ------------------------------------------------------------ the reference's own tests
def test_ref_pca_zero_component(pd):  # src/pca.rs:863-875
    pca = pd.PcaBuilder.new(0).build()
    y = pca.fit_transform(np.zeros((0, 5), dtype=np.float32))
    assert y.shape == (0, 0)
    y = pca.fit_transform(X3.astype(np.float32))
    assert y.shape == (3, 0)


def test_ref_pca_single_sample(pd):  # src/pca.rs:878-883
    y = pd.Pca.new(1).fit_transform(np.array([[1.0, 1.0]], dtype=np.float32))
    assert y.shape == (1, 1) and y[0, 0] == 0.0


def test_ref_pca(pd):  # src/pca.rs:886-906
    pca = pd.Pca.new(1)
    assert pca.n_components() == 1
    y = pca.fit_transform(X3)
    assert abs(abs(y[0, 0]) - 5) < 1e-10 and abs(y[1, 0]) < 1e-10 and abs(abs(y[2, 0]) - 5) < 1e-10
    z = pca.inverse_transform(y)
    assert np.allclose(z, X3, atol=1e-10, rtol=0)
    pca = pd.Pca.new(1)
    pca.fit(X3)
    assert pca.n_components() == 1
    # sign golden is a rounding tie in the reference: magnitudes only
    assert np.allclose(np.abs(pca.components()), [[0.6, 0.8]], atol=1e-10, rtol=0)
    y = pca.transform(X3)
    assert abs(abs(y[0, 0]) - 5) < 1e-10 and abs(y[1, 0]) < 1e-10 and abs(abs(y[2, 0]) - 5) < 1e-10


def test_ref_pca_without_centering(pd):  # src/pca.rs:909-916
    y = pd.PcaBuilder.new(1).centering(False).build().fit_transform(X3)
    assert abs(y[0, 0]) < 1e-10 and abs(y[1, 0] - 5) < 1e-10 and abs(abs(y[2, 0]) - 10) < 1e-10


def test_ref_pca_explained_variance_ratio(pd):  # src/pca.rs:919-933
    pca = pd.Pca.new(2)
    pca.fit(X6)
    r = pca.explained_variance_ratio()
    assert r[0] > 0.99244 and r[1] < 0.00756


def test_ref_readme_example(pd):  # README.md:37-48
    x = np.array([[0.0, 0], [1, 1], [2, 2]])
    pca = pd.PcaBuilder.new(2).build()
    pca.fit(x)
    assert np.allclose(pca.singular_values(), [2.0, 0.0], atol=1e-7)
    assert np.allclose(pca.explained_variance_ratio(), [1.0, 0.0], atol=1e-12)
    y = pca.transform(x)
    assert np.allclose(np.abs(y[:, 0]), [np.sqrt(2), 0, np.sqrt(2)], atol=1e-10)


def test_ref_randomized_pca(pd):  # src/pca.rs:950-970
    pca = pd.RandomizedPca.with_seed(1, RNG_SEED)
    assert pca.n_components() == 1
    pca.fit(X3)
    y = pca.transform(X3)
    assert abs(abs(y[0, 0]) - 5) < 1e-10 and abs(y[1, 0]) < 1e-10 and abs(abs(y[2, 0]) - 5) < 1e-10
    z = pca.inverse_transform(y)
    assert np.allclose(z, X3, atol=1e-10, rtol=0)
    pca = pd.RandomizedPca.with_rng(1, pd.Pcg.from_entropy())
    y = pca.fit_transform(X3)
    assert abs(abs(y[0, 0]) - 5) < 1e-10 and abs(y[1, 0]) < 1e-10 and abs(abs(y[2, 0]) - 5) < 1e-10


def test_ref_randomized_pca_explained_variance_ratio(pd):  # src/pca.rs:973-987
    pca = pd.RandomizedPca.with_rng(2, pd.Pcg.from_entropy())
    pca.fit(X6)
    r = pca.explained_variance_ratio()
    assert r[0] > 0.99244 and r[1] < 0.00756


def test_ref_randomized_vs_exact(pd):  # src/pca.rs:990-1027
    rng = pd.Pcg.new(RNG_SEED)
    x = rng.standard_normal((100, 80))
    pca, pca_rand = pd.Pca.new(2), pd.RandomizedPca.with_rng(2, rng)
    pca.fit(x)
    pca_rand.fit(x)
    assert np.allclose(pca.explained_variance_ratio(), pca_rand.explained_variance_ratio(), rtol=0.05)
    assert np.allclose(pca.singular_values(), pca_rand.singular_values(), rtol=0.05)


XI = np.array([[0.0, 0.0], [1.0, 1.0], [1.0, -1.0]])


@pytest.mark.parametrize("lim_variant", [0, 1])
def test_ref_fast_ica_fit_transform(pd, lim_variant):  # src/ica.rs:408-420
    ica = pd.FastIca.with_seed(RNG_SEED)
    ica.lim_variant = lim_variant
    ica.fit(XI)
    assert ica.n_iter == 1
    a = ica.transform(XI)
    ica = pd.FastIca.with_seed(RNG_SEED)
    ica.lim_variant = lim_variant
    b = ica.fit_transform(XI)
    assert ica.n_iter == 1
    assert np.allclose(a, b, atol=1e-14, rtol=0)


def test_ref_ica_par_single_iter(pd):  # src/ica.rs:435-444 (x is components x samples there)
    x = np.array([[-0.5, 0.5], [-0.3, 0.3]])
    w = np.array([[1.0, 2], [3, 4]])
    y, n = pd.ica_par(np.ascontiguousarray(x.T), 0.5, 1, w)
    assert np.allclose(y, [[0.51449576, -0.85749293], [-0.85749293, -0.51449576]], atol=1e-8, rtol=0)
    assert n == 1


@pytest.mark.parametrize("lim_variant", [0, 1])
def test_ref_ica_par_multi_iter(pd, lim_variant):  # src/ica.rs:447-456
    x = np.array([[1.0, -1], [0, 0]])
    w = np.array([[1.0, 2], [3, 4]])
    y, n = pd.ica_par(np.ascontiguousarray(x.T), 1e-4, 200, w, lim_variant=lim_variant)
    assert np.allclose(y, [[-0.00172682, 0.99999851], [0.99999851, 0.00172682]], atol=1e-8, rtol=0)
    assert n == 6


def test_ref_symmetric_decorrelation(pd):  # src/ica.rs:471-478
    w = pd.symmetric_decorrelation(np.array([[33.0, 24], [48, 57]]))
    assert np.allclose(w, [[0.96623494, -0.25766265], [0.25766265, 0.96623494]], rtol=1e-8)


def test_error_messages(pd):  # src/pca.rs:199-204,736-741,798-803; src/ica.rs:124-128
    with pytest.raises(pd.InvalidInput, match="every dimension should be at least 3"):
        pd.Pca.new(3).fit(X3)
    with pytest.raises(pd.InvalidInput, match="every dimension should be at least 3"):
        pd.RandomizedPca.with_seed(3, 1).fit(X3)
    pca = pd.Pca.new(1)
    pca.fit(X3)
    with pytest.raises(pd.InvalidInput, match="# of columns should be 2"):
        pca.transform(np.zeros((2, 3)))
    with pytest.raises(pd.InvalidInput, match="# of columns should be 1"):
        pca.inverse_transform(np.zeros((2, 2)))
    ica = pd.FastIca.with_seed(1)
    ica.fit(XI)
    with pytest.raises(pd.InvalidInput, match="too many columns"):
        ica.transform(np.zeros((2, 3)))


# ------------------------------------------------------------ small solver
@pytest.mark.parametrize("m,ln", [(1, 1), (2, 2), (5, 3), (7, 19), (64, 64), (74, 1024), (130, 130), (257, 300)])
def test_small_svd(pd, m, ln):  # (257, 300): persistent cooperative multi-CTA engine (does not fit shared memory)
    rng = np.random.default_rng(m * 1000 + ln)
    a = rng.standard_normal((m, ln)) * np.logspace(0, -3, m)[:, None]
    u, s, vt = pd.small_svd(a)
    s_ref = np.linalg.svd(a, compute_uv=False)
    r = min(m, ln)
    assert np.allclose(s[:r], s_ref[:r], rtol=1e-11, atol=1e-13 * s_ref[0])
    assert np.allclose(s[r:], 0, atol=1e-12 * s_ref[0])
    assert np.allclose(u @ u.T, np.eye(m), atol=1e-12)
    assert np.allclose((u * s) @ vt, a, atol=1e-12 * s_ref[0] * 10)
    assert np.allclose(vt[:r] @ vt[:r].T, np.eye(r), atol=1e-10)


def test_symmetric_decorrelation_random(pd):
    rng = np.random.default_rng(5)
    for m in (3, 8, 64):
        w = rng.standard_normal((m, m))
        got = pd.symmetric_decorrelation(w)
        assert np.allclose(got, oica.symmetric_decorrelation(w, "textbook"), atol=1e-10)
        assert np.allclose(got @ got.T, np.eye(m), atol=1e-12)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n,d", [(1, 1), (3, 2), (1000, 7), (4097, 100), (513, 260)])
def test_colmean_gram(pd, dtype, n, d):
    x = synth.gaussian(n, d, seed=n + d, dtype=dtype)
    mean, gram = pd.colmean_gram(x)
    x64 = x.astype(np.float64)
    mref = x.mean(axis=0, dtype=np.float64).astype(dtype).astype(np.float64)
    tol = 1e-12 if dtype == np.float64 else 1e-6
    assert np.allclose(mean, mref, rtol=tol, atol=tol)
    xc = (x - mean.astype(dtype)).astype(np.float64)
    gref = xc.T @ xc
    gtol = 1e-12 if dtype == np.float64 else 2e-5
    assert np.allclose(gram, gref, rtol=gtol, atol=gtol * np.abs(gref).max())
    del x64


# ------------------------------------------------------------ exact PCA vs oracle
def _check_pca(model, ref, x, tol_s, tol_c, k):
    s, sr = np.asarray(model.singular_values(), np.float64), np.asarray(ref.singular_values(), np.float64)
    big = sr > 1e-4 * sr[0]
    # relative tolerance on every singular value that is not numerically zero; the Gram route
    # resolves exact zeros (rank-deficient data) only to sqrt(eps) * sigma_max
    assert rel(s[big], sr[big]) < tol_s
    assert np.all(np.abs(s[~big] - sr[~big]) < 1e-7 * sr[0])
    ev, evr = model.explained_variance_ratio(), ref.explained_variance_ratio()
    assert rel(ev[big], evr[big]) < 2 * tol_s and np.all(np.abs(ev[~big] - evr[~big]) < 1e-13)
    assert abs(model._total_variance - ref.total_variance) < tol_s * ref.total_variance
    assert np.allclose(model.mean(), ref.means, rtol=tol_s, atol=tol_s)
    kk = int(np.sum(big))
    cm = opca.sign_normalize_rows(model.components()[:kk])
    cr = opca.sign_normalize_rows(ref.components[:kk])
    assert np.max(np.abs(cm - cr)) < tol_c


@pytest.mark.parametrize("n,d,k", [(10000, 100, 10), (2000, 37, 5), (300, 160, 8), (50, 64, 50)])
def test_pca_f64_vs_oracle(pd, n, d, k):
    x = synth.lowrank_noise(n, d, rank=min(d, 24), decay=0.8, seed=n + d)
    ref = opca.Pca(k, economy=True)
    yr = ref.fit_transform(x)
    m = pd.Pca.new(k)
    y = m.fit_transform(x)
    _check_pca(m, ref, x, 1e-10, 1e-7, k)
    # u-based svd_flip makes signs comparable directly (for the non-null components)
    kk = int(np.sum(ref.singular_values() > 1e-4 * ref.singular_values()[0]))
    assert np.allclose(m.components()[:kk], ref.components[:kk], atol=1e-7)
    assert np.allclose(y, yr, atol=1e-8 * np.abs(yr).max())
    assert np.allclose(m.transform(x), ref.transform(x), atol=1e-8 * np.abs(yr).max())
    z = m.inverse_transform(y)
    assert np.allclose(z, ref.inverse_transform(yr), atol=1e-8 * np.abs(x).max())


def test_pca_c1_config(pd):  # BASELINE.json configs[0]
    x = synth.gaussian(10000, 100, seed=1)
    ref = opca.Pca(10, economy=True)
    ref.fit(x)
    m = pd.Pca.new(10)
    m.fit(x)
    _check_pca(m, ref, x, 1e-10, 1e-6, 10)


@pytest.mark.parametrize("n,d,k", [(5000, 128, 12), (1001, 33, 4)])
def test_pca_f32_vs_oracle(pd, n, d, k):
    x = synth.lowrank_noise(n, d, rank=min(d, 24), decay=0.8, seed=7, dtype=np.float32)
    ref = opca.Pca(k, economy=True)
    ref.fit(x.astype(np.float64))
    m = pd.Pca.new(k)
    y = m.fit_transform(x)
    assert y.dtype == np.float32 and m.components().dtype == np.float32
    assert rel(m.singular_values(), ref.singular_values()) < 1e-4
    assert rel(m.explained_variance_ratio(), ref.explained_variance_ratio()) < 1e-4
    ang = opca.principal_angles(m.components(), ref.components)
    assert ang.max() < 2e-3


def test_pca_no_centering_vs_oracle(pd):
    x = synth.lowrank_noise(3000, 40, rank=10, seed=3)
    ref = opca.Pca(4, centering=False, economy=True)
    ref.fit(x)
    m = pd.PcaBuilder.new(4).centering(False).build()
    m.fit(x)
    assert rel(m.singular_values(), ref.singular_values()) < 1e-10
    assert np.all(m.mean() == 0)


# ------------------------------------------------------------ randomized PCA vs oracle
@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-9), (np.float32, 1e-4)])
@pytest.mark.parametrize("n,d,k,q", [(20000, 256, 16, 4), (3000, 75, 8, 7), (2, 5, 1, 2), (40, 300, 5, 3)])
def test_rpca_vs_oracle(pd, dtype, tol, n, d, k, q):
    x = synth.lowrank_noise(n, d, rank=min(d, 40), decay=0.8, noise=0.01, seed=11, dtype=dtype)
    omega = Mcg128Xsl64.from_seed_u128(RNG_SEED).normal_matrix(d, k + 10, dtype)
    ref = opca.RandomizedPca(k, n_iter=q)
    yr = ref.fit_transform(x.astype(np.float64), omega.astype(np.float64))
    m = pd.RandomizedPcaBuilder.new(k).seed(RNG_SEED).n_power_iter(q).build()
    y = m.fit_transform(x)
    # the product drew the same Omega from its own RNG
    assert rel(m.singular_values(), ref.singular_values()) < tol
    assert rel(m.explained_variance_ratio(), ref.explained_variance_ratio()) < tol
    ang = opca.principal_angles(m.components(), ref.components)
    assert ang.max() < (1e-6 if dtype == np.float64 else 5e-3)
    ctol = 1e-6 if dtype == np.float64 else 5e-3
    assert np.max(np.abs(opca.sign_normalize_rows(m.components()) - opca.sign_normalize_rows(ref.components))) < ctol
    assert np.allclose(np.abs(y), np.abs(yr), atol=ctol * np.abs(yr).max() * 10)
    # second fit on the same model draws a different Omega (rng state advances, src/pca.rs:532)
    s1 = m.rng.state()
    m.fit(x)
    assert m.rng.state() != s1
    assert rel(m.singular_values(), ref.singular_values()) < max(tol, 1e-6)


def test_rpca_f32_wide_sketch(pd):
    """k + 10 > 80 columns: the X^T Y kernel has no room for its chain-cutting accumulator buffers there and runs
    the long-chain mode; singular values must still be inside the f32 tolerance on a well-separated spectrum."""
    n, d, k, q = 30_011, 512, 100, 2
    x = synth.lowrank_noise(n, d, rank=300, decay=0.97, noise=1e-3, seed=4, dtype=np.float32)
    omega = Mcg128Xsl64.from_seed_u128(RNG_SEED).normal_matrix(d, k + 10, np.float32)
    ref = opca.RandomizedPca(k, n_iter=q)
    ref.fit(x.astype(np.float64), omega.astype(np.float64))
    m = pd.RandomizedPcaBuilder.new(k).seed(RNG_SEED).n_power_iter(q).build()
    m.fit(x)
    assert rel(m.singular_values(), ref.singular_values()) < 1e-4
    assert rel(m.explained_variance_ratio(), ref.explained_variance_ratio()) < 1e-4
    assert opca.principal_angles(m.components()[:50], ref.components[:50]).max() < 5e-3


def test_rpca_f32_noise_floor_accuracy(pd):
    """Trailing components next to the noise floor are the ones a biased accumulation damages first: the
    tcgen05 engine must keep every singular value inside the 1e-4 tolerance there (the tensor core adds into
    its fp32 accumulator with truncation; the X^T Y passes cut their TMEM chains every 4 K blocks for this)."""
    n, d, k, q = 100_000, 256, 32, 4
    x = synth.lowrank_noise(n, d, rank=40, decay=0.8, noise=0.01, seed=11, dtype=np.float32)
    omega = Mcg128Xsl64.from_seed_u128(RNG_SEED).normal_matrix(d, k + 10, np.float32)
    ref = opca.RandomizedPca(k, n_iter=q)
    ref.fit(x.astype(np.float64), omega.astype(np.float64))
    m = pd.RandomizedPcaBuilder.new(k).seed(RNG_SEED).n_power_iter(q).build()
    m.fit(x)
    sr = ref.singular_values()
    err = np.abs(m.singular_values().astype(np.float64) - sr) / sr
    assert err.max() < 1e-4, err
    assert np.median(err) < 5e-6, err
    assert rel(m.explained_variance_ratio(), ref.explained_variance_ratio()) < 1e-4


def test_rpca_roundtrip_f32(pd):
    x = synth.lowrank_noise(5000, 64, rank=8, noise=0.0, seed=2, dtype=np.float32)
    m = pd.RandomizedPca.with_seed(8, 99)
    y = m.fit_transform(x)
    z = m.inverse_transform(y)
    assert np.allclose(z, x, atol=2e-3)
    assert np.allclose(m.transform(x), y, atol=2e-3 * np.abs(y).max())


# ------------------------------------------------------------ FastICA vs oracle
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,d", [(20000, 4), (30000, 8)])
def test_fastica_vs_oracle(pd, dtype, n, d):
    x, a = synth.mixed_sources(n, d, seed=d, dtype=dtype)
    w_init = Mcg128Xsl64.from_seed_u128(RNG_SEED).normal_matrix(d, d, dtype)
    ref = oica.FastIca(symdec="textbook", lim="rowrow")
    sr = ref.fit_transform(x.astype(np.float64), w_init.astype(np.float64))
    ica = pd.FastIca.with_seed(RNG_SEED)
    s = ica.fit_transform(x)
    # The whitening eigenvectors are defined up to sign (LAPACK backends differ among themselves),
    # so the two runs start from sign-flipped coordinates: trajectories differ, the fixed point is
    # the same up to the convergence tolerance (tol = 1e-4 on the rows of W).
    assert ica.n_iter < 200 and abs(ica.n_iter - ref.n_iter) <= 4
    matched, defect = oica.match_rows(ica.components, ref.components)
    assert defect < 1e-6  # 1 - |cos| of matched unmixing rows
    assert oica.amari_index(ica.components, a) < 0.05
    assert np.allclose(ica.transform(x), s, atol=1e-10 if dtype == np.float64 else 1e-3)
    assert np.allclose(np.asarray(s).std(axis=0), np.asarray(sr).std(axis=0)[0], rtol=0.05)


@pytest.mark.parametrize("n,d", [(5000, 3), (20000, 6), (8000, 16)])
def test_ica_par_trajectory_vs_oracle(pd, n, d):
    """Same whitened input, same w_init: the fixed-point trajectory must agree with the oracle
    iteration by iteration (identical n_iter, W to 1e-9)."""
    x, _ = synth.mixed_sources(n, d, seed=d + 1)
    xc = (x - x.mean(axis=0)).T
    u, s, _ = np.linalg.svd(xc, full_matrices=False)
    x1 = ((u / s).T @ xc) * np.sqrt(n)
    w_init = Mcg128Xsl64.from_seed_u128(RNG_SEED).normal_matrix(d, d)
    wr, nr = oica.ica_par(x1, 1e-4, 200, w_init)
    w, ni = pd.ica_par(np.ascontiguousarray(x1.T), 1e-4, 200, w_init)
    assert ni == nr
    assert np.allclose(w, wr, atol=1e-9)
    for fun, name in [(pd.EXP, "exp"), (pd.CUBE, "cube")]:
        w2, n2 = pd.ica_par(np.ascontiguousarray(x1.T), 1e-4, 200, w_init, fun=fun)
        assert np.allclose(w2 @ w2.T, np.eye(d), atol=1e-10) and n2 <= 200


def test_fastica_tight_tolerance_matches_oracle(pd):
    """Converged to 1e-12 both runs reach the same fixed point whatever the starting signs."""
    x, a = synth.mixed_sources(20000, 5, seed=9)
    w_init = Mcg128Xsl64.from_seed_u128(RNG_SEED).normal_matrix(5, 5)
    ref = oica.FastIca(tol=1e-12, max_iter=1000)
    ref.fit(x, w_init)
    ica = pd.FastIca(pd.Pcg.from_seed(RNG_SEED), tol=1e-12, max_iter=1000)
    ica.fit(x)
    matched, defect = oica.match_rows(ica.components, ref.components)
    assert defect < 1e-12
    assert np.allclose(matched, ref.components, atol=1e-7 * np.abs(ref.components).max())


def test_fastica_d3_literal_reference_diverges(pd):
    """for d >= 3 the reference's literal code does not orthogonalise; the product
    implements the textbook algorithm and converges where the literal restatement does not."""
    x, a = synth.mixed_sources(20000, 3, seed=3)
    w_init = Mcg128Xsl64.from_seed_u128(RNG_SEED).normal_matrix(3, 3)
    ica = pd.FastIca.with_seed(RNG_SEED)
    ica.fit(x)
    assert ica.n_iter < 200
    assert oica.amari_index(ica.components, a) < 0.05
    lit = oica.FastIca(symdec="literal", lim="rowcol")
    with np.errstate(all="ignore"):
        try:
            lit.fit(x, w_init)
            bad = (not np.all(np.isfinite(lit.components))) or lit.n_iter == 200
        except Exception:
            bad = True
    assert bad


# ------------------------------------------------------------ device tensors / edge shapes
def test_torch_device_io(pd):
    import torch
    x = synth.lowrank_noise(4000, 48, rank=6, seed=4, dtype=np.float32)
    xd = torch.from_numpy(x).cuda()
    m = pd.Pca.new(3)
    y = m.fit_transform(xd)
    assert isinstance(y, torch.Tensor) and y.is_cuda and y.shape == (4000, 3)
    m2 = pd.Pca.new(3)
    y2 = m2.fit_transform(x)
    torch.cuda.synchronize()
    assert np.allclose(y.cpu().numpy(), y2, atol=1e-4 * np.abs(y2).max())
    z = m.inverse_transform(y)
    assert z.is_cuda and z.shape == (4000, 48)


def test_non_contiguous_rejected(pd):
    x = np.asfortranarray(synth.gaussian(10, 4))
    with pytest.raises(pd.InvalidInput):
        pd.Pca.new(1).fit(x)


def test_launch_counter(pd):
    ctx = pd.default_context()
    before = ctx.launch_count()
    pd.Pca.new(1).fit(X3)
    assert ctx.launch_count() > before


# ------------------------------------------------------------ tcgen05 engine vs SIMT engine vs numpy
@pytest.mark.parametrize("n,d,k", [(5000, 96, 7), (777, 100, 16), (4096, 1024, 74), (3001, 36, 1),
                                   (2048, 260, 128), (1500, 64, 33)])
def test_tc_xb_engines(pd, n, d, k):
    """transform = (x - mean) C^T through the tcgen05 3xTF32 kernel and through the FFMA kernel."""
    rng = np.random.default_rng(n + d + k)
    x = (rng.standard_normal((n, d)) * 3 + rng.uniform(-2, 2, d)).astype(np.float32)
    comps = rng.standard_normal((k, d)).astype(np.float32)
    mean = x.mean(axis=0).astype(np.float32)
    ref = (x.astype(np.float64) - mean.astype(np.float64)) @ comps.astype(np.float64).T
    m = pd.Pca.new(k)
    m._components, m._means = comps, mean
    ctx = pd.default_context()
    outs = {}
    for eng in (0, 1):
        ctx.set_f32_engine(eng)
        outs[eng] = np.asarray(m.transform(x), np.float64)
    ctx.set_f32_engine(1)
    scale = np.abs(ref).max()
    assert np.max(np.abs(outs[0] - ref)) < 2e-6 * scale * np.sqrt(d)
    assert np.max(np.abs(outs[1] - ref)) < 2e-6 * scale * np.sqrt(d)


@pytest.mark.parametrize("n,d,l", [(5000, 96, 7), (1030, 100, 16), (20000, 1024, 74), (3001, 36, 1),
                                   (2048, 260, 128), (9000, 512, 42)])
def test_tc_atb_engines(pd, n, d, l):
    rng = np.random.default_rng(n + d + l)
    x = (rng.standard_normal((n, d)) * 2 + rng.uniform(-1, 1, d)).astype(np.float32)
    l4 = ((l + 3) // 4) * 4
    y = np.zeros((n, l4), dtype=np.float32)
    y[:, :l] = rng.standard_normal((n, l)).astype(np.float32)
    mean = x.mean(axis=0).astype(np.float32)
    ref = (x.astype(np.float64) - mean.astype(np.float64)).T @ y.astype(np.float64)
    ctx = pd.default_context()
    outs = {}
    for eng in (0, 1):
        ctx.set_f32_engine(eng)
        outs[eng] = pd.xty(x, y, mean)
    ctx.set_f32_engine(1)
    scale = np.abs(ref).max() + np.sqrt(n)
    assert np.max(np.abs(outs[0] - ref)) < 1e-5 * scale
    assert np.max(np.abs(outs[1] - ref)) < 1e-5 * scale


def test_tc_atb_precise_vs_long_chains(pd, monkeypatch):
    """The tensor core adds into its TMEM accumulator with truncation: the default (precise) mode cuts the
    accumulation chains and must be an order of magnitude closer to the f64 product than 1024-row chains."""
    rng = np.random.default_rng(7)
    n, d, l = 60000, 256, 74
    x = (rng.standard_normal((n, d)) + 0.5).astype(np.float32)       # positive-ish partial sums: worst case for the bias
    y = np.zeros((n, 76), dtype=np.float32)
    y[:, :l] = (rng.standard_normal((n, l)) + 0.5).astype(np.float32)
    ref = x.astype(np.float64).T @ y.astype(np.float64)
    errs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("PETAL_TC_PRECISE", mode)
        got = pd.xty(x, y, None)
        errs[mode] = np.max(np.abs(got - ref) / np.abs(ref).max())
    monkeypatch.delenv("PETAL_TC_PRECISE")
    assert errs["1"] < 2e-6, errs
    assert errs["1"] < errs["0"], errs


@pytest.mark.parametrize("fun", [0, 1, 2])  # logcosh, exp, cube
@pytest.mark.parametrize("n,d", [(5_000, 8), (20_000, 64), (50_000, 12), (40_001, 64)])  # 1, 2, 3 tiles per CTA; ragged last tile
def test_fastica_one_pass_kernel_matches_three_kernel_path(pd, monkeypatch, fun, n, d):
    """The fused tcgen05 pass (U, g(U), sum g', H^T in one read of X) against tc_xb + nonlin + tc_atb."""
    x, _ = synth.mixed_sources(n, d, seed=d + fun, dtype=np.float32)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("PETAL_ICA_ONEPASS", mode)
        ica = pd.FastIcaBuilder.new().seed(RNG_SEED).fun(fun).build()
        ica.fit(x)
        res[mode] = (np.asarray(ica.components, np.float64), ica.n_iter)
    monkeypatch.delenv("PETAL_ICA_ONEPASS")
    (c1, it1), (c0, it0) = res["1"], res["0"]
    if it1 < 200 and it0 < 200:
        assert abs(it1 - it0) <= 1
        _, defect = oica.match_rows(c1, c0)
        assert defect < 1e-6
    # a run that does not converge (cube on sub-gaussian mixtures can cycle) must at least stay orthonormal in the
    # whitened coordinates on both paths: checked through finite outputs here
    assert np.isfinite(c1).all() and np.isfinite(c0).all()


def test_two_gpu_row_sharding(pd):
    """Row-sharded collective fits on 2 GPUs (NCCL inside the library) against the oracle."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517",
                        os.path.join(root, "tests", "dist_gpu_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("n,d", [(1000, 64), (4097, 100), (3000, 260), (513, 129)])
def test_dmma_gram_engines(pd, n, d):
    """Centred f64 Gram through the DMMA (mma.sync f64) kernel and through the DFMA kernel."""
    x = synth.gaussian(n, d, seed=n + d)
    ctx = pd.default_context()
    outs = {}
    for eng in (0, 1):
        ctx.set_f64_engine(eng)
        outs[eng] = pd.colmean_gram(x)[1]
    ctx.set_f64_engine(1)
    mean = x.mean(axis=0)
    ref = (x - mean).T @ (x - mean)
    for eng in (0, 1):
        assert np.allclose(outs[eng], ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
    y = synth.gaussian(n, 96, seed=3)
    for eng in (0, 1):
        ctx.set_f64_engine(eng)
        z = pd.xty(x, y, mean)
        assert np.allclose(z, (x - mean).T @ y, rtol=1e-12, atol=1e-11 * n)
    ctx.set_f64_engine(1)


# ------------------------------------------------------------ serde (reference tests with feature "serde")
def test_ref_pca_serialize(pd):  # src/pca.rs:935-947
    pca = pd.Pca.new(1)
    x = np.array([[1.0, 1.0]], dtype=np.float32)
    pca.fit(x)
    back = pd.Pca.from_json(pca.to_json(), np.float32)
    assert np.allclose(back.components(), pca.components(), atol=1e-12)
    assert np.allclose(back.mean(), pca.mean())
    assert np.allclose(back.transform(x), pca.transform(x))


def test_ref_randomized_pca_serialize(pd):  # src/pca.rs:1029-1041 (read back as a Pca, like the reference does)
    pca = pd.RandomizedPca.with_seed(1, RNG_SEED)
    x = np.array([[1.0, 1.0]], dtype=np.float32)
    pca.fit(x)
    back = pd.Pca.from_json(pca.to_json(), np.float32)
    assert np.allclose(back.components(), pca.components(), atol=1e-12)
    assert np.allclose(back.mean(), pca.mean())
    again = pd.RandomizedPca.from_json(pca.to_json(), np.float32)
    assert again.rng.state() == pca.rng.state()


def test_ref_fast_ica_serialize(pd):  # src/ica.rs:422-432
    x = np.array([[0.0, 0.0], [1.0, 1.0], [1.0, -1.0]])
    ica = pd.FastIca.new()
    ica.fit(x)
    back = pd.FastIca.from_json(ica.to_json())
    assert np.allclose(back.components, ica.components, atol=1e-12)
    assert np.allclose(back.means, ica.means)
    assert np.allclose(back.transform(x), ica.transform(x))
