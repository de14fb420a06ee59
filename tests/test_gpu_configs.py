"""GPU parity tests at the shapes BASELINE.json's configs name (reduced rows where the oracle needs it),
every one through the C ABI (ctypes -> libpetal_b200.so), compared with the CPU oracle.

  c1  exact Pca f64 10000 x 100                      tests/test_gpu_parity.py::test_pca_c1_config
  c2  RandomizedPca f32  *  x 1024, k=64, q=4 / 7    test_c2_shape_rpca_f32_vs_oracle        (200 000 rows)
  c3  FastIca logcosh f32  *  x 64                   test_c3_shape_fastica_f32_vs_oracle     (200 000 rows)
  c4  exact Pca f64  *  x 4096                       test_c4_shape_pca_f64_vs_oracle         (20 000 rows)
  c5  RandomizedPca f32  *  x 256, k=32, q=4         test_c5_shape_rpca_f32_vs_oracle        (400 000 rows)

The c2 / c5 fits run the same kernel instantiations as the benchmark (panel-major Y, n_pad = 80 / 48,
precise chains, in-kernel B_lo splitter): the launch log is checked for them.
"""
import numpy as np
import pytest

import bench
from oracle import ica as oica
from oracle import pca as opca
from oracle.rng import Mcg128Xsl64
from tests import synth

pytestmark = pytest.mark.gpu

RNG_SEED = 1_234_567_891_011_121_314


@pytest.fixture(scope="module")
def pd():
    import petal_decomposition_b200 as pd_
    return pd_


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def _profiled_fit(pd, model, x):
    """Fits on the DEVICE-resident copy of x, like bench.py's `value` arm: these tests are about the kernel
    instantiations of the benchmarked pass sequence (a host-fed randomized PCA replaces the power iterations'
    passes by products with the Gram matrix taken during the ingest, see tests/test_gpu_streaming.py)."""
    import torch
    ctx = pd.default_context()
    ctx.set_profiling(True)
    ctx.profile()
    model.fit(torch.from_numpy(x).cuda())
    prof = ctx.profile()
    ctx.set_profiling(False)
    return prof


# ------------------------------------------------------------------------------------ c2 / c5
@pytest.mark.parametrize("q", [4, 7], ids=["c2_q4", "c2_q7"])
def test_c2_shape_rpca_f32_vs_oracle(pd, q):
    """configs[1] at 200 000 rows: d = 1024, k = 64 (l = 74 -> n_pad = 80), same synthetic family as bench.py.
    The only path on which tc_gemm_kernel<1,80,1,2> (panel X^T Y, precise) and the n_pad = 80 panel tc_xb run."""
    n, d, k = 200_000, 1024, 64
    x = bench.make_x_host(n, d, "f32", "rpca")
    omega = Mcg128Xsl64.from_seed_u128(RNG_SEED).normal_matrix(d, k + 10, np.float32)
    ref = opca.RandomizedPca(k, n_iter=q)
    ref.fit(x.astype(np.float64), omega.astype(np.float64))
    m = pd.RandomizedPcaBuilder.new(k).seed(RNG_SEED).n_power_iter(q).build()
    prof = _profiled_fit(pd, m, x)
    assert prof["tc_atb_f32"]["count"] == q + 1 and prof["tc_xb_f32"]["count"] == q + 1, prof
    sr = ref.singular_values()
    err = np.abs(m.singular_values().astype(np.float64) - sr) / sr
    assert err.max() < 1e-4, err
    assert rel(m.explained_variance_ratio(), ref.explained_variance_ratio()) < 2e-4
    assert abs(m._total_variance - ref.total_variance) < 1e-5 * ref.total_variance
    assert np.allclose(m.mean(), ref.means, atol=1e-5)
    # the 40 leading directions are separated from the noise floor (s_i = 10 * 0.9^i against noise 0.1)
    ang = opca.principal_angles(m.components()[:40], ref.components[:40])
    assert ang.max() < 5e-3, ang.max()
    # closed form of the synthetic spectrum: sigma_i^2 ~ n (s_i^2 + noise^2) for the separated part
    s_true = np.sqrt(n * (bench.spectrum(128)[:20] ** 2 + 0.01))
    assert rel(m.singular_values()[:20], s_true) < 2e-2


def test_c5_shape_rpca_f32_vs_oracle(pd):
    """configs[4] per-GPU shape at 400 000 rows: d = 256, k = 32 (l = 42 -> n_pad = 48), q = 4."""
    n, d, k, q = 400_000, 256, 32, 4
    x = bench.make_x_host(n, d, "f32", "rpca")
    omega = Mcg128Xsl64.from_seed_u128(RNG_SEED).normal_matrix(d, k + 10, np.float32)
    ref = opca.RandomizedPca(k, n_iter=q)
    ref.fit(x.astype(np.float64), omega.astype(np.float64))
    m = pd.RandomizedPcaBuilder.new(k).seed(RNG_SEED).n_power_iter(q).build()
    y = m.fit_transform(x)
    sr = ref.singular_values()
    err = np.abs(m.singular_values().astype(np.float64) - sr) / sr
    assert err.max() < 1e-4, err
    assert rel(m.explained_variance_ratio(), ref.explained_variance_ratio()) < 2e-4
    ang = opca.principal_angles(m.components()[:24], ref.components[:24])
    assert ang.max() < 5e-3, ang.max()
    # scores: fit_transform == transform on the fitted model (f32 accuracy), signs per svd_flip
    assert np.allclose(m.transform(x), y, atol=2e-3 * np.abs(y).max())


@pytest.mark.parametrize("d,q,gram", [(256, 4, True), (256, 7, True), (320, 4, True), (384, 4, False), (256, 1, False)])
def test_rpca_device_gram_route_vs_oracle(pd, d, q, gram):
    """X in HBM, f32: when the Gram matrix costs fewer passes over X than the power iterations it replaces (c5: d = 256,
    q = 4 -> 4 window passes), Z <- Xc^T (Xc B) = G B runs on the small side and only the last pair of products is
    streamed.  Launch counts say which route ran; the oracle says that both give the reference's result."""
    import torch
    n, k = 200_000, 32
    x = bench.make_x_host(n, d, "f32", "rpca")
    omega = Mcg128Xsl64.from_seed_u128(RNG_SEED).normal_matrix(d, k + 10, np.float32)
    ref = opca.RandomizedPca(k, n_iter=q)
    ref.fit(x.astype(np.float64), omega.astype(np.float64))
    xd = torch.from_numpy(x).cuda()
    ctx = pd.default_context()
    results = {}
    for on in (1, 0):
        ctx.set_host_gram(on)
        try:
            m = pd.RandomizedPcaBuilder.new(k).seed(RNG_SEED).n_power_iter(q).build()
            ctx.set_profiling(True)
            ctx.profile()
            m.fit(xd, omega)
            prof = ctx.profile()
        finally:
            ctx.set_profiling(False)
            ctx.set_host_gram(1)
        n_atb = sum(v["count"] for kn, v in prof.items() if kn.startswith("tc_atb_f32"))
        n_xb = sum(v["count"] for kn, v in prof.items() if kn.startswith("tc_xb_f32"))
        if on and gram:
            assert n_atb == -(-d // 64) + 1 and n_xb == 1, prof
        else:
            assert n_atb == q + 1 and n_xb == q + 1, prof
        sr = ref.singular_values()
        assert np.max(np.abs(m.singular_values().astype(np.float64) - sr) / sr) < 1e-4
        assert rel(m.explained_variance_ratio(), ref.explained_variance_ratio()) < 2e-4
        assert np.allclose(m.mean(), ref.means, atol=1e-5)
        assert abs(m._total_variance - ref.total_variance) < 1e-5 * ref.total_variance
        assert opca.principal_angles(m.components()[:20], ref.components[:20]).max() < 5e-3
        results[on] = m.singular_values().astype(np.float64)
    assert np.max(np.abs(results[1] - results[0]) / results[0]) < 2e-5


# ------------------------------------------------------------------------------------ c4
def test_c4_shape_pca_f64_vs_oracle(pd):
    """configs[3] at 20 000 rows: exact Pca f64, d = 4096 - CholeskyQR2 (two DMMA Gram passes, blocked Cholesky),
    R = R2 R1, block one-sided Jacobi SVD of R (4096 x 4096) - against the oracle's economy gesvd."""
    n, d, k = 20_000, 4096, 64
    import os
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c4_shape_20000x4096.npz"))
    x = bench.make_x_host(n, d, "f64", "pca")
    # the fixture is the oracle's result on exactly this matrix (tests/golden/make_c4_fixture.py; the economy gesvd of
    # 20000 x 4096 costs minutes of CPU time, so it is not recomputed on the GPU box)
    assert abs(float(np.sum(x[::997, ::13])) - float(fx["x_checksum"][0])) < 1e-6 * abs(float(fx["x_checksum"][0]))
    sref_all, tv_ref, comps_ref = fx["singular"], float(fx["total_variance"][0]), fx["components16"]
    m = pd.Pca.new(k)
    prof = _profiled_fit(pd, m, x)
    assert "jacobi_block" in prof and "cholesky_blocked" in prof, prof
    assert rel(m.singular_values(), sref_all[:k]) < 1e-10
    assert rel(m.explained_variance_ratio(), sref_all[:k] ** 2 / tv_ref) < 1e-10
    assert abs(m._total_variance - tv_ref) < 1e-10 * tv_ref
    assert np.allclose(m.mean(), fx["means"], atol=1e-12)
    cm = opca.sign_normalize_rows(m.components()[:16])
    cr = opca.sign_normalize_rows(comps_ref)
    assert np.max(np.abs(cm - cr)) < 1e-7
    # the whole spectrum (all 4096 singular values), not only the k reported ones
    mfull = pd.Pca.new(d)
    mfull.fit(x)
    sfull = mfull.singular_values()
    assert np.max(np.abs(sfull - sref_all)) < 1e-12 * sref_all[0]
    assert rel(sfull, sref_all) < 1e-10


def test_pca_f64_graded_spectrum_accuracy(pd):
    """Singular values spread over six orders of magnitude.  A backward-stable SVD (the reference's gesvd, the
    oracle) resolves sigma_j to ~eps * sigma_1 absolute; an eigen-decomposition of the Gram matrix only to
    eps * sigma_1^2 / sigma_j (1e-4 relative at sigma_j = 1e-6 sigma_1).  CholeskyQR2 + Jacobi SVD of R must be in the
    first class: every sigma_j >= 1e-6 sigma_1 within 2e-13 sigma_1 of the oracle's, and within 1e-10 relative down
    to 1e-3 sigma_1."""
    rng = np.random.default_rng(12)
    n, d = 6000, 160
    u, _ = np.linalg.qr(rng.standard_normal((n, d)))
    v, _ = np.linalg.qr(rng.standard_normal((d, d)))
    s = np.logspace(0, -6, d) * 50.0
    x = (u * s) @ v.T + rng.uniform(-1, 1, d)
    x -= x.mean(axis=0)  # keep the constructed spectrum: centring must not mix it
    ref = opca.Pca(d, economy=True)
    ref.fit(x)
    m = pd.Pca.new(d)
    m.fit(x)
    sg, sr = m.singular_values(), ref.singular_values()
    assert np.max(np.abs(sg - sr)) < 2e-13 * sr[0], np.max(np.abs(sg - sr)) / sr[0]
    big = sr > 1e-3 * sr[0]
    assert rel(sg[big], sr[big]) < 1e-10
    assert rel(sg, sr) < 1e-6          # eps * sigma_1 / sigma_j at the small end (the Gram route: 1e-4)
    cm = opca.sign_normalize_rows(m.components()[big])
    cr = opca.sign_normalize_rows(ref.components[big])
    assert np.max(np.abs(cm - cr)) < 1e-7


@pytest.mark.parametrize("m_,ln", [(300, 64), (2048, 2048)])
def test_block_jacobi_svd(pd, monkeypatch, m_, ln):
    """The block engine (pairs of 32-row blocks, DMMA Gram + 64 x 64 Jacobi + DMMA update) against LAPACK;
    PETAL_JACOBI_BLOCK_MIN lowers its threshold so that a small ragged case (m not a multiple of 64, m > len:
    rank-deficient rows) runs through it too."""
    rng = np.random.default_rng(m_ + ln)
    monkeypatch.setenv("PETAL_JACOBI_BLOCK_MIN", "256")
    if m_ == 300:
        a = rng.standard_normal((m_, ln)) @ np.diag(np.logspace(0, -3, ln))   # 300 x 64: padded to 320 rows, rank 64
    else:
        a = rng.standard_normal((m_, ln)) * np.logspace(0, -4, ln)
    u, s, vt = pd.small_svd(a)
    sref = np.linalg.svd(a, compute_uv=False)
    kk = len(sref)
    assert np.max(np.abs(s[:kk] - sref)) < 1e-12 * sref[0]
    assert np.allclose((u[:, :kk] * s[:kk]) @ vt[:kk], a, atol=1e-11 * sref[0])
    assert np.allclose(u.T @ u, np.eye(u.shape[0]), atol=1e-10)


# ------------------------------------------------------------------------------------ c3
def test_c3_shape_fastica_f32_vs_oracle(pd):
    """configs[2] at 200 000 rows: d = nc = 64 f32 - the one-pass tcgen05 kernel + the fused update kernel
    against oracle.FastIca (textbook symmetric decorrelation, row.row test) in f64."""
    n, d = 200_000, 64
    x, a = synth.mixed_sources(n, d, seed=1, dtype=np.float32)
    w_init = Mcg128Xsl64.from_seed_u128(RNG_SEED).normal_matrix(d, d, np.float32)
    ref = oica.FastIca(symdec="textbook", lim="rowrow")
    ref.fit(x.astype(np.float64), w_init.astype(np.float64))
    ica = pd.FastIca.with_seed(RNG_SEED)
    prof = _profiled_fit(pd, ica, x)
    assert "ica_fused_f32" in prof and prof["ica_fused_f32"]["count"] >= ica.n_iter, prof
    assert ica.n_iter < 200 and ref.n_iter < 200
    # the two runs start from sign-flipped whitening coordinates (eigenvector signs are not defined): trajectories
    # differ, the fixed point is the same up to the convergence tolerance (1e-4 on the rows of W)
    assert abs(ica.n_iter - ref.n_iter) <= 6, (ica.n_iter, ref.n_iter)
    _, defect = oica.match_rows(ica.components, ref.components)
    assert defect < 1e-4, defect
    am, amr = oica.amari_index(ica.components, a), oica.amari_index(ref.components, a)
    assert am < 0.05 and abs(am - amr) < 5e-3, (am, amr)
    s = ica.transform(x[:5000])
    sr = ref.transform(x[:5000].astype(np.float64))
    sm, _ = oica.match_rows(np.asarray(s, np.float64).T, sr.T)
    assert np.max(np.abs(sm - sr.T)) < 5e-2 * np.abs(sr).max()


@pytest.mark.parametrize("fun,name", [(0, "logcosh"), (1, "exp"), (2, "cube")])
def test_c3_one_pass_iterates_vs_oracle(pd, fun, name):
    """Fixed-point iterates of the one-pass tcgen05 kernel + fused update kernel (d = nc = 64, f32 data) against the
    oracle's ica_par on the SAME whitened f32 data and w_init: no whitening sign ambiguity is left, so W must agree
    step by step to f32 accuracy.  The first steps are ill-conditioned (Gd = E[g(Wx) x^T] - diag(E g') W is a small
    difference of O(1) terms: smallest singular value ~2e-4 here), so "f32 accuracy" is calibrated by the oracle
    itself run in float32: the GPU path must stay within a small factor of that deviation."""
    n, d = 120_000, 64
    x, _ = synth.mixed_sources(n, d, seed=3, dtype=np.float64)
    xc = (x - x.mean(axis=0)).T
    u, s, _ = np.linalg.svd(xc, full_matrices=False)
    x1t = np.ascontiguousarray((((u / s).T @ xc) * np.sqrt(n)).T.astype(np.float32))   # n x d, white
    w_init = Mcg128Xsl64.from_seed_u128(RNG_SEED).normal_matrix(d, d)
    ctx = pd.default_context()
    for iters in (1, 3):
        wr, _ = oica.ica_par(x1t.T.astype(np.float64), 0.0, iters, w_init, fun=name)
        w32, _ = oica.ica_par(np.ascontiguousarray(x1t.T), np.float32(0.0), iters, w_init.astype(np.float32), fun=name)
        dev32 = float(np.max(np.abs(w32 - wr)))
        ctx.set_profiling(True)
        ctx.profile()
        w, ni = pd.ica_par(x1t, 0.0, iters, w_init, fun=fun)
        prof = ctx.profile()
        ctx.set_profiling(False)
        assert "ica_fused_f32" in prof, prof
        assert ni == iters
        err = float(np.max(np.abs(w - wr)))
        assert err < max(5.0 * dev32, 1e-4), (iters, err, dev32)
        assert np.allclose(w @ w.T, np.eye(d), atol=1e-9)


# ------------------------------------------------------------------------------------ nonlinearities
def test_ref_logcosh_golden_on_gpu(pd):  # src/ica.rs:459-468
    x = np.array([[1.0, 2.0], [3.0, 4.0]])
    g, gp = pd.logcosh(x)
    want = np.array([[0.76159416, 0.96402758], [0.99505475, 0.99932930]])
    assert np.max(np.abs(g - want) / want) < 1e-8
    assert abs(gp[0] - 0.24531258) / 0.24531258 < 1e-6 and abs(gp[1] - 0.00560349) / 0.00560349 < 1e-6
    for engine in (0, 1):  # f32: the generic kernel and the one-pass kernel's epilogue function
        g32, gp32 = pd.logcosh(x.astype(np.float32), engine=engine)
        assert np.max(np.abs(g32 - want)) < 5e-7, (engine, g32)
        assert np.max(np.abs(gp32 - np.array([0.24531258, 0.00560349]))) < 1e-6, (engine, gp32)


def test_approx_tanh_bound(pd):
    """The one-pass kernel's tanh (ex2.approx + rcp.approx) against np.tanh over [-20, 20] and near 0."""
    u = np.concatenate([np.linspace(-20, 20, 400_001), np.linspace(-1e-3, 1e-3, 20_001),
                        np.array([0.0, -0.0, 1e-30, -1e-30, 88.0, -88.0, 1e10, -1e10])]).astype(np.float32)
    g, gp = pd.logcosh(u[None, :], engine=1)
    want = np.tanh(u.astype(np.float64))
    err = np.abs(g[0].astype(np.float64) - want)
    assert err.max() < 5e-7, (err.max(), u[np.argmax(err)])
    assert np.all(np.abs(g[0]) <= 1.0) and np.all(np.sign(g[0]) * np.sign(u) >= 0)
    assert abs(float(gp[0]) - float(np.mean(1 - want ** 2))) < 1e-6
    for fun, f in [(1, oica.exp_fun), (2, oica.cube_fun)]:
        uu = np.linspace(-6, 6, 100_001).astype(np.float32)
        for engine in (0, 1):
            g, gp = pd.logcosh(uu[None, :], fun=fun, engine=engine)
            gw, gpw = f(uu.astype(np.float64)[None, :])
            assert np.max(np.abs(g - gw)) < 2e-6 * max(1.0, np.abs(gw).max()), (fun, engine)
            assert abs(float(gp[0]) - float(gpw[0])) < 1e-5 * max(1.0, abs(float(gpw[0]))), (fun, engine)


@pytest.mark.parametrize("name,fun", [("logcosh", 0), ("exp", 1), ("cube", 2)])
@pytest.mark.parametrize("n,d", [(5000, 3), (20000, 6), (8000, 16)])
def test_ica_par_trajectory_all_contrast_functions(pd, name, fun, n, d):
    """f64 ica_par, same whitened input and w_init: after a fixed number of steps W must equal the oracle's
    (exp / cube restate sklearn's _exp / _cube, decomposition/_fastica.py:160-168)."""
    x, _ = synth.mixed_sources(n, d, seed=d + 1)
    xc = (x - x.mean(axis=0)).T
    u, s, _ = np.linalg.svd(xc, full_matrices=False)
    x1 = ((u / s).T @ xc) * np.sqrt(n)
    w_init = Mcg128Xsl64.from_seed_u128(RNG_SEED).normal_matrix(d, d)
    for iters in (1, 4):
        wr, nr = oica.ica_par(x1, 0.0, iters, w_init, fun=name)
        w, ni = pd.ica_par(np.ascontiguousarray(x1.T), 0.0, iters, w_init, fun=fun)
        assert ni == nr == iters
        assert np.allclose(w, wr, atol=1e-9), np.max(np.abs(w - wr))
    wr, nr = oica.ica_par(x1, 1e-4, 200, w_init, fun=name)
    w, ni = pd.ica_par(np.ascontiguousarray(x1.T), 1e-4, 200, w_init, fun=fun)
    if nr < 200:  # converged runs: same iteration count, same W
        assert ni == nr
        assert np.allclose(w, wr, atol=1e-8)


# ------------------------------------------------------------------------------------ inverse_transform
@pytest.mark.parametrize("n,d,k", [(70001, 1000, 64), (4096, 1024, 36), (5000, 130, 32), (3000, 64, 64)])
def test_inverse_transform_f32_vs_oracle(pd, n, d, k):
    """inverse_transform = Y C + mean (src/pca.rs:788-811) on the tcgen05 engine: 128-column output windows
    (d > 128: several windows, ragged last one), bias added in the epilogue."""
    rng = np.random.default_rng(n + d + k)
    y = (rng.standard_normal((n, k)) * np.linspace(5, 0.5, k)).astype(np.float32)
    comps, _ = np.linalg.qr(rng.standard_normal((d, k)))
    comps = np.ascontiguousarray(comps.T.astype(np.float32))
    mean = rng.uniform(-2, 2, d).astype(np.float32)
    ref = opca.inverse_transform(y.astype(np.float64), comps.astype(np.float64), mean.astype(np.float64), True)
    m = pd.Pca.new(k)
    m._components, m._means = comps, mean
    out = {}
    ctx = pd.default_context()
    for eng in (0, 1):
        ctx.set_f32_engine(eng)
        out[eng] = np.asarray(m.inverse_transform(y), np.float64)
    ctx.set_f32_engine(1)
    scale = np.abs(ref).max()
    for eng in (0, 1):
        assert out[eng].shape == (n, d)
        assert np.max(np.abs(out[eng] - ref)) < 2e-6 * scale * np.sqrt(k), eng
    # and without centering (no bias)
    m2 = pd.PcaBuilder.new(k).centering(False).build()
    m2._components, m2._means = comps, np.zeros(d, np.float32)
    ref2 = y.astype(np.float64) @ comps.astype(np.float64)
    assert np.max(np.abs(np.asarray(m2.inverse_transform(y), np.float64) - ref2)) < 2e-6 * scale * np.sqrt(k)


# ------------------------------------------------------------------------------------ error paths
def test_did_not_converge_is_reported(pd, monkeypatch):  # src/linalg.rs:84,115
    rng = np.random.default_rng(0)
    a = rng.standard_normal((48, 48))
    monkeypatch.setenv("PETAL_JACOBI_MAX_SWEEPS", "1")
    with pytest.raises(pd.LinalgError) as e:
        pd.small_svd(a)
    assert "did not converge" in str(e.value)
    big = rng.standard_normal((300, 300))   # cooperative multi-CTA engine
    with pytest.raises(pd.LinalgError) as e:
        pd.small_svd(big)
    assert "did not converge" in str(e.value)
    monkeypatch.delenv("PETAL_JACOBI_MAX_SWEEPS")
    u, s, vt = pd.small_svd(a)  # the context is usable again
    assert np.allclose(s, np.linalg.svd(a, compute_uv=False), atol=1e-11)


def test_unsupported_torch_inputs_rejected(pd):
    import torch
    x16 = torch.zeros((64, 8), dtype=torch.float16, device="cuda")
    with pytest.raises(pd.InvalidInput):
        pd.Pca.new(2).fit(x16)
    xi = torch.zeros((64, 8), dtype=torch.int32, device="cuda")
    with pytest.raises(pd.InvalidInput):
        pd.RandomizedPca.with_seed(2, 1).fit(xi)
    with pytest.raises(pd.InvalidInput):
        pd.Pca.new(2).fit(torch.zeros((64, 8), dtype=torch.float32))  # CPU tensor: numpy is the host path


def test_fastica_fewer_samples_than_features(pd):
    """n < d (nc = n): rank-deficient whitening; the fused update kernel stages K1 (nc x d) with a pitch that
    covers d > nc + 4 columns. Must return finite numbers or a LinalgError, never garbage."""
    rng = np.random.default_rng(5)
    x = rng.laplace(size=(20, 50))
    ica = pd.FastIca(pd.Pcg.from_seed(3), max_iter=20)
    try:
        ica.fit(x)
    except pd.LinalgError:
        return
    assert ica.components.shape == (20, 50) and np.isfinite(ica.components).all()
    x32 = rng.laplace(size=(30, 64)).astype(np.float32)
    ica = pd.FastIca(pd.Pcg.from_seed(3), max_iter=20)
    try:
        ica.fit(x32)
    except pd.LinalgError:
        return
    assert np.isfinite(ica.components).all()
