"""Out-of-core / streaming ingest: a host X reaches HBM in row chunks on a copy stream and the
fit consumes the chunks as they land - kept resident when X fits, re-streamed through a two-slot ring when it does
not.  Every mode must give the fit of the device-resident X (same kernels, same order of the rows; only the f64
accumulation order of the partial sums changes) and, through it, the oracle's.  Chunk sizes are forced small so
that each traversal really has several chunks, a short tail that is absorbed, and slot reuse in the ring."""
import numpy as np
import pytest

from oracle import ica as oica
from oracle import pca as opca
from oracle.rng import Mcg128Xsl64
from tests import synth

pytestmark = pytest.mark.gpu

RNG_SEED = 1_234_567_891_011_121_314
RESIDENT, RING = 1, 2


@pytest.fixture(scope="module")
def pd():
    import petal_decomposition_b200 as m
    return m


@pytest.fixture()
def staging(pd):
    ctx = pd.default_context()

    def set_(mode, chunk_bytes, gram=0):
        ctx.set_host_staging(mode, chunk_bytes)
        ctx.set_host_gram(gram)  # 0: the plain pass sequence, whose trip counts the tests below assert
        return ctx

    yield set_
    ctx.set_host_staging(0, 1 << 30)
    ctx.set_host_gram(1)


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def _dev(x):
    import torch
    return torch.from_numpy(x).cuda()


def _signed_close(a, b, atol):
    s = np.sign(np.sum(a * b, axis=1, keepdims=True))
    return np.allclose(a, s * b, atol=atol, rtol=0)


@pytest.mark.parametrize("mode", [RESIDENT, RING], ids=["resident", "ring"])
@pytest.mark.parametrize("n,d,k,q", [(60_000, 256, 16, 3), (3 * 8192 + 500, 128, 20, 2), (40_000, 256, 22, 0)])
def test_rpca_f32_host_chunks_match_device_fit(pd, staging, mode, n, d, k, q):
    """panel-major tcgen05 path (folded mean for q >= 1; k + 10 = 32 leaves no padding column: separate mean pass)."""
    x = synth.lowrank_noise(n, d, rank=40, seed=5, dtype=np.float32)
    omega = Mcg128Xsl64.from_seed_u128(RNG_SEED).normal_matrix(d, k + 10, np.float32)
    base = pd.RandomizedPcaBuilder.new(k).seed(RNG_SEED).n_power_iter(q).build()
    yb = base.fit_transform(_dev(x), omega).cpu().numpy()
    ctx = staging(mode, 8192 * d * 4)  # 8192-row chunks
    m = pd.RandomizedPcaBuilder.new(k).seed(RNG_SEED).n_power_iter(q).build()
    y = m.fit_transform(x, omega)
    st = ctx.host_stream_stats()
    assert st["out_of_core"] == (mode == RING)
    folded = q >= 1 and (k + 10) % 16 != 0
    trips = (q + 1) + (0 if folded else 1)
    if mode == RING:
        assert st["traversals"] == trips
        head = min(n, 8192) * d * 4 if folded else 0
        assert st["h2d_bytes"] == trips * x.nbytes + head
    else:
        assert st["h2d_bytes"] <= x.nbytes + 8192 * d * 4
    assert rel(m.singular_values(), base.singular_values()) < 2e-6
    assert np.allclose(m.mean(), base.mean(), atol=1e-6)
    assert abs(m._total_variance - base._total_variance) < 1e-5 * base._total_variance
    assert opca.principal_angles(m.components()[: k // 2], base.components()[: k // 2]).max() < 2e-4
    assert np.allclose(np.abs(y), np.abs(yb), atol=2e-3 * np.abs(yb).max())
    # and the oracle
    ref = opca.RandomizedPca(k, n_iter=q)
    ref.fit(x.astype(np.float64), omega.astype(np.float64))
    assert rel(m.singular_values()[: k // 2], ref.singular_values()[: k // 2]) < 1e-4


@pytest.mark.parametrize("mode", [RESIDENT, RING], ids=["resident", "ring"])
def test_rpca_f64_host_chunks_vs_oracle(pd, staging, mode):
    """row-major (SIMT / DMMA) path: q + 3 traversals (mean; q + 1 products; C' = Xc^T Y1)."""
    n, d, k, q = 30_000, 96, 12, 2
    x = synth.lowrank_noise(n, d, rank=30, seed=6)
    omega = Mcg128Xsl64.from_seed_u128(RNG_SEED).normal_matrix(d, k + 10, np.float64)
    ref = opca.RandomizedPca(k, n_iter=q)
    ref.fit(x, omega)
    ctx = staging(mode, 4096 * d * 8)
    m = pd.RandomizedPcaBuilder.new(k).seed(RNG_SEED).n_power_iter(q).build()
    m.fit(x, omega)
    assert ctx.host_stream_stats()["out_of_core"] == (mode == RING)
    if mode == RING:
        assert ctx.host_stream_stats()["traversals"] == q + 3
    assert rel(m.singular_values(), ref.singular_values()) < 1e-9
    assert np.allclose(m.mean(), ref.means, atol=1e-12)
    assert opca.principal_angles(m.components(), ref.components).max() < 1e-6


@pytest.mark.parametrize("mode", [RESIDENT, RING], ids=["resident", "ring"])
@pytest.mark.parametrize("dtype,n,d,k,tol", [(np.float64, 30_000, 128, 10, 1e-10), (np.float64, 5 * 4096 + 77, 40, 40, 1e-10),
                                             (np.float32, 50_000, 64, 8, 1e-4)])
def test_pca_host_chunks_vs_oracle(pd, staging, mode, dtype, n, d, k, tol):
    """exact PCA: one trip for (provisional mean, column sums, Gram) with the rank-one shift of the Gram matrix, one
    for the second CholeskyQR2 pass (f64), one for the scores."""
    x = synth.lowrank_noise(n, d, rank=min(d, 32), seed=7, dtype=dtype)
    ref = opca.Pca(k, economy=True)
    yr = ref.fit_transform(x.astype(np.float64))
    ctx = staging(mode, 4096 * d * x.itemsize)
    m = pd.Pca.new(k)
    y = m.fit_transform(x)
    st = ctx.host_stream_stats()
    if mode == RING:
        assert st["traversals"] == (3 if dtype == np.float64 else 2), st
        assert st["h2d_bytes"] == st["traversals"] * x.nbytes + min(n, 8192) * d * x.itemsize
    assert rel(m.singular_values(), ref.singular_values()) < tol
    assert rel(m.explained_variance_ratio(), ref.explained_variance_ratio()) < 2 * tol
    assert np.allclose(m.mean(), ref.means, atol=1e-6 if dtype == np.float32 else 1e-13)
    atol = (1e-4 if dtype == np.float32 else 1e-8)
    assert _signed_close(m.components(), ref.components, atol)
    assert np.allclose(y, yr, atol=atol * np.abs(yr).max() * 10)


@pytest.mark.parametrize("mode", [RESIDENT, RING], ids=["resident", "ring"])
@pytest.mark.parametrize("dtype,d", [(np.float64, 8), (np.float32, 8), (np.float32, 64)])
def test_fastica_host_chunks_match_device_fit(pd, staging, mode, dtype, d):
    """FastICA: whitening statistics in one trip, then (ring) one trip per fixed-point iteration through the generic
    three-kernel pass; d = 64 f32 takes the one-pass tcgen05 kernel when X is resident."""
    n = 40_000
    x, _ = synth.mixed_sources(n, d, seed=3, dtype=dtype)
    w0 = Mcg128Xsl64.from_seed_u128(RNG_SEED).normal_matrix(d, d, dtype)
    base = pd.FastIca.with_seed(RNG_SEED)
    sb = base.fit_transform(_dev(x), w0).cpu().numpy()
    ctx = staging(mode, 4096 * d * x.itemsize)
    m = pd.FastIca.with_seed(RNG_SEED)
    s = m.fit_transform(x, w0)
    st = ctx.host_stream_stats()
    assert st["out_of_core"] == (mode == RING)
    assert abs(m.n_iter - base.n_iter) <= (0 if dtype == np.float64 else 2)
    if mode == RING:
        assert st["traversals"] >= 1 + m.n_iter + 1
    tol = 1e-8 if dtype == np.float64 else 5e-3
    assert np.allclose(m.means, base.means, atol=1e-6 if dtype == np.float32 else 1e-13)
    assert oica.match_rows(m.components, base.components)[1] < tol
    assert s.shape == sb.shape
    if dtype == np.float64:
        assert np.allclose(s, sb, atol=1e-7)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_transform_and_inverse_transform_host_chunks(pd, staging, dtype):
    n, d, k = 50_000, 256, 24
    x = synth.lowrank_noise(n, d, rank=30, seed=9, dtype=dtype)
    m = pd.Pca.new(k)
    m.fit(_dev(x))
    comps, mu = m.components().astype(np.float64), m.mean().astype(np.float64)
    staging(RING, 4096 * d * x.itemsize)
    y = m.transform(x)
    yr = (x.astype(np.float64) - mu) @ comps.T
    tol = 1e-4 if dtype == np.float32 else 1e-11
    assert np.allclose(y, yr, atol=tol * np.abs(yr).max())
    z = m.inverse_transform(y)  # host output drained chunk by chunk
    zr = y.astype(np.float64) @ comps + mu
    assert z.shape == (n, d)
    assert np.allclose(z, zr, atol=tol * np.abs(zr).max())


def test_auto_mode_keeps_small_inputs_resident(pd, staging):
    ctx = staging(0, 1 << 30)
    x = synth.lowrank_noise(20_000, 64, rank=10, seed=1, dtype=np.float32)
    pd.RandomizedPca.with_seed(4, 1).fit(x)
    st = ctx.host_stream_stats()
    assert st["out_of_core"] is False and st["h2d_bytes"] <= 2 * x.nbytes


@pytest.mark.parametrize("mode", [RESIDENT, RING], ids=["resident", "ring"])
@pytest.mark.parametrize("dtype,n,d,k,q,tol", [(np.float32, 60_000, 256, 16, 4, 1e-4), (np.float32, 40_000, 1024, 64, 4, 1e-4),
                                               (np.float64, 30_000, 128, 12, 3, 1e-9), (np.float32, 50_000, 160, 22, 7, 1e-4)])
def test_rpca_host_gram_mode_vs_oracle(pd, staging, mode, dtype, n, d, k, q, tol):
    """Host-fed randomized PCA with the power iterations on the Gram matrix taken during the ingest (the default for a
    host X): 2 traversals of X whatever q - (mean, column sums, Gram) and the final (Y = Xc B_q, C' = Xc^T Y) pair, plus
    the C' = Xc^T Y1 pass of the row-major path - and the oracle's singular values, variance ratios and subspace."""
    x = synth.lowrank_noise(n, d, rank=40, seed=13, dtype=dtype)
    omega = Mcg128Xsl64.from_seed_u128(RNG_SEED).normal_matrix(d, k + 10, dtype)
    ref = opca.RandomizedPca(k, n_iter=q)
    yr = ref.fit_transform(x.astype(np.float64), omega.astype(np.float64))
    ctx = staging(mode, 8192 * d * x.itemsize, gram=1)
    m = pd.RandomizedPcaBuilder.new(k).seed(RNG_SEED).n_power_iter(q).build()
    y = m.fit_transform(x, omega)
    st = ctx.host_stream_stats()
    panel = dtype == np.float32
    assert st["traversals"] == (2 if panel else 3), st
    if mode == RING:
        assert st["h2d_bytes"] == st["traversals"] * x.nbytes + min(n, 8192) * d * x.itemsize
    assert rel(m.singular_values(), ref.singular_values()) < tol
    assert rel(m.explained_variance_ratio(), ref.explained_variance_ratio()) < 2 * tol
    assert np.allclose(m.mean(), ref.means, atol=1e-6 if dtype == np.float32 else 1e-12)
    assert abs(m._total_variance - ref.total_variance) < (1e-5 if dtype == np.float32 else 1e-10) * ref.total_variance
    h = min(k, 12)
    assert opca.principal_angles(m.components()[:h], ref.components[:h]).max() < (2e-3 if dtype == np.float32 else 1e-6)
    assert np.allclose(y[:, :h], yr[:, :h], atol=(2e-3 if dtype == np.float32 else 1e-6) * np.abs(yr).max())


def test_rpca_host_gram_mode_without_centering(pd, staging):
    n, d, k, q = 30_000, 96, 10, 3
    x = synth.lowrank_noise(n, d, rank=30, seed=17, dtype=np.float32)
    omega = Mcg128Xsl64.from_seed_u128(RNG_SEED).normal_matrix(d, k + 10, np.float32)
    ref = opca.RandomizedPca(k, n_iter=q, centering=False)
    ref.fit(x.astype(np.float64), omega.astype(np.float64))
    ctx = staging(RING, 4096 * d * 4, gram=1)
    m = pd.RandomizedPcaBuilder.new(k).seed(RNG_SEED).n_power_iter(q).centering(False).build()
    m.fit(x, omega)
    assert ctx.host_stream_stats()["traversals"] == 2
    assert rel(m.singular_values(), ref.singular_values()) < 1e-4
    assert np.all(m.mean() == 0)
    assert abs(m._total_variance - ref.total_variance) < 1e-5 * ref.total_variance
