"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol the header
declares, the host RNG matches the oracle's restatement bit for bit, and the host mirror's
validation logic behaves like the reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from petal_decomposition_b200 import build, _cabi
    build.build_library()
    return _cabi.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "petal_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(petal_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 30
    from petal_decomposition_b200 import _cabi
    for n in sorted(names):
        assert hasattr(lib, n), f"libpetal_b200.so does not export {n}"
        assert n in _cabi.SYMBOLS, f"ctypes binding misses {n}"
    assert set(_cabi.SYMBOLS) == names


def test_rng_matches_oracle(lib):
    import petal_decomposition_b200 as pd
    from oracle.rng import Mcg128Xsl64
    seed = 1_234_567_891_011_121_314
    a = pd.Pcg.from_seed(seed).standard_normal((40, 30))
    b = Mcg128Xsl64.from_seed_u128(seed).normal_matrix(40, 30)
    assert np.array_equal(a, b)
    r1, r2 = pd.Pcg.new(seed), Mcg128Xsl64(seed)
    assert [r1.next_u64() for _ in range(8)] == [r2.next_u64() for _ in range(8)]
    assert r1.state() == r2.state
    f = pd.Pcg.from_seed(7).standard_normal((5, 5), np.float32)
    g = Mcg128Xsl64.from_seed_u128(7).normal_matrix(5, 5, np.float32)
    assert f.dtype == np.float32 and np.array_equal(f, g)


def test_rng_golden_fixture(lib):
    """tests/golden/rng_stream.json pins the restated host RNG stream (oracle and the library's C++ copy)."""
    import json
    import struct
    import petal_decomposition_b200 as pd
    from oracle.rng import Mcg128Xsl64
    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rng_stream.json")))
    for seed_s, g in fx["seeds"].items():
        seed = int(seed_s)
        o = Mcg128Xsl64.from_seed_u128(seed)
        assert [str(o.next_u64()) for _ in range(16)] == g["next_u64"]
        want = np.array([struct.unpack(">d", bytes.fromhex(h))[0] for h in g["standard_normal_f64_hex"]])
        o = Mcg128Xsl64.from_seed_u128(seed)
        assert np.array_equal(np.array([o.standard_normal() for _ in range(32)]), want)
        assert str(o.state) == g["state_after_32_normals"]
        lib_draws = pd.Pcg.from_seed(seed).standard_normal((32,))
        assert np.array_equal(lib_draws, want)


def test_rng_tail_and_moments(lib):
    import petal_decomposition_b200 as pd
    x = pd.Pcg.from_seed(3).standard_normal((400000,))
    assert abs(x.mean()) < 0.01 and abs(x.std() - 1) < 0.01
    assert np.abs(x).max() > 3.7  # the ziggurat tail branch is exercised


def test_no_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import petal_decomposition_b200 as pd
    with pytest.raises(pd.LinalgError, match="no CPU fallback"):
        pd.Context(0)
    with pytest.raises(pd.DecompositionError):
        pd.Pca.new(1).fit(np.zeros((3, 2)))


def test_host_validation_without_gpu(lib):
    import petal_decomposition_b200 as pd
    m = pd.Pca.new(1)
    m._components = np.array([[0.6, 0.8]])
    m._means = np.zeros(2)
    with pytest.raises(pd.InvalidInput, match="# of columns should be 2"):
        m.transform(np.zeros((2, 3)))
    with pytest.raises(pd.InvalidInput, match="# of columns should be 1"):
        m.inverse_transform(np.zeros((2, 2)))
    ica = pd.FastIca.with_seed(1)
    ica.means = np.zeros(2)
    with pytest.raises(pd.InvalidInput, match="too many columns"):
        ica.transform(np.zeros((2, 3)))
    with pytest.raises(pd.InvalidInput):
        pd.Pca.new(1).fit(np.zeros(3))
    assert str(pd.InvalidInput("x")) == "invalid matrix: x"
    assert str(pd.LinalgError("y")).startswith("linear algerba operation failed")
    # builders mirror the reference's signatures
    assert pd.RandomizedPcaBuilder.with_rng(pd.Pcg.new(1), 3).build().n_components() == 3
    assert pd.RandomizedPca.with_rng(3, pd.Pcg.new(1)).n_components() == 3
    assert pd.PcaBuilder.new(2).centering(False).build()._centering is False


def test_shard_rows():
    from petal_decomposition_b200.dist import shard_rows
    for n, w in [(10, 3), (7, 8), (100, 4), (0, 2)]:
        spans = [shard_rows(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


def test_cpp_host_mirror_compiles(lib, tmp_path):
    """include/petal_decomposition.hpp (the C++ host mirror of the reference API) compiles and links
    against the C ABI; without a GPU it must fail loudly with the reference's error wording."""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = tmp_path / "host_mirror_test"
    libdir = os.path.join(ROOT, "petal_decomposition_b200")
    r = subprocess.run(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-o", str(exe),
                        os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"), "-L", libdir, "-lpetal_b200",
                        f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True).stdout
    import torch
    if torch.cuda.is_available():
        assert out.split()[:3] == ["5", "0", "5"] or abs(float(out.split()[0]) - 5) < 1e-9
    else:
        assert "linear algerba operation failed" in out and "no CPU fallback" in out


# ------------------------------------------------------------------ serde wire format (host logic, no GPU)
def test_serde_wire_format_round_trip(lib):
    """Field names / order of the reference structs (src/pca.rs:41-51,317-329; src/ica.rs:41-50), ndarray's
    {"v","dim","data"} arrays, rand_pcg's {"state": u128}."""
    import json
    import petal_decomposition_b200 as pd
    m = pd.Pca.new(1)
    m._components = np.array([[0.70710677, 0.70710677]], dtype=np.float32)
    m._means = np.array([1, 1], np.float32)
    m._singular = np.array([0.0], np.float32)
    m._total_variance, m._n_samples = 0.0, 1
    text = m.to_json()
    assert text == ('{"components":{"v":1,"dim":[1,2],"data":[0.70710677,0.70710677]},"n_samples":1,'
                    '"means":{"v":1,"dim":[2],"data":[1.0,1.0]},"total_variance":0.0,'
                    '"singular":{"v":1,"dim":[1],"data":[0.0]},"centering":true}')
    back = pd.Pca.from_json(text, np.float32)
    assert back.components().dtype == np.float32 and np.array_equal(back.components(), m.components())
    assert np.array_equal(back.mean(), m.mean()) and back.n_components() == 1

    r = pd.RandomizedPca.with_seed(1, 1_234_567_891_011_121_314)
    r._components, r._means, r._singular = m._components, m._means, m._singular
    doc = json.loads(r.to_json())
    assert list(doc) == ["rng", "components", "n_samples", "means", "total_variance", "singular", "centering"]
    assert doc["rng"] == {"state": r.rng.state()} and doc["rng"]["state"] % 2 == 1
    r2 = pd.RandomizedPca.from_json(r.to_json(), np.float32)
    assert r2.rng.state() == r.rng.state() and r2.rng.next_u64() == r.rng.next_u64()
    # the reference's randomized_pca_serialize test reads the RandomizedPca document into a Pca (src/pca.rs:1037)
    assert np.array_equal(pd.Pca.from_json(r.to_json(), np.float32).components(), m.components())

    ica = pd.FastIca.with_seed(7)
    ica.components = np.array([[1.5, -2.0], [0.25, 4.0]])
    ica.means = np.array([0.5, -0.5])
    ica.n_iter = 3
    doc = json.loads(ica.to_json())
    assert list(doc) == ["rng", "components", "means", "n_iter"] and doc["n_iter"] == 3
    i2 = pd.FastIca.from_json(ica.to_json())
    assert np.array_equal(i2.components, ica.components) and np.array_equal(i2.means, ica.means) and i2.n_iter == 3
    with pytest.raises(pd.InvalidInput):
        pd.Pca.from_json(text.replace('"dim":[1,2]', '"dim":[3,2]'), np.float32)


def test_folded_mean_algebra():
    """The rank-one corrections petal_rpca_fit applies when the column means are folded into the first two
    range-finder passes: products taken with a provisional mean mu~ plus the column sums of
    X - mu~ give the exactly centred quantities."""
    rng = np.random.default_rng(3)
    n, d, l = 500, 12, 5
    x = rng.standard_normal((n, d)) * rng.uniform(0.5, 3, d) + rng.uniform(-4, 4, d)
    omega = rng.standard_normal((d, l))
    mu_t = x[:64].mean(axis=0)                      # provisional mean from a row sample
    xt = x - mu_t
    yt = xt @ omega                                 # X Omega pass with mu~
    zt = xt.T @ np.hstack([yt, np.ones((n, 1))])    # X^T [Y | 1] pass: last column = column sums of X - mu~
    c = zt[:, l]
    delta = c / n
    w, u = omega.T @ delta, omega.T @ c
    z = zt[:, :l] - np.outer(c, w) - np.outer(delta, u) + n * np.outer(delta, w)
    tv = np.sum(xt * xt) - 2 * delta @ c + n * delta @ delta
    xc = x - x.mean(axis=0)
    assert np.allclose(mu_t + delta, x.mean(axis=0), rtol=0, atol=1e-13)
    assert np.allclose(z, xc.T @ (xc @ omega), rtol=1e-12, atol=1e-9)
    assert np.isclose(tv, np.sum(xc * xc), rtol=1e-12)


def test_gram_shift_algebra():
    """The identity behind the one-trip (mean, Gram) ingest of a host X (petal_b200.cu::mean_and_gram): with a provisional
    mean mu~ (first rows), c = sum(x - mu~) and delta = mu - mu~,
        (X - mu)^T (X - mu) = G~ - c delta^T - delta c^T + n delta delta^T,   G~ = (X - mu~)^T (X - mu~),
    also when mu is the mean ROUNDED to the data type (float32), which is the vector the later passes subtract."""
    rng = np.random.default_rng(5)
    n, d = 5000, 12
    x = (rng.standard_normal((n, d)) * rng.uniform(0.5, 3.0, d) + rng.uniform(-5, 5, d)).astype(np.float32).astype(np.float64)
    mu0 = x[:300].mean(axis=0).astype(np.float32).astype(np.float64)       # provisional mean, type T
    mu = x.mean(axis=0).astype(np.float32).astype(np.float64)              # the mean that is subtracted, type T
    g0 = (x - mu0).T @ (x - mu0)
    c = x.sum(axis=0) - n * mu0
    delta = mu - mu0
    g = g0 - np.outer(c, delta) - np.outer(delta, c) + n * np.outer(delta, delta)
    ref = (x - mu).T @ (x - mu)
    assert np.allclose(g, ref, rtol=0, atol=1e-9 * np.abs(ref).max())
    # and the power iteration identity of the Gram mode: Xc^T (Xc B) = G B
    b = rng.standard_normal((d, 4))
    assert np.allclose(ref @ b, (x - mu).T @ ((x - mu) @ b), rtol=1e-10)


def test_gram_route_spans_the_reference_range():
    """The Gram route of randomized PCA (petal_b200.cu::rpca_fit): Z <- Xc^T (Xc B) evaluated as (Xc^T Xc) B.  In
    exact arithmetic the iterates span the same spaces as the reference's range finder (src/pca.rs:707-715, restated
    in oracle.pca.randomized_range_finder with its LU normalisation - any invertible column mixing leaves the span
    alone), so the final projection sees the same Q up to rounding and sigma agrees."""
    from oracle import pca as opca
    from tests import synth
    x = synth.lowrank_noise(3000, 40, rank=12, decay=0.7, noise=0.05, seed=2)
    xc = x - x.mean(axis=0)
    k, q = 6, 4
    omega = np.random.default_rng(9).standard_normal((40, k + 10))
    q_ref = opca.randomized_range_finder(xc, k + 10, q, omega)  # n x l, orthonormal
    g = xc.T @ xc
    b = omega
    for _ in range(q):
        b, _ = np.linalg.qr(g @ b)                               # orth(G B): d x l
    q_gram, _ = np.linalg.qr(xc @ b)                             # the one streamed product pair starts here
    # compare through the quantities the fit reports: singular values of Q^T Xc and the leading right subspace
    s_ref = np.linalg.svd(q_ref.T @ xc, compute_uv=False)
    u, s_gram, vt_gram = np.linalg.svd(q_gram.T @ xc, full_matrices=False)
    assert np.allclose(s_gram[:k], s_ref[:k], rtol=1e-9)
    vt_ref = np.linalg.svd(q_ref.T @ xc, full_matrices=False)[2]
    assert opca.principal_angles(vt_gram[:k], vt_ref[:k]).max() < 1e-6


def test_host_options_validate_without_gpu(lib):
    import petal_decomposition_b200 as pd
    with pytest.raises(pd.InvalidInput, match="parallel.*deflation"):
        pd.FastIca(pd.Pcg.new(1), algorithm="symmetric")
    assert pd.FastIcaBuilder.new().seed(3).algorithm(pd.DEFLATION).build().algorithm == "deflation"
    assert pd.FastIca.with_seed(1).algorithm == "parallel"  # the reference's scheme is the default
