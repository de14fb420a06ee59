"""Seeded synthetic inputs shared by the parity tests (numpy, host side)."""
import numpy as np


def lowrank_noise(n, d, rank, decay=0.9, noise=0.1, seed=0, dtype=np.float64, offset=True):
    """X = Z diag(s) V^T + noise*N(0,1) + per-feature offsets."""
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((n, rank))
    s = 10.0 * decay ** np.arange(rank)
    v, _ = np.linalg.qr(rng.standard_normal((d, rank)))
    x = (z * s) @ v.T + noise * rng.standard_normal((n, d))
    if offset:
        x += rng.uniform(-1, 1, size=d)
    return np.ascontiguousarray(x.astype(dtype))


def gaussian(n, d, seed=0, dtype=np.float64, offset=True):
    """i.i.d. N(0,1) + per-feature offset (config c1)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d))
    if offset:
        x += rng.uniform(-1, 1, size=d)
    return np.ascontiguousarray(x.astype(dtype))


def mixed_sources(n, d, seed=0, dtype=np.float64):
    """d independent non-Gaussian sources (Laplace / uniform / signed |N|^1.5) mixed by a
    well-conditioned matrix + offsets (config c3). Returns (x, mixing)."""
    rng = np.random.default_rng(seed)
    s = np.empty((n, d))
    for j in range(d):
        kind = j % 3
        if kind == 0:
            s[:, j] = rng.laplace(size=n) / np.sqrt(2.0)
        elif kind == 1:
            s[:, j] = rng.uniform(-np.sqrt(3), np.sqrt(3), size=n)
        else:
            g = rng.standard_normal(n)
            v = np.sign(g) * np.abs(g) ** 1.5
            s[:, j] = v / v.std()
    q1, _ = np.linalg.qr(rng.standard_normal((d, d)))
    q2, _ = np.linalg.qr(rng.standard_normal((d, d)))
    a = q1 @ np.diag(np.linspace(1.0, 5.0, d)) @ q2.T
    x = s @ a.T + rng.uniform(-1, 1, size=d)
    return np.ascontiguousarray(x.astype(dtype)), a
