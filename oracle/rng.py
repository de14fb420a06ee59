"""ORACLE (test infrastructure only) - host RNG restatement.

Restates the random stream the reference draws Omega / w_init from:
`rand_pcg::Mcg128Xsl64` (reference src/pca.rs:9-12,356-358; src/ica.rs:8-11,75-77)
sampled through `rand_distr::StandardNormal` (reference src/pca.rs:701-705,
src/ica.rs:210-214).  Those crates are NOT vendored under /root/reference
(rand_pcg 0.9 / rand 0.9 / rand_distr 0.5 per Cargo.toml:56-58); this file restates
their published algorithms:

* PCG XSL-RR 128/64 MCG (O'Neill 2014): state *= 0x2360ED051FC65DA44385DF649FCCF645;
  out = rotr64(hi ^ lo, state >> 122).
* Ziggurat normal sampler (Marsaglia & Tsang 2000 / Doornik 2005), 256 layers,
  R = 3.6541528853610088, V = 0.00492867323399, tables built exactly as rand's
  `utils/ziggurat_tables.py` does (double arithmetic).

PARITY STATUS: the stream is pinned by ONE reference test bit only
(src/ica.rs:412 `n_iter == 1`, see tests/test_oracle_golden.py); the table entries
are regenerated, so they may differ from the shipped tables in the last ulp.
"Omega bit-parity unpinned" - which is why the C ABI takes Omega / w_init from
the caller.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may
import this package.
"""
from __future__ import annotations

import math

import numpy as np

MASK64 = (1 << 64) - 1
MASK128 = (1 << 128) - 1
PCG_MULT = 0x2360ED051FC65DA44385DF649FCCF645

ZIG_NORM_R = 3.6541528853610088
ZIG_NORM_V = 0.00492867323399
TABLE_LEN = 256


def _norm_f(x: float) -> float:
    return math.exp(-x * x / 2.0)


def _norm_f_inv(y: float) -> float:
    return math.sqrt(-2.0 * math.log(y))


def ziggurat_tables():
    """X / F tables, same recurrence as rand's generator script."""
    xvec = [0.0] * (TABLE_LEN + 1)
    xvec[0] = ZIG_NORM_V / _norm_f(ZIG_NORM_R)
    xvec[1] = ZIG_NORM_R
    for i in range(2, TABLE_LEN):
        last = xvec[i - 1]
        xvec[i] = _norm_f_inv(ZIG_NORM_V / last + _norm_f(last))
    xvec[TABLE_LEN] = 0.0
    fvec = [_norm_f(x) for x in xvec]
    return xvec, fvec


ZIG_NORM_X, ZIG_NORM_F = ziggurat_tables()


def _bits_to_f64(bits: int) -> float:
    return np.frombuffer(np.uint64(bits).tobytes(), dtype=np.float64)[0].item()


class Mcg128Xsl64:
    """rand_pcg::Mcg128Xsl64 (a.k.a. Pcg64Mcg)."""

    def __init__(self, state: int):
        # Pcg64Mcg::new(state): state | 1   (used by reference tests, src/pca.rs:991)
        self.state = (state | 1) & MASK128

    @classmethod
    def from_seed_u128(cls, seed: int) -> "Mcg128Xsl64":
        """`Pcg::from_seed(seed.to_be_bytes())` (reference src/pca.rs:357, src/ica.rs:76).

        from_seed reads the 16 bytes as a little-endian u128, so the effective
        state is byteswap128(seed) | 1.
        """
        be = int(seed).to_bytes(16, "big")
        return cls(int.from_bytes(be, "little"))

    def next_u64(self) -> int:
        self.state = (self.state * PCG_MULT) & MASK128
        s = self.state
        rot = s >> 122
        xsl = ((s >> 64) ^ s) & MASK64
        return ((xsl >> rot) | (xsl << ((64 - rot) & 63))) & MASK64

    # --- rand 0.9 float conversions -------------------------------------------------
    def _standard_f64(self) -> float:  # rng.random::<f64>()
        return (self.next_u64() >> 11) * (1.0 / (1 << 53))

    def _open01(self) -> float:  # rng.sample(Open01)
        v = _bits_to_f64((self.next_u64() >> 12) | (1023 << 52))
        return v - (1.0 - 2.0 ** -53)

    def standard_normal(self) -> float:
        """rand_distr::StandardNormal for f64 (ziggurat, symmetric)."""
        X, F = ZIG_NORM_X, ZIG_NORM_F
        while True:
            bits = self.next_u64()
            i = bits & 0xFF
            u = _bits_to_f64((bits >> 12) | (1024 << 52)) - 3.0
            x = u * X[i]
            if abs(x) < X[i + 1]:
                return x
            if i == 0:
                return self._tail(u)
            if F[i + 1] + (F[i] - F[i + 1]) * self._standard_f64() < _norm_f(x):
                return x

    def _tail(self, u: float) -> float:
        x, y = 1.0, 0.0
        while -2.0 * y < x * x:
            x_ = self._open01()
            y_ = self._open01()
            x = math.log(x_) / ZIG_NORM_R
            y = math.log(y_)
        return x - ZIG_NORM_R if u < 0.0 else ZIG_NORM_R - x

    def normal_matrix(self, rows: int, cols: int, dtype=np.float64) -> np.ndarray:
        """`Array2::from_shape_fn((rows, cols), |_| A::from_f64(rng.sample(StandardNormal)))`
        - row-major visiting order, one f64 draw per element, then cast."""
        out = np.empty(rows * cols, dtype=np.float64)
        for i in range(rows * cols):
            out[i] = self.standard_normal()
        return out.reshape(rows, cols).astype(dtype)
