"""Generates tests/golden/c4_shape_20000x4096.npz: the oracle's exact-PCA result (economy gesvd, oracle/pca.py) on the
c4-shaped synthetic matrix bench.make_x_host(20000, 4096, "f64", "pca").  The economy SVD of 20000 x 4096 takes minutes
of CPU time, which the GPU test must not spend on the GPU box; the matrix itself is regenerated from its seed there.
Run from the repository root:  python tests/golden/make_c4_fixture.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from oracle import pca as opca  # noqa: E402

n, d = 20_000, 4096
x = bench.make_x_host(n, d, "f64", "pca")
ref = opca.Pca(d, economy=True)
ref.fit(x)
out = os.path.join(ROOT, "tests", "golden", "c4_shape_20000x4096.npz")
np.savez_compressed(out, singular=ref.singular_values(), total_variance=np.array([ref.total_variance]),
                    components16=ref.components[:16], means=ref.means,
                    x_checksum=np.array([float(np.sum(x[::997, ::13]))]), numpy_version=np.array([np.__version__]))
print(out, os.path.getsize(out))
