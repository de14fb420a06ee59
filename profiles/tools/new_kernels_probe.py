"""Small driver for the ncu captures of the kernels added late in r02 (profiles/collect_ncu_new.sh):
  defl : deflation FastICA on 1M x 64 f32, 3 iterations of the first 2 components (ica_defl_pass kernel)
  inv  : inverse_transform f32 1M x 64 -> 1024 (tc_xb on 128-column output windows)"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import petal_decomposition_b200 as pd  # noqa: E402

what = sys.argv[1]
if what == "defl":
    x = torch.randn(1_000_000, 64, device="cuda", dtype=torch.float32).abs_().pow_(1.5)
    w0 = np.random.default_rng(0).standard_normal((64, 64)).astype(np.float32)
    m = pd.FastIca(pd.Pcg.from_seed(1), max_iter=3, tol=0.0, algorithm=pd.DEFLATION)
    m.fit(x, w0)
    print("deflation n_iter", m.n_iter)
else:
    n, d, k = 1_000_000, 1024, 64
    x = torch.randn(n, d, device="cuda", dtype=torch.float32) + 0.5
    m = pd.RandomizedPcaBuilder.new(k).seed(1).n_power_iter(1).build()
    m.fit(x[:200000])
    y = m.transform(x)
    z = m.inverse_transform(y)
    torch.cuda.synchronize()
    print("inverse_transform", tuple(z.shape))
