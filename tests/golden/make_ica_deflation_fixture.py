"""Generates tests/golden/ica_deflation.json: golden input/output vectors of sklearn's own deflation FastICA
(`sklearn.decomposition._fastica._ica_def` with `_logcosh` / `_exp` / `_cube`, scikit-learn as installed in the build
container) on small whitened inputs.  The oracle restatement (oracle/ica.py::ica_def) and, through it, the CUDA path
are pinned against these.  sklearn does not travel to the GPU box; the vectors do.

Run from the repo root:  python tests/golden/make_ica_deflation_fixture.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import sklearn  # noqa: E402
from sklearn.decomposition import _fastica as sk  # noqa: E402

from tests import synth  # noqa: E402


def whiten(x):
    xc = (x - x.mean(axis=0)).T
    u, s, _ = np.linalg.svd(xc, full_matrices=False)
    k = (u / s).T
    return (k @ xc) * np.sqrt(x.shape[0])


def main():
    cases = []
    for seed, (n, d) in enumerate([(400, 3), (600, 5), (500, 4)]):
        x, _ = synth.mixed_sources(n, d, seed=20 + seed)
        x1 = whiten(x)
        w_init = np.random.default_rng(100 + seed).standard_normal((d, d))
        for fun, g, args in (("logcosh", sk._logcosh, {"alpha": 1.0}), ("exp", sk._exp, {}), ("cube", sk._cube, {})):
            w, n_iter = sk._ica_def(x1.copy(), tol=1e-6, g=g, fun_args=args, max_iter=200, w_init=w_init.copy())
            cases.append({"fun": fun, "tol": 1e-6, "max_iter": 200, "x1": x1.tolist(), "w_init": w_init.tolist(),
                          "w": np.asarray(w).tolist(), "n_iter": int(n_iter)})
    out = {"generator": "sklearn.decomposition._fastica._ica_def", "sklearn_version": sklearn.__version__, "cases": cases}
    path = os.path.join(ROOT, "tests", "golden", "ica_deflation.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print(path, len(cases), "cases", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
