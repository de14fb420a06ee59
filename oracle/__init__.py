"""ORACLE - test infrastructure only.

CPU restatement (numpy/scipy on OpenBLAS LAPACK) of petal-decomposition's fit/transform
hot path, pinned against the reference's own unit tests (tests/test_oracle_golden.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
arm may import it; the product (petal_decomposition_b200) never does.
"""
from . import ica, pca, rng  # noqa: F401
