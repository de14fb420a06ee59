"""petal_decomposition_b200 - B200-native (sm_100a) fit/transform hot path of petal-decomposition.

Public API mirrors the Rust crate (reference src/lib.rs:17-18): PcaBuilder/Pca,
RandomizedPcaBuilder/RandomizedPca, FastIcaBuilder/FastIca, DecompositionError.
All arithmetic runs in libpetal_b200.so (hand-written CUDA kernels behind the C ABI declared in
include/petal_b200.h); importing this package fails loudly if that library has not been built.
"""
from . import _cabi

_cabi.load()  # no CPU fallback: the CUDA library must exist

from .api import (CUBE, DEFLATION, EXP, LOGCOSH, PARALLEL, Context, DecompositionError, FastIca, FastIcaBuilder,  # noqa: E402,F401
                  InvalidInput, LinalgError, Pca, PcaBuilder, Pcg, RandomizedPca, RandomizedPcaBuilder,
                  colmean_gram, default_context, ica_def, ica_par, logcosh, set_default_context, small_svd,
                  symmetric_decorrelation, xty)

__all__ = [
    "Pca", "PcaBuilder", "RandomizedPca", "RandomizedPcaBuilder", "FastIca", "FastIcaBuilder",
    "DecompositionError", "InvalidInput", "LinalgError", "Pcg", "Context", "default_context",
    "set_default_context", "ica_par", "ica_def", "logcosh", "symmetric_decorrelation", "small_svd", "colmean_gram", "xty",
    "LOGCOSH", "EXP", "CUBE", "PARALLEL", "DEFLATION",
]
