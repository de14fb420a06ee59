"""ctypes binding of libpetal_b200.so (include/petal_b200.h). Fails loudly when the library is
missing: there is no CPU / PyTorch fallback for the product path."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PETAL_B200_LIB", os.path.join(HERE, "libpetal_b200.so"))  # override: A/B testing of builds

PETAL_OK, PETAL_INVALID_INPUT, PETAL_LINALG_ERROR = 0, 1, 2
COMM_ID_BYTES = 128

c_i64, c_int, c_dbl, c_vp = C.c_int64, C.c_int, C.c_double, C.c_void_p

# every symbol include/petal_b200.h declares: name -> (restype, argtypes)
_TYPED = {
    "petal_pca_fit": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "petal_rpca_fit": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_int, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp,
                               c_vp, c_vp]),
    "petal_transform": (c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp]),
    "petal_inverse_transform": (c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp]),
    "petal_fastica_fit": (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, c_dbl, c_i64, c_int, c_vp, c_vp, c_vp,
                                  C.POINTER(c_i64), C.POINTER(c_dbl), c_vp]),
    "petal_fastica_deflation_fit": (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, c_dbl, c_i64, c_vp, c_vp, c_vp,
                                            C.POINTER(c_i64), C.POINTER(c_dbl), c_vp]),
    "petal_colmean_gram": (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_vp]),
    "petal_xty": (c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_i64, c_vp]),
    "petal_ica_nonlin": (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, c_int, c_vp]),
}
SYMBOLS = {
    "petal_ctx_create": (c_int, [c_int, C.POINTER(c_vp)]),
    "petal_ctx_destroy": (None, [c_vp]),
    "petal_last_error": (C.c_char_p, [c_vp]),
    "petal_last_global_error": (C.c_char_p, []),
    "petal_ctx_set_stream": (c_int, [c_vp, c_vp]),
    "petal_ctx_synchronize": (c_int, [c_vp]),
    "petal_ctx_trim": (c_int, [c_vp]),
    "petal_ctx_launch_count": (c_i64, [c_vp]),
    "petal_ctx_set_f32_engine": (c_int, [c_vp, c_int]),
    "petal_ctx_set_f64_engine": (c_int, [c_vp, c_int]),
    "petal_ctx_set_host_staging": (c_int, [c_vp, c_int, c_i64]),
    "petal_ctx_set_host_gram": (c_int, [c_vp, c_int]),
    "petal_ctx_host_stream_stats": (c_int, [c_vp, C.POINTER(c_i64), C.POINTER(c_i64), C.POINTER(c_int)]),
    "petal_ctx_set_profiling": (c_int, [c_vp, c_int]),
    "petal_ctx_profile_json": (c_i64, [c_vp, C.c_char_p, c_i64]),
    "petal_comm_unique_id": (c_int, [c_vp]),
    "petal_comm_init": (c_int, [c_vp, c_vp, c_int, c_int]),
    "petal_rng_from_seed": (c_vp, [C.c_uint64, C.c_uint64]),
    "petal_rng_from_state": (c_vp, [C.c_uint64, C.c_uint64]),
    "petal_rng_free": (None, [c_vp]),
    "petal_rng_next_u64": (C.c_uint64, [c_vp]),
    "petal_rng_get_state": (None, [c_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "petal_rng_normal_f64": (None, [c_vp, c_vp, c_i64]),
    "petal_rng_normal_f32": (None, [c_vp, c_vp, c_i64]),
    "petal_ica_par_f32": (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, c_dbl, c_i64, c_int, c_vp, c_vp,
                                  C.POINTER(c_i64), C.POINTER(c_dbl)]),
    "petal_ica_par_f64": (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, c_dbl, c_i64, c_int, c_vp, c_vp,
                                  C.POINTER(c_i64), C.POINTER(c_dbl)]),
    "petal_ica_defl_f32": (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, c_dbl, c_i64, c_vp, c_vp,
                                   C.POINTER(c_i64), C.POINTER(c_dbl)]),
    "petal_ica_defl_f64": (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, c_dbl, c_i64, c_vp, c_vp,
                                   C.POINTER(c_i64), C.POINTER(c_dbl)]),
    "petal_symmetric_decorrelation_f64": (c_int, [c_vp, c_vp, c_i64, c_vp]),
    "petal_probe_dmma_tflops": (c_int, [c_vp, c_int, c_vp]),
    "petal_small_svd_f64": (c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp]),
}
for _name, _sig in _TYPED.items():
    for _sfx in ("f32", "f64"):
        SYMBOLS[f"{_name}_{_sfx}"] = _sig

_lib = None


def load() -> C.CDLL:
    """Loads the shared library (once) and binds every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing - build it with `python -m petal_decomposition_b200.build` "
            "(or __graft_entry__.build()). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
