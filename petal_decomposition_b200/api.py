"""Host-side mirror of petal-decomposition's public API (reference src/lib.rs:17-28) on top of the
C ABI of libpetal_b200.so.

Same type and method names, argument meaning and error behaviour as the Rust crate:
`PcaBuilder`/`Pca` (src/pca.rs:41-283), `RandomizedPcaBuilder`/`RandomizedPca` (src/pca.rs:317-663),
`FastIcaBuilder`/`FastIca` (src/ica.rs:41-317), `DecompositionError::{InvalidInput, LinalgError}`
(src/lib.rs:22-28).  The Rust generic `A` is the dtype of the array passed to `fit` (float32 or
float64).  Arrays may be numpy arrays (host) or torch CUDA tensors (already resident in HBM);
outputs come back in the same kind.  Fitted model state (components, mean, singular values) is
kept on the host as numpy arrays, like the reference's `Array2`/`Array1` fields.

This module contains no arithmetic: every number is produced by the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

from . import _cabi

try:  # torch is plumbing only (device tensors, streams, torch.distributed)
    import torch
except Exception:  # pragma: no cover
    torch = None


class DecompositionError(Exception):
    """reference src/lib.rs:22-28."""


class InvalidInput(DecompositionError):
    def __init__(self, msg):
        super().__init__(f"invalid matrix: {msg}")
        self.reason = msg


class LinalgError(DecompositionError):
    def __init__(self, msg):
        super().__init__(f"linear algerba operation failed: {msg}")  # (sic) src/lib.rs:26
        self.reason = msg


# ---------------------------------------------------------------------------------------------
# context
# ---------------------------------------------------------------------------------------------
class Context:
    """Owns a petal_ctx (one GPU, one stream). One per process per device."""

    def __init__(self, device: int | None = None, adopt_torch_stream: bool = True):
        self.lib = _cabi.load()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
            if torch is not None and torch.cuda.is_available():
                device = torch.cuda.current_device() if "LOCAL_RANK" not in os.environ else device
        h = C.c_void_p()
        st = self.lib.petal_ctx_create(int(device), C.byref(h))
        if st != _cabi.PETAL_OK:
            raise LinalgError(self.lib.petal_last_global_error().decode())
        self.handle = h
        self.device = int(device)
        self.rank, self.world = 0, 1
        if adopt_torch_stream and torch is not None and torch.cuda.is_available():
            with torch.cuda.device(self.device):
                s = torch.cuda.current_stream().cuda_stream
            self.check(self.lib.petal_ctx_set_stream(self.handle, C.c_void_p(s)))

    def check(self, status: int):
        if status == _cabi.PETAL_OK:
            return
        msg = self.lib.petal_last_error(self.handle).decode()
        if status == _cabi.PETAL_INVALID_INPUT:
            raise InvalidInput(msg)
        raise LinalgError(msg)

    def synchronize(self):
        self.check(self.lib.petal_ctx_synchronize(self.handle))

    def trim(self):
        """Returns the pooled workspaces to the driver."""
        self.check(self.lib.petal_ctx_trim(self.handle))

    def launch_count(self) -> int:
        return int(self.lib.petal_ctx_launch_count(self.handle))

    def set_profiling(self, enable: bool):
        self.check(self.lib.petal_ctx_set_profiling(self.handle, int(bool(enable))))

    def profile(self) -> dict:
        """Per-kernel CUDA-event timings collected since the last call (and clears them)."""
        import json
        buf = C.create_string_buffer(1 << 16)
        self.lib.petal_ctx_profile_json(self.handle, buf, len(buf))
        return json.loads(buf.value.decode() or "{}")

    def set_f64_engine(self, engine: int) -> int:
        return int(self.lib.petal_ctx_set_f64_engine(self.handle, int(engine)))

    def set_f32_engine(self, engine: int) -> int:
        return int(self.lib.petal_ctx_set_f32_engine(self.handle, int(engine)))

    RESIDENT_IF_FITS, RESIDENT, OUT_OF_CORE = 0, 1, 2

    def set_host_staging(self, mode: int = -1, chunk_bytes: int = 0) -> int:
        """How a host (numpy) X reaches HBM: 0 = resident copy when it fits, else out-of-core; 1 = always resident;
        2 = always out-of-core (X re-streamed through a two-slot ring by every pass).  chunk_bytes: H2D chunk size."""
        return int(self.lib.petal_ctx_set_host_staging(self.handle, int(mode), int(chunk_bytes)))

    def set_host_gram(self, enable: int = -1) -> int:
        """Randomized PCA on a host X: power iterations on the Gram matrix accumulated during the ingest (default on)."""
        return int(self.lib.petal_ctx_set_host_gram(self.handle, int(enable)))

    def host_stream_stats(self) -> dict:
        """H2D bytes, traversals of X and the mode of the last call that was given a host X."""
        b, t, r = C.c_int64(0), C.c_int64(0), C.c_int(0)
        self.check(self.lib.petal_ctx_host_stream_stats(self.handle, C.byref(b), C.byref(t), C.byref(r)))
        return {"h2d_bytes": int(b.value), "traversals": int(t.value), "out_of_core": bool(r.value)}

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = C.create_string_buffer(unique_id, _cabi.COMM_ID_BYTES)
        self.check(self.lib.petal_comm_init(self.handle, buf, int(rank), int(world)))
        self.rank, self.world = int(rank), int(world)

    def probe_dmma_tflops(self, ctas_per_sm: int = 2) -> float:
        """FP64 tensor-pipe issue-rate probe (register-only DMMA loop), TFLOP/s."""
        out = C.c_double(0.0)
        self.check(self.lib.petal_probe_dmma_tflops(self.handle, int(ctas_per_sm), C.byref(out)))
        return float(out.value)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.petal_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


_default_ctx: Context | None = None
_ctx_lock = threading.Lock()


def default_context() -> Context:
    global _default_ctx
    with _ctx_lock:
        if _default_ctx is None:
            _default_ctx = Context()
        return _default_ctx


def set_default_context(ctx: Context | None):
    global _default_ctx
    _default_ctx = ctx


# ---------------------------------------------------------------------------------------------
# RNG: rand_pcg::Mcg128Xsl64 + rand_distr::StandardNormal, drawn by the C library on the host
# ---------------------------------------------------------------------------------------------
class Pcg:
    """`rand_pcg::Mcg128Xsl64` (reference src/pca.rs:11-12, src/ica.rs:10-11)."""

    def __init__(self, handle):
        self._lib = _cabi.load()
        self._h = handle

    @classmethod
    def from_seed(cls, seed: int) -> "Pcg":
        """`Pcg::from_seed(seed.to_be_bytes())` (src/pca.rs:357, src/ica.rs:76)."""
        lib = _cabi.load()
        seed = int(seed) & ((1 << 128) - 1)
        return cls(C.c_void_p(lib.petal_rng_from_seed(seed >> 64, seed & ((1 << 64) - 1))))

    @classmethod
    def new(cls, state: int) -> "Pcg":
        """`Pcg64Mcg::new(state)` (src/pca.rs:991)."""
        lib = _cabi.load()
        state = int(state) & ((1 << 128) - 1)
        return cls(C.c_void_p(lib.petal_rng_from_state(state >> 64, state & ((1 << 64) - 1))))

    @classmethod
    def from_entropy(cls) -> "Pcg":
        """`Pcg::from_rng(&mut rand::rng())` / `rand::rng().random()` seeds (src/pca.rs:342-347,580-583)."""
        return cls.from_seed(int.from_bytes(os.urandom(16), "little"))

    def to_json_obj(self) -> dict:
        """rand_pcg's serde form of `Mcg128Xsl64`: {"state": u128}."""
        return {"state": self.state()}

    @classmethod
    def from_json_obj(cls, obj) -> "Pcg":
        lib = _cabi.load()
        st = int(obj["state"]) & ((1 << 128) - 1)
        # the generator keeps the state verbatim (an odd value, as every constructor leaves it)
        return cls(C.c_void_p(lib.petal_rng_from_state(st >> 64, st & ((1 << 64) - 1))))

    def next_u64(self) -> int:
        return int(self._lib.petal_rng_next_u64(self._h))

    def state(self) -> int:
        hi, lo = C.c_uint64(), C.c_uint64()
        self._lib.petal_rng_get_state(self._h, C.byref(hi), C.byref(lo))
        return (hi.value << 64) | lo.value

    def standard_normal(self, shape, dtype=np.float64) -> np.ndarray:
        """`from_shape_fn(shape, |_| A::from_f64(rng.sample(StandardNormal)))` - row-major order."""
        out = np.empty(shape, dtype=dtype)
        fn = self._lib.petal_rng_normal_f32 if out.dtype == np.float32 else self._lib.petal_rng_normal_f64
        fn(self._h, out.ctypes.data_as(C.c_void_p), out.size)
        return out

    def __del__(self):  # pragma: no cover
        try:
            if self._h:
                self._lib.petal_rng_free(self._h)
                self._h = None
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------
# serde wire format of fitted models (reference: `#[derive(Serialize, Deserialize)]` on the model structs,
# src/pca.rs:36-51,309-329, src/ica.rs:33-50, with ndarray's and rand_pcg's own serde impls) as serde_json
# writes it, so that a model fitted here loads in the CPU crate and vice versa:
#   Array2 / Array1 -> {"v": 1, "dim": [rows, cols] | [len], "data": [row-major elements]}   (ndarray array_serde)
#   Mcg128Xsl64     -> {"state": <u128 as a JSON integer>}                                   (rand_pcg, feature serde)
#   struct fields in declaration order; unknown fields are ignored on input like serde does by default
#   (the reference's own test reads a RandomizedPca document into a Pca, src/pca.rs:1029-1041).
# ---------------------------------------------------------------------------------------------
def _json_scalar(v, dtype):
    if np.dtype(dtype) == np.float32:
        return float(str(np.float32(v)))  # shortest digits that round-trip as f32, like serde_json's f32 output
    return float(v)


def _json_array(a: np.ndarray):
    a = np.asarray(a)
    return {"v": 1, "dim": list(a.shape), "data": [_json_scalar(v, a.dtype) for v in a.reshape(-1)]}


def _array_from_json(obj, dtype, ndim):
    if not isinstance(obj, dict) or obj.get("v") != 1:
        raise InvalidInput("unsupported ndarray serde version")
    dim = [int(x) for x in obj["dim"]]
    if len(dim) != ndim:
        raise InvalidInput(f"expected a {ndim}-dimensional array")
    data = np.asarray(obj["data"], dtype=dtype)
    if data.size != int(np.prod(dim)):
        raise InvalidInput("data and dimension must match in size")  # ndarray's own message
    return np.ascontiguousarray(data.reshape(dim))


# ---------------------------------------------------------------------------------------------
# array plumbing
# ---------------------------------------------------------------------------------------------
_SUFFIX = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}


class _Arr:
    """A 2-D row-major array on the host (numpy) or on the device (torch CUDA tensor)."""

    def __init__(self, x):
        self.obj = x
        if torch is not None and isinstance(x, torch.Tensor):
            if x.dim() != 2:
                raise InvalidInput("input must be two-dimensional")
            if not x.is_contiguous():  # reference: assert!(a.is_standard_layout()) src/linalg.rs:75
                raise InvalidInput("input must be in the standard (row-major, contiguous) layout")
            self.is_torch = True
            # explicit lookup: np.dtype(None) would silently be float64 and the library would then read
            # n*d*8 bytes from a smaller buffer
            if x.dtype not in (torch.float32, torch.float64):
                raise InvalidInput("only float32 and float64 are supported")
            if x.device.type != "cuda":
                raise InvalidInput("torch inputs must be CUDA tensors (pass numpy arrays for host data)")
            self.dtype = np.dtype(np.float32 if x.dtype == torch.float32 else np.float64)
            self.ptr = C.c_void_p(x.data_ptr())
            self.shape = tuple(x.shape)
            self.device = x.device
        else:
            x = np.asarray(x)
            if x.ndim != 2:
                raise InvalidInput("input must be two-dimensional")
            if not x.flags["C_CONTIGUOUS"]:
                raise InvalidInput("input must be in the standard (row-major, contiguous) layout")
            self.obj = x
            self.is_torch = False
            self.dtype = x.dtype
            self.ptr = C.c_void_p(x.ctypes.data)
            self.shape = x.shape
            self.device = None
        if self.dtype not in _SUFFIX:
            raise InvalidInput("only float32 and float64 are supported")
        self.suffix = _SUFFIX[self.dtype]

    def empty_like_kind(self, shape):
        """Allocates an output of the same kind (host numpy / device torch) and dtype."""
        if self.is_torch:
            t = torch.empty(shape, dtype=self.obj.dtype, device=self.device)
            return t, C.c_void_p(t.data_ptr() if t.numel() else 0)
        a = np.empty(shape, dtype=self.dtype)
        return a, C.c_void_p(a.ctypes.data if a.size else 0)


def _np_ptr(a: np.ndarray | None):
    if a is None or a.size == 0:
        return C.c_void_p(0)
    return C.c_void_p(a.ctypes.data)


def _ctx_for(ctx: Context | None) -> Context:
    return ctx if ctx is not None else default_context()


def _check_device(ctx: Context, a: "_Arr"):
    """A device tensor must live on the context's GPU (the library dereferences the pointer there)."""
    if a.is_torch and a.device.index is not None and a.device.index != ctx.device:
        raise InvalidInput(f"input is on cuda:{a.device.index} but the context is bound to cuda:{ctx.device}")


def _replicate_from_rank0(ctx: Context, m: np.ndarray) -> np.ndarray:
    """Row-sharded fits keep Omega / w_init replicated: every rank must use rank 0's draw (each rank's own
    entropy-seeded generator would otherwise give different matrices)."""
    if ctx.world <= 1:
        return m
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(m))
    if dist.get_backend() == "nccl":
        t = t.cuda(ctx.device)
    dist.broadcast(t, src=0)
    return t.cpu().numpy()


# ---------------------------------------------------------------------------------------------
# shared transform / inverse_transform (reference src/pca.rs:726-750, 788-811)
# ---------------------------------------------------------------------------------------------
def _transform(ctx: Context | None, x, components: np.ndarray, means: np.ndarray, centering: bool):
    a = _Arr(x)
    if a.shape[1] != means.shape[0]:
        raise InvalidInput(f"# of columns should be {means.shape[0]}")
    ctx = _ctx_for(ctx)
    _check_device(ctx, a)
    comps = np.ascontiguousarray(components, dtype=a.dtype)
    mu = np.ascontiguousarray(means, dtype=a.dtype) if centering else None
    n, d = a.shape
    k = comps.shape[0]
    out, optr = a.empty_like_kind((n, k))
    fn = getattr(ctx.lib, f"petal_transform_{a.suffix}")
    ctx.check(fn(ctx.handle, a.ptr, n, d, _np_ptr(comps), k, _np_ptr(mu), optr))
    return out


def _inverse_transform(ctx: Context | None, y, components: np.ndarray, means: np.ndarray, centering: bool):
    a = _Arr(y)
    if a.shape[1] != components.shape[0]:
        raise InvalidInput(f"# of columns should be {components.shape[0]}")
    ctx = _ctx_for(ctx)
    _check_device(ctx, a)
    comps = np.ascontiguousarray(components, dtype=a.dtype)
    mu = np.ascontiguousarray(means, dtype=a.dtype) if centering else None
    n, k = a.shape
    d = comps.shape[1]
    out, optr = a.empty_like_kind((n, d))
    fn = getattr(ctx.lib, f"petal_inverse_transform_{a.suffix}")
    ctx.check(fn(ctx.handle, a.ptr, n, k, _np_ptr(comps), d, _np_ptr(mu), optr))
    return out


class _PcaBase:
    """Fields of the reference's `Pca<A>` / `RandomizedPca<A, R>` (src/pca.rs:41-51,317-329)."""

    def __init__(self, n_components: int, centering: bool, ctx: Context | None):
        self._k = int(n_components)
        self._centering = bool(centering)
        self._ctx = ctx
        self._components = np.zeros((self._k, 0))  # Array2::zeros((n_components, 0)), src/pca.rs:60
        self._means = np.zeros(0)
        self._singular = np.zeros(0)
        self._total_variance = 0.0
        self._n_samples = 0

    # accessors, src/pca.rs:78-105 / 392-419
    def components(self) -> np.ndarray:
        return self._components

    def mean(self) -> np.ndarray:
        return self._means

    def n_components(self) -> int:
        return self._components.shape[0]

    def singular_values(self) -> np.ndarray:
        return self._singular

    def explained_variance_ratio(self) -> np.ndarray:
        return (self._singular * self._singular) / self._singular.dtype.type(self._total_variance)

    def transform(self, x):
        return _transform(self._ctx, x, self._components, self._means, self._centering)

    def inverse_transform(self, y):
        return _inverse_transform(self._ctx, y, self._components, self._means, self._centering)

    def fit(self, x) -> None:
        self._inner_fit(x, want_scores=False)

    def fit_transform(self, x):
        return self._inner_fit(x, want_scores=True)

    def _store(self, a: _Arr, comps, mean, sing, tv):
        self._components = comps
        self._means = mean
        self._singular = sing
        self._total_variance = float(tv[0])
        self._n_samples = a.shape[0]

    # ---- serde (field order of the reference structs, src/pca.rs:41-51 / 317-329) ----
    def _json_fields(self) -> dict:
        dt = self._components.dtype if self._components.dtype in _SUFFIX else np.dtype(np.float64)
        return {"components": _json_array(self._components.astype(dt, copy=False)), "n_samples": int(self._n_samples),
                "means": _json_array(np.asarray(self._means, dtype=dt)),
                "total_variance": _json_scalar(self._total_variance, dt),
                "singular": _json_array(np.asarray(self._singular, dtype=dt)), "centering": bool(self._centering)}

    def _load_json_fields(self, obj: dict, dtype):
        self._components = _array_from_json(obj["components"], dtype, 2)
        self._n_samples = int(obj["n_samples"])
        self._means = _array_from_json(obj["means"], dtype, 1)
        self._total_variance = float(obj["total_variance"])
        self._singular = _array_from_json(obj["singular"], dtype, 1)
        self._centering = bool(obj["centering"])
        self._k = self._components.shape[0]

    def to_json(self) -> str:
        """`serde_json::to_string(&model)` of the reference (tests src/pca.rs:935-947, 1029-1041)."""
        import json
        return json.dumps(self._json_fields(), separators=(",", ":"))


class Pca(_PcaBase):
    """Principal component analysis - reference `Pca<A>` (src/pca.rs:41-231)."""

    def __init__(self, n_components: int, centering: bool = True, ctx: Context | None = None):
        super().__init__(n_components, centering, ctx)

    @classmethod
    def new(cls, n_components: int) -> "Pca":
        return cls(n_components)

    @classmethod
    def from_json(cls, text: str, dtype=np.float64, ctx: Context | None = None) -> "Pca":
        """`serde_json::from_str::<Pca<A>>` with A = dtype; extra fields (a RandomizedPca's "rng") are ignored."""
        import json
        m = cls(0, ctx=ctx)
        m._load_json_fields(json.loads(text), np.dtype(dtype))
        return m

    def _inner_fit(self, x, want_scores: bool):
        a = _Arr(x)
        ctx = _ctx_for(self._ctx)
        _check_device(ctx, a)
        n, d = a.shape
        k = self._k
        if ctx.world == 1 and n == 0:  # src/pca.rs:207-211 (mean_axis of zero rows is None)
            if min(n, d) < k:
                raise InvalidInput(f"every dimension should be at least {k}")
            return a.empty_like_kind((0, k))[0] if want_scores else None
        comps = np.empty((k, d), dtype=a.dtype)
        mean = np.empty(d, dtype=a.dtype)
        sing = np.empty(k, dtype=a.dtype)
        tv = np.zeros(1, dtype=a.dtype)
        scores, sptr = a.empty_like_kind((n, k)) if want_scores else (None, C.c_void_p(0))
        fn = getattr(ctx.lib, f"petal_pca_fit_{a.suffix}")
        ctx.check(fn(ctx.handle, a.ptr, n, d, k, int(self._centering), _np_ptr(comps), _np_ptr(mean),
                     _np_ptr(sing), _np_ptr(tv), sptr))
        self._store(a, comps, mean, sing, tv)
        return scores


class PcaBuilder:
    """reference src/pca.rs:246-283."""

    def __init__(self, n_components: int):
        self._k = n_components
        self._centering = True

    @classmethod
    def new(cls, n_components: int) -> "PcaBuilder":
        return cls(n_components)

    def centering(self, centering: bool) -> "PcaBuilder":
        self._centering = centering
        return self

    def build(self, ctx: Context | None = None) -> Pca:
        return Pca(self._k, self._centering, ctx)


class RandomizedPca(_PcaBase):
    """Randomized PCA - reference `RandomizedPca<A, R>` (src/pca.rs:317-550).

    `n_oversamples` (10) and `n_power_iter` (7) default to the constants hard-coded at
    src/pca.rs:679-680; they are exposed because BASELINE.json's configs name 4 power iterations."""

    def __init__(self, n_components: int, rng: Pcg | None = None, centering: bool = True,
                 n_oversamples: int = 10, n_power_iter: int = 7, ctx: Context | None = None):
        super().__init__(n_components, centering, ctx)
        self.rng = rng if rng is not None else Pcg.from_entropy()
        self.n_oversamples = int(n_oversamples)
        self.n_power_iter = int(n_power_iter)

    @classmethod
    def new(cls, n_components: int) -> "RandomizedPca":
        return cls(n_components)

    @classmethod
    def with_seed(cls, n_components: int, seed: int) -> "RandomizedPca":
        return cls(n_components, Pcg.from_seed(seed))

    @classmethod
    def with_rng(cls, n_components: int, rng: Pcg) -> "RandomizedPca":
        return cls(n_components, rng)

    def _json_fields(self) -> dict:
        out = {"rng": self.rng.to_json_obj()}  # first field of the struct, src/pca.rs:322
        out.update(super()._json_fields())
        return out

    @classmethod
    def from_json(cls, text: str, dtype=np.float64, ctx: Context | None = None) -> "RandomizedPca":
        import json
        obj = json.loads(text)
        m = cls(0, Pcg.from_json_obj(obj["rng"]), ctx=ctx)
        m._load_json_fields(obj, np.dtype(dtype))
        return m

    def _inner_fit(self, x, want_scores: bool, omega: np.ndarray | None = None):
        a = _Arr(x)
        ctx = _ctx_for(self._ctx)
        _check_device(ctx, a)
        n, d = a.shape
        k = self._k
        if ctx.world == 1 and n == 0:  # src/pca.rs:521-525
            if min(n, d) < k:
                raise InvalidInput(f"every dimension should be at least {k}")
            return a.empty_like_kind((0, k))[0] if want_scores else None
        if ctx.world == 1 and min(n, d) < k:  # before the RNG is touched, src/pca.rs:513-518
            raise InvalidInput(f"every dimension should be at least {k}")
        l = k + self.n_oversamples
        if omega is None:  # src/pca.rs:701-705: d x l draws, row-major, f64 -> A
            omega = self.rng.standard_normal((d, l), a.dtype)
        omega = _replicate_from_rank0(ctx, np.ascontiguousarray(omega, dtype=a.dtype))
        if omega.shape != (d, l):
            raise InvalidInput(f"omega should be {d} x {l}")
        comps = np.empty((k, d), dtype=a.dtype)
        mean = np.empty(d, dtype=a.dtype)
        sing = np.empty(k, dtype=a.dtype)
        tv = np.zeros(1, dtype=a.dtype)
        scores, sptr = a.empty_like_kind((n, k)) if want_scores else (None, C.c_void_p(0))
        fn = getattr(ctx.lib, f"petal_rpca_fit_{a.suffix}")
        ctx.check(fn(ctx.handle, a.ptr, n, d, k, int(self._centering), self.n_oversamples, self.n_power_iter,
                     _np_ptr(omega), _np_ptr(comps), _np_ptr(mean), _np_ptr(sing), _np_ptr(tv), sptr))
        self._store(a, comps, mean, sing, tv)
        return scores

    def fit(self, x, omega=None) -> None:
        self._inner_fit(x, False, omega)

    def fit_transform(self, x, omega=None):
        return self._inner_fit(x, True, omega)


class RandomizedPcaBuilder:
    """reference src/pca.rs:564-663 (note the argument order of `with_rng`, src/pca.rs:643)."""

    def __init__(self, n_components: int, rng: Pcg | None = None):
        self._k = n_components
        self._rng = rng
        self._centering = True
        self._n_oversamples = 10
        self._n_power_iter = 7

    @classmethod
    def new(cls, n_components: int) -> "RandomizedPcaBuilder":
        return cls(n_components)

    @classmethod
    def with_rng(cls, rng: Pcg, n_components: int) -> "RandomizedPcaBuilder":
        return cls(n_components, rng)

    def seed(self, seed: int) -> "RandomizedPcaBuilder":
        self._rng = Pcg.from_seed(seed)
        return self

    def centering(self, centering: bool) -> "RandomizedPcaBuilder":
        self._centering = centering
        return self

    def n_power_iter(self, n: int) -> "RandomizedPcaBuilder":  # extension (reference: 7)
        self._n_power_iter = n
        return self

    def n_oversamples(self, n: int) -> "RandomizedPcaBuilder":  # extension (reference: 10)
        self._n_oversamples = n
        return self

    def build(self, ctx: Context | None = None) -> RandomizedPca:
        return RandomizedPca(self._k, self._rng, self._centering, self._n_oversamples, self._n_power_iter, ctx)


LOGCOSH, EXP, CUBE = 0, 1, 2
PARALLEL, DEFLATION = "parallel", "deflation"


class FastIca:
    """Independent component analysis - reference `FastIca<A, R>` (src/ica.rs:41-222).

    Public surface of the reference: `fit`, `transform`, `fit_transform` only; `components`,
    `means`, `n_iter` are private fields there (read by its in-module tests) and plain attributes
    here.  tol / max_iter default to the constants at src/ica.rs:216."""

    def __init__(self, rng: Pcg | None = None, fun: int = LOGCOSH, tol: float = 1e-4, max_iter: int = 200,
                 lim_variant: int = 0, ctx: Context | None = None, algorithm: str = PARALLEL):
        if algorithm not in (PARALLEL, DEFLATION):
            raise InvalidInput("algorithm must be 'parallel' or 'deflation'")
        self.rng = rng if rng is not None else Pcg.from_entropy()
        self.fun, self.tol, self.max_iter, self.lim_variant = fun, tol, max_iter, lim_variant
        # 'parallel' = the reference's symmetric scheme (src/ica.rs:319-361); 'deflation' = one component at a time
        # (extension; sklearn `_ica_def`)
        self.algorithm = algorithm
        self._ctx = ctx
        self.components = np.zeros((0, 0))  # src/ica.rs:66-72
        self.means = np.zeros(0)
        self.n_iter = 0
        self.final_lim = float("nan")

    @classmethod
    def new(cls) -> "FastIca":
        return cls()

    @classmethod
    def with_seed(cls, seed: int) -> "FastIca":
        return cls(Pcg.from_seed(seed))

    @classmethod
    def with_rng(cls, rng: Pcg) -> "FastIca":
        return cls(rng)

    def _inner_fit(self, x, want_sources: bool, w_init: np.ndarray | None = None):
        a = _Arr(x)
        ctx = _ctx_for(self._ctx)
        _check_device(ctx, a)
        n, d = a.shape
        if ctx.world == 1 and n == 0:  # src/ica.rs:174-176
            return a.empty_like_kind((0, min(n, d)))[0] if want_sources else None
        if ctx.world > 1:
            import torch.distributed as dist
            t = torch.tensor([n], dtype=torch.int64)
            if dist.get_backend() == "nccl":
                t = t.cuda(ctx.device)
            dist.all_reduce(t)
            n_total = int(t.item())
        else:
            n_total = n
        nc = min(n_total, d)  # src/ica.rs:173
        if w_init is None:  # src/ica.rs:210-214
            w_init = self.rng.standard_normal((nc, nc), a.dtype)
        w_init = _replicate_from_rank0(ctx, np.ascontiguousarray(w_init, dtype=a.dtype))
        comps = np.empty((nc, d), dtype=a.dtype)
        mean = np.empty(d, dtype=a.dtype)
        n_iter, lim = C.c_int64(0), C.c_double(0.0)
        sources, sptr = a.empty_like_kind((n, nc)) if want_sources else (None, C.c_void_p(0))
        if self.algorithm == DEFLATION:
            fn = getattr(ctx.lib, f"petal_fastica_deflation_fit_{a.suffix}")
            ctx.check(fn(ctx.handle, a.ptr, n, d, self.fun, float(self.tol), int(self.max_iter), _np_ptr(w_init),
                         _np_ptr(comps), _np_ptr(mean), C.byref(n_iter), C.byref(lim), sptr))
        else:
            fn = getattr(ctx.lib, f"petal_fastica_fit_{a.suffix}")
            ctx.check(fn(ctx.handle, a.ptr, n, d, self.fun, float(self.tol), int(self.max_iter),
                         int(self.lim_variant), _np_ptr(w_init), _np_ptr(comps), _np_ptr(mean), C.byref(n_iter),
                         C.byref(lim), sptr))
        self.components, self.means = comps, mean
        self.n_iter, self.final_lim = int(n_iter.value), float(lim.value)
        return sources

    def fit(self, x, w_init=None) -> None:
        self._inner_fit(x, False, w_init)

    def fit_transform(self, x, w_init=None):
        return self._inner_fit(x, True, w_init)

    def to_json(self) -> str:
        """serde form of `FastIca { rng, components, means, n_iter }` (src/ica.rs:41-50; test :422-432)."""
        import json
        dt = self.components.dtype if self.components.dtype in _SUFFIX else np.dtype(np.float64)
        return json.dumps({"rng": self.rng.to_json_obj(), "components": _json_array(self.components.astype(dt, copy=False)),
                           "means": _json_array(np.asarray(self.means, dtype=dt)), "n_iter": int(self.n_iter)},
                          separators=(",", ":"))

    @classmethod
    def from_json(cls, text: str, dtype=np.float64, ctx: Context | None = None) -> "FastIca":
        import json
        obj = json.loads(text)
        m = cls(Pcg.from_json_obj(obj["rng"]), ctx=ctx)
        m.components = _array_from_json(obj["components"], np.dtype(dtype), 2)
        m.means = _array_from_json(obj["means"], np.dtype(dtype), 1)
        m.n_iter = int(obj["n_iter"])
        return m

    def transform(self, x):
        a = _Arr(x)
        if a.shape[1] != self.means.shape[0]:  # src/ica.rs:124-128
            raise InvalidInput("too many columns")
        return _transform(self._ctx, x, self.components, self.means, True)


class FastIcaBuilder:
    """reference src/ica.rs:244-317."""

    def __init__(self, rng: Pcg | None = None):
        self._rng = rng
        self._fun = LOGCOSH
        self._algorithm = PARALLEL

    @classmethod
    def new(cls) -> "FastIcaBuilder":
        return cls()

    @classmethod
    def with_rng(cls, rng: Pcg) -> "FastIcaBuilder":
        return cls(rng)

    def seed(self, seed: int) -> "FastIcaBuilder":
        self._rng = Pcg.from_seed(seed)
        return self

    def fun(self, fun: int) -> "FastIcaBuilder":  # extension: EXP / CUBE contrast functions
        self._fun = fun
        return self

    def algorithm(self, algorithm: str) -> "FastIcaBuilder":  # extension: 'deflation' (the reference is 'parallel')
        self._algorithm = algorithm
        return self

    def build(self, ctx: Context | None = None) -> FastIca:
        return FastIca(self._rng, fun=self._fun, ctx=ctx, algorithm=self._algorithm)


# ---------------------------------------------------------------------------------------------
# building blocks exposed for the parity tests
# ---------------------------------------------------------------------------------------------
def ica_par(x1t: np.ndarray, tol: float, max_iter: int, w_init: np.ndarray, fun: int = LOGCOSH,
            lim_variant: int = 0, ctx: Context | None = None):
    """reference `ica_par` (src/ica.rs:319-361); x1t is the whitened data as samples x components
    (float64, or float32 to run the f32 engines); W is f64 either way."""
    ctx = _ctx_for(ctx)
    x1t = np.asarray(x1t)
    x1t = np.ascontiguousarray(x1t, dtype=np.float32 if x1t.dtype == np.float32 else np.float64)
    w_init = np.ascontiguousarray(w_init, dtype=np.float64)
    n, nc = x1t.shape
    w = np.empty((nc, nc))
    n_iter, lim = C.c_int64(0), C.c_double(0.0)
    fn = getattr(ctx.lib, f"petal_ica_par_{_SUFFIX[x1t.dtype]}")
    ctx.check(fn(ctx.handle, _np_ptr(x1t), n, nc, fun, float(tol), int(max_iter), int(lim_variant), _np_ptr(w_init),
                 _np_ptr(w), C.byref(n_iter), C.byref(lim)))
    return w, int(n_iter.value)


def ica_def(x1t: np.ndarray, tol: float, max_iter: int, w_init: np.ndarray, fun: int = LOGCOSH,
            ctx: Context | None = None):
    """Deflation FastICA on whitened data (sklearn `_ica_def`, _fastica.py:65-100); x1t is samples x components."""
    ctx = _ctx_for(ctx)
    x1t = np.asarray(x1t)
    x1t = np.ascontiguousarray(x1t, dtype=np.float32 if x1t.dtype == np.float32 else np.float64)
    w_init = np.ascontiguousarray(w_init, dtype=np.float64)
    n, nc = x1t.shape
    w = np.empty((nc, nc))
    n_iter, lim = C.c_int64(0), C.c_double(0.0)
    fn = getattr(ctx.lib, f"petal_ica_defl_{_SUFFIX[x1t.dtype]}")
    ctx.check(fn(ctx.handle, _np_ptr(x1t), n, nc, fun, float(tol), int(max_iter), _np_ptr(w_init), _np_ptr(w),
                 C.byref(n_iter), C.byref(lim)))
    return w, int(n_iter.value)


def logcosh(wx: np.ndarray, fun: int = LOGCOSH, engine: int = 0, ctx: Context | None = None):
    """reference `logcosh` (src/ica.rs:383-398): wx is components x samples; returns (g(wx), row means of g'(wx)).
    engine 1 applies the device function of the one-pass tcgen05 kernel's epilogue (f32 only)."""
    ctx = _ctx_for(ctx)
    wx = np.asarray(wx)
    if wx.dtype not in _SUFFIX:
        raise InvalidInput("only float32 and float64 are supported")
    nc, n = wx.shape
    u = np.array(wx.T, order="C", copy=True)  # samples x components (the streaming pass's layout); replaced by g(u)
    gsum = np.zeros(nc)
    fn = getattr(ctx.lib, f"petal_ica_nonlin_{_SUFFIX[wx.dtype]}")
    ctx.check(fn(ctx.handle, _np_ptr(u), n, nc, int(fun), int(engine), _np_ptr(gsum)))
    return np.ascontiguousarray(u.T), (gsum / n).astype(wx.dtype)


def symmetric_decorrelation(w: np.ndarray, ctx: Context | None = None) -> np.ndarray:
    """reference `symmetric_decorrelation` (src/ica.rs:363-381)."""
    ctx = _ctx_for(ctx)
    w = np.ascontiguousarray(w, dtype=np.float64)
    out = np.empty_like(w)
    ctx.check(ctx.lib.petal_symmetric_decorrelation_f64(ctx.handle, _np_ptr(w), w.shape[0], _np_ptr(out)))
    return out


def small_svd(a: np.ndarray, ctx: Context | None = None):
    """One-sided Jacobi SVD of a small matrix (rows orthogonalised): returns u, s, vt."""
    ctx = _ctx_for(ctx)
    a = np.ascontiguousarray(a, dtype=np.float64)
    m, ln = a.shape
    u, s, vt = np.empty((m, m)), np.empty(m), np.empty((m, ln))
    ctx.check(ctx.lib.petal_small_svd_f64(ctx.handle, _np_ptr(a), m, ln, _np_ptr(u), _np_ptr(s), _np_ptr(vt)))
    return u, s, vt


def colmean_gram(x, centering: bool = True, ctx: Context | None = None):
    """Column means and centred Gram matrix of x (f64 results on the host)."""
    ctx = _ctx_for(ctx)
    a = _Arr(x)
    n, d = a.shape
    mean, gram = np.zeros(d), np.zeros((d, d))
    fn = getattr(ctx.lib, f"petal_colmean_gram_{a.suffix}")
    ctx.check(fn(ctx.handle, a.ptr, n, d, int(centering), _np_ptr(mean), _np_ptr(gram)))
    return mean, gram


def xty(x, y, mean=None, ctx: Context | None = None) -> np.ndarray:
    """(x - mean)^T y as an f64 host matrix (the X^T*Q pass of the range finder)."""
    a, b = _Arr(x), _Arr(y)
    ctx = _ctx_for(ctx)
    n, d = a.shape
    l = b.shape[1]
    mu = None if mean is None else np.ascontiguousarray(mean, dtype=a.dtype)
    out = np.zeros((d, l))
    fn = getattr(ctx.lib, f"petal_xty_{a.suffix}")
    ctx.check(fn(ctx.handle, a.ptr, n, d, _np_ptr(mu), b.ptr, l, _np_ptr(out)))
    return out
