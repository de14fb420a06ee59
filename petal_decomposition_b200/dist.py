"""Multi-GPU plumbing: one process per GPU (torchrun), rows sharded across ranks.

torch.distributed is used only to ship the NCCL unique id of the library's own communicator
(and for barriers in bench.py); the data-path collectives (all-reduce of column sums, Gram
matrices, X^T*Q partials, FastICA k x k sums) run inside libpetal_b200 on its stream.
"""
from __future__ import annotations

import os

import numpy as np

from . import _cabi
from .api import Context, set_default_context

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def shard_rows(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Rank r owns rows [r*n/N, (r+1)*n/N). Returns (start, stop)."""
    return (rank * n_total) // world, ((rank + 1) * n_total) // world


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0, device=None) -> bytes:
    """Broadcasts a byte string from `src` over the default torch.distributed group
    (works with gloo on CPU and nccl on GPU)."""
    buf = torch.zeros(nbytes, dtype=torch.uint8)
    if dist.get_rank() == src:
        buf = torch.tensor(list(payload), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        buf = buf.cuda(device)
    dist.broadcast(buf, src=src)
    return bytes(buf.cpu().numpy().tobytes())


def init_distributed(backend: str | None = None) -> Context:
    """Initialises torch.distributed from the torchrun environment (if needed), creates the
    library context on LOCAL_RANK's GPU and its NCCL communicator, and installs it as the
    default context. Returns the context."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"),
                                rank=rank, world_size=world)
    ctx = Context(local)
    if world > 1:
        lib = _cabi.load()
        uid = None
        if rank == 0:
            import ctypes as C
            buf = C.create_string_buffer(_cabi.COMM_ID_BYTES)
            st = lib.petal_comm_unique_id(buf)
            if st != _cabi.PETAL_OK:
                raise RuntimeError(lib.petal_last_global_error().decode())
            uid = buf.raw
        uid = broadcast_bytes(uid, _cabi.COMM_ID_BYTES, 0, local)
        ctx.comm_init(uid, rank, world)
    set_default_context(ctx)
    return ctx
