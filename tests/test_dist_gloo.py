"""World-size-2 gloo tests (CPU): the host-side plumbing of the row-sharded multi-GPU path -
rank row spans, shipping the communicator id over torch.distributed, and the algebra of combining
per-rank partial statistics (column sums, centred Gram via the global mean, X^T*Q partials), which is
what the NCCL all-reduces inside libpetal_b200 implement on the GPU.  The per-rank arithmetic is done
with numpy here (test stand-in only); the product's kernels are exercised by the -m gpu tests and by
tests/dist_gpu_check.py under torchrun."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from petal_decomposition_b200.dist import broadcast_bytes, shard_rows
        from oracle import pca as opca
        from tests import synth

        # 1. communicator id shipping (128 opaque bytes from rank 0)
        payload = bytes(range(128)) if rank == 0 else None
        got = broadcast_bytes(payload, 128, src=0)
        assert got == bytes(range(128))

        # 2. row sharding + all-reduce algebra == single-process oracle
        x = synth.lowrank_noise(1001, 24, rank=6, seed=3)
        r0, r1 = shard_rows(x.shape[0], rank, world)
        xs = x[r0:r1]
        cnt = torch.tensor([float(xs.shape[0])], dtype=torch.float64)
        s = torch.from_numpy(xs.sum(axis=0))
        dist.all_reduce(cnt)
        dist.all_reduce(s)
        mean = s.numpy() / cnt.item()
        g = torch.from_numpy((xs - mean).T @ (xs - mean))
        dist.all_reduce(g)
        lam = np.linalg.eigvalsh(g.numpy())[::-1]
        ref = opca.Pca(4, economy=True)
        ref.fit(x)
        assert np.allclose(np.sqrt(lam[:4]), ref.singular_values(), rtol=1e-10)
        assert np.isclose(np.trace(g.numpy()), ref.total_variance, rtol=1e-12)
        # X^T Q partials of the range finder
        rng = np.random.default_rng(1)
        omega = rng.standard_normal((24, 8))
        y = (xs - mean) @ omega
        z = torch.from_numpy((xs - mean).T @ y)
        dist.all_reduce(z)
        zfull = (x - mean).T @ ((x - mean) @ omega)
        assert np.allclose(z.numpy(), zfull, rtol=1e-10, atol=1e-8)
        # svd_flip combine rule: first rank with the strictly larger |.| wins
        loc = np.array([[3.0, 5, -1], [3.0, 2, 1]])[rank]
        gathered = [torch.zeros(3, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(loc))
        best = gathered[0].numpy().copy()
        for r in range(1, world):
            if gathered[r][0] > best[0]:
                best = gathered[r].numpy().copy()
        assert np.array_equal(best, [3.0, 5, -1])
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, f"fail: {e!r}"))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
