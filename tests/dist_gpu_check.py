"""Multi-GPU parity check, run under torchrun (one process per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_gpu_check.py
Each rank fits its row shard collectively (NCCL all-reduce inside libpetal_b200); rank 0 compares the
result with the oracle on the full data."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import petal_decomposition_b200 as pd
    from petal_decomposition_b200.dist import init_distributed, shard_rows
    from oracle import ica as oica
    from oracle import pca as opca
    from oracle.rng import Mcg128Xsl64
    from tests import synth

    ctx = init_distributed()
    rank, world = ctx.rank, ctx.world
    seed = 1_234_567_891_011_121_314
    ok = True

    def report(name, cond, detail=""):
        nonlocal ok
        if rank == 0:
            print(f"[dist check] {name}: {'ok' if cond else 'FAIL'} {detail}", flush=True)
        ok = ok and bool(cond)

    # exact PCA f64
    x = synth.lowrank_noise(6000, 96, rank=20, decay=0.8, seed=5)
    r0, r1 = shard_rows(x.shape[0], rank, world)
    m = pd.Pca.new(6)
    y = m.fit_transform(np.ascontiguousarray(x[r0:r1]))
    ref = opca.Pca(6, economy=True)
    yr = ref.fit_transform(x)
    err = np.max(np.abs(m.singular_values() - ref.singular_values()) / ref.singular_values())
    report("pca f64 sigma", err < 1e-10, f"rel err {err:.2e}")
    report("pca f64 scores (signs incl.)", np.allclose(y, yr[r0:r1], atol=1e-8 * np.abs(yr).max()))

    # randomized PCA f32 (tcgen05 engine)
    x32 = synth.lowrank_noise(40000, 256, rank=40, decay=0.8, noise=0.01, seed=11, dtype=np.float32)
    r0, r1 = shard_rows(x32.shape[0], rank, world)
    omega = Mcg128Xsl64.from_seed_u128(seed).normal_matrix(256, 26, np.float32)
    rp = pd.RandomizedPcaBuilder.new(16).seed(seed).n_power_iter(4).build()
    rp.fit(np.ascontiguousarray(x32[r0:r1]))
    rref = opca.RandomizedPca(16, n_iter=4)
    rref.fit(x32.astype(np.float64), omega.astype(np.float64))
    err = np.max(np.abs(rp.singular_values() - rref.singular_values()) / rref.singular_values())
    report("rpca f32 sigma", err < 1e-4, f"rel err {err:.2e}")
    ang = opca.principal_angles(rp.components(), rref.components).max()
    report("rpca f32 subspace", ang < 5e-3, f"max principal angle {ang:.2e}")

    # FastICA f64
    xi, a = synth.mixed_sources(30000, 6, seed=6)
    r0, r1 = shard_rows(xi.shape[0], rank, world)
    ica = pd.FastIca.with_seed(seed)
    ica.fit(np.ascontiguousarray(xi[r0:r1]))
    oref = oica.FastIca()
    oref.fit(xi, Mcg128Xsl64.from_seed_u128(seed).normal_matrix(6, 6))
    _, defect = oica.match_rows(ica.components, oref.components)
    report("fastica f64 unmixing rows", defect < 1e-6, f"defect {defect:.2e} n_iter {ica.n_iter}/{oref.n_iter}")

    # randomized PCA f32: fit_transform scores of the local shard (signs decided across ranks, src/pca.rs:815-850)
    rp2 = pd.RandomizedPcaBuilder.new(16).seed(seed).n_power_iter(4).build()
    r0, r1 = shard_rows(x32.shape[0], rank, world)
    y32 = rp2.fit_transform(np.ascontiguousarray(x32[r0:r1]))
    yr32 = rref.fit_transform(x32.astype(np.float64), omega.astype(np.float64))[r0:r1]
    sc = np.abs(yr32).max()
    report("rpca f32 scores (signs incl.)", np.allclose(y32[:, :8], yr32[:, :8], atol=2e-3 * sc),
           f"max abs diff {np.max(np.abs(y32[:, :8] - yr32[:, :8])) / sc:.2e} of max |score|")

    # FastICA f32, d = nc = 64: one-pass tcgen05 kernel per rank + 16.6 KB all-reduce per iteration
    x64c, a64 = synth.mixed_sources(120000, 64, seed=1, dtype=np.float32)
    r0, r1 = shard_rows(x64c.shape[0], rank, world)
    ica32 = pd.FastIca.with_seed(seed)
    s32 = ica32.fit_transform(np.ascontiguousarray(x64c[r0:r1]))
    oref32 = oica.FastIca()
    oref32.fit(x64c.astype(np.float64), Mcg128Xsl64.from_seed_u128(seed).normal_matrix(64, 64, np.float32).astype(np.float64))
    _, defect = oica.match_rows(ica32.components, oref32.components)
    am = oica.amari_index(ica32.components, a64)
    report("fastica f32 d=64 unmixing rows", defect < 1e-4 and am < 0.05 and ica32.n_iter < 200,
           f"defect {defect:.2e} amari {am:.3f} n_iter {ica32.n_iter}/{oref32.n_iter}")
    report("fastica f32 sources shape", s32.shape == (r1 - r0, 64) and np.isfinite(np.asarray(s32)).all())

    # uneven shards straddling the kernels' row thresholds (1024 rows): every rank must take the same path,
    # otherwise ranks issue different collectives and hang
    nt = 1024 * world - 1
    xs = synth.lowrank_noise(nt, 64, rank=12, decay=0.7, noise=0.01, seed=21, dtype=np.float32)
    r0, r1 = shard_rows(nt, rank, world)
    om = Mcg128Xsl64.from_seed_u128(seed).normal_matrix(64, 18, np.float32)
    rs = pd.RandomizedPcaBuilder.new(8).seed(seed).n_power_iter(3).build()
    rs.fit(np.ascontiguousarray(xs[r0:r1]))
    rr = opca.RandomizedPca(8, n_iter=3)
    rr.fit(xs.astype(np.float64), om.astype(np.float64))
    err = np.max(np.abs(rs.singular_values() - rr.singular_values()) / rr.singular_values())
    report("rpca f32 uneven shards around 1024 rows", err < 1e-4, f"rel err {err:.2e}")
    xi2, _ = synth.mixed_sources(nt, 8, seed=8, dtype=np.float32)
    ic2 = pd.FastIca.with_seed(seed)
    ic2.fit(np.ascontiguousarray(xi2[r0:r1]))
    or2 = oica.FastIca()
    or2.fit(xi2.astype(np.float64), Mcg128Xsl64.from_seed_u128(seed).normal_matrix(8, 8, np.float32).astype(np.float64))
    _, defect = oica.match_rows(ic2.components, or2.components)
    report("fastica f32 uneven shards around 1024 rows", defect < 1e-3, f"defect {defect:.2e} n_iter {ic2.n_iter}/{or2.n_iter}")

    # (the fits above ran with the default for host shards: power iterations on the all-reduced ingest-time Gram matrix)
    ctx.set_host_gram(0)
    rp0 = pd.RandomizedPcaBuilder.new(16).seed(seed).n_power_iter(4).build()
    r0, r1 = shard_rows(x32.shape[0], rank, world)
    rp0.fit(np.ascontiguousarray(x32[r0:r1]))
    err = np.max(np.abs(rp0.singular_values() - rref.singular_values()) / rref.singular_values())
    d01 = np.max(np.abs(rp0.singular_values() - rp.singular_values()) / rref.singular_values())
    report("rpca f32 plain pass sequence (host Gram off)", err < 1e-4 and d01 < 1e-4 and ctx.host_stream_stats()["traversals"] == 5,
           f"rel err {err:.2e}, vs Gram mode {d01:.2e}, traversals {ctx.host_stream_stats()['traversals']}")
    # host shards streamed out of core (two ring slots, 2048-row chunks) on every rank: same collectives, same results
    ctx.set_host_staging(2, 2048 * 256 * 4)
    r0, r1 = shard_rows(x32.shape[0], rank, world)
    rp3 = pd.RandomizedPcaBuilder.new(16).seed(seed).n_power_iter(4).build()
    y3 = rp3.fit_transform(np.ascontiguousarray(x32[r0:r1]))
    st = ctx.host_stream_stats()
    err = np.max(np.abs(rp3.singular_values() - rref.singular_values()) / rref.singular_values())
    ctx.set_host_gram(1)
    rp4 = pd.RandomizedPcaBuilder.new(16).seed(seed).n_power_iter(4).build()
    rp4.fit(np.ascontiguousarray(x32[r0:r1]))
    st4 = ctx.host_stream_stats()
    err4 = np.max(np.abs(rp4.singular_values() - rref.singular_values()) / rref.singular_values())
    report("rpca f32 out-of-core shards, Gram mode", err4 < 1e-4 and st4["out_of_core"] and st4["traversals"] == 2,
           f"rel err {err4:.2e} traversals {st4['traversals']}")
    report("rpca f32 out-of-core shards", err < 1e-4 and st["out_of_core"] and st["traversals"] == 5,
           f"rel err {err:.2e} traversals {st['traversals']}")
    report("rpca f32 out-of-core scores", np.allclose(y3[:, :8], yr32[:, :8], atol=2e-3 * sc))
    ctx.set_host_staging(2, 1024 * 96 * 8)
    r0, r1 = shard_rows(x.shape[0], rank, world)
    m3 = pd.Pca.new(6)
    y3 = m3.fit_transform(np.ascontiguousarray(x[r0:r1]))
    err = np.max(np.abs(m3.singular_values() - ref.singular_values()) / ref.singular_values())
    report("pca f64 out-of-core shards", err < 1e-10 and np.allclose(y3, yr[r0:r1], atol=1e-8 * np.abs(yr).max()),
           f"rel err {err:.2e} traversals {ctx.host_stream_stats()['traversals']}")
    ctx.set_host_staging(0, 1 << 30)

    # deflation FastICA f64 over row shards: whitening read back (max_iter = 0, w_init = I), the oracle's deflation run in
    # those coordinates (see tests/test_gpu_deflation.py)
    r0, r1 = shard_rows(xi.shape[0], rank, world)
    xloc = np.ascontiguousarray(xi[r0:r1])
    mk = pd.FastIca(pd.Pcg.from_seed(seed), max_iter=0, algorithm=pd.DEFLATION)
    mk.fit(xloc, np.eye(6))
    w6 = Mcg128Xsl64.from_seed_u128(seed).normal_matrix(6, 6)
    x1 = (mk.components @ (xi - mk.means).T) * np.sqrt(xi.shape[0])
    w_ref, it_ref = oica.ica_def(x1, 1e-4, 200, w6, "logcosh")
    md = pd.FastIca(pd.Pcg.from_seed(seed), algorithm=pd.DEFLATION)
    md.fit(xloc, w6)
    cref = w_ref @ mk.components
    sgn = np.sign(np.sum(md.components * cref, axis=1, keepdims=True))
    derr = np.max(np.abs(md.components - sgn * cref)) / np.abs(cref).max()
    report("fastica deflation f64 over shards", derr < 1e-7 and md.n_iter == it_ref, f"err {derr:.2e} n_iter {md.n_iter}/{it_ref}")

    flag = torch.tensor([1 if ok else 0], device=f"cuda:{ctx.device}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("[dist check] ALL OK" if flag.item() == 1 else "[dist check] FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
