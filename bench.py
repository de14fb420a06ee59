#!/usr/bin/env python
"""bench.py - fit throughput of the B200 hot path on BASELINE.json's headline configuration.

Workload (N=1): configs[1] = RandomizedPca f32, 10,000,000 x 1024, k=64, 4 power iterations,
seeded Omega.  One "step" = one complete `RandomizedPca.fit` over the rank's row shard.
Multi-GPU (torchrun, one process per GPU): rows are sharded, every rank holds its own 10M x 1024
shard (weak scaling); the fit is collective (NCCL all-reduce of the small replicated matrices).

  value  : samples/s with X resident in HBM (device-timed, CUDA events, max over ranks)
  e2e    : samples/s through the public API with a pinned HOST X (H2D inside the call) and host outputs; the library
           streams the host X in 1 GiB chunks and consumes them as they land (`--host-staging 2`: out of core, X never
           resident; `--no-host-gram`: plain pass sequence instead of power iterations on the ingest-time Gram matrix);
           `e2e.sigma_max_rel_diff_vs_device_fit` compares the host-fed model with the device-resident one
  roofline, cpu_baseline, clocks, gpu_launches: see the task contract / DESIGN.md

`--impl reference` times the CPU restatement of the reference (oracle/, numpy+OpenBLAS LAPACK,
all host threads) on a bounded row sample of the same workload (the Rust crate itself cannot be
built in this image: no cargo/rustc).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEED = 1_234_567_891_011_121_314
CONFIGS = {
    # name: (algorithm, dtype, n, d, k, q)
    "c2": ("rpca", "f32", 10_000_000, 1024, 64, 4),
    "c2q7": ("rpca", "f32", 10_000_000, 1024, 64, 7),  # the crate's default number of power iterations
    "c1": ("pca", "f64", 10_000, 100, 10, 0),
    "c3": ("ica", "f32", 1_000_000, 64, 64, 0),
    "c4s": ("pca", "f64", 2_000_000, 512, 64, 0),
    "c4": ("pca", "f64", 2_000_000, 4096, 64, 0),       # configs[3] at its named size (65.5 GB)
    "c5s": ("rpca", "f32", 40_000_000, 256, 32, 4),
    "c5": ("rpca", "f32", 100_000_000, 256, 32, 4),     # configs[4] at its named size (102.4 GB in total)
}
# configs whose row count is the TOTAL over all ranks (strong scaling: the shard shrinks as N grows)
STRONG_DEFAULT = {"c4", "c5"}
METRIC = "fit samples/s (exact/randomized PCA, FastICA) @1/2/4/8 B200; % HBM roofline"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--rows", type=int, default=0, help="override rows per GPU (testing only)")
    ap.add_argument("--engine", type=int, default=-1, help="f32 engine: 0 SIMT, 1 tcgen05 (default: library default)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--host-staging", type=int, default=0, choices=[0, 1, 2],
                    help="e2e arm: how the host X reaches HBM - 0 resident copy when it fits (chunks consumed as they land), "
                         "1 always resident, 2 out-of-core (X re-streamed through a two-slot ring by every traversal)")
    ap.add_argument("--no-host-gram", action="store_true",
                    help="e2e arm, randomized PCA: plain pass sequence instead of power iterations on the ingest-time Gram matrix")
    ap.add_argument("--host-chunk-mb", type=int, default=0, help="e2e arm: H2D chunk size (default: the library's 1 GiB)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-rows", type=int, default=0)
    ap.add_argument("--scaling", default="auto", choices=["auto", "weak", "strong"],
                    help="weak: every rank holds the config's rows; strong: the config's rows are split over the ranks "
                         "(auto: strong for c4 / c5, weak otherwise)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# synthetic data: low-rank + noise + offsets; same generator family on both arms
# ---------------------------------------------------------------------------------------------
def spectrum(rank):
    return 10.0 * 0.9 ** np.arange(rank)


def lowrank_rank(d):
    return 512 if d >= 4096 else min(d, 128)   # c4 is "as c2 with rank 512"


def make_x_device(n, d, dtype, rank_seed, device, algorithm, aux=None):
    """Generates the rank's shard on the device in row chunks (torch is data plumbing here).
    `aux` (dict) receives what the closed-form parity check needs (mixing matrix / spectrum)."""
    import torch
    tdt = torch.float32 if dtype == "f32" else torch.float64
    g = torch.Generator(device=device)
    g.manual_seed(20240607 + rank_seed)
    gs = torch.Generator(device=device)
    gs.manual_seed(777)  # shared structure (V, offsets, mixing) identical on every rank
    x = torch.empty((n, d), dtype=tdt, device=device)
    chunk = max(1, min(n, (1 << 28) // max(d, 1)))
    if algorithm == "ica":
        a = torch.randn((d, d), generator=gs, device=device, dtype=torch.float64)
        q1, _ = torch.linalg.qr(a)
        q2, _ = torch.linalg.qr(torch.randn((d, d), generator=gs, device=device, dtype=torch.float64))
        mix = (q1 * torch.linspace(1.0, 5.0, d, device=device, dtype=torch.float64)) @ q2.T
        off = torch.rand(d, generator=gs, device=device, dtype=torch.float64) * 2 - 1
        if aux is not None:
            aux["mixing"] = mix.cpu().numpy()
        for r0 in range(0, n, chunk):
            r1 = min(n, r0 + chunk)
            m = r1 - r0
            u = torch.rand((m, d), generator=g, device=device, dtype=torch.float64)
            lap = -torch.sign(u - 0.5) * torch.log1p(-2 * torch.abs(u - 0.5) + 1e-300) / np.sqrt(2.0)
            uni = (u * 2 - 1) * np.sqrt(3.0)
            gn = torch.randn((m, d), generator=g, device=device, dtype=torch.float64)
            sg = torch.sign(gn) * torch.abs(gn) ** 1.5 / 1.4
            kind = (torch.arange(d, device=device) % 3)[None, :]
            s = torch.where(kind == 0, lap, torch.where(kind == 1, uni, sg))
            x[r0:r1] = (s @ mix.T + off).to(tdt)
        return x
    rank = lowrank_rank(d)
    if aux is not None:
        aux["spectrum"] = spectrum(rank)
        aux["noise"] = 0.1
    v, _ = torch.linalg.qr(torch.randn((d, rank), generator=gs, device=device, dtype=torch.float32))
    sv = (v * torch.tensor(spectrum(rank), device=device, dtype=torch.float32)).T.contiguous()  # rank x d
    off = torch.rand(d, generator=gs, device=device, dtype=torch.float32) * 2 - 1
    for r0 in range(0, n, chunk):
        r1 = min(n, r0 + chunk)
        z = torch.randn((r1 - r0, rank), generator=g, device=device, dtype=torch.float32)
        blk = z @ sv
        blk += 0.1 * torch.randn((r1 - r0, d), generator=g, device=device, dtype=torch.float32)
        blk += off
        x[r0:r1] = blk.to(tdt)
        del z, blk
    return x


def make_x_host(n, d, dtype, algorithm):
    """numpy version of the same distribution family for the CPU arm (bounded row sample)."""
    rng = np.random.default_rng(20240607)
    npdt = np.float32 if dtype == "f32" else np.float64
    if algorithm == "ica":
        from tests import synth
        return synth.mixed_sources(n, d, seed=1, dtype=npdt)[0]
    rank = lowrank_rank(d)
    v, _ = np.linalg.qr(rng.standard_normal((d, rank)))
    x = (rng.standard_normal((n, rank), dtype=np.float32) * spectrum(rank).astype(np.float32)) @ v.T.astype(np.float32)
    x += 0.1 * rng.standard_normal((n, d), dtype=np.float32)
    x += rng.uniform(-1, 1, size=d).astype(np.float32)
    return np.ascontiguousarray(x.astype(npdt))


# ---------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU arm (oracle) - bounded sample
# ---------------------------------------------------------------------------------------------
def cpu_fit_once(algorithm, x, k, q):
    from oracle import ica as oica
    from oracle import pca as opca
    from oracle.rng import Mcg128Xsl64
    d = x.shape[1]
    if algorithm == "rpca":
        rng = np.random.default_rng(5)
        omega = rng.standard_normal((d, k + 10)).astype(x.dtype)  # values irrelevant for timing
        m = opca.RandomizedPca(k, n_iter=q)
        m.fit(x, omega)
        return m
    if algorithm == "pca":
        m = opca.Pca(k, economy=True)  # economy SVD: the reference's full n x n U cannot be afforded
        m.fit(x)
        return m
    w_init = np.random.default_rng(5).standard_normal((d, d)).astype(x.dtype)
    m = oica.FastIca()
    m.fit(x, w_init)
    return m


def cpu_sample_rows(algorithm, d, requested):
    if requested:
        return requested
    # sized for roughly 10 s of CPU work on 16 host threads (the whole default run must stay within minutes)
    if algorithm == "rpca":
        return max(1000, int(1_000_000 * 1024 / d))
    if algorithm == "pca":
        return max(1000, int(4_000_000 / min(d, 1024)))  # d > 1024: timed at d = 1024, see run_cpu
    return 1_000_000


CPU_PCA_MAX_D = 1024  # the reference's gesvd needs > 15 min per fit at d = 4096 (measured: 14 s at 4000 x 1024, 124 s at 3000 x 2048)


def run_cpu(algorithm, dtype, d, k, q, rows, steps, warmup, threads=None):
    """Times the oracle restatement; BLAS pools pinned to `threads` (default: every host core) for the duration.
    Exact PCA beyond d = 1024 is timed at d = 1024 and scaled by (1024 / d)^2 (the per-sample cost of the SVD grows
    with d^2): the reference's LAPACK gesvd at d = 4096 does not fit the bench's time budget at any row count."""
    scale = 1.0
    if algorithm == "pca" and d > CPU_PCA_MAX_D:
        scale = (CPU_PCA_MAX_D / d) ** 2
        d = CPU_PCA_MAX_D
        rows = min(rows, 4000)
        k = min(k, d)
        val, dt, used = run_cpu(algorithm, dtype, d, k, q, rows, steps, warmup, threads)
        return val * scale, dt, used
    x = make_x_host(rows, d, dtype, algorithm)
    with all_host_threads(threads):
        for _ in range(min(warmup, 1)):
            cpu_fit_once(algorithm, x, k, q)
        t0 = time.perf_counter()
        for _ in range(steps):
            cpu_fit_once(algorithm, x, k, q)
        dt = (time.perf_counter() - t0) / steps
        used = blas_threads()
    return rows / dt, dt, used


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class all_host_threads:
    """Pins the BLAS pools to every host core for the CPU legs.  torchrun exports OMP_NUM_THREADS=1 for
    nproc > 1, which would otherwise make the N >= 2 reference arms run on one thread and inflate the ratio."""

    def __init__(self, limit=None):
        self.limit = limit or host_cores()
        self.ctx = None

    def __enter__(self):
        try:
            from threadpoolctl import threadpool_limits
            self.ctx = threadpool_limits(limits=self.limit)
            self.ctx.__enter__()
        except Exception:
            self.ctx = None
        return self

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)
        return False


def bind_to_gpu_numa_node(index):
    """Best effort: run this rank's host threads (and first-touch its pinned staging buffer) on the CPUs next to its
    GPU, so that 4 - 8 ranks do not stage 41 GB each through one memory controller."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {i * 64 + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= set(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else cpus
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# ---------------------------------------------------------------------------------------------
# closed-form check of the fitted model of the timed run (the synthetic structure is known)
# ---------------------------------------------------------------------------------------------
def parity_check(algorithm, model, aux, n_total, k):
    try:
        if algorithm == "ica":
            w = np.asarray(model.components, np.float64)
            a = aux["mixing"]
            p = np.abs(w @ a)
            dd = p.shape[0]
            r = (p / p.max(axis=1, keepdims=True)).sum(axis=1) - 1.0
            c = (p / p.max(axis=0, keepdims=True)).sum(axis=0) - 1.0
            amari = float((r.sum() + c.sum()) / (2.0 * dd * (dd - 1)))
            return {"what": "Amari index of the unmixing matrix against the known mixing (0 = perfect)", "value": amari,
                    "tol": 0.05, "ok": bool(amari < 0.05)}
        kk = min(int(k), 20, len(aux["spectrum"]))
        s_true = np.sqrt(n_total * (aux["spectrum"][:kk] ** 2 + aux["noise"] ** 2))
        s = np.asarray(model.singular_values(), np.float64)[:kk]
        err = float(np.max(np.abs(s - s_true) / s_true))
        tol = max(1e-3, 4.0 / np.sqrt(n_total))
        return {"what": f"leading {kk} singular values against the closed form sqrt(n (s_i^2 + noise^2)) of the synthetic "
                        "spectrum (sampling error ~ 1/sqrt(n))", "max_rel_err": err, "tol": tol, "ok": bool(err < tol)}
    except Exception as e:  # never let the check break the bench line
        return {"what": "closed-form check", "error": str(e)[:200], "ok": False}


# ---------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    algorithm, dtype, n_cfg, d, k, q = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    strong = args.scaling == "strong" or (args.scaling == "auto" and args.config in STRONG_DEFAULT)
    n_cfg = args.rows or n_cfg
    if strong:
        n = (rank + 1) * n_cfg // world - rank * n_cfg // world   # this rank's shard of the fixed total
        n_total = n_cfg
    else:
        n = n_cfg
        n_total = n_cfg * world
    workload = {
        "c2": "RandomizedPca f32 10Mx1024 k=64 q=4 (configs[1])",
        "c2q7": "RandomizedPca f32 10Mx1024 k=64 q=7 (configs[1] with the reference's default power iterations)",
        "c1": "Pca f64 10000x100 k=10 (configs[0])",
        "c3": "FastIca logcosh f32 1Mx64 (configs[2])", "c4s": "Pca f64 2Mx512 k=64 (configs[3] at d=512)",
        "c4": "Pca f64 2Mx4096 k=64 (configs[3])",
        "c5s": "RandomizedPca f32 40Mx256 k=32 q=4 (configs[4] per-GPU shard)",
        "c5": "RandomizedPca f32 100Mx256 k=32 q=4 (configs[4], rows split over the ranks)"}[args.config]
    config = {"workload": workload, "algorithm": algorithm, "rows_per_gpu": n, "rows_total": n_total, "features": d,
              "n_components": k, "power_iterations": q, "oversamples": 10, "sharding": f"rows x{world}",
              "l2": "inputs larger than L2 (no flush needed)" if n * d * (4 if dtype == "f32" else 8) > (256 << 20)
              else "L2 flushed between steps"}
    scaling = "strong" if strong else "weak"

    if args.impl == "reference":
        if rank != 0:
            return 0
        rows = cpu_sample_rows(algorithm, d, args.cpu_rows)
        val, dt, cores = run_cpu(algorithm, dtype, d, k, q, rows, max(1, args.steps), args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                "scaling": scaling, "vs_baseline": None, "dtype": dtype, "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port",
                                 "sample": (f"{rows} rows x {d} (oracle restatement of the reference, numpy/OpenBLAS, BLAS pool "
                                            f"pinned to {cores} threads; Rust crate not buildable here)")
                                 if not (algorithm == "pca" and d > CPU_PCA_MAX_D) else
                                 (f"{min(rows, 4000)} rows x {CPU_PCA_MAX_D} features scaled by ({CPU_PCA_MAX_D}/{d})^2 to d = {d} "
                                  f"(LAPACK gesvd at d = {d} needs > 15 min per fit on this host); oracle restatement, "
                                  f"numpy/OpenBLAS, {cores} threads")},
                "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    import petal_decomposition_b200 as pd
    from petal_decomposition_b200.dist import init_distributed

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa_cpus = bind_to_gpu_numa_node(local) if world > 1 else None
    ctx = init_distributed()
    if args.engine >= 0:
        ctx.set_f32_engine(args.engine)
    dev = torch.device("cuda", local)
    esize = 4 if dtype == "f32" else 8

    aux = {}
    x = make_x_device(n, d, dtype, rank, dev, algorithm, aux)
    torch.cuda.synchronize()

    def make_model():
        if algorithm == "rpca":
            return pd.RandomizedPcaBuilder.new(k).seed(SEED).n_power_iter(q).build()
        if algorithm == "pca":
            return pd.Pca.new(k)
        return pd.FastIca.with_seed(SEED)

    flush = None
    if n * d * esize <= (256 << 20):
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(inp):
        m = make_model()
        m.fit(inp)
        return m

    for _ in range(args.warmup):
        step(x)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.set_profiling(True)
    ctx.profile()
    launches0 = ctx.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    model = None
    for i in range(args.steps):
        if flush is not None:
            flush.fill_(i & 0xFF)
        ev[i][0].record()
        model = step(x)
        ev[i][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = ctx.launch_count() - launches0
    prof = ctx.profile()
    ctx.set_profiling(False)
    clocks = sampler.stop() if rank == 0 else None
    tm = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    dev_ms_max = float(tm.item())
    ms_per_step = dev_ms_max / args.steps
    value = n_total / (ms_per_step * 1e-3)
    n_iter_info = getattr(model, "n_iter", None)
    parity = parity_check(algorithm, model, aux, n_total, k) if rank == 0 else None

    # ---- e2e: host (pinned) input, H2D inside the call, host outputs ----
    e2e = None
    if not args.no_e2e:
        try:
            n_e2e = n
            try:
                xh_t = torch.empty((n_e2e, d), dtype=x.dtype, pin_memory=True)
            except Exception:
                n_e2e = max(1, n // 8)
                xh_t = torch.empty((n_e2e, d), dtype=x.dtype, pin_memory=True)
            xh_t.copy_(x[:n_e2e])
            torch.cuda.synchronize()
            xh = xh_t.numpy()
            del x
            torch.cuda.empty_cache()
            ctx.trim()
            ctx.set_host_staging(args.host_staging, args.host_chunk_mb << 20)
            if args.no_host_gram:
                ctx.set_host_gram(0)
            for _ in range(min(args.warmup, 1)):
                step(xh)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                m = step(xh)
            barrier()
            dt = time.perf_counter() - t0
            te = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dt = float(te.item()) / args.steps
            d2h = (m.components().nbytes + m.mean().nbytes + m.singular_values().nbytes + esize) \
                if algorithm != "ica" else (m.components.nbytes + m.means.nbytes)
            rows_e2e_total = n_e2e * world if not strong else (n_total if n_e2e == n else n_e2e * world)
            e2e = {"value": rows_e2e_total / dt, "unit": "samples/s", "h2d_bytes_per_step": int(n_e2e * d * esize),
                   "d2h_bytes_per_step": int(d2h), "rows_per_gpu": n_e2e, "ms_per_step": dt * 1e3,
                   "numa_bound_cpus": numa_cpus,
                   "timing": "wall clock around the public API call (includes H2D of X from pinned host memory)"}
            if algorithm != "ica" and n_e2e == n and model is not None:
                # the host-fed fit (chunked ingest, power iterations on the ingest-time Gram matrix) against the
                # device-resident fit of the same X: same model up to rounding
                s_dev = np.asarray(model.singular_values(), np.float64)
                s_e2e = np.asarray(m.singular_values(), np.float64)
                e2e["sigma_max_rel_diff_vs_device_fit"] = float(np.max(np.abs(s_e2e - s_dev) / s_dev))
            st = ctx.host_stream_stats()
            e2e["host_staging"] = {"mode": "out-of-core ring" if st["out_of_core"] else "resident copy, chunks consumed as they land",
                                   "traversals_of_x": st["traversals"], "h2d_bytes_moved": st["h2d_bytes"],
                                   "power_iterations_on_ingest_gram": bool(ctx.set_host_gram(-1)) and algorithm == "rpca",
                                   "chunk_bytes": (args.host_chunk_mb << 20) or (1 << 30)}
            if st["out_of_core"]:
                e2e["h2d_bytes_per_step"] = int(st["h2d_bytes"])
        except Exception as e:  # host memory too small etc.
            e2e = {"value": None, "unit": "samples/s", "error": str(e)[:200]}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0

    # ---- roofline of the dominant kernel (largest total device time in the timed region) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    roof = None
    if prof:
        stream_k = {kname: v for kname, v in prof.items() if v.get("work", 0) > 0}
        if stream_k:
            top = max(stream_k, key=lambda kn: stream_k[kn]["total_ms"])
            v = stream_k[top]
            all_ms = {kn: round(vv["total_ms"] / args.steps, 4) for kn, vv in prof.items()}
            if top in ("atb_dmma_f64", "gemm_dmma_f64", "syrk_dmma_f64") and v["total_ms"] / v["count"] > 1.0:
                # FP64 tensor pipe: the Gram passes of exact PCA at d >= ~256 are compute-bound (AI ~ d/8 flop/B);
                # flops syrk-counted as SURVEY 8(d): d (d + 1) per sample and pass
                # rows are summed over the launches through `work` (= rows x d x 8 bytes for a symmetric Gram launch)
                flops = (v["work"] / (d * 8.0)) * d * (d + 1)
                achieved = flops / (v["total_ms"] * 1e-3) / 1e12
                fp64_peak, fp64_src = 40.0, "nominal 40 TFLOP/s (B200 FP64 tensor)"
                try:
                    pj = json.load(open(os.path.join(ROOT, "profiles", "r02_fp64_dmma_peak.json")))
                    fp64_peak, fp64_src = float(pj["tflops"]), "measured DMMA issue-rate probe (profiles/r02_fp64_dmma_peak.json)"
                except Exception:
                    pass
                roof = {"bound": "tensor", "kernel": top, "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                        "frac": achieved / fp64_peak, "traffic": None, "peak_source": fp64_src,
                        "flops_convention": "syrk-counted d(d+1) per sample per pass", "launches": v["count"],
                        "avg_ms": v["total_ms"] / v["count"], "kernel_share_of_step": v["total_ms"] / dev_ms,
                        "all_kernels_ms": all_ms}
            else:
                achieved = v["work"] / (v["total_ms"] * 1e-3) / 1e9
                traffic = None
                try:
                    # ncu's DRAM bytes were captured on the c2 shape with fewer rows: what carries over to this
                    # launch is the measured traffic / algorithmic ratio (profiles/roofline_traffic.json)
                    tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
                    ratio = tj.get(top, {}).get("ratio")
                    if ratio is not None:
                        traffic = float(ratio) * v["work"] / v["count"]
                except Exception:
                    pass
                roof = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                        "frac": achieved / hbm_peak, "frac_of_nominal_8000_GBps": achieved / 8000.0, "traffic": traffic,
                        "traffic_source": "ncu dram bytes / algorithmic bytes ratio (profiles/roofline_traffic.json) x this launch's algorithmic bytes",
                        "peak_source": peak_kind,
                        "launches": v["count"], "avg_ms": v["total_ms"] / v["count"],
                        "algorithmic_bytes_per_launch": v["work"] / v["count"],
                        "kernel_share_of_step": v["total_ms"] / dev_ms,
                        "all_kernels_ms": all_ms}
                if algorithm == "rpca":
                    # whole fit against the reference's pass count (src/pca.rs:707-715: X Omega, q x {X^T P, X P}, Q^T X)
                    ref_bytes = (2 * q + 2) * float(n) * d * esize
                    roof["fit_vs_reference_passes"] = {
                        "bytes": ref_bytes, "achieved_GBps": ref_bytes / (dev_ms / args.steps * 1e-3) / 1e9,
                        "frac": ref_bytes / (dev_ms / args.steps * 1e-3) / 1e9 / hbm_peak,
                        "note": "(2q+2) d s bytes per sample, the reference algorithm's X traffic, over the whole fit time"}

    cpu = None
    if not args.no_cpu:
        rows = cpu_sample_rows(algorithm, d, args.cpu_rows)
        val, dt, cores = run_cpu(algorithm, dtype, d, k, q, rows, 1, 1)
        cpu = {"value": val, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"{rows} rows x {d}, one fit ({dt:.1f} s), oracle restatement on numpy/OpenBLAS, BLAS pool pinned to {cores} threads"}
        if algorithm == "pca" and d > CPU_PCA_MAX_D:
            cpu["sample"] = (f"min({rows}, 4000) rows x {CPU_PCA_MAX_D} features, one fit ({dt:.1f} s), scaled by "
                             f"({CPU_PCA_MAX_D}/{d})^2 to d = {d}: LAPACK gesvd at d = {d} needs > 15 min per fit on this host; "
                             f"oracle restatement on numpy/OpenBLAS, {cores} threads")
        # the crate's own GEMMs (matrixmultiply) are single-threaded: also time the restatement on one BLAS thread,
        # on a quarter of the sample so the default run stays short
        try:
            rows1 = max(rows // 4, min(rows, 4 * d))
            val1, dt1, _ = run_cpu(algorithm, dtype, d, k, q, rows1, 1, 0, threads=1)
            cpu["value_1_thread"] = val1
            cpu["sample_1_thread"] = f"{rows1} rows x {d}, one fit ({dt1:.1f} s), 1 BLAS thread"
        except Exception as e:
            cpu["value_1_thread"] = None
            cpu["sample_1_thread"] = f"unavailable: {e}"[:120]

    line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": dtype, "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
            "parity_check": parity, "wall_ms_per_step": t_wall / args.steps * 1e3}
    if n_iter_info is not None:
        line["config"]["ica_iterations"] = n_iter_info
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
