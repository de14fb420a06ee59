"""Summarises the ncu CSV pages collected by profiles/collect_ncu.sh (in gpurun_out/) into the tracked files
profiles/<round>_ncu_tc_kernels.txt, profiles/<round>_ncu_launches_c2_2Mrows.txt, profiles/<round>_ncu_dmma_kernels.txt,
profiles/<round>_ncu_launches_block_jacobi.txt and profiles/roofline_traffic.json.
Usage: python profiles/summarize_ncu.py [gpurun_out] [round]"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, sys.argv[1] if len(sys.argv) > 1 else "gpurun_out")
OUT = os.path.join(ROOT, "profiles")
ROUND = sys.argv[2] if len(sys.argv) > 2 else "r02"


def raw(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, units = rows[hi], rows[hi + 1]
    return hdr, dict(zip(hdr, units)), [dict(zip(hdr, r)) for r in rows[hi + 2:] if len(r) == len(hdr)]


def fnum(x):
    return float(x.replace(",", ""))


def tobytes(v, u):
    return fnum(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


def toms(v, u):
    return fnum(v) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0,
                      "second": 1e3}[u]


WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic"]

WANT += ["sm__cycles_elapsed.avg.per_second", "smsp__issue_active.avg.pct_of_peak_sustained_active",
         "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
         "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
lines = ["# ncu --set full --clock-control none (profiles/collect_ncu.sh), raw page; " + ROUND,
         "# tc_gemm_kernel<ATB, NP, PANEL, MODE>: <0,0,1,0> = tc_xb fast, <0,0,1,2> = tc_xb precise, <1,80,1,2> = tc_atb precise"]
agg = collections.defaultdict(list)
for path, title in [("tc_full_raw.csv", "tc_gemm kernels, randomized PCA 2M x 1024 f32, l = 74 (first 11 launches of a fit)"),
                    ("ica_full_raw.csv", "ica_fused kernel, FastICA 1M x 64 f32")]:
    if not os.path.exists(os.path.join(SRC, path)):
        continue
    hdr, u, data = raw(os.path.join(SRC, path))
    lines.append("# " + title)
    for d in data:
        lines.append("---")
        for k in WANT:
            if k in d:
                lines.append("  %-72s %s %s" % (k, d[k], u.get(k, "")))
        m = re.search(r"(tc_gemm_kernel<[^>]*>|ica_fused_kernel<[^>]*>)", d["Kernel Name"])
        b = tobytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]) + tobytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
        agg[m.group(1)].append((b, toms(d["gpu__time_duration.sum"], u["gpu__time_duration.sum"]),
                                fnum(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"])))
open(os.path.join(OUT, ROUND + "_ncu_tc_kernels.txt"), "w").write("\n".join(lines) + "\n")

n2 = 2_000_000
alg = n2 * (1024 + 74) * 4.0


def mean(v, i):
    return sum(x[i] for x in v) / len(v)


def entry(kernel, note, alg_bytes, scale):
    v = agg[kernel]
    b = mean(v, 0)
    return {"ncu_kernel": kernel + " " + note, "launches_captured": len(v), "dram_bytes_per_launch_captured": b,
            "algorithmic_bytes_captured": alg_bytes, "ratio": b / alg_bytes, "dram_bytes_per_launch": b * scale,
            "ncu_ms": mean(v, 1), "algorithmic_GBps_under_ncu": alg_bytes / mean(v, 1) / 1e6,
            "tensor_pipe_active_pct": mean(v, 2)}


xb_key = [k for k in agg if k.startswith("tc_gemm_kernel<0, 0, 1, 0>") or k.startswith("tc_gemm_kernel<0, 0, 1, 3>")][0]
traffic = {
    "_round": ROUND,
    "_note": "DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch from one `ncu --set full` capture "
             "(profiles/collect_ncu.sh -> profiles/summarize_ncu.py); tc kernels at 2M x 1024 f32 rows, l = 74 "
             "(dram_bytes_per_launch scaled x5 to the 10M-row bench workload). Keys are bench.py's kernel names; bench.py "
             "uses `ratio`.",
    "tc_xb_f32": entry(xb_key, "(fast mode; the precise last pass moves the same bytes)", alg, 5),
    "tc_atb_f32": entry("tc_gemm_kernel<1, 80, 1, 2>", "(precise mode)", alg, 5),
}
ica = [k for k in agg if k.startswith("ica_fused")]
if ica:
    traffic["ica_fused_f32"] = entry(ica[0], "(logcosh), 1M x 64 f32", 1_000_000 * 64 * 4.0, 1)
json.dump(traffic, open(os.path.join(OUT, "roofline_traffic.json"), "w"), indent=1)
for k, v in agg.items():
    print(k, len(v), "dram GB %.3f" % (mean(v, 0) / 1e9), "ms %.4f" % mean(v, 1), "tensor %.1f%%" % mean(v, 2))

# launch list
rows = list(csv.reader(open(os.path.join(SRC, "launches_c2_2Mrows.csv"))))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
data = [dict(zip(hdr, r)) for r in rows[hi + 1:] if len(r) == len(hdr)]
la = collections.OrderedDict()
for d in data:
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", d["Kernel Name"])
    a = la.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += toms(d["Metric Value"], d["Metric Unit"])
tot = sum(a[1] for a in la.values())
out = ["# ncu --metrics gpu__time_duration.sum --clock-control none (profiles/collect_ncu.sh): python bench.py --rows 2000000 --steps 1 --warmup 1",
       "# randomized PCA f32 2M x 1024, k = 64, q = 4: two fits (warm-up + timed) + torch's data generation kernels; " + ROUND,
       "# kernel, launches, total ms, share of all captured GPU time (cold-cache, serialised: compare shares, not absolutes)"]
for k, a in sorted(la.items(), key=lambda kv: -kv[1][1]):
    out.append("%-110s %4d %9.3f ms %5.1f %%" % (k[:110], a[0], a[1], 100 * a[1] / tot))
open(os.path.join(OUT, ROUND + "_ncu_launches_c2_2Mrows.txt"), "w").write("\n".join(out) + "\n")


# FP64 tensor-path kernels (exact PCA at c4s = 2M x 512): one big launch each
dl = ["# ncu --set full --clock-control none (profiles/collect_ncu.sh), raw page; " + ROUND,
      "# exact Pca f64 2M x 512 (bench.py --config c4s): first Gram launch (atb_dmma_kernel) and first pass-2 GEMM (gemm_nn_dmma_kernel)"]
for path in ("dmma_gram_raw.csv", "dmma_gemm_raw.csv"):
    fp = os.path.join(SRC, path)
    if not os.path.exists(fp):
        continue
    hdr, u, data = raw(fp)
    for d in data:
        dl.append("---")
        for k in WANT:
            if k in d:
                dl.append("  %-72s %s %s" % (k, d[k], u.get(k, "")))
if len(dl) > 2:
    open(os.path.join(OUT, ROUND + "_ncu_dmma_kernels.txt"), "w").write("\n".join(dl) + "\n")

# launch list of the block Jacobi engine (2048 x 2048 SVD)
fp = os.path.join(SRC, "launches_block_jacobi.csv")
if os.path.exists(fp):
    rows = list(csv.reader(open(fp)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    data = [dict(zip(hdr, r)) for r in rows[hi + 1:] if len(r) == len(hdr)]
    la = collections.OrderedDict()
    for d in data:
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", d["Kernel Name"])
        a = la.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += toms(d["Metric Value"], d["Metric Unit"])
    tot = sum(a[1] for a in la.values())
    out = ["# ncu --metrics gpu__time_duration.sum --clock-control none: one-sided block Jacobi SVD of a 2048 x 2048 matrix "
           "(tests/test_gpu_configs.py::test_block_jacobi_svd[2048-2048]); " + ROUND,
           "# kernel, launches, total ms, share (serialised launches: compare shares)"]
    for k, a in sorted(la.items(), key=lambda kv: -kv[1][1]):
        out.append("%-110s %5d %9.3f ms %5.1f %%" % (k[:110], a[0], a[1], 100 * a[1] / tot))
    open(os.path.join(OUT, ROUND + "_ncu_launches_block_jacobi.txt"), "w").write("\n".join(out) + "\n")
