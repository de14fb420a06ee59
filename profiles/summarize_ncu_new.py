"""Summarises the ncu raw pages collected by profiles/collect_ncu_new.sh (gpurun_out/defl_full_raw.csv,
inv_full_raw.csv) into profiles/<round>_ncu_new_kernels.txt.  Usage: python profiles/summarize_ncu_new.py [gpurun_out] [round]"""
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, sys.argv[1] if len(sys.argv) > 1 else "gpurun_out")
ROUND = sys.argv[2] if len(sys.argv) > 2 else "r02"

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"]


def raw(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, units = rows[hi], rows[hi + 1]
    return dict(zip(hdr, units)), [dict(zip(hdr, r)) for r in rows[hi + 2:] if len(r) == len(hdr)]


lines = ["# ncu --set full --clock-control none (profiles/collect_ncu_new.sh, profiles/tools/new_kernels_probe.py), raw page; " + ROUND,
         "# cold-cache, serialised replays: the bench / event-timed figures in DESIGN.md are the throughput numbers"]
for path, title in [("defl_full_raw.csv", "defl_pass_kernel<float, 1, 1> = deflation FastICA one-pass kernel, 1M x 64 f32 (16 lanes per row, 16 B loads); "
                                          "algorithmic bytes per launch 256 MB"),
                    ("inv_full_raw.csv", "tc_gemm_kernel<0,0,0,2> = transform 1M x 1024 -> 64 (precise tc_xb); tc_gemm_kernel<0,0,0,0> = one 128-column "
                                         "output window of inverse_transform 1M x 64 -> 1024 (reads the 256 MB score matrix, writes 512 MB)")]:
    p = os.path.join(SRC, path)
    if not os.path.exists(p):
        continue
    u, data = raw(p)
    lines.append("# " + title)
    for d in data:
        lines.append("---")
        for k in WANT:
            if k in d:
                lines.append("  %-72s %s %s" % (k, d[k], u.get(k, "")))
open(os.path.join(ROOT, "profiles", ROUND + "_ncu_new_kernels.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:6]))
