"""bench.py contract pieces that can be checked without a GPU: the reference arm (CPU restatement timed on the host
cores) prints one JSON line with the keys the driver reads, and the committed bench lines carry the required objects."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_reference(config, rows):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", config,
                          "--steps", "1", "--warmup", "0", "--cpu-rows", str(rows)],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().split("\n")[-1])


def test_reference_arm_line_c1():
    j = _run_reference("c1", 2000)
    assert j["impl"] == "reference"
    assert j["unit"] == "samples/s" and j["higher_is_better"] is True and j["n_gpus"] == 1
    assert j["value"] > 0 and j["ms_per_step"] > 0
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"]["value"] == j["value"]
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in j["config"] and "model" not in j["config"]


def test_reference_arm_line_c3():
    j = _run_reference("c3", 20000)
    assert j["impl"] == "reference" and j["value"] > 0
    assert j["config"]["algorithm"] == "ica"


def test_committed_bench_lines_have_the_contract_keys():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_bench_*.json")))
    assert files, "no committed bench lines under profiles/"
    for f in files:
        j = json.loads(open(f).read().strip().split("\n")[-1])
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                    "scaling", "dtype", "data", "config"):
            assert key in j, (f, key)
        if j.get("impl") == "reference":
            continue
        assert j["gpu_launches"] > 0, f
        assert "clocks" in j and "roofline" in j, f
        r = j["roofline"]
        assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s"), f
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9, f


def test_roofline_traffic_file():
    t = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    for k in ("tc_xb_f32", "tc_atb_f32", "ica_fused_f32"):
        assert 0.9 < t[k]["ratio"] < 1.5, (k, t[k]["ratio"])
