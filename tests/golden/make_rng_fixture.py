"""Generates tests/golden/rng_stream.json: the first outputs of the restated host RNG (Mcg128Xsl64 + ziggurat
StandardNormal) for the seeds the reference's tests and doctests use.  The reference crate cannot be
built in this image (no Rust toolchain), so these values pin the *restatement* against itself across rounds and
against the C++ copy in libpetal_b200 (tests/test_host_cpu.py); the only bit the reference's own tests pin on this
stream is the sign check at src/ica.rs:412,417 (tests/test_oracle_golden.py).

    python tests/golden/make_rng_fixture.py
"""
import json
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.rng import Mcg128Xsl64, ziggurat_tables  # noqa: E402


def f64_hex(x):
    return struct.pack(">d", float(x)).hex()


out = {"_about": "restated rand_pcg::Mcg128Xsl64 + rand_distr::StandardNormal streams (see make_rng_fixture.py)", "seeds": {}}
for seed in (1_234_567_891_011_121_314, 1234567891011121314 + 1, 0, 7, 2 ** 127 + 12345):
    r = Mcg128Xsl64.from_seed_u128(seed)
    u = [r.next_u64() for _ in range(16)]
    r = Mcg128Xsl64.from_seed_u128(seed)
    z = [r.standard_normal() for _ in range(32)]
    out["seeds"][str(seed)] = {"next_u64": [str(v) for v in u], "standard_normal_f64_hex": [f64_hex(v) for v in z],
                               "state_after_32_normals": str(r.state)}
x, f = ziggurat_tables()
out["ziggurat"] = {"X0": f64_hex(x[0]), "X1": f64_hex(x[1]), "X2": f64_hex(x[2]), "X255": f64_hex(x[255]), "F1": f64_hex(f[1]),
                   "F255": f64_hex(f[255])}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "rng_stream.json"), "w"), indent=1)
print("wrote rng_stream.json")
