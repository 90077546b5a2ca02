"""Generates tests/golden/oracle_fixtures.npz: seeded inputs and the ORACLE's outputs for a
handful of small shapes of every hot-path entry point, committed so that the GPU parity tests
also run against frozen vectors (and so a later oracle change is noticed).

    python tests/golden/make_oracle_fixtures.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from algo_dsp_b200 import siggen as G  # noqa: E402

out = {}
cases = [("ols_k300_n5000", 300, 5000), ("ols_k4097_n20000", 4097, 20000), ("ols_k1_n100", 1, 100), ("ols_k900_n100", 900, 100)]
for name, K, n in cases:
    h, x = G.decaying_ir(K, seed=7), G.white(n, seed=1)
    out[name + "_h"], out[name + "_x"] = h, x
    out[name + "_y"] = O.overlap_save(h, 0, x)
h, x = G.test_kernel(64), G.white(4096, seed=2)
out["direct_k64_h"], out["direct_k64_x"], out["direct_k64_y"] = h, x, O.convolve(x, h)
a, b = G.white(3000, seed=3), G.white(700, seed=4)
out["corr_a"], out["corr_b"], out["corr_y"] = a, b, O.correlate(a, b)
out["corr_peak"] = np.array(O.find_peak(out["corr_y"]), dtype=np.float64)
h, x = G.exp_kernel(1024), G.white(4096, seed=5)
p = O.Partitioned(h, 6, 13)
out["part_h"], out["part_x"], out["part_y"] = h, x, p.process_block(x)
h32, x32 = G.decaying_ir(500).astype(np.float32), G.white(6000, seed=6).astype(np.float32)
out["ols32_h"], out["ols32_x"], out["ols32_y"] = h32, x32, O.overlap_save(h32, 0, x32, dtype=np.float32)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_fixtures.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path))
