"""Expands the formula-defined inputs of tests/golden/reference_kats.json (the Go tests build
them with math.Sin/Cos/Exp loops)."""
import numpy as np


def gen(rec):
    k = rec["kind"]
    n = rec["n"]
    i = np.arange(n, dtype=np.float64)
    if k == "sine":            # math.Sin(2*pi*i/period)
        return np.sin(2 * np.pi * i / rec["period"])
    if k == "sine_period":     # example_test.go:110-112
        return np.sin(2 * np.pi * i / float(rec["period"]))
    if k == "cos":             # conv_test.go:248-251
        return np.cos(2 * np.pi * i / rec["period"])
    if k == "mod10":           # float64(i % 10)
        return (np.arange(n) % 10).astype(np.float64)
    if k == "exp_decay":       # math.Exp(-i/tau)
        return np.exp(-i / rec["tau"])
    raise KeyError(k)


def pcg_uniform(n, seed=42):
    """Stand-in for makePartitionedTestSignal (partitioned_test.go:23-32: PCG(42,0) uniform
    in [-1,1)); Go's PCG stream is not reproduced, distribution and determinism are."""
    return np.random.default_rng(seed).random(n) * 2.0 - 1.0
