"""Deterministic synthetic inputs shared by the oracle and the GPU path.

The formulas follow the reference's dsp/signal generators (generate.go:157-250); the uniform
stream is numpy's PCG64 (Go's math/rand v1 table is not reproducible here), so "identical
inputs" means: one generator, seeded, feeds both sides."""
import numpy as np


def white(n, seed=1, amp=1.0):
    """(u*2-1)*amp -- generate.go:199-202."""
    return (np.random.default_rng(seed).random(n) * 2.0 - 1.0) * amp


def pink(n, seed=1, amp=1.0):
    """Voss-McCartney, 5 bands -- generate.go:210-245."""
    pA = np.array([0.23980, 0.18727, 0.16380, 0.194685, 0.214463])
    pSUM = np.array([0.00198, 0.01478, 0.06378, 0.23378, 0.91578])
    rng = np.random.default_rng(seed)
    u = rng.random((n, 2))
    val = u[:, 1] * 2 - 1
    band = np.searchsorted(pSUM, u[:, 0], side="left")  # first b with ur1 <= pSUM[b]; 5 = none
    out = np.zeros(n)
    contrib = np.zeros(5)
    # vectorised hold-last-value per band
    for b in range(5):
        hit = band == b
        idx = np.where(hit, np.arange(n), -1)
        last = np.maximum.accumulate(idx)
        v = np.where(last >= 0, val[np.maximum(last, 0)] * pA[b], 0.0)
        out += v
    return out * amp


def log_sweep(n, f0=20.0, f1=20000.0, fs=48000.0, amp=1.0):
    """sin(2*pi*f0*(exp(k t)-1)/k), k = ln(f1/f0)/T -- generate.go:157-185."""
    t = np.arange(n) / fs
    k = np.log(f1 / f0) / (n / fs)
    return amp * np.sin(2 * np.pi * f0 * (np.expm1(k * t) / k))


def decaying_ir(K, seed=7):
    """h[i] = (u_i*2-1) * 10^(-3 i / K): -60 dB at the last tap."""
    u = np.random.default_rng(seed).random(K) * 2 - 1
    return u * 10.0 ** (-3.0 * np.arange(K) / K)


def exp_kernel(K, r=0.99):
    """makeImpulseKernel -- partitioned_test.go:11-20."""
    return r ** np.arange(K)


def test_kernel(n):
    """makeTestKernel -- conv_bench_test.go:296-312 (Hann-windowed sinc)."""
    i = np.arange(n, dtype=np.float64)
    x = i - (n - 1) / 2
    with np.errstate(divide="ignore", invalid="ignore"):
        k = np.where(x == 0, 1.0, np.sin(np.pi * x / 4) / (np.pi * x / 4))
    return k * 0.5 * (1 - np.cos(2 * np.pi * i / (n - 1)))


def rel_l2(y, ref):
    y = np.asarray(y, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    d = np.linalg.norm(y - ref)
    r = np.linalg.norm(ref)
    return d / r if r > 0 else d
