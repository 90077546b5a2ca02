"""dsp/signal generators: the library's host twins against the numpy restatement of generate.go (oracle/),
both on the CPU; the device kernels against both on the GPU (bit-exact for everything the hash and +,-,* produce, and for
the sweeps because host twin and kernel share csrc/siggen_core.h)."""
import ctypes as C

import numpy as np
import pytest

from algo_dsp_b200 import _lib as L, siggen as G
from oracle import siggen_oracle as SO

# known-answer vector of the uniform stream (seed 42, stream 0, indices 0..4), computed with Python integers from the
# definition in oracle/siggen_oracle.py; pins host twin, device kernel and numpy restatement to the same stream
KAT_SEED42 = [SO._mix64_int((SO.hash_key(42) + (i + 1) * SO.GOLD) & SO.M64) >> 11 for i in range(5)]


def test_uniform_stream_known_answer_and_range():
    u = G.uniform(5, seed=42)
    assert [int(v * 2.0 ** 53) for v in u] == KAT_SEED42
    assert np.array_equal(u, SO.uniform(5, seed=42))
    big = G.uniform(200000, seed=3)
    assert big.min() >= 0.0 and big.max() < 1.0 and abs(big.mean() - 0.5) < 5e-3 and abs(big.var() - 1 / 12) < 2e-3
    # index0: any shard of the stream equals the slice of the whole
    assert np.array_equal(G.uniform(1000, seed=3, index0=12345), big[12345:13345])
    # streams of neighbouring seeds are uncorrelated
    assert abs(np.corrcoef(G.uniform(100000, seed=10), G.uniform(100000, seed=11))[0, 1]) < 0.02


def test_white_and_decaying_ir_bit_exact_vs_numpy():
    assert np.array_equal(G.white(50000, seed=9, amp=0.75), SO.white(50000, seed=9, amp=0.75))
    assert np.array_equal(G.white(777, seed=9, index0=40000), SO.white(50000, seed=9)[40000:40777])
    h, ho = G.decaying_ir(96000), SO.decaying_ir(96000)
    assert G.rel_l2(h, ho) <= 1e-15 and abs(abs(h[-1]) / abs(2 * G.uniform(96000, seed=7)[-1] - 1) - 1e-3) < 1e-6   # -60 dB at the last tap


def test_pink_matches_the_reference_loop():
    n = 30000
    p = G.pink(n, seed=5, amp=0.5)
    assert np.array_equal(p, SO.pink(n, seed=5, amp=0.5))       # same additions in the same order: exact
    # a shard started in the middle of the stream carries the right band state in
    assert np.array_equal(G.pink(5000, seed=5, amp=0.5, index0=20000), p[20000:25000])
    assert np.max(np.abs(p)) <= 0.5 and np.std(p) > 0.05


def test_sweeps_against_numpy_and_long_double():
    n = 1 << 18
    for host, ref in ((G.log_sweep, SO.log_sweep), (G.linear_sweep, SO.linear_sweep)):
        y = host(n, 20.0, 20000.0, 48000.0, 0.8)
        truth = ref(n, 20.0, 20000.0, 48000.0, 0.8, dtype=np.longdouble)
        libm = ref(n, 20.0, 20000.0, 48000.0, 0.8)
        # phase up to ~1e5 rad: one rounding of the phase is ~1e-11; both float64 evaluations sit that close to the truth
        assert np.max(np.abs(y - truth)) <= 2e-10 and np.max(np.abs(libm - truth)) <= 2e-10
        assert np.array_equal(host(1000, 20.0, 20000.0, 48000.0, 0.8, index0=5000, total=n), y[5000:6000])
    # k == 0 (f1 == f0) is a plain sine (generate.go:174-176)
    s = G.log_sweep(4800, 1000.0, 1000.0, 48000.0, 1.0)
    assert np.max(np.abs(s - np.sin(2 * np.pi * 1000.0 * np.arange(4800) / 48000.0))) <= 1e-11


def test_delay_hash():
    d = [G.delay_of(p) for p in range(64)]
    assert d == [SO.delay_of(p) for p in range(64)] and all(0 <= v < 4096 for v in d) and len(set(d)) > 48


# ---------------------------------------------------------------- device kernels
@pytest.mark.gpu
@pytest.mark.parametrize("prec", [L.F64, L.F32])
def test_device_generators_equal_host_twins(conv, prec):
    ctx = conv.Context(0)
    dt = np.float64 if prec == L.F64 else np.float32
    rows, n = 3, 70001
    cast = (lambda a: a) if prec == L.F64 else (lambda a: a.astype(np.float32))
    buf = G.DeviceArray(ctx, rows, n, dt)
    G.uniform_device(ctx, buf.ptr, n, rows, buf.stride, seed0=4, seed_step=10, index0=99, prec=prec)
    got = buf.get()
    for r in range(rows):
        assert np.array_equal(got[r], cast(G.uniform(n, seed=4 + 10 * r, index0=99)))
    G.white_device(ctx, buf.ptr, n, rows, buf.stride, amp=0.3, seed0=1, seed_step=1, index0=0, prec=prec)
    got = buf.get()
    for r in range(rows):
        assert np.array_equal(got[r], cast(G.white(n, seed=1 + r, amp=0.3)))
    G.pink_device(ctx, buf.ptr, n, rows, buf.stride, amp=0.9, seed0=100, seed_step=1, index0=0, prec=prec)
    got = buf.get()
    for r in range(rows):
        assert np.array_equal(got[r], cast(G.pink(n, seed=100 + r, amp=0.9)))
    # a shard deep inside the stream: the carry-in search has to look back across tile boundaries
    G.pink_device(ctx, buf.ptr, 5000, 1, buf.stride, amp=0.9, seed0=100, index0=65432, prec=prec)
    assert np.array_equal(buf.get(0, 1, 0, 5000)[0], cast(G.pink(5000, seed=100, amp=0.9, index0=65432)))
    G.decaying_ir_device(ctx, buf.ptr, n, 1, buf.stride, prec=prec)
    assert np.array_equal(buf.get(0, 1)[0], cast(G.decaying_ir(n)))
    G.log_sweep_device(ctx, buf.ptr, n, prec=prec)
    assert np.array_equal(buf.get(0, 1)[0], cast(G.log_sweep(n)))
    G.linear_sweep_device(ctx, buf.ptr, 3000, index0=1000, total=n, prec=prec)
    assert np.array_equal(buf.get(0, 1, 0, 3000)[0], cast(G.linear_sweep(3000, index0=1000, total=n)))
    buf.free()


@pytest.mark.gpu
def test_device_delay_mix_normalize_remove_dc(conv):
    ctx = conv.Context(0)
    n, rows = 40000, 5
    src = G.DeviceArray(ctx, 1, n)
    G.log_sweep_device(ctx, src.ptr, n)
    sweep = G.log_sweep(n)
    out = G.DeviceArray(ctx, rows, n)
    dl = G.DeviceArray(ctx, 1, rows, np.int64)
    G.delay_mix_device(ctx, out.ptr, n, rows, out.stride, src.ptr, noise_amp=0.01, seed0=1000, delay_seed=0, delay_mod=4096, delays_ptr=dl.ptr)
    got, delays = out.get(), dl.get()[0]
    for r in range(rows):
        d = G.delay_of(r)
        assert delays[r] == d
        ref = np.zeros(n)
        ref[d:] = sweep[: n - d]
        ref = ref + G.white(n, seed=1000 + r, amp=0.01)
        assert np.array_equal(got[r], ref)
    # Normalize / RemoveDC against the reference formulas
    nrm = G.DeviceArray(ctx, rows, n)
    G.normalize_device(ctx, out.ptr, n, rows, out.stride, 0.5, nrm.ptr, nrm.stride)
    gn = nrm.get()
    for r in range(rows):
        assert np.array_equal(gn[r], SO.normalize(got[r], 0.5)) and abs(np.max(np.abs(gn[r])) - 0.5) < 1e-15
    G.remove_dc_device(ctx, out.ptr, n, rows, out.stride, nrm.ptr, nrm.stride)
    gd = nrm.get()
    for r in range(rows):
        assert np.max(np.abs(gd[r] - SO.remove_dc(got[r]))) <= 1e-14 and abs(gd[r].mean()) < 1e-15
    # all-zero rows normalise to zeros (generate.go:272-274)
    z = G.DeviceArray(ctx, 1, 100)
    G.white_device(ctx, z.ptr, 100, amp=0.0)
    G.normalize_device(ctx, z.ptr, 100, 1, z.stride, 1.0, z.ptr, z.stride)
    assert not z.get().any()
