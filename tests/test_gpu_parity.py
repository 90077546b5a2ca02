"""Parity of the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs,
against the committed golden fixtures, and -- at BASELINE.json's full sizes -- through
size-independent properties.  Tolerances are the north star's: fp64 <= 1e-12 relative L2,
fp32 <= 1e-5 relative L2."""
import os

import numpy as np
import pytest

from algo_dsp_b200 import siggen as G

pytestmark = pytest.mark.gpu

TOL64 = 1e-12
TOL32 = 1e-5
FIX = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_fixtures.npz"))


def rel(y, ref):
    return G.rel_l2(y, ref)


# ---------------------------------------------------------------- golden fixtures
@pytest.mark.parametrize("name", ["ols_k300_n5000", "ols_k4097_n20000", "ols_k1_n100", "ols_k900_n100"])
def test_fixture_overlap_save(conv, name):
    h, x, y = FIX[name + "_h"], FIX[name + "_x"], FIX[name + "_y"]
    assert rel(conv.NewOverlapSave(h, 0).Process(x), y) <= TOL64
    assert rel(conv.NewOverlapAdd(h, 0).Process(x), y) <= TOL64
    assert rel(conv.OverlapSaveConvolve(x, h), y) <= TOL64


def test_fixture_direct_correlate_partitioned(conv):
    assert rel(conv.Convolve(FIX["direct_k64_x"], FIX["direct_k64_h"]), FIX["direct_k64_y"]) <= TOL64
    c = conv.Correlate(FIX["corr_a"], FIX["corr_b"])
    assert rel(c, FIX["corr_y"]) <= TOL64
    idx, val = conv.FindPeak(c)
    assert idx == int(FIX["corr_peak"][0]) and abs(val - FIX["corr_peak"][1]) <= 1e-9 * abs(val)
    p = conv.NewPartitionedConvolution(FIX["part_h"], 6, 13)
    out = np.zeros(len(FIX["part_x"]))
    p.ProcessBlock(FIX["part_x"], out)
    assert rel(out, FIX["part_y"]) <= TOL64
    y32 = conv.NewOverlapSave(FIX["ols32_h"], 0, dtype=np.float32).Process(FIX["ols32_x"])
    assert y32.dtype == np.float32 and rel(y32, FIX["ols32_y"]) <= TOL32


# ---------------------------------------------------------------- oracle on seeded inputs
OLS_SHAPES = [(1, 1), (1, 777), (2, 3), (3, 1000), (64, 4096), (65, 1000), (100, 99), (257, 5000), (1000, 30000), (1025, 10000),
              (2049, 100000), (5000, 50000), (9000, 9000), (20000, 100000), (33000, 10), (50000, 300000), (131073, 300000)]


@pytest.mark.parametrize("K,n", OLS_SHAPES)
def test_overlap_save_and_add_vs_oracle(conv, oracle, K, n):
    h, x = G.decaying_ir(K, seed=7), G.white(n, seed=K + n)
    ref = oracle.overlap_save(h, 0, x)
    ols = conv.NewOverlapSave(h, 0)
    assert (ols.FFTSize(), ols.StepSize()) == oracle.ols_sizes(K, 0)       # getters report the reference's values
    y = ols.Process(x)
    assert len(y) == n + K - 1 and rel(y, ref) <= TOL64
    ola = conv.NewOverlapAdd(h, 0)
    assert (ola.BlockSize(), ola.FFTSize()) == oracle.ola_sizes(K, 0)
    assert rel(ola.Process(x), oracle.overlap_add(h, 0, x)) <= TOL64
    # Process is stateless across calls (overlap_save.go:136-138): a second call gives the same answer
    assert np.array_equal(ols.Process(x), y)


def test_config1_ols_96k_taps(conv, oracle):
    """BASELINE config 1: 10 s white noise @48 kHz, 96 000-tap decaying IR, mono f64."""
    h, x = G.decaying_ir(96000), G.white(480000, seed=1)
    ols = conv.NewOverlapSave(h, 0)
    assert (ols.FFTSize(), ols.StepSize(), ols.KernelLen()) == (262144, 166145, 96000)
    y = ols.Process(x)
    assert len(y) == 575999
    assert rel(y, oracle.overlap_save(h, 0, x)) <= TOL64
    y32 = conv.NewOverlapSave(h, 0, dtype=np.float32).Process(x)
    assert rel(y32, oracle.overlap_save(h, 0, x)) <= TOL32


def test_user_fft_sizes_and_errors(conv, oracle):
    h, x = G.decaying_ir(300), G.white(5000, seed=3)
    for f in (1024, 4096, 256):            # 256 < 2K is silently raised (overlap_save.go:71-73)
        c = conv.NewOverlapSave(h, f)
        assert (c.FFTSize(), c.StepSize()) == oracle.ols_sizes(300, f)
        assert rel(c.Process(x), oracle.overlap_save(h, f, x)) <= TOL64
    for b in (32, 100, 4096):              # OLA block size need not be a power of two
        c = conv.NewOverlapAdd(h, b)
        assert (c.BlockSize(), c.FFTSize()) == oracle.ola_sizes(300, b)
        assert rel(c.Process(x), oracle.overlap_add(h, b, x)) <= TOL64
    with pytest.raises(conv.ConvError) as ei:
        conv.NewOverlapSave(h, 1000)
    assert conv.errors_is(ei.value, conv.ErrInvalidBlockSize)
    with pytest.raises(conv.ConvError) as ei:
        conv.NewOverlapSave([], 0)
    assert conv.errors_is(ei.value, conv.ErrEmptyKernel)
    with pytest.raises(conv.ConvError) as ei:
        conv.NewOverlapSave(h, 0).Process([])
    assert conv.errors_is(ei.value, conv.ErrEmptyInput)
    with pytest.raises(conv.ConvError) as ei:
        conv.NewOverlapSave(h, 0).ProcessTo(np.zeros(5), x)
    assert conv.errors_is(ei.value, conv.ErrLengthMismatch)


@pytest.mark.parametrize("n,m", [(10, 1), (1000, 3), (4096, 16), (5000, 15), (100000, 64), (64, 100000), (70, 65), (3, 3)])
def test_convolve_auto_select_vs_oracle(conv, oracle, n, m):
    """Convolve: swap so the longer operand is the signal, direct iff the shorter has <= 64 taps."""
    a, b = G.white(n, seed=n), G.white(m, seed=m + 1)
    ref = oracle.convolve(a, b)
    assert rel(conv.Convolve(a, b), ref) <= TOL64
    assert rel(conv.Direct(a, b), oracle.direct(a, b)) <= TOL64
    for mode in (conv.ModeFull, conv.ModeSame, conv.ModeValid):
        got, want = conv.ConvolveMode(a, b, mode), oracle.convolve_mode(a, b, mode)
        assert len(got) == len(want) and rel(got, want) <= TOL64


def test_direct_long_kernel_and_f32(conv, oracle):
    a, b = G.white(3000, seed=1), G.white(700, seed=2)
    assert rel(conv.Direct(a, b), oracle.direct(a, b)) <= TOL64          # Direct is legal for any kernel length
    a32, b32 = a.astype(np.float32), G.test_kernel(64).astype(np.float32)
    assert rel(conv.Direct32(a32, b32), oracle.direct(a32, b32, np.float32)) <= TOL32
    assert rel(conv.Convolve32(a32, b.astype(np.float32)), oracle.convolve(a, b)) <= TOL32


@pytest.mark.parametrize("m", [63, 64, 65, 128, 255, 256, 257, 1000, 1024, 1025, 1300])
def test_direct_kernel_families_by_tap_count(conv, oracle, m):
    """Up to 64 taps, 65..256, 257..1024 (taps as kernel parameters) and beyond (taps in shared memory), shared and per-channel
    kernels, both precisions; a tile boundary inside the result (1024 outputs per CTA) and a signal shorter than the kernel."""
    ch, n = 3, 2500
    x = np.stack([G.white(n, seed=10 + c) for c in range(ch)])
    k = G.white(m, seed=99)
    y = conv.DirectBatch(x, k)
    assert y.shape == (ch, n + m - 1)
    for c in range(ch):
        assert rel(y[c], oracle.direct(x[c], k)) <= TOL64
    ks = np.stack([k * (c + 1) for c in range(ch)])               # per-channel kernels: the shared-memory-tap kernel
    y2 = conv.DirectBatch(x, ks)
    assert np.array_equal(y2[0], y[0])                            # same accumulation order in both kernels: bit identical
    assert rel(y2[2], oracle.direct(x[2], ks[2])) <= TOL64
    short = G.white(40, seed=5)
    assert rel(conv.Direct(short, k), oracle.direct(short, k)) <= TOL64
    assert rel(conv.Direct32(x[1].astype(np.float32), k.astype(np.float32)), oracle.direct(x[1], k, np.float32)) <= TOL32


def test_direct_batch_config2_shape(conv, oracle):
    """Config 2 shape (channels x samples, 64-tap Hann-windowed sinc), reduced channel count."""
    ch, n = 8, 1 << 16
    x = np.stack([G.white(n, seed=1 + c) for c in range(ch)])
    k = G.test_kernel(64)
    y = conv.DirectBatch(x, k)
    assert y.shape == (ch, n + 63)
    for c in (0, ch - 1):
        assert rel(y[c], oracle.direct(x[c], k)) <= TOL64
    # per-channel kernels
    ks = np.stack([k * (c + 1) for c in range(ch)])
    y2 = conv.DirectBatch(x, ks)
    assert rel(y2[3], oracle.direct(x[3], ks[3])) <= TOL64


@pytest.mark.parametrize("n,m", [(11, 5), (3000, 700), (700, 3000), (100000, 50), (65536, 65536)])
def test_correlation_family_vs_oracle(conv, oracle, n, m):
    a, b = G.white(n, seed=n + 5), G.white(m, seed=m + 6)
    ref = oracle.correlate(a, b)
    got = conv.Correlate(a, b)
    assert len(got) == n + m - 1 and rel(got, ref) <= TOL64
    assert rel(conv.CorrelateFFT(a, b), oracle.correlate_fft(a, b)) <= TOL64
    assert rel(conv.CorrelateNormalized(a, b), oracle.correlate_normalized(a, b)) <= TOL64
    assert conv.FindPeak(got)[0] == oracle.find_peak(ref)[0]
    if n * m <= 3_000_000:
        assert rel(conv.CorrelateDirect(a, b), oracle.correlate_direct(a, b)) <= TOL64
    for mode in (conv.ModeSame, conv.ModeValid):
        assert rel(conv.CorrelateMode(a, b, mode), oracle.correlate_mode(a, b, mode)) <= TOL64
    assert rel(conv.AutoCorrelateNormalized(a), oracle.auto_correlate_normalized(a)) <= TOL64


def test_find_peak_rules(conv):
    assert conv.FindPeak([]) == (-1, 0.0)                       # correlate.go:201-203
    assert conv.FindPeak([3, 7, 7, 1]) == (1, 7.0)              # first maximum wins (strict >)
    assert conv.FindPeak([-5, -2, -9]) == (1, -2.0)             # signed, not absolute
    x = np.zeros(1 << 20)
    x[[12345, 999999]] = 2.5
    assert conv.FindPeak(x) == (12345, 2.5)
    assert conv.FindPeak([1.0]) == (0, 1.0)


def test_correlate_batch_peak_lag_config4_shape(conv, oracle):
    """Config 4 shape, reduced: sweep/response pairs, delay recovered exactly from the peak lag."""
    n, pairs = 1 << 16, 4
    sweep = G.log_sweep(n)
    a = np.zeros((pairs, n))
    delays = [0, 17, 1000, 4095]
    for p, d in enumerate(delays):
        a[p, d:] = sweep[: n - d]
        a[p] += G.white(n, seed=1000 + p, amp=0.01)
    b = np.tile(sweep, (pairs, 1))
    out, pi, pv = conv.CorrelateBatch(a, b)
    for p, d in enumerate(delays):
        assert conv.LagFromIndex(int(pi[p]), n) == d
    assert rel(out[2], oracle.correlate(a[2], b[2])) <= TOL64
    _, pi2, _ = conv.CorrelateBatch(a, b, want_output=False)
    assert np.array_equal(pi, pi2)


@pytest.mark.parametrize("n,m,pairs", [(5000, 5000, 9), (3000, 700, 40), (40000, 40000, 5), (1 << 18, 1 << 18, 3)])
def test_correlate_batch_groups_of_pairs(conv, oracle, n, m, pairs):
    """Many pairs per call: groups of pairs share launches and two pairs share an inverse transform; odd counts leave
    the last transform half empty.  Every pair must match the single-pair result."""
    a = np.stack([G.white(n, seed=p) for p in range(pairs)])
    b = np.stack([G.pink(m, seed=50 + p) for p in range(pairs)])
    out, pi, pv = conv.CorrelateBatch(a, b)
    for p in range(pairs):
        ref = oracle.correlate(a[p], b[p])
        assert rel(out[p], ref) <= TOL64
        assert int(pi[p]) == oracle.find_peak(ref)[0]


# ---------------------------------------------------------------- partitioned / streaming semantics
@pytest.mark.parametrize("K,n,mn,mx,chunk", [(64, 512, 4, 10, 0), (1024, 4096, 6, 13, 0), (8192, 16384, 6, 13, 1000),
                                              (3000, 9000, 7, 13, 128), (96000, 200000, 7, 13, 48000)])
def test_partitioned_vs_oracle(conv, oracle, K, n, mn, mx, chunk):
    """ProcessBlock (partitioned.go:348-396): conv delayed by 2^minOrder, arbitrary call sizes."""
    h, x = G.exp_kernel(K, 0.9999 if K > 10000 else 0.99), G.white(n, seed=K)
    ref = oracle.Partitioned(h, mn, mx).process_block(x)
    p = conv.NewPartitionedConvolution(h, mn, mx)
    assert p.Latency() == 1 << mn and p.KernelLen() == K
    if chunk == 0:
        out = np.zeros(n)
        p.ProcessBlock(x, out)
    else:
        parts = []
        for i in range(0, n, chunk):
            o = np.zeros(len(x[i:i + chunk]))
            p.ProcessBlock(x[i:i + chunk], o)
            parts.append(o)
        out = np.concatenate(parts)
    assert rel(out, ref) <= TOL64
    p.Reset()
    out2 = np.zeros(n)
    p.ProcessBlock(x, out2)
    assert rel(out2, ref) <= TOL64


def test_partitioned_f32(conv, oracle):
    h, x = G.exp_kernel(2000).astype(np.float32), G.white(8000, seed=9).astype(np.float32)
    ref = oracle.Partitioned(h.astype(np.float64), 6, 13).process_block(x.astype(np.float64))
    p = conv.NewPartitionedConvolution32(h, 6, 13)
    out = np.zeros(len(x), np.float32)
    p.ProcessBlock(x, out)
    assert rel(out, ref) <= TOL32


# ---------------------------------------------------------------- full-size properties (BASELINE sizes)
def test_full_size_batch_properties(conv, oracle):
    """Bench workload shape (96k taps, many channels): spot-check channels against the oracle and
    check linearity / DC-gain / impulse identities on the whole batch."""
    K, n, ch = 96000, 480000, 16
    h = G.decaying_ir(K)
    x = np.stack([G.white(n, seed=1 + c) for c in range(ch)])
    ols = conv.NewOverlapSave(h, 0)
    y = ols.ProcessBatch(x)
    assert y.shape == (ch, n + K - 1)
    for c in (0, 7, ch - 1):
        assert rel(y[c], oracle.overlap_save(h, 0, x[c])) <= TOL64
    # DC gain: sum(y) = sum(x) * sum(h)
    assert np.allclose(y.sum(axis=1), x.sum(axis=1) * h.sum(), rtol=1e-9, atol=1e-6)
    # linearity across channels: conv(2*x0 - 3*x1) == 2*y0 - 3*y1
    z = ols.Process(2 * x[0] - 3 * x[1])
    assert rel(z, 2 * y[0] - 3 * y[1]) <= 1e-12
    # impulse at t0 returns the IR shifted by t0
    d = np.zeros(n)
    d[12345] = 1.0
    yi = ols.Process(d)
    assert rel(yi[12345:12345 + K], h) <= 1e-12 and np.max(np.abs(yi[:12345])) <= 1e-13


def test_large_batch_groups_and_streams(conv, oracle, monkeypatch):
    """Large batches are cut into L2-sized groups of block pairs rotated over worker streams; the
    grouping must not change the results (also: odd block counts, fp32, host-chunked pipeline)."""
    K, n, ch = 20000, 150000, 48
    h = G.decaying_ir(K)
    x = np.stack([G.white(n, seed=50 + c) for c in range(ch)])
    ols = conv.NewOverlapSave(h, 0)
    monkeypatch.setenv("ADSP_PIPE_CHUNK_MB", "4096")   # one chunk: the whole batch in a single engine call
    y_one = ols.ProcessBatch(x)
    monkeypatch.setenv("ADSP_SCRATCH_MB", "8")         # tiny scratch budget -> many small groups
    y_small = ols.ProcessBatch(x)
    monkeypatch.delenv("ADSP_SCRATCH_MB")
    monkeypatch.setenv("ADSP_PIPE_CHUNK_MB", "16")     # host pipeline in many chunks
    y_chunks = ols.ProcessBatch(x)
    # two real blocks share one complex transform, so a block's rounding depends (at the 1e-16 level)
    # on its partner: regrouping is not bit-identical, only far inside the tolerance
    assert rel(y_small, y_one) <= 1e-14 and rel(y_chunks, y_one) <= 1e-14
    for c in (0, 17, ch - 1):
        assert rel(y_one[c], oracle.overlap_save(h, 0, x[c])) <= TOL64
    y2 = conv.NewOverlapSave(h, 0).ProcessBatch(x[:33])
    assert rel(y2, y_one[:33]) <= 1e-14
    y32 = conv.NewOverlapSave(h, 0, dtype=np.float32).ProcessBatch(x[:33].astype(np.float32))
    assert rel(y32[5], oracle.overlap_save(h, 0, x[5])) <= TOL32


def test_long_kernel_config5_shape(conv, oracle):
    """Kernel longer than half of the preferred 2^20 transform (config 5 shape, reduced): one 2^22-point
    transform; and, forced back to 2^20, the sum over IR partitions."""
    K, n = 600000, 700000
    h, x = G.decaying_ir(K), G.white(n, seed=5)
    ref = oracle.overlap_save(h, 0, x)
    c = conv.NewOverlapSave(h, 0)
    assert c.internal_geometry()["fft_n"] == 1 << 22 and c.internal_geometry()["partitions"] == 1
    assert rel(c.Process(x), ref) <= TOL64


def test_long_kernel_partition_sum(conv, oracle, monkeypatch):
    monkeypatch.setenv("ADSP_MAX_FFT", str(1 << 20))
    K, n = 600000, 700000
    h, x = G.decaying_ir(K), G.white(n, seed=5)
    c = conv.NewOverlapSave(h, 0)
    assert c.internal_geometry()["partitions"] >= 2
    assert rel(c.Process(x), oracle.overlap_save(h, 0, x)) <= TOL64
    monkeypatch.setenv("ADSP_MAX_FFT", str(1 << 13))          # many partitions, small transforms
    K, n = 30000, 50000
    h, x = G.decaying_ir(K), G.white(n, seed=6)
    c = conv.NewOverlapSave(h, 0)
    assert c.internal_geometry()["partitions"] >= 7
    assert rel(c.Process(x), oracle.overlap_save(h, 0, x)) <= TOL64
    y32 = conv.NewOverlapSave(h, 0, dtype=np.float32).Process(x)
    assert rel(y32, oracle.overlap_save(h, 0, x)) <= TOL32


def test_time_block_sharding_with_halo(conv, oracle):
    """Config 5 decomposition: a long signal cut into time shards with a K-1 halo reproduces the
    unsharded result exactly where shards abut."""
    K, n, shards = 5000, 400000, 4
    h, x = G.decaying_ir(K), G.white(n, seed=11)
    full = conv.NewOverlapSave(h, 0).Process(x)
    S = -(-n // shards)
    c = conv.NewOverlapSave(h, 0)
    out = np.zeros(n + K - 1)
    for s in range(shards):
        lo, hi = s * S, min((s + 1) * S, n)
        seg = x[max(0, lo - (K - 1)):hi]
        ys = c.Process(seg)
        skip = lo - max(0, lo - (K - 1))
        last = s == shards - 1
        take = (hi - lo) + (K - 1 if last else 0)
        out[lo:lo + take] = ys[skip:skip + take]
    assert rel(out, full) <= 1e-13
    assert rel(out, oracle.overlap_save(h, 0, x)) <= TOL64


# ---------------------------------------------------------------- mixed-radix transform lengths (P * 2^k, P in {3,5,7,9})
@pytest.mark.parametrize("N,n1", [(16 * 3 * 256, 48), (16 * 5 * 256, 80), (16 * 7 * 512, 112), (16 * 9 * 256, 144), (16 * 9 * 1024, 144),
                                   (16 * 3 * 2048, 48), (9 << 15, 144), (3 << 16, 96), (5 << 16, 160), (7 << 16, 224), (9 << 16, 288),
                                   (3 << 17, 96), (5 << 17, 160), (7 << 17, 224)])
def test_forced_odd_transform_lengths_overlap_save(conv, oracle, monkeypatch, N, n1):
    """Overlap-save blocks (with discard) on every mixed-radix column shape (N1 = 16*P and 32*P); forced
    because the planner only picks these lengths where they win."""
    K = max(2, N // 5)
    n = int(2.6 * N)
    monkeypatch.setenv("ADSP_FFT_N", str(N))
    h, x = G.decaying_ir(K, seed=n1), G.white(n, seed=N)
    ref = oracle.overlap_save(h, 0, x)
    ols = conv.NewOverlapSave(h, 0)
    geom = ols.internal_geometry()
    assert geom["fft_n"] == N and geom["n1"] == n1
    assert rel(ols.Process(x), ref) <= TOL64
    y32 = conv.NewOverlapSave(h, 0, dtype=np.float32).Process(x)
    assert rel(y32, ref) <= TOL32


def test_mixed_radix_16p_columns_at_2_16(conv, oracle, monkeypatch):
    """P * 2^16 through the N1 = 16*P kernels with 4096-point rows (the shape the 32*P kernels replaced)."""
    monkeypatch.setenv("ADSP_MR_NO_2P", "1")
    K, n = 96000, 480000
    h, x = G.decaying_ir(K), G.white(n, seed=1)
    plan = conv.NewOverlapSave(h, 0)
    assert rel(plan.Process(x), oracle.overlap_save(h, 0, x)) <= TOL64


@pytest.mark.parametrize("K,n", [(2000, 9000), (3000, 17000), (96000, 480000), (30000, 199000), (50000, 390000), (7000, 49000)])
def test_single_zero_padded_block_lengths(conv, oracle, K, n):
    """Shapes whose full result fits one P*2^k transform run as a single zero-padded block
    (planner option B); results and lengths are unchanged (overlap_save.go:146-251)."""
    h, x = G.decaying_ir(K, seed=11), G.white(n, seed=K)
    ref = oracle.overlap_save(h, 0, x)
    plan = conv.NewOverlapSave(h, 0)
    cover = plan.describe_cover(n)
    assert len(cover) == 1 and cover[0]["single_block"] and cover[0]["fft_n"] >= n + K - 1
    if (K, n) == (96000, 480000):
        assert cover[0]["fft_n"] == 9 * 65536          # the bench shape: 589 824 points instead of 2^19 + 2^18
    y = plan.Process(x)
    assert len(y) == n + K - 1 and rel(y, ref) <= TOL64
    # batch of channels (odd count: the last block pair is half empty)
    xb = np.stack([G.white(n, seed=K + c) for c in range(3)])
    yb = conv.NewOverlapSave(h, 0).ProcessBatch(xb)
    for c in range(3):
        assert rel(yb[c], oracle.overlap_save(h, 0, xb[c])) <= TOL64


def test_one_process_many_plans_channel_and_time_sharding(conv, oracle):
    """adsp_plans_process_batch / adsp_plans_process_long: one host call drives several plans (one per context; here
    three contexts on the same device stand in for three GPUs).  Same results as a single plan, no collective."""
    K, n, channels = 3000, 40000, 7
    h = G.decaying_ir(K)
    ctxs = [conv.Context(0) for _ in range(3)]
    plans = [conv.OverlapSave(h, 0, ctx=c) for c in ctxs]
    x = np.stack([G.white(n, seed=c) for c in range(channels)])
    y = conv.ProcessBatchMulti(plans, x)
    for c in range(channels):
        assert rel(y[c], oracle.overlap_save(h, 0, x[c])) <= TOL64
    xl = G.white(300000, seed=9)
    yl = conv.ProcessLongMulti(plans, xl)
    assert len(yl) == len(xl) + K - 1 and rel(yl, oracle.overlap_save(h, 0, xl)) <= TOL64
    short = G.white(50, seed=1)                          # fewer samples than shards * alignment: trailing plans idle
    assert rel(conv.ProcessLongMulti(plans, short), oracle.overlap_save(h, 0, short)) <= TOL64


def test_tma_fed_persistent_column_kernels_parity():
    """The opt-in persistent TMA-fed mixed-radix column kernels (ADSP_MRP=1, conv_kernels_mrp.cuh: cp.async.bulk.tensor tile
    loads, mbarrier completion) give the same results as the default kernels and the oracle.  The switch is read once per
    process, so the check runs in a subprocess."""
    import subprocess
    import sys
    code = r'''
import numpy as np, sys
sys.path.insert(0, %r)
from algo_dsp_b200 import conv, siggen as G
from oracle import oracle as O
O.build()
worst = 0.0
# single zero-padded mixed-radix blocks: (K, n, channels) -> 9*2^16 (M=18), 3*2^12 .. (small M), odd channel count, signal
# ending on / off a row boundary
for K, n, ch in ((96000, 480000, 5), (3000, 17000, 3), (20000, 60000, 2), (9000, 2048 * 12, 4), (700, 11000, 1)):
    h = G.decaying_ir(K)
    x = np.stack([G.white(n, seed=10 + c) for c in range(ch)])
    plan = conv.OverlapSave(h, 0)
    y = plan.ProcessBatch(x)
    for c in (0, ch - 1):
        worst = max(worst, G.rel_l2(y[c], O.overlap_save(h, 0, x[c])))
    y32 = conv.OverlapSave(h, 0, dtype=np.float32).ProcessBatch(x.astype(np.float32))
    assert G.rel_l2(y32[0], O.overlap_save(h, 0, x[0])) <= 1e-5
    print(K, n, ch, plan.describe_cover(n))
print("WORST", worst)
assert worst <= 1e-12
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, ADSP_MRP="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "WORST" in r.stdout
