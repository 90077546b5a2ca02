"""BASELINE.json configs 2-5 at their FULL sizes on one B200, through the C ABI, with inputs generated on the device by the
library's dsp/signal generators and parity against the CPU oracle on sampled channels / windows (the oracle
gets the very same samples from the generators' bit-identical host twins, so nothing large crosses PCIe).
Tolerances: fp64 <= 1e-12 relative L2, fp32 <= 1e-5 (north star).  Whole file: about two minutes on a B200."""
import ctypes as C

import numpy as np
import pytest

from algo_dsp_b200 import _lib as L, siggen as G
from algo_dsp_b200.shard import time_shards

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

TOL64, TOL32 = 1e-12, 1e-5


@pytest.fixture(scope="module")
def ctx(conv):
    c = conv.Context(0)
    yield c
    c.close()


def _free():
    torch.cuda.empty_cache()


def test_config2_direct_64tap_1024ch_x_2e20(conv, oracle, ctx):
    """1024 channels x 2^20 white noise (seed 1 + channel), 64-tap Hann-windowed sinc: Convolve picks Direct (conv.go:209)."""
    ch, n, m = 1024, 1 << 20, 64
    x = torch.empty((ch, n), device="cuda", dtype=torch.float64)
    G.white_device(ctx, x.data_ptr(), n, ch, n, amp=1.0, seed0=1, seed_step=1)
    k = G.test_kernel(m)
    kd = torch.tensor(k, device="cuda")
    ol = n + m - 1
    y = torch.empty((ch, ol), device="cuda", dtype=torch.float64)
    n0 = ctx.launch_count()
    st = L.load().adsp_direct_batch_device(ctx.handle, x.data_ptr(), n, n, kd.data_ptr(), m, 0, ch, y.data_ptr(), ol, L.F64)
    assert st == L.OK
    ctx.sync()
    assert ctx.launch_count() - n0 == 1                      # one direct kernel over the whole batch, no FFT
    for c in (0, 517, ch - 1):
        xc = G.white(n, seed=1 + c)
        assert np.array_equal(x[c, :4096].cpu().numpy(), xc[:4096])          # device generator == host twin
        assert G.rel_l2(y[c].cpu().numpy(), oracle.convolve(xc, k)) <= TOL64
    # linearity across the whole batch (size-independent property): sum of outputs == conv(sum of inputs)
    ysum = y.sum(dim=0).cpu().numpy()
    xsum = x.sum(dim=0).cpu().numpy()
    assert G.rel_l2(ysum, oracle.convolve(xsum, k)) <= 1e-11
    del x, y
    _free()


def test_config3_reverb_64ch_x_5min_288k_taps(conv, oracle, ctx):
    """64 channels x 14.4 M samples of pink noise (seed 100 + channel) through a 288 000-tap decaying IR."""
    ch, n, K = 64, 14_400_000, 288_000
    x = torch.empty((ch, n), device="cuda", dtype=torch.float64)
    G.pink_device(ctx, x.data_ptr(), n, ch, n, amp=1.0, seed0=100, seed_step=1)
    h = G.decaying_ir(K)
    plan = conv.OverlapSave(h, 0, ctx=ctx)
    ol = n + K - 1
    ostr = (ol + 31) // 32 * 32
    y = torch.empty((ch, ostr), device="cuda", dtype=torch.float64)
    plan.process_device(x.data_ptr(), n, ch, n, y.data_ptr(), ostr)
    plan.sync()
    W = 2_000_000
    for c in (0, 41, ch - 1):
        # first 2 M outputs: inputs [0, W)
        xs = G.pink(W, seed=100 + c)
        assert np.array_equal(x[c, :8192].cpu().numpy(), xs[:8192])
        ref = oracle.overlap_save(h, 0, xs)[:W]
        assert G.rel_l2(y[c, :W].cpu().numpy(), ref) <= TOL64
        # last 2 M outputs including the K-1 tail: inputs [ol - W - (K-1), n)
        lo = ol - W - (K - 1)
        xt = G.pink(n - lo, seed=100 + c, index0=lo)
        ref = oracle.overlap_save(h, 0, xt)[K - 1:]
        assert len(ref) == W and G.rel_l2(y[c, ol - W:ol].cpu().numpy(), ref) <= TOL64
    plan.Close()
    del x, y
    _free()


def test_config4_correlate_1024_pairs_x_2e20_peak_lags(conv, oracle, ctx):
    """1024 sweep/response pairs: b = log sweep 20 Hz..20 kHz, a_p = b delayed by d_p = hash(p) mod 4096 plus white noise at
    -40 dB (seed 1000 + p); Correlate + FindPeak + LagFromIndex must return d_p exactly for every pair."""
    pairs, n = 1024, 1 << 20
    b = torch.empty((1, n), device="cuda", dtype=torch.float64)
    G.log_sweep_device(ctx, b.data_ptr(), n)
    a = torch.empty((pairs, n), device="cuda", dtype=torch.float64)
    G.delay_mix_device(ctx, a.data_ptr(), n, pairs, n, b.data_ptr(), noise_amp=0.01, seed0=1000, delay_seed=0, delay_mod=4096)
    ol = 2 * n - 1
    out = torch.empty((pairs, ol), device="cuda", dtype=torch.float64)
    pi = torch.empty(pairs, device="cuda", dtype=torch.int64)
    pv = torch.empty(pairs, device="cuda", dtype=torch.float64)
    st = L.load().adsp_correlate_batch_device(ctx.handle, a.data_ptr(), n, n, b.data_ptr(), n, 0, pairs, out.data_ptr(), ol, pi.data_ptr(), pv.data_ptr(),
                                              L.F64)
    assert st == L.OK
    ctx.sync()
    lags = pi.cpu().numpy() - (n - 1)                                  # LagFromIndex, correlate.go:221
    want = np.array([G.delay_of(p) for p in range(pairs)])
    assert np.array_equal(lags, want)
    sweep = G.log_sweep(n)
    for p in (3, pairs - 1):
        ap = np.zeros(n)
        ap[want[p]:] = sweep[: n - want[p]]
        ap = ap + G.white(n, seed=1000 + p, amp=0.01)
        assert np.array_equal(a[p, :4096].cpu().numpy(), ap[:4096])
        ref = oracle.correlate(ap, sweep)
        got = out[p].cpu().numpy()
        assert G.rel_l2(got, ref) <= TOL64
        idx, val = oracle.find_peak(ref)
        assert idx == int(pi[p]) and abs(val - float(pv[p])) <= 1e-9 * abs(val)
    # every pair with its OWN copy of b (b_stride = n): the per-pair path (packed transform + mirror-bin product), where the
    # call above (b_stride = 0, one b for all pairs) ran as a batched convolution with one cached spectrum
    bb = b.expand(pairs, n).contiguous()
    out2 = torch.empty((pairs, ol), device="cuda", dtype=torch.float64)
    pi3 = torch.empty(pairs, device="cuda", dtype=torch.int64)
    pv3 = torch.empty(pairs, device="cuda", dtype=torch.float64)
    st = L.load().adsp_correlate_batch_device(ctx.handle, a.data_ptr(), n, n, bb.data_ptr(), n, n, pairs, out2.data_ptr(), ol, pi3.data_ptr(), pv3.data_ptr(),
                                              L.F64)
    assert st == L.OK
    ctx.sync()
    assert torch.equal(pi3, pi)
    for p in (3, pairs - 1):
        assert G.rel_l2(out2[p].cpu().numpy(), out[p].cpu().numpy()) <= TOL64
    pi4 = torch.empty(pairs, device="cuda", dtype=torch.int64)
    pv4 = torch.empty(pairs, device="cuda", dtype=torch.float64)
    st = L.load().adsp_correlate_batch_device(ctx.handle, a.data_ptr(), n, n, bb.data_ptr(), n, n, pairs, None, 0, pi4.data_ptr(), pv4.data_ptr(), L.F64)
    assert st == L.OK
    ctx.sync()
    assert torch.equal(pi4, pi3) and torch.equal(pv4, pv3)
    del bb, out2
    # peaks only (out == NULL): same indices and values
    pi2 = torch.empty(pairs, device="cuda", dtype=torch.int64)
    pv2 = torch.empty(pairs, device="cuda", dtype=torch.float64)
    st = L.load().adsp_correlate_batch_device(ctx.handle, a.data_ptr(), n, n, b.data_ptr(), n, 0, pairs, None, 0, pi2.data_ptr(), pv2.data_ptr(), L.F64)
    assert st == L.OK
    ctx.sync()
    assert torch.equal(pi, pi2) and torch.equal(pv, pv2)
    del a, out
    _free()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_config5_single_signal_2e31_x_2e20_time_sharded(conv, oracle, ctx, dtype):
    """One 2^31-sample white-noise signal (seed 1) through a 2^20-tap IR: unsharded, and as the 8 time-block shards of the
    8-GPU decomposition (K-1 halo) generated shard by shard from the index-hash PRNG; windows against the oracle:
    the first and last 2^22 outputs and every shard boundary +- 2^20."""
    n, K, world = 1 << 31, 1 << 20, 8
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    prec = L.F64 if dtype == np.float64 else L.F32
    tol = TOL64 if dtype == np.float64 else TOL32
    esz = np.dtype(dtype).itemsize
    ol = n + K - 1
    h = G.decaying_ir(K)
    plan = conv.OverlapSave(h, 0, ctx=ctx, dtype=dtype)
    x = torch.empty(n, device="cuda", dtype=tdt)
    G.white_device(ctx, x.data_ptr(), n, 1, n, amp=1.0, seed0=1, prec=prec)
    y = torch.empty(ol + 32, device="cuda", dtype=tdt)
    plan.process_device(x.data_ptr(), n, 1, n, y.data_ptr(), ol + 32)
    plan.sync()
    # the 8 shards: each generates ITS OWN input (block + halo) from the stream index, as each GPU of the 8-GPU run does
    shards = time_shards(n, K, world)
    y2 = torch.zeros(ol, device="cuda", dtype=tdt)
    seg_max = max(s.in_hi - s.in_lo for s in shards)
    xs = torch.empty(seg_max, device="cuda", dtype=tdt)
    tmp = torch.empty(seg_max + K - 1 + 32, device="cuda", dtype=tdt)
    for s in shards:
        if s.out_hi <= s.out_lo:
            continue
        seg = s.in_hi - s.in_lo
        G.white_device(ctx, xs.data_ptr(), seg, 1, seg, amp=1.0, seed0=1, index0=s.in_lo, prec=prec)
        plan.process_device(xs.data_ptr(), seg, 1, seg, tmp.data_ptr(), seg + K - 1 + 32)
        plan.sync()
        y2[s.out_lo:s.out_hi] = tmp[s.skip:s.skip + (s.out_hi - s.out_lo)]
    torch.cuda.synchronize()
    W = 1 << 22
    wins = [(0, W), (ol - W, ol)] + [(s.out_lo - (1 << 20), s.out_lo + (1 << 20)) for s in shards[1:] if s.out_hi > s.out_lo]
    assert len(wins) == 9
    worst = 0.0
    for a, b in wins:
        lo = max(0, a - (K - 1))
        xin = G.white(min(b, n) - lo, seed=1, index0=lo).astype(dtype).astype(np.float64)   # host twin of the device generator
        ref = oracle.overlap_save(h.astype(dtype).astype(np.float64), 0, xin)[a - lo:a - lo + (b - a)]
        for yy in (y, y2):
            worst = max(worst, G.rel_l2(yy[a:b].cpu().numpy().astype(np.float64), ref))
    assert worst <= tol
    plan.Close()
    del x, y, y2, xs, tmp
    _free()
