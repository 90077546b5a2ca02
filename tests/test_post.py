"""The step after the dsp/conv path: measure/ir, measure/sweep, dsp/filter/fir, dsp/resample.
CPU part: the oracle restatement (oracle/post_oracle.py) against the known answers and properties the reference's own tests
hold (file:line in every test).  GPU part: the library (C ABI, csrc/post.cu) against that oracle."""
import ctypes as C
import math

import numpy as np
import pytest

from algo_dsp_b200 import _lib as L, siggen as G
from oracle import post_oracle as PO


def exp_decay(sr, rt60, dur):
    """makeExponentialDecay, measure/ir/ir_test.go:10-22: amplitude falls 60 dB in rt60 seconds."""
    n = int(sr * dur)
    tau = rt60 / (3 * math.log(10))
    return np.exp(-np.arange(n) / sr / tau)


def sine(f, sr, n):
    return np.sin(2 * np.pi * f * np.arange(n) / sr)


# ---------------------------------------------------------------- oracle vs the reference's tests (CPU)
def test_oracle_schroeder_properties():
    """TestSchroederIntegral, ir_test.go:62-105."""
    ir = exp_decay(48000.0, 1.0, 3.0)[:48000]          # a third of the test's length keeps the pure-Python loop short
    s = PO.schroeder_integral(ir)
    assert len(s) == len(ir) and abs(s[0]) <= 0.01
    assert np.all(np.diff(s) <= 0.001)
    assert s[len(s) // 4] < -5 and s[len(s) // 2] < s[len(s) // 4]
    assert not PO.schroeder_integral(np.zeros(10)).any()      # no energy: raw sums (ir.go:115-118)


def test_oracle_find_impulse_start_kats():
    """TestFindImpulseStart, ir_test.go:344-410."""
    ir = np.zeros(1000); ir[0] = 1.0
    assert PO.find_impulse_start(ir) == 0
    ir = np.zeros(10000); ir[5000] = 1.0; ir[5001] = 0.5
    assert PO.find_impulse_start(ir) == 5000 and PO.find_peak(ir) == 5000
    ir = np.zeros(10000)
    ir[:5000] = 0.001 * (np.arange(5000) % 2 * 2 - 1)
    ir[5000] = 1.0
    assert PO.find_impulse_start(ir) == 5000


def test_oracle_fir_kats():
    """TestProcessSample_MovingAverage / _Differentiator / TestProcessBlock_MatchesSample, filter_test.go:60-112."""
    f = PO.Fir([1 / 3, 1 / 3, 1 / 3])
    assert np.allclose([f.process_sample(1.0) for _ in range(5)], [1 / 3, 2 / 3, 1, 1, 1], atol=1e-12)
    f = PO.Fir([1, -1])
    assert np.allclose([f.process_sample(x) for x in (0, 1, 3, 6, 10)], [0, 1, 2, 3, 4], atol=1e-12)
    coeffs, x = [0.25, 0.5, 0.25], [1, 0.5, -0.3, 0.7, 0, -1, 0.2, 0.8]
    f1, f2 = PO.Fir(coeffs), PO.Fir(coeffs)
    ref = [f1.process_sample(v) for v in x]
    blk = list(map(float, x))
    f2.process_block(blk)
    assert np.allclose(blk, ref, atol=1e-12)
    # the >= 32-tap branch dots the coefficients with the window stored oldest first (filter.go:93-94)
    h = np.arange(1, 41, dtype=float)
    f3 = PO.Fir(h)
    imp = [1.0] + [0.0] * 39
    f3.process_block(imp)
    assert np.allclose(imp, h[::-1])


def test_oracle_resample_kats():
    """resample_test.go:20-111, resample_design_test.go:8-18."""
    assert (PO.Resampler(320, 294).up, PO.Resampler(320, 294).down) == (160, 147)
    assert PO.approximate_ratio(48000 / 44100) == (160, 147) and PO.approximate_ratio(0.5) == (1, 2)
    r = PO.Resampler(3, 2)
    x = sine(1000, 48000, 257)
    want = r.predict_output_len(len(x))
    assert len(r.process(x)) == want
    r1, r2 = PO.Resampler(160, 147), PO.Resampler(160, 147)
    x = sine(1000, 44100, 2048)
    whole = r1.process(x)
    chunked = np.concatenate([r2.process(x[i:i + 257]) for i in range(0, len(x), 257)])
    assert len(whole) == len(chunked) and np.max(np.abs(whole - chunked)) <= 1e-12
    assert abs(len(whole) - round(len(x) * 48000 / 44100)) <= 1


def test_oracle_logsweep_kats():
    """TestLogSweepGenerate / Short / InverseFilter / DeconvolveIdentity, sweep_test.go:35-169."""
    s = PO.logsweep_generate(20, 20000, 1, 48000)
    assert len(s) == 48000 and np.max(np.abs(s)) <= 1.001 and abs(s[0]) <= 1e-10
    assert len(PO.logsweep_generate(100, 1000, 0.1, 8000)) == 800
    inv = PO.logsweep_inverse_filter(100, 4000, 0.5, 16000)
    assert len(inv) == 8000 and np.max(np.abs(inv)) > 0
    sw = PO.logsweep_generate(100, 4000, 0.25, 16000)
    ir = PO.deconvolve_with_inverse(sw, PO.logsweep_inverse_filter(100, 4000, 0.25, 16000))
    pk = np.max(np.abs(ir))
    assert 10 * np.log10(pk * pk / np.mean(ir * ir)) >= 15


def test_host_twins_of_the_sweep_match_the_oracle():
    """The library's LogSweep.Generate / InverseFilter host twins (same arithmetic as the kernels) against the oracle and a
    long-double evaluation: the phase reaches ~2e4 rad here, so one rounding of it is ~2e-12."""
    from algo_dsp_b200 import post
    for args in ((20.0, 20000.0, 1.0, 48000.0), (100.0, 4000.0, 0.5, 16000.0)):
        sw = post.LogSweep(*args)
        assert sw.samples() == PO.logsweep_samples(args[2], args[3])
        truth = PO.logsweep_generate(*args, dtype=np.longdouble)
        assert np.max(np.abs(sw.Generate() - truth)) <= 5e-11 and np.max(np.abs(PO.logsweep_generate(*args) - truth)) <= 5e-11
        tinv = PO.logsweep_inverse_filter(*args, dtype=np.longdouble)
        assert np.max(np.abs(sw.InverseFilter() - tinv)) <= 5e-11 * np.max(np.abs(tinv)) + 1e-18
    with pytest.raises(Exception):
        post.LogSweep(0, 100, 1, 48000).Generate()
    with pytest.raises(Exception):
        post.LogSweep(200, 100, 1, 48000).Generate()
    assert post.approximateRatio(48000 / 44100) == (160, 147)


# ---------------------------------------------------------------- the library on the GPU vs the oracle
@pytest.mark.gpu
def test_gpu_schroeder_and_impulse_start(conv):
    from algo_dsp_b200 import post
    an = post.Analyzer(48000.0)
    ir = exp_decay(48000.0, 1.0, 0.5) * G.white(24000, seed=3)
    got = an.SchroederIntegral(ir)
    ref = PO.schroeder_integral(ir)
    assert np.max(np.abs(got - ref)) <= 1e-9 and abs(got[0]) <= 1e-12
    assert np.all(np.diff(got) <= 1e-9)
    assert not an.SchroederIntegral(np.zeros(100)).any()
    for n in (1, 7, 2047, 2048, 2049, 100001):          # tile boundaries of the three-pass scan
        x = G.white(n, seed=n)
        assert np.max(np.abs(an.SchroederIntegral(x) - PO.schroeder_integral(x))) <= 1e-9
    # the reference's FindImpulseStart cases (ir_test.go:344-410)
    z = np.zeros(1000); z[0] = 1.0
    assert an.FindImpulseStart(z) == 0
    z = np.zeros(10000); z[5000] = 1.0; z[5001] = 0.5
    assert an.FindImpulseStart(z) == 5000 and an.findPeak(z) == 5000
    z = np.zeros(10000); z[:5000] = 0.001 * (np.arange(5000) % 2 * 2 - 1); z[5000] = 1.0
    assert an.FindImpulseStart(z) == 5000
    x = G.white(300000, seed=9) * np.linspace(0.01, 1, 300000)
    assert an.FindImpulseStart(x) == PO.find_impulse_start(x) and an.findPeak(x) == PO.find_peak(x)
    assert an.FindImpulseStart(np.zeros(50)) == 0 and an.findPeak(np.zeros(50)) == 0
    for bad in (an.SchroederIntegral, an.FindImpulseStart):
        with pytest.raises(conv.ConvError) as e:
            bad([])
        assert conv.errors_is(e.value, conv.ErrEmptyImpulseResponse)


@pytest.mark.gpu
def test_gpu_ir_rows_on_device(conv):
    """Batched device entry points: Schroeder curve, impulse start and abs-peak of every row of a correlation-sized batch."""
    torch = pytest.importorskip("torch")
    ctx = conv.default_context()
    rows, n = 5, 70000
    x = np.stack([G.white(n, seed=r) * np.exp(-np.arange(n) / (3000.0 * (r + 1))) for r in range(rows)])
    x[:, :100 * 3] *= 0.01
    xd = torch.tensor(x, device="cuda")
    out = torch.empty((rows, n), device="cuda", dtype=torch.float64)
    idx = torch.empty(rows, device="cuda", dtype=torch.int64)
    pk = torch.empty(rows, device="cuda", dtype=torch.int64)
    lib = L.load()
    assert lib.adsp_ir_schroeder_device(ctx.handle, xd.data_ptr(), n, rows, n, out.data_ptr(), n) == L.OK
    assert lib.adsp_ir_find_impulse_start_device(ctx.handle, xd.data_ptr(), n, rows, n, C.c_double(0.1), idx.data_ptr()) == L.OK
    assert lib.adsp_ir_find_peak_device(ctx.handle, xd.data_ptr(), n, rows, n, pk.data_ptr()) == L.OK
    ctx.sync()
    for r in range(rows):
        assert np.max(np.abs(out[r].cpu().numpy() - PO.schroeder_integral(x[r]))) <= 1e-9
        assert int(idx[r]) == PO.find_impulse_start(x[r]) and int(pk[r]) == PO.find_peak(x[r])


@pytest.mark.gpu
def test_gpu_logsweep_deconvolve_recovers_a_known_ir(conv, oracle):
    """TestLogSweepDeconvolveKnownIR, sweep_test.go:171-239: sweep through a short IR, deconvolved, peaks where the IR does;
    and the result equals the oracle's zero-padded FFT product (sweep.go:182-239)."""
    from algo_dsp_b200 import post
    sw = post.LogSweep(100, 4000, 0.25, 16000)
    sweep = sw.Generate()
    ident = sw.Deconvolve(sweep)
    assert len(ident) == 2 * len(sweep) - 1
    pk = np.max(np.abs(ident))
    assert 10 * np.log10(pk * pk / np.mean(ident * ident)) >= 15
    # the reference's case: duration 0.5 s, IR = delta + 0.3 * delta[100]; the recovered IR has a secondary peak 80..120
    # samples behind the main one with 0.15 .. 0.5 of its amplitude (sweep_test.go:174-238)
    sw2 = post.LogSweep(100, 4000, 0.5, 16000)
    sweep2 = sw2.Generate()
    h = np.zeros(200); h[0] = 1.0; h[100] = 0.3
    resp = oracle.convolve(sweep2, h)
    got = sw2.Deconvolve(resp)
    ref = PO.deconvolve_with_inverse(resp, sw2.InverseFilter())
    assert len(got) == len(resp) + len(sweep2) - 1 and G.rel_l2(got, ref) <= 1e-12
    pidx = int(np.argmax(np.abs(got)))
    ratio = np.max(np.abs(got[pidx + 80:pidx + 120])) / abs(got[pidx])
    assert 0.15 <= ratio <= 0.5
    with pytest.raises(conv.ConvError) as e:
        sw.Deconvolve([])
    assert conv.errors_is(e.value, conv.ErrEmptyInput)
    # device generators == host twins
    torch = pytest.importorskip("torch")
    ctx = conv.default_context()
    d = torch.empty(sw.samples(), device="cuda", dtype=torch.float64)
    assert L.load().adsp_logsweep_generate_device(ctx.handle, d.data_ptr(), *sw._args()) == L.OK
    ctx.sync()
    assert np.array_equal(d.cpu().numpy(), sweep)
    assert L.load().adsp_logsweep_inverse_filter_device(ctx.handle, d.data_ptr(), *sw._args()) == L.OK
    ctx.sync()
    assert np.array_equal(d.cpu().numpy(), sw.InverseFilter())


@pytest.mark.gpu
@pytest.mark.parametrize("ntaps", [1, 3, 31, 32, 64, 65, 257])
def test_gpu_fir_process_block_matches_the_reference_loop(conv, ntaps):
    from algo_dsp_b200 import post
    h = G.white(ntaps, seed=ntaps)                      # asymmetric on purpose: both tap orders of filter.go are exercised
    x = G.white(5000, seed=1)
    ref = PO.Fir(h)
    f = post.New(h)
    assert f.Order() == ntaps - 1
    pos = 0
    for blk in (1, 7, 300, ntaps, 1000, 3692 - ntaps):  # blocks shorter and longer than the delay line
        a = x[pos:pos + blk].copy()
        b = list(a)
        f.ProcessBlock(a)
        ref.process_block(b)
        assert np.max(np.abs(a - np.array(b))) <= 1e-12 * max(1.0, np.max(np.abs(b)))
        pos += blk
    f.Reset()
    a = np.zeros(2 * ntaps + 5); a[0] = 1.0
    f.ProcessBlock(a)                                   # impulse response after Reset
    want = h if ntaps < 32 else h[::-1]
    assert np.allclose(a[:ntaps], want, atol=1e-15) and not a[ntaps:].any()
    # several channels at once, each row its own delay line
    f2 = post.New(h, channels=3)
    blk = np.stack([G.white(700, seed=40 + c) for c in range(3)])
    refs = [PO.Fir(h) for _ in range(3)]
    for lo in (0, 350):
        part = np.ascontiguousarray(blk[:, lo:lo + 350])
        f2.ProcessBlock(part)
        for c in range(3):
            b = list(blk[c, lo:lo + 350])
            refs[c].process_block(b)
            assert np.max(np.abs(part[c] - np.array(b))) <= 1e-12 * max(1.0, np.max(np.abs(b)))


@pytest.mark.gpu
@pytest.mark.parametrize("ntaps", [2, 64, 200, 1024, 1025])
def test_gpu_fir_long_blocks_in_place_on_device_rows(conv, ntaps):
    """Blocks of several 8192-sample segments filtered IN PLACE on device rows (up to 1024 taps: descending tile walk with
    saved segment halos; beyond: the copy path), streaming across calls == one pass over the whole signal."""
    torch = pytest.importorskip("torch")
    from algo_dsp_b200 import post
    ch, blocks = 3, (20000, 9000, 100, 8192, 1)
    total = sum(blocks)
    x = np.stack([G.white(total, seed=70 + c) for c in range(ch)])
    h = G.white(ntaps, seed=ntaps + 1)
    c = h[::-1] if ntaps >= 32 else h                   # filter.go:93-94, see test above
    want = np.stack([np.convolve(x[r], c)[:total] for r in range(ch)])
    f = post.New(h, channels=ch)
    d = torch.tensor(x, device="cuda")
    stride = d.stride(0)
    pos = 0
    for blk in blocks:
        f.process_block_device(d.data_ptr() + pos * 8, blk, stride)     # rows stay where they are: stride = the full row
        pos += blk
    conv.default_context().sync()
    got = d.cpu().numpy()
    for r in range(ch):
        assert G.rel_l2(got[r], want[r]) <= 1e-13
    # host call, same filter state machine: continue the stream with one more block and compare with a fresh single pass
    f.Reset()
    y = x.copy()
    f.ProcessBlock(y)
    assert np.array_equal(y, got)                       # host and device entry points run the same kernels
    f.Close()


@pytest.mark.gpu
def test_gpu_resampler_matches_the_reference_loop(conv):
    from algo_dsp_b200 import post
    # ratios, reduction and design (resample_test.go:20-30, resample_design_test.go:8-18)
    r = post.NewRational(320, 294)
    assert r.Ratio() == (160, 147)
    assert post.NewForRates(44100, 48000).Ratio() == (160, 147)
    for q in (post.QualityFast, post.QualityBalanced, post.QualityBest):
        ro = PO.Resampler(3, 2, q)
        rg = post.NewRational(3, 2, quality=q)
        assert rg.TapsPerPhase() == ro.max_phase_len and np.max(np.abs(rg.Prototype() - np.array(ro.taps))) <= 1e-15
    with pytest.raises(conv.ConvError):
        post.NewRational(0, 1)
    # streaming: chunked == whole == oracle, bit for bit (same tap order, unfused multiply-add)
    for up, down, sr in ((160, 147, 44100), (147, 160, 48000), (2, 1, 48000), (1, 2, 48000), (3, 2, 48000)):
        x = sine(1000, sr, 6000) + 0.1 * G.white(6000, seed=up)
        ro, rg, rc = PO.Resampler(up, down), post.NewRational(up, down), post.NewRational(up, down)
        want = rg.PredictOutputLen(len(x))
        whole = rg.Process(x)
        ref = ro.process(x)
        assert len(whole) == want == len(ref) and np.array_equal(whole, ref)
        chunks = []
        for i in range(0, len(x), 257):
            chunks.append(rc.Process(x[i:i + 257]))
        assert np.array_equal(np.concatenate(chunks), whole)
        assert abs(len(whole) - round(len(x) * up / down)) <= 1
    # quality modes: passband droop and stopband attenuation of 2:1 decimation (resample_design_test.go:20-64)
    for q, max_pass, min_stop in ((post.QualityFast, 0.7, 20), (post.QualityBalanced, 0.35, 35), (post.QualityBest, 0.2, 50)):
        rms = lambda v: math.sqrt(float(np.mean(v * v)))
        ip, is_ = sine(2000, 48000, 32768), sine(17000, 48000, 32768)
        op = post.NewRational(1, 2, quality=q).Process(ip)
        os_ = post.NewRational(1, 2, quality=q).Process(is_)
        assert abs(20 * math.log10(rms(op[2048:]) / rms(ip[4096:]))) <= max_pass
        assert -20 * math.log10(rms(os_[2048:]) / rms(is_[4096:])) >= min_stop
    # channels as rows
    x2 = np.stack([sine(500 * (c + 1), 48000, 3000) for c in range(4)])
    y2 = post.NewRational(3, 2, channels=4).Process(x2)
    for c in range(4):
        assert np.array_equal(y2[c], PO.Resampler(3, 2).process(x2[c]))


@pytest.mark.gpu
@pytest.mark.parametrize("up,down", [(160, 147), (147, 160), (2, 1), (1, 4), (3, 2), (640, 441), (7, 1000)])
def test_gpu_resampler_device_rows_streaming(conv, up, down):
    """Device rows read where they are (no work-row copy), several tiles per call, blocks shorter than the history, and the
    host entry point on the same stream of samples: all bit identical to one whole-signal host call, which the test above pins
    to the reference loop."""
    torch = pytest.importorskip("torch")
    from algo_dsp_b200 import post
    ch, blocks = 3, (30000, 19990, 10, 3, 5000)
    total = sum(blocks)
    x = np.stack([G.white(total, seed=90 + c) for c in range(ch)])
    whole = post.NewRational(up, down, channels=ch).Process(x)
    ref1 = PO.Resampler(up, down).process(x[1][:3000])
    assert np.array_equal(whole[1][: len(ref1)], ref1)
    r = post.NewRational(up, down, channels=ch)
    d = torch.tensor(x, device="cuda")
    cap = whole.shape[1] + 64
    o = torch.zeros((ch, cap), device="cuda", dtype=torch.float64)
    pos = opos = 0
    for i, blk in enumerate(blocks):
        if i == 3:                                      # one block through the host entry point in the middle of the stream
            y = r.Process(np.ascontiguousarray(x[:, pos:pos + blk]))
            got = 0 if y.size == 0 else y.shape[1]
            if got:
                o[:, opos:opos + got] = torch.tensor(y, device="cuda")
        else:
            got = r.process_device(d.data_ptr() + pos * 8, blk, d.stride(0), o.data_ptr() + opos * 8, cap - opos, o.stride(0))
        pos += blk
        opos += got
    conv.default_context().sync()
    assert opos == whole.shape[1]
    assert np.array_equal(o[:, :opos].cpu().numpy(), whole)

