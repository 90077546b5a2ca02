"""Randomised parity sweep through the public Python mirror (-> C ABI -> CUDA) against the CPU oracle: random lengths, kernel
sizes, block splits and batch shapes for every family on the path.  Not part of the test suite (the suite pins fixed shapes);
a soak run for a GPU box:
    python tests/tools/fuzz_parity.py [seconds=120] [seed=1] [big]
Prints one line per family with the case count and the worst relative L2 error; exits 1 on the first violation."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from algo_dsp_b200 import conv, post, siggen as G
from oracle import oracle as O, post_oracle as PO

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
TOL64, TOL32 = 1e-12, 1e-5
worst, count, worst_case = {}, {}, {}


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    d = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / d) if d > 0 else float(np.linalg.norm(a - b))


def note(fam, err, tol, what):
    if err >= worst.get(fam, -1.0):
        worst[fam], worst_case[fam] = err, what
    count[fam] = count.get(fam, 0) + 1
    if not err <= tol:
        print(f"FAIL {fam}: rel L2 {err:.3e} > {tol:g} for {what}", flush=True)
        sys.exit(1)


BIG = len(sys.argv) > 3 and sys.argv[3] == "big"     # draw every size from the top decade and a half of its range


def logint(lo, hi):
    if BIG:
        lo = max(lo, hi // 30)
    return int(round(np.exp(rng.uniform(np.log(lo), np.log(hi)))))


def sig(n):
    return rng.uniform(-1, 1, n)


def case_convolve():
    n, m = logint(1, 300000), logint(1, 40000)
    a, b = sig(n), sig(m)
    note("Convolve", rel(conv.Convolve(a, b), O.convolve(a, b)), TOL64, (n, m))


def case_ols():
    K, n = logint(1, 60000), logint(1, 400000)
    h, x = sig(K), sig(n)
    fam = rng.choice(["OverlapSave", "OverlapAdd", "OverlapSave32"])
    if fam == "OverlapSave":
        note(fam, rel(conv.NewOverlapSave(h, 0).Process(x), O.overlap_save(h, 0, x)), TOL64, (K, n))
    elif fam == "OverlapAdd":
        note(fam, rel(conv.NewOverlapAdd(h, 0).Process(x), O.overlap_add(h, 0, x)), TOL64, (K, n))
    else:
        h32, x32 = h.astype(np.float32), x.astype(np.float32)
        got = conv.NewOverlapSave(h32, 0, dtype=np.float32).Process(x32)
        note(fam, rel(got, O.overlap_save(h32.astype(np.float64), 0, x32.astype(np.float64))), TOL32, (K, n))


def case_correlate():
    n, m = logint(1, 200000), logint(1, 200000)
    a, b = sig(n), sig(m)
    ref = O.correlate(a, b)
    got = conv.Correlate(a, b)
    note("Correlate", rel(got, ref), TOL64, (n, m))
    gi, _ = conv.FindPeak(got)
    ri, _ = O.find_peak(ref)
    if gi != ri and abs(ref[gi] - ref[ri]) > 1e-9 * abs(ref[ri]):
        print(f"FAIL FindPeak: {gi} vs {ri} for {(n, m)}")
        sys.exit(1)


def case_correlate_batch():
    pairs, n, m = int(rng.integers(2, 9)), logint(65, 60000), logint(65, 60000)
    a = rng.uniform(-1, 1, (pairs, n))
    shared = bool(rng.integers(0, 2))
    b = rng.uniform(-1, 1, (pairs, m))
    if shared:
        b[:] = b[0]
    out, pi, pv = conv.CorrelateBatch(a, b)
    p = int(rng.integers(0, pairs))
    ref = O.correlate(a[p], b[p])
    note("CorrelateBatch", rel(out[p], ref), TOL64, (pairs, n, m, shared))
    ri, rv = O.find_peak(ref)
    if int(pi[p]) != ri and abs(ref[int(pi[p])] - rv) > 1e-9 * abs(rv):
        print(f"FAIL CorrelateBatch peak: {int(pi[p])} vs {ri}")
        sys.exit(1)


def case_direct():
    ch, n, m = int(rng.integers(1, 6)), logint(1, 20000), logint(1, 1500)
    x = rng.uniform(-1, 1, (ch, n))
    per_ch = bool(rng.integers(0, 2))
    k = rng.uniform(-1, 1, (ch, m)) if per_ch else sig(m)
    y = conv.DirectBatch(x, k)
    c = int(rng.integers(0, ch))
    note("Direct", rel(y[c], O.direct(x[c], k[c] if per_ch else k)), TOL64, (ch, n, m, per_ch))


def case_fir():
    ch, taps = int(rng.integers(1, 4)), logint(1, 1400)
    h = sig(taps)
    c = h[::-1] if taps >= 32 else h
    f = post.New(h, channels=ch)
    total, pos = logint(1, 60000), 0
    x = rng.uniform(-1, 1, (ch, total))
    got = np.empty_like(x)
    while pos < total:
        blk = min(total - pos, logint(1, 30000))
        part = x[:, pos:pos + blk].copy()               # ProcessBlock works in place: never hand it a view of x
        f.ProcessBlock(part)
        got[:, pos:pos + blk] = part
        pos += blk
    r = int(rng.integers(0, ch))
    note("fir.ProcessBlock", rel(got[r], np.convolve(x[r], c)[:total]), 1e-12, (ch, taps, total))
    f.Close()


def case_resample():
    up, down = int(rng.integers(1, 200)), int(rng.integers(1, 200))
    ch, total = int(rng.integers(1, 4)), logint(1, 40000)
    x = rng.uniform(-1, 1, (ch, total))
    whole = post.NewRational(up, down, channels=ch).Process(x)
    r = post.NewRational(up, down, channels=ch)
    parts, pos = [], 0
    while pos < total:
        blk = min(total - pos, logint(1, 20000))
        y = r.Process(np.ascontiguousarray(x[:, pos:pos + blk]))
        if y.size:
            parts.append(y.reshape(ch, -1))
        pos += blk
    cat = np.concatenate(parts, axis=1) if parts else np.empty((ch, 0))
    whole2 = np.asarray(whole).reshape(ch, -1) if np.asarray(whole).size else np.empty((ch, 0))
    if cat.shape != whole2.shape or not np.array_equal(cat, whole2):
        print(f"FAIL resample: chunked != whole for {(up, down, ch, total)}")
        sys.exit(1)
    small = min(total, 1500)
    ref = PO.Resampler(up, down).process(x[0][:small])
    if not np.array_equal(whole2[0][:len(ref)], ref):
        print(f"FAIL resample: != reference loop for {(up, down, small)}")
        sys.exit(1)
    count["resample.Process"] = count.get("resample.Process", 0) + 1
    worst["resample.Process"] = 0.0


def case_partitioned():
    K, mn = logint(1, 30000), int(rng.integers(5, 9))
    h = sig(K)
    total = logint(1, 40000)
    x = sig(total)
    p = conv.NewPartitionedConvolution(h, mn, 13)
    ref = O.Partitioned(h, mn, 13)
    pos, got, want = 0, [], []
    while pos < total:
        blk = min(total - pos, logint(1, 9000))
        out = np.empty(blk)
        p.ProcessBlock(np.ascontiguousarray(x[pos:pos + blk]), out)
        got.append(out)
        want.append(ref.process_block(x[pos:pos + blk]))
        pos += blk
    note("Partitioned.ProcessBlock", rel(np.concatenate(got), np.concatenate(want)), TOL64, (K, mn, total))


def case_ols_batch():
    ch, K, n = int(rng.integers(1, 12)), logint(1, 40000), logint(1, 200000)
    h = sig(K)
    x = rng.uniform(-1, 1, (ch, n))
    y = conv.NewOverlapSave(h, 0).ProcessBatch(x)
    c = int(rng.integers(0, ch))
    note("OverlapSave.ProcessBatch", rel(y[c], O.overlap_save(h, 0, x[c])), TOL64, (ch, K, n))


def case_streaming():
    K = logint(1, 5000)
    B = 1 << int(rng.integers(4, 13))
    ols = bool(rng.integers(0, 2))
    h = sig(K)
    s = (conv.StreamingOverlapSave if ols else conv.StreamingOverlapAdd)(h, B)
    ref = O.Streaming(h, B, ols)
    B = s.BlockSize()
    worst_e = 0.0
    for _ in range(int(rng.integers(1, 6))):
        blk = sig(B)
        worst_e = max(worst_e, rel(s.ProcessBlock(blk), ref.process_block(blk)))
    note("Streaming.ProcessBlock", worst_e, 1e-11, (K, B, ols))


def case_deconvolve():
    n, m = logint(8, 50000), logint(1, 2000)
    x, k = sig(n), sig(m)
    k[0] += 2.0                                            # well conditioned
    y = O.convolve(x, k)
    got = conv.Deconvolve(y, k, conv.DeconvOptions(conv.DeconvRegularized, 1e-9))
    ref = O.deconvolve(y, k, O.DECONV_REGULARIZED, 1e-9)
    note("Deconvolve", rel(got, ref), 1e-9, (n, m))


def case_device_rows():
    """Device entry points with row strides larger than the rows: OLS plan, block FIR, resampler, reverb wet/dry in place."""
    import torch
    ctx = conv.default_context()
    ch, n = int(rng.integers(1, 7)), logint(1, 60000)
    pad_in, pad_out = int(rng.integers(0, 70)), int(rng.integers(0, 70))
    x = rng.uniform(-1, 1, (ch, n))
    xd = torch.zeros((ch, n + pad_in), device="cuda", dtype=torch.float64)
    xd[:, :n] = torch.tensor(x, device="cuda")
    which = int(rng.integers(0, 4))
    if which == 0:
        K = logint(1, 20000)
        h = sig(K)
        plan = conv.NewOverlapSave(h, 0)
        ol = n + K - 1
        yd = torch.full((ch, ol + pad_out), 7.0, device="cuda", dtype=torch.float64)
        plan.process_device(xd.data_ptr(), n, ch, xd.stride(0), yd.data_ptr(), yd.stride(0))
        plan.sync()
        c = int(rng.integers(0, ch))
        note("OverlapSave.process_device", rel(yd[c, :ol].cpu().numpy(), O.overlap_save(h, 0, x[c])), TOL64, (ch, K, n, pad_in, pad_out))
        assert pad_out == 0 or bool((yd[:, ol:] == 7.0).all()), "wrote beyond the row"
    elif which == 1:
        taps = logint(1, 1400)
        h = sig(taps)
        f = post.New(h, channels=ch)
        f.process_block_device(xd.data_ptr(), n, xd.stride(0))
        ctx.sync()
        c = int(rng.integers(0, ch))
        cc = h[::-1] if taps >= 32 else h
        note("fir.process_block_device", rel(xd[c, :n].cpu().numpy(), np.convolve(x[c], cc)[:n]), TOL64, (ch, taps, n, pad_in))
        assert pad_in == 0 or bool((xd[:, n:] == 0).all()), "wrote beyond the row"
        f.Close()
    elif which == 2:
        up, down = int(rng.integers(1, 200)), int(rng.integers(1, 200))
        r = post.NewRational(up, down, channels=ch)
        want = np.asarray(post.NewRational(up, down, channels=ch).Process(x)).reshape(ch, -1) if n else np.empty((ch, 0))
        cap = want.shape[1] + pad_out
        od = torch.full((ch, max(cap, 1)), 7.0, device="cuda", dtype=torch.float64)
        got = r.process_device(xd.data_ptr(), n, xd.stride(0), od.data_ptr(), cap, od.stride(0))
        ctx.sync()
        if got != want.shape[1] or not np.array_equal(od[:, :got].cpu().numpy(), want):
            print(f"FAIL resample.process_device: {(up, down, ch, n)}")
            sys.exit(1)
        count["resample.process_device"] = count.get("resample.process_device", 0) + 1
        worst["resample.process_device"] = 0.0
    else:
        K, mn = logint(1, 20000), int(rng.integers(5, 9))
        h = sig(K)
        wet, dry = float(rng.uniform(0, 1)), float(rng.uniform(0, 1))
        rv = conv.NewConvolutionReverb(h, mn, channels=ch)
        rv.SetWetDry(wet, dry)
        rv.process_in_place_device(xd.data_ptr(), n, xd.stride(0))
        ctx.sync()
        c = int(rng.integers(0, ch))
        wet_sig = np.concatenate([np.zeros(rv.Latency()), O.convolve(x[c], h)])[:n]
        ref = dry * x[c] + wet * wet_sig                    # convolution.go:76-80
        note("reverb device rows", rel(xd[c, :n].cpu().numpy(), ref), 1e-11, (ch, K, mn, n, pad_in))


cases = [case_device_rows, case_ols_batch, case_streaming, case_deconvolve, case_convolve, case_ols, case_correlate, case_correlate_batch, case_direct, case_fir, case_resample, case_partitioned]
t0 = time.time()
i = 0
while time.time() - t0 < budget:
    cases[i % len(cases)]()
    i += 1
for fam in sorted(count):
    print(f"{fam:28s} {count[fam]:5d} cases, worst rel L2 {worst[fam]:.2e} at {worst_case.get(fam)}")
print(f"fuzz_parity: ok ({i} cases in {time.time() - t0:.0f} s)")
