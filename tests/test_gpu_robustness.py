"""Robustness of the C-ABI library on the GPU: strided batches, plan reuse across signal lengths
(graph cache, lazily built remainder transforms), concurrent one-shot calls, several contexts."""
import ctypes as C
import threading

import numpy as np
import pytest

from algo_dsp_b200 import _lib as L, siggen as G

pytestmark = pytest.mark.gpu


def test_strided_batch_through_the_abi(conv, oracle):
    """adsp_plan_process_batch with in/out strides larger than the row lengths."""
    K, n, ch, istr = 700, 9000, 5, 9100
    h = G.decaying_ir(K)
    xs = np.zeros((ch, istr))
    for c in range(ch):
        xs[c, :n] = G.white(n, seed=c)
        xs[c, n:] = 1e30          # poison: must never be read as signal
    ostr = n + K - 1 + 13
    out = np.full((ch, ostr), -7.0)
    plan = conv.NewOverlapSave(h, 0)
    st = L.load().adsp_plan_process_batch(plan._h, xs.ctypes.data_as(C.c_void_p), n, ch, istr, out.ctypes.data_as(C.c_void_p), ostr)
    assert st == L.OK
    for c in range(ch):
        assert G.rel_l2(out[c, : n + K - 1], oracle.overlap_save(h, 0, xs[c, :n])) <= 1e-12
        assert np.all(out[c, n + K - 1:] == -7.0)     # padding of the output rows untouched


def test_plan_reuse_across_lengths(conv, oracle):
    """One convolver, many signal lengths: remainder transforms are built lazily and cached."""
    K = 3000
    h = G.decaying_ir(K)
    plan = conv.NewOverlapSave(h, 0)
    for n in (1, 17, 2999, 3000, 3001, 40000, 20479, 100000, 40000, 1):
        x = G.white(n, seed=n)
        y = plan.Process(x)
        assert len(y) == n + K - 1 and G.rel_l2(y, oracle.overlap_save(h, 0, x)) <= 1e-12


def test_repeated_device_calls_replay_graph(conv, oracle):
    """The same device-resident call three times (eager, captured, replayed) gives identical results."""
    torch = pytest.importorskip("torch")
    K, n, ch = 20000, 100000, 24
    h = G.decaying_ir(K)
    plan = conv.NewOverlapSave(h, 0)
    x = torch.tensor(np.stack([G.white(n, seed=c) for c in range(ch)]), device="cuda")
    ol = n + K - 1
    outs = []
    y = torch.zeros((ch, ol), device="cuda", dtype=torch.float64)
    for _ in range(4):
        y.zero_()
        plan.process_device(x.data_ptr(), n, ch, n, y.data_ptr(), ol)
        plan.sync()
        outs.append(y.cpu().numpy().copy())
    for o in outs[1:]:
        assert np.array_equal(o, outs[0])
    assert G.rel_l2(outs[0][5], oracle.overlap_save(h, 0, x[5].cpu().numpy())) <= 1e-12


def test_concurrent_one_shot_calls(conv, oracle):
    """Package-level functions are goroutine-safe in the reference (pooled instances); here they
    serialise on the context: concurrent callers must all get correct answers."""
    jobs = [(G.white(5000 + 37 * i, seed=i), G.decaying_ir(100 + 50 * i)) for i in range(8)]
    res = [None] * len(jobs)

    def work(i):
        res[i] = conv.Convolve(*jobs[i])

    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for (x, h), y in zip(jobs, res):
        assert G.rel_l2(y, oracle.convolve(x, h)) <= 1e-12


def test_concurrent_mixed_families_on_one_context(conv, oracle):
    """Eight host threads, each hammering a different family (plans, direct, correlation + peak, block FIR, resampler,
    streaming reverb, device generator + plan) on the SAME context: every entry point takes the context lock, so all
    answers must equal the single-threaded ones."""
    from algo_dsp_b200 import post
    h = G.decaying_ir(3000)
    x = G.white(40000, seed=11)
    fir_h = G.white(200, seed=12)
    errs = []

    def check(name, got, want, tol=1e-12):
        e = G.rel_l2(got, want)
        if not e <= tol:
            errs.append((name, e))

    def t_ols():
        p = conv.NewOverlapSave(h, 0)
        for _ in range(6):
            check("ols", p.Process(x), oracle.overlap_save(h, 0, x))

    def t_direct():
        k = G.white(300, seed=13)
        for _ in range(6):
            check("direct", conv.Direct(x[:9000], k), oracle.direct(x[:9000], k))

    def t_corr():
        for _ in range(6):
            c = conv.Correlate(x[:20000], x[500:4500])
            check("corr", c, oracle.correlate(x[:20000], x[500:4500]))
            if conv.FindPeak(c)[0] != oracle.find_peak(oracle.correlate(x[:20000], x[500:4500]))[0]:
                errs.append(("peak", 0))

    def t_fir():
        want = np.convolve(x, fir_h[::-1])[: len(x)]
        for _ in range(6):
            f = post.New(fir_h)
            y = x.copy()
            f.ProcessBlock(y[:15000]); f.ProcessBlock(y[15000:])
            check("fir", y, want)
            f.Close()

    def t_resample():
        want = post.NewRational(160, 147).Process(x)
        for _ in range(6):
            r = post.NewRational(160, 147)
            got = np.concatenate([r.Process(x[:17000]), r.Process(x[17000:])])
            if not np.array_equal(got, want):
                errs.append(("resample", 0))

    def t_reverb():
        want = np.concatenate([np.zeros(128), oracle.convolve(x, h)])[: len(x)]
        for _ in range(3):
            p = conv.NewPartitionedConvolution(h, 7, 13)
            out = np.empty(len(x))
            for a in range(0, len(x), 5000):
                p.ProcessBlock(np.ascontiguousarray(x[a:a + 5000]), out[a:a + 5000])
            check("partitioned", out, want, 1e-11)

    def t_convolve():
        for i in range(6):
            a, b = G.white(7000 + i, seed=20 + i), G.white(50 + i, seed=30 + i)
            check("convolve", conv.Convolve(a, b), oracle.convolve(a, b))

    ts = [threading.Thread(target=f) for f in (t_ols, t_direct, t_corr, t_fir, t_resample, t_reverb, t_convolve, t_ols)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs


def test_two_contexts_are_independent(conv, oracle):
    a, b = conv.Context(0), conv.Context(0)
    h, x = G.decaying_ir(5000), G.white(60000, seed=3)
    ya = conv.OverlapSave(h, 0, ctx=a).Process(x)
    yb = conv.OverlapSave(h, 0, ctx=b).Process(x)
    assert np.array_equal(ya, yb) and G.rel_l2(ya, oracle.overlap_save(h, 0, x)) <= 1e-12
    assert a.launch_count() > 0 and b.launch_count() > 0


def test_two_contexts_concurrently(conv, oracle):
    """Two contexts on the same device share no lock: their calls really overlap (plans, host pipeline with copy threads,
    block FIR, device generators)."""
    from algo_dsp_b200 import post
    h, x = G.decaying_ir(4000), G.white(300000, seed=5)
    want = oracle.overlap_save(h, 0, x)
    fh = G.white(100, seed=6)
    fwant = np.convolve(x, fh[::-1])[: len(x)]
    errs = []

    def work(seed):
        try:
            c = conv.Context(0)
            p = conv.OverlapSave(h, 0, ctx=c)
            xs = np.stack([x, x[::-1].copy(), x * 0.5, x])          # 9.6 MB: the staged host pipeline with copy threads
            for _ in range(4):
                y = p.ProcessBatch(xs)
                if G.rel_l2(y[0], want) > 1e-12 or G.rel_l2(y[2], 0.5 * want) > 1e-12:
                    errs.append(("ols", seed))
                f = post.New(fh, ctx=c)
                z = x.copy()
                f.ProcessBlock(z)
                if G.rel_l2(z, fwant) > 1e-12:
                    errs.append(("fir", seed))
                f.Close()
                w = G.DeviceArray(c, 1, 4096)
                G.white_device(c, w.ptr, 4096, 1, 4096, amp=1.0, seed0=seed)
                if not np.array_equal(w.get()[0], G.white(4096, seed=seed)):
                    errs.append(("gen", seed))
            p.Close()
            c.close()
        except Exception as e:                                        # noqa: BLE001 - reported below
            errs.append((repr(e), seed))

    ts = [threading.Thread(target=work, args=(s,)) for s in (1, 2, 3)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs


def test_nan_and_inf_propagate_like_ieee(conv):
    """No clamping or flushing: a NaN sample poisons exactly the outputs it touches in the direct path."""
    x = np.ones(100)
    x[40] = np.nan
    y = conv.Direct(x, [1.0, 2.0, 3.0])
    assert np.isnan(y[40:43]).all() and not np.isnan(np.delete(y, [40, 41, 42])).any()


def test_graph_replay_survives_scratch_reallocation(conv, oracle):
    """a captured graph bakes the shared scratch pointer in.  Small call x3 (eager, capture, replay),
    then a larger call on the same context grows the scratch (cudaFree + cudaMalloc), then the small call again: the
    stale graph must be dropped, not replayed against freed memory."""
    torch = pytest.importorskip("torch")
    ctx = conv.Context(0)
    K = 20000
    h = G.decaying_ir(K)
    plan = conv.OverlapSave(h, 0, ctx=ctx)
    n, ch = 100000, 2
    x = torch.tensor(np.stack([G.white(n, seed=c) for c in range(ch)]), device="cuda")
    ol = n + K - 1
    y = torch.zeros((ch, ol), device="cuda", dtype=torch.float64)
    ref = oracle.overlap_save(h, 0, x[1].cpu().numpy())

    def small():
        y.zero_()
        plan.process_device(x.data_ptr(), n, ch, n, y.data_ptr(), ol)
        plan.sync()
        assert G.rel_l2(y[1].cpu().numpy(), ref) <= 1e-12

    for _ in range(3):
        small()
    # grow the context's scratch: many more channels of a longer transform through another plan of the same context
    big = conv.OverlapSave(G.decaying_ir(90000), 0, ctx=ctx)
    nb, chb = 400000, 48
    xb = torch.rand((chb, nb), device="cuda", dtype=torch.float64)
    yb = torch.zeros((chb, nb + 90000 - 1), device="cuda", dtype=torch.float64)
    big.process_device(xb.data_ptr(), nb, chb, nb, yb.data_ptr(), nb + 90000 - 1)
    big.sync()
    for _ in range(3):
        small()
    # and a one-shot correlate (another user of the shared scratch) between replays
    conv.Correlate(G.white(300000, seed=5), G.white(200000, seed=6), ctx=ctx)
    small()
    big.Close()
    plan.Close()


@pytest.mark.parametrize("in_pinned,out_pinned", [(False, False), (True, False), (False, True), (True, True)])
def test_pageable_and_pinned_host_buffers_through_the_pipeline(conv, oracle, monkeypatch, in_pinned, out_pinned):
    """The chunked host pipeline (stage-in | H2D | kernels | D2H | stage-out) with pageable and pinned caller memory in
    every combination; chunk size forced small so that several chunks and slot reuse happen at a test-sized batch."""
    monkeypatch.setenv("ADSP_STAGE_PIPE_CHUNK_MB", "1")
    monkeypatch.setenv("ADSP_PIPE_CHUNK_MB", "1")
    K, n, ch = 3000, 40000, 112         # (n + out_len) * 8 B * ch = 74 MB > the 32 MB small-call limit; 1 MB chunks -> 1 row per chunk... many chunks
    h = G.decaying_ir(K)
    ol = n + K - 1
    xs = conv.pinned_empty((ch, n)) if in_pinned else np.empty((ch, n))
    for c in range(ch):
        xs[c] = G.white(n, seed=c)
    out = conv.pinned_empty((ch, ol)) if out_pinned else np.empty((ch, ol))
    out[:] = -7.0
    assert bool(L.load().adsp_host_ptr_is_pinned(C.c_void_p(xs.ctypes.data))) == in_pinned
    assert bool(L.load().adsp_host_ptr_is_pinned(C.c_void_p(out.ctypes.data))) == out_pinned
    ctx = conv.Context(0)
    plan = conv.OverlapSave(h, 0, ctx=ctx)
    for _ in range(2):                  # second pass reuses the pinned slots and events
        st = L.load().adsp_plan_process_batch(plan._h, xs.ctypes.data_as(C.c_void_p), n, ch, n, out.ctypes.data_as(C.c_void_p), ol)
        assert st == L.OK
    for c in (0, 1, 2, 55, ch - 2, ch - 1):
        assert G.rel_l2(out[c], oracle.overlap_save(h, 0, xs[c])) <= 1e-12
    prof = ctx.host_profile_get()
    assert (prof["staged_in_bytes"] > 0) == (not in_pinned) and (prof["staged_out_bytes"] > 0) == (not out_pinned)
    plan.Close()


def test_small_call_profile_and_staged_one_shot(conv, oracle):
    """Config-1-sized mono call from pageable memory: staged through the pinned slots, phase breakdown recorded."""
    ctx = conv.Context(0)
    h, x = G.decaying_ir(96000), G.white(480000, seed=1)
    plan = conv.OverlapSave(h, 0, ctx=ctx)
    plan.Process(x)
    ctx.host_profile(True)
    y = plan.Process(x)
    prof = ctx.host_profile_get()
    ctx.host_profile(False)
    assert G.rel_l2(y, oracle.overlap_save(h, 0, x)) <= 1e-12
    assert prof["total_ms"] > 0 and prof["kernels_ms"] > 0 and prof["upload_ms"] > 0 and prof["download_ms"] > 0
    assert prof["staged_in_bytes"] >= x.nbytes and ctx.stage_threads() >= 1
    # one-shot entry points take the same staged path
    a, b = G.white(700000, seed=2), G.decaying_ir(5000)
    assert G.rel_l2(conv.Convolve(a, b, ctx=ctx), oracle.convolve(a, b)) <= 1e-12
    plan.Close()


def test_partitioned_nan_stays_local(conv):
    """a non-finite input sample must not leak through delay-line slots that do not belong to a firing
    (0 * NaN through a padded tap).  Block transforms spread a NaN over the blocks that contain it (the reference's
    per-partition FFTs do the same, partitioned.go:139-186), so the poisoned band is block granular: it must cover
    [i0 + L, i0 + L + K) and stay within a few of the largest internal blocks (2048 samples) around it."""
    K, L_ord = 40000, 7
    h = G.decaying_ir(K)
    n = 120000
    x = G.white(n, seed=4)
    i0 = 30000
    x[i0] = np.nan
    p = conv.NewPartitionedConvolution(h, L_ord, 13)
    y = np.zeros(n)
    for a in range(0, n, 8192):
        p.ProcessBlock(x[a:a + 8192], y[a:a + 8192])
    lat = 1 << L_ord
    bad = np.isnan(y)
    assert bad[i0 + lat: i0 + lat + K].all()
    assert not bad[: i0 + lat - 4096].any() and not bad[i0 + lat + K + 8192:].any()


def test_host_register_makes_caller_memory_dma_able(conv, oracle):
    """adsp_host_register pins ordinary caller memory in place: the call then takes the direct DMA path (nothing staged)."""
    ctx = conv.Context(0)
    h, x = G.decaying_ir(20000), G.white(300000, seed=8)
    y = np.empty(x.size + h.size - 1)
    lib = L.load()
    assert lib.adsp_host_ptr_is_pinned(C.c_void_p(x.ctypes.data)) == 0
    assert lib.adsp_host_register(C.c_void_p(x.ctypes.data), x.nbytes) == L.OK
    assert lib.adsp_host_register(C.c_void_p(y.ctypes.data), y.nbytes) == L.OK
    try:
        assert lib.adsp_host_ptr_is_pinned(C.c_void_p(x.ctypes.data)) == 1
        plan = conv.OverlapSave(h, 0, ctx=ctx)
        before = ctx.host_profile_get()
        plan.ProcessTo(y, x)
        after = ctx.host_profile_get()
        assert after["staged_in_bytes"] == before["staged_in_bytes"] and after["staged_out_bytes"] == before["staged_out_bytes"]
        assert G.rel_l2(y, oracle.overlap_save(h, 0, x)) <= 1e-12
        plan.Close()
    finally:
        assert lib.adsp_host_unregister(C.c_void_p(x.ctypes.data)) == L.OK
        assert lib.adsp_host_unregister(C.c_void_p(y.ctypes.data)) == L.OK
