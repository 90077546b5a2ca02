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


def test_two_contexts_are_independent(conv, oracle):
    a, b = conv.Context(0), conv.Context(0)
    h, x = G.decaying_ir(5000), G.white(60000, seed=3)
    ya = conv.OverlapSave(h, 0, ctx=a).Process(x)
    yb = conv.OverlapSave(h, 0, ctx=b).Process(x)
    assert np.array_equal(ya, yb) and G.rel_l2(ya, oracle.overlap_save(h, 0, x)) <= 1e-12
    assert a.launch_count() > 0 and b.launch_count() > 0


def test_nan_and_inf_propagate_like_ieee(conv):
    """No clamping or flushing: a NaN sample poisons exactly the outputs it touches in the direct path."""
    x = np.ones(100)
    x[40] = np.nan
    y = conv.Direct(x, [1.0, 2.0, 3.0])
    assert np.isnan(y[40:43]).all() and not np.isnan(np.delete(y, [40, 41, 42])).any()
