import os, sys, time
sys.path.insert(0, "/root/repo")
import torch
from algo_dsp_b200 import conv, siggen as G, _lib as L
n = 1 << 20
ctx = conv.default_context(); lib = L.load()
b = torch.empty((1, n), device="cuda", dtype=torch.float64)
G.log_sweep_device(ctx, b.data_ptr(), n)
for pairs in (64, 256, 1024):
    a = torch.empty((pairs, n), device="cuda", dtype=torch.float64)
    G.delay_mix_device(ctx, a.data_ptr(), n, pairs, n, b.data_ptr())
    out = torch.empty((pairs, 2 * n - 1), device="cuda", dtype=torch.float64)
    pi = torch.empty(pairs, device="cuda", dtype=torch.int64); pv = torch.empty(pairs, device="cuda", dtype=torch.float64)
    def run(peaks=True):
        st = lib.adsp_correlate_batch_device(ctx.handle, a.data_ptr(), n, n, b.data_ptr(), n, 0, pairs, out.data_ptr(), 2 * n - 1,
                                             pi.data_ptr() if peaks else None, pv.data_ptr() if peaks else None, 0)
        assert st == 0
    for peaks in (True, False):
        run(peaks); ctx.sync()
        t0 = time.perf_counter(); run(peaks); t1 = time.perf_counter(); ctx.sync(); t2 = time.perf_counter()
        print(f"pairs={pairs} peaks={peaks}: enqueue {1e3*(t1-t0):.1f} ms, total {1e3*(t2-t0):.1f} ms -> {pairs/(t2-t0):.0f} pairs/s", flush=True)
    del a, out
