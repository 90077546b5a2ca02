"""Diagnostics: (1) per-call wall time of repeated adsp_correlate_batch_device calls (shared b / per-pair b); (2) latency of small
pinned H2D / D2H copies issued back to back and after idle gaps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from algo_dsp_b200 import conv, siggen as G, _lib as L
n = 1 << 20
ctx = conv.default_context(); lib = L.load()
b = torch.empty((1, n), device="cuda", dtype=torch.float64)
G.log_sweep_device(ctx, b.data_ptr(), n)
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
a = torch.empty((pairs, n), device="cuda", dtype=torch.float64)
G.delay_mix_device(ctx, a.data_ptr(), n, pairs, n, b.data_ptr())
out = torch.empty((pairs, 2 * n - 1), device="cuda", dtype=torch.float64)
pi = torch.empty(pairs, device="cuda", dtype=torch.int64); pv = torch.empty(pairs, device="cuda", dtype=torch.float64)
bb = b.expand(pairs, n).contiguous()
ctx.sync(); torch.cuda.synchronize()
for name, bp, bs in (("shared", b.data_ptr(), 0), ("per-pair", bb.data_ptr(), n)):
    ts = []
    for i in range(6):
        t0 = time.perf_counter()
        st = lib.adsp_correlate_batch_device(ctx.handle, a.data_ptr(), n, n, bp, n, bs, pairs, out.data_ptr(), 2 * n - 1, pi.data_ptr(), pv.data_ptr(), 0)
        ctx.sync()
        ts.append(round((time.perf_counter() - t0) * 1e3, 2))
    print(name, "pairs", pairs, "ms per call:", ts, flush=True)
# small-copy latency
for mb in (4.6, 0.5):
    nb = int(mb * (1 << 20))
    h = torch.empty(nb, dtype=torch.uint8).pin_memory()
    d = torch.empty(nb, dtype=torch.uint8, device="cuda")
    for gap in (0.0, 0.002, 0.02):
        ts_h, ts_d = [], []
        for i in range(12):
            time.sleep(gap)
            t0 = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
            h.copy_(d, non_blocking=True); torch.cuda.synchronize(); t2 = time.perf_counter()
            ts_h.append((t1 - t0) * 1e3); ts_d.append((t2 - t1) * 1e3)
        ts_h.sort(); ts_d.sort()
        print(f"copy {mb} MB, idle gap {gap*1e3:.0f} ms: H2D median {ts_h[6]:.3f} ms, D2H median {ts_d[6]:.3f} ms", flush=True)
