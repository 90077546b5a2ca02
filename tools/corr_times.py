"""Per-kernel device time of the batched FFT correlation (BASELINE config 4 shape): python tools/corr_times.py [pairs] [log2n] [with_out]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from algo_dsp_b200 import conv, siggen as G, _lib as L
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 20)
with_out = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
ctx = conv.default_context(); lib = L.load()
b = torch.empty((1, n), device="cuda", dtype=torch.float64)
G.log_sweep_device(ctx, b.data_ptr(), n)
a = torch.empty((pairs, n), device="cuda", dtype=torch.float64)
G.delay_mix_device(ctx, a.data_ptr(), n, pairs, n, b.data_ptr())
out = torch.empty((pairs, 2 * n - 1), device="cuda", dtype=torch.float64) if with_out else None
pi = torch.empty(pairs, device="cuda", dtype=torch.int64); pv = torch.empty(pairs, device="cuda", dtype=torch.float64)
def run():
    st = lib.adsp_correlate_batch_device(ctx.handle, a.data_ptr(), n, n, b.data_ptr(), n, 0, pairs, out.data_ptr() if with_out else None, 2 * n - 1 if with_out else 0,
                                         pi.data_ptr(), pv.data_ptr(), 0)
    assert st == 0, L.last_error()
for _ in range(2): run()
ctx.sync()
import time
t0 = time.perf_counter()
for _ in range(3): run()
ctx.sync()
ms = (time.perf_counter() - t0) / 3 * 1e3        # wall clock: the call is synchronous
ctx.kernel_timing(True); ctx.kernel_times(reset=True)
run(); ctx.sync()
kt = ctx.kernel_times()
lags_ok = all(int(pi[p]) - (n - 1) == G.delay_of(p) for p in range(pairs))
print(f"{os.environ.get('LABEL','')} pairs={pairs} n=2^{n.bit_length()-1} out={with_out}: {ms:.3f} ms  {pairs/ms*1e3:.0f} pairs/s  lags_ok={lags_ok}  per-kernel ms (one call): "
      + str({k: (round(v[0], 3), v[1]) for k, v in kt.items() if v[1]}), flush=True)
