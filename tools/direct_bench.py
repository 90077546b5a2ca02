"""Config 2 only: direct m-tap FIR over [channels] x 2^20 samples (device resident): python tools/direct_bench.py [channels=512] [m=64]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from algo_dsp_b200 import conv, siggen as G, _lib as L
ctx = conv.default_context(); lib = L.load()
st = torch.cuda.ExternalStream(ctx.stream())
ch = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n, m = 1 << 20, (int(sys.argv[2]) if len(sys.argv) > 2 else 64)
x = torch.rand((ch, n), device="cuda", dtype=torch.float64) * 2 - 1
k = torch.tensor(G.test_kernel(m), device="cuda")
y = torch.empty((ch, n + m - 1), device="cuda", dtype=torch.float64)
fn = lambda: lib.adsp_direct_batch_device(ctx.handle, x.data_ptr(), n, n, k.data_ptr(), m, 0, ch, y.data_ptr(), n + m - 1, 0)
for _ in range(2): fn()
ctx.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(5): fn()
e1.record(st); ctx.sync(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
ref = np.convolve(x[5].cpu().numpy(), G.test_kernel(m))
print(f"{os.environ.get('LABEL',''):24s} direct {m}-tap {ch}x2^20 f64: {ms:.3f} ms  {ch*(n+m-1)/ms/1e6:.1f} Gs/s  hbm_frac={ch*(n+m-1)*16/ms/1e6/6555.8:.3f} "
      f"fp64_frac={ch*(n+m-1)*m/ms/1e6/(148*64*1.965):.3f} relL2={G.rel_l2(y[5].cpu().numpy(), ref):.1e}", flush=True)
