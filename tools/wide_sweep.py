"""Wide shapes: config 3 (ch x 14.4 M, 288 000 taps) and config 5 (1 x 2^28 slice, 2^20 taps), device resident.
    [ADSP_FFT_N=...] python tools/wide_sweep.py [c3|c5] [channels=16] [f64|f32]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from algo_dsp_b200 import conv, siggen as G, _lib as L
which = sys.argv[1] if len(sys.argv) > 1 else "c3"
ch = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dt = np.float32 if (len(sys.argv) > 3 and sys.argv[3] == "f32") else np.float64
tdt = torch.float64 if dt == np.float64 else torch.float32
ctx = conv.default_context()
st = torch.cuda.ExternalStream(ctx.stream())
if which == "c3":
    n, K = 14_400_000, 288_000
else:
    n, K, ch = 1 << 28, 1 << 20, 1
h = G.decaying_ir(K)
plan = conv.OverlapSave(h, 0, ctx=ctx, dtype=dt)
x = (torch.rand((ch, n), device="cuda", dtype=tdt) * 2 - 1)
ol = n + K - 1
ostr = (ol + 31) // 32 * 32
y = torch.empty((ch, ostr), device="cuda", dtype=tdt)
fn = lambda: plan.process_device(x.data_ptr(), n, ch, n, y.data_ptr(), ostr)
for _ in range(2): fn()
plan.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
R = 4
for _ in range(R): fn()
e1.record(st); plan.sync(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / R
print(json.dumps({"label": os.environ.get("LABEL", ""), "which": which, "ch": ch, "dtype": np.dtype(dt).name, "ms": round(ms, 3),
                  "gsamples_s": round(ch * n / ms / 1e6, 2), "plan": plan.internal_geometry()}), flush=True)
