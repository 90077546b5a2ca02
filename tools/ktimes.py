import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from algo_dsp_b200 import conv, siggen as G
K, n, ch = 96000, 480000, 256
ctx = conv.default_context()
plan = conv.NewOverlapSave(G.decaying_ir(K), 0)
x = torch.rand((ch, n), device="cuda", dtype=torch.float64) * 2 - 1
ol = n + K - 1; ostr = (ol + 31) // 32 * 32
y = torch.empty((ch, ostr), device="cuda", dtype=torch.float64)
for _ in range(3): plan.process_device(x.data_ptr(), n, ch, n, y.data_ptr(), ostr)
plan.sync()
ctx.kernel_timing(True); ctx.kernel_times(reset=True)
for _ in range(10): plan.process_device(x.data_ptr(), n, ch, n, y.data_ptr(), ostr)
plan.sync()
kt = ctx.kernel_times()
print(os.environ.get("LABEL", ""), {k: (round(v[0], 2), v[1]) for k, v in kt.items() if v[1]})
