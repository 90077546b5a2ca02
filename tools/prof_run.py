"""Tiny fixed workload for ncu captures: python tools/prof_run.py K n channels [iters] [dtype]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from algo_dsp_b200 import conv, siggen as G
K, n, ch = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dtype = np.float32 if (len(sys.argv) > 5 and sys.argv[5] == "f32") else np.float64
tdt = torch.float64 if dtype == np.float64 else torch.float32
plan = conv.NewOverlapSave(G.decaying_ir(K), 0, dtype=dtype)
x = torch.rand((ch, n), device="cuda", dtype=tdt) * 2 - 1
ol = n + K - 1; os_ = (ol + 31) // 32 * 32
y = torch.empty((ch, os_), device="cuda", dtype=tdt)
torch.cuda.synchronize()
for _ in range(iters):
    plan.process_device(x.data_ptr(), n, ch, n, y.data_ptr(), os_)
plan.sync()
print("ok", plan.internal_geometry(), float(y[0, :8].abs().sum()))
