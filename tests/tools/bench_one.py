"""One timing of the bench workload under the current env (ADSP_LIB_PATH etc.): prints ms and Gs/s; checks one channel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from algo_dsp_b200 import conv, siggen as G
K = int(os.environ.get("BK", 96000)); n = int(os.environ.get("BN", 480000)); ch = int(os.environ.get("BCH", 256))
dt = np.float32 if os.environ.get("BF32") else np.float64
tdt = torch.float32 if os.environ.get("BF32") else torch.float64
ctx = conv.default_context()
h = G.decaying_ir(K)
plan = conv.NewOverlapSave(h, 0, dtype=dt)
x = torch.rand((ch, n), device="cuda", dtype=tdt) * 2 - 1
ol = n + K - 1; ostr = (ol + 31) // 32 * 32
y = torch.empty((ch, ostr), device="cuda", dtype=tdt)
st = torch.cuda.ExternalStream(ctx.stream())
for _ in range(3): plan.process_device(x.data_ptr(), n, ch, n, y.data_ptr(), ostr)
plan.sync()
iters = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time
e0.record(st)
th0 = time.perf_counter()
for _ in range(iters): plan.process_device(x.data_ptr(), n, ch, n, y.data_ptr(), ostr)
host_ms = (time.perf_counter() - th0) * 1e3 / iters
e1.record(st); plan.sync(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
err = ""
if os.environ.get("BCHECK"):
    from oracle import oracle as O
    ref = O.overlap_save(h, 0, x[ch - 1].cpu().numpy().astype(np.float64))
    err = f" relL2={G.rel_l2(y[ch-1,:ol].cpu().numpy().astype(np.float64), ref):.2e}"
print(f"{os.environ.get('LABEL',''):45s} host-enqueue {host_ms:6.3f} ms | {ms:7.3f} ms {ch*ol/ms/1e6:7.1f} Gs/s {plan.internal_geometry()}{err}", flush=True)
