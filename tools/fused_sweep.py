import os, sys, json, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from algo_dsp_b200 import conv, siggen as G
K, n, ch = 96000, 480000, 256
ctx = conv.default_context()
x = torch.rand((ch, n), device="cuda", dtype=torch.float64) * 2 - 1
ol = n + K - 1; ostr = (ol + 31) // 32 * 32
y = torch.empty((ch, ostr), device="cuda", dtype=torch.float64)
st = torch.cuda.ExternalStream(ctx.stream())
def run(env, iters=8, label=""):
    for k, v in env.items(): os.environ[k] = str(v)
    plan = conv.NewOverlapSave(G.decaying_ir(K), 0)
    for _ in range(2): plan.process_device(x.data_ptr(), n, ch, n, y.data_ptr(), ostr)
    plan.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(iters): plan.process_device(x.data_ptr(), n, ch, n, y.data_ptr(), ostr)
    e1.record(st); plan.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{label:40s} {ms:.3f} ms  {ch*ol/ms/1e6:.1f} Gs/s", flush=True)
    if env.get("ADSP_FUSED_STATS"):
        pass
    for k in env: os.environ.pop(k, None)
    plan.Close()
run({"ADSP_NO_FUSED": 1}, label="three-kernel")
run({}, label="fused dynamic tickets")
run({"ADSP_FUSED_FLAGS": 2}, label="fused static tickets")
run({"ADSP_FUSED_FLAGS": 1}, label="fused dynamic DRY (sched only)")
run({"ADSP_FUSED_FLAGS": 3}, label="fused static DRY (sched only)")
os.environ["ADSP_FUSED_STATS"] = "1"
for fl in (0, 2):
    os.environ["ADSP_FUSED_FLAGS"] = str(fl)
    plan = conv.NewOverlapSave(G.decaying_ir(K), 0)
    plan.process_device(x.data_ptr(), n, ch, n, y.data_ptr(), ostr); plan.sync()
    plan.process_device(x.data_ptr(), n, ch, n, y.data_ptr(), ostr); plan.sync(); plan.Close()
os.environ.pop("ADSP_FUSED_STATS"); os.environ.pop("ADSP_FUSED_FLAGS")
for lag2 in (4, 6, 8):
    run({"ADSP_FUSED_LAG_X2": lag2, "ADSP_FUSED_FLAGS": 2}, label=f"fused static lag={lag2/2}")
