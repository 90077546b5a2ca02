"""Where does the host-API latency of a config-1-sized mono call go?  Raw library copies, then the profiled call."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from algo_dsp_b200 import conv, siggen as G, _lib as L
lib = L.load(); ctx = conv.Context(0)
nb = 575999 * 8
hp = conv.pinned_empty(575999)
dptr = C.c_void_p(); lib.adsp_device_alloc(ctx.handle, nb, C.byref(dptr))
def tm(fn, k=20):
    fn(); ts = []
    for _ in range(k):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort(); return round(ts[len(ts) // 2], 3), round(ts[0], 3), round(ts[-1], 3)
print("adsp_memcpy_h2d 4.6 MB pinned   (median, min, max ms):", tm(lambda: lib.adsp_memcpy_h2d(ctx.handle, dptr, hp.ctypes.data_as(C.c_void_p), nb)))
print("adsp_memcpy_d2h 4.6 MB pinned   :", tm(lambda: lib.adsp_memcpy_d2h(ctx.handle, hp.ctypes.data_as(C.c_void_p), dptr, nb)))
pg = np.empty(575999)
print("adsp_memcpy_d2h 4.6 MB pageable (runtime bounce):", tm(lambda: lib.adsp_memcpy_d2h(ctx.handle, pg.ctypes.data_as(C.c_void_p), dptr, nb)))
h, x = G.decaying_ir(96000), G.white(480000, seed=1)
plan = conv.OverlapSave(h, 0, ctx=ctx)
y = np.empty(575999)
print("plan.ProcessTo pageable:", tm(lambda: plan.ProcessTo(y, x)))
xp, yp = conv.pinned_empty(480000), conv.pinned_empty(575999); xp[:] = x
print("plan.ProcessTo pinned  :", tm(lambda: plan.ProcessTo(yp, xp)))
def pinned_then_read():
    plan.ProcessTo(yp, xp); return float(yp.sum())
print("plan.ProcessTo pinned, caller reads the output between calls:", tm(pinned_then_read))
ctx.host_profile(True)
for name, a, b in (("pageable", x, y), ("pinned", xp, yp)):
    plan.ProcessTo(b, a); plan.ProcessTo(b, a)
    print("profile", name, {k: round(v, 4) for k, v in ctx.host_profile_get().items() if k.endswith("_ms")})
ctx.host_profile(False)
import torch
xd = torch.tensor(x[None, :], device="cuda"); yd = torch.empty((1, 576000), device="cuda", dtype=torch.float64)
def dev():
    plan.process_device(xd.data_ptr(), 480000, 1, 480000, yd.data_ptr(), 576000); plan.sync()
print("device-resident call + sync:", tm(dev))
os.environ["ADSP_GRAPHS"] = "0"
