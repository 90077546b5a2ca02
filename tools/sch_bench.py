"""Schroeder integral + impulse start over 64 rows x 2^21 samples, device resident: python tools/sch_bench.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from algo_dsp_b200 import conv, siggen as G, _lib as L
ctx = conv.default_context(); lib = L.load()
st = torch.cuda.ExternalStream(ctx.stream())
rows, n = 64, 1 << 21
x = torch.empty((rows, n), device="cuda", dtype=torch.float64)
G.white_device(ctx, x.data_ptr(), n, rows, n, amp=1.0, seed0=500, seed_step=1)
y = torch.empty((rows, n), device="cuda", dtype=torch.float64)
idx = torch.empty(rows, device="cuda", dtype=torch.int64)
def timeit(fn, iters=10):
    for _ in range(3): fn()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(iters): fn()
    e1.record(st); ctx.sync(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
ms = timeit(lambda: lib.adsp_ir_schroeder_device(ctx.handle, x.data_ptr(), n, rows, n, y.data_ptr(), n))
mo = timeit(lambda: lib.adsp_ir_find_impulse_start_device(ctx.handle, x.data_ptr(), n, rows, n, C.c_double(0.1), idx.data_ptr()))
print(f"schroeder {ms:.3f} ms ({rows*n*16/ms/1e6:.0f} GB/s algorithmic)  impulse start {mo:.3f} ms ({rows*n*8/mo/1e6:.0f} GB/s)", flush=True)
