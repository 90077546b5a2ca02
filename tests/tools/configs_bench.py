"""Timings of the other BASELINE configs (device-resident, reduced batch where noted)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, ctypes as C
from algo_dsp_b200 import conv, siggen as G, _lib as L
ctx = conv.default_context(); lib = L.load()
st = torch.cuda.ExternalStream(ctx.stream())
def timeit(fn, iters=5):
    for _ in range(2): fn()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(iters): fn()
    e1.record(st); ctx.sync(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
# config 2: direct 64-tap FIR, 1024 ch x 2^20
ch, n, m = 1024, 1 << 20, 64
x = torch.rand((ch, n), device="cuda", dtype=torch.float64) * 2 - 1
k = torch.tensor(G.test_kernel(m), device="cuda")
y = torch.empty((ch, n + m - 1), device="cuda", dtype=torch.float64)
ms = timeit(lambda: lib.adsp_direct_batch_device(ctx.handle, x.data_ptr(), n, n, k.data_ptr(), m, 0, ch, y.data_ptr(), n + m - 1, 0))
ref = np.convolve(x[5].cpu().numpy(), G.test_kernel(m))
print(f"config2 direct 64-tap 1024x2^20 f64: {ms:.3f} ms  {ch*(n+m-1)/ms/1e6:.1f} Gs/s  hbm_frac={ch*(n+m-1)*16/ms/1e6/6555.8:.3f} relL2={G.rel_l2(y[5].cpu().numpy(), ref):.1e}", flush=True)
# same through the FFT single-kernel path (what an FFT would give for 64 taps)
plan = conv.NewOverlapSave(G.test_kernel(m), 0)
y2 = torch.empty((ch, n + m - 1), device="cuda", dtype=torch.float64)
ms = timeit(lambda: plan.process_device(x.data_ptr(), n, ch, n, y2.data_ptr(), n + m - 1))
print(f"   (64-tap via FFT path)             : {ms:.3f} ms  {ch*(n+m-1)/ms/1e6:.1f} Gs/s", flush=True)
del x, y, y2; torch.cuda.empty_cache()
# config 3: 64 ch x 14.4M, 288k taps (full size: 7.4 GB in, 7.5 GB out)
ch, n, K = 64, 14_400_000, 288_000
x = torch.rand((ch, n), device="cuda", dtype=torch.float64) * 2 - 1
ol = n + K - 1; ostr = (ol + 31) // 32 * 32
y = torch.empty((ch, ostr), device="cuda", dtype=torch.float64)
h = G.decaying_ir(K)
plan = conv.NewOverlapSave(h, 0)
ms = timeit(lambda: plan.process_device(x.data_ptr(), n, ch, n, y.data_ptr(), ostr), iters=3)
from oracle import oracle as O
seg = x[3, :2_000_000].cpu().numpy(); ref = O.overlap_save(h, 0, seg)[:2_000_000]
print(f"config3 reverb 64ch x 14.4M x 288k taps: {ms:.2f} ms  {ch*ol/ms/1e6:.1f} Gs/s hbm_frac={ch*ol*16/ms/1e6/6555.8:.3f} geom={plan.internal_geometry()} relL2(first 2M)={G.rel_l2(y[3,:2_000_000].cpu().numpy(), ref):.1e}", flush=True)
del x, y; torch.cuda.empty_cache()
# config 4 at full size: 1024 sweep/response pairs x 2^20 samples, FFT correlate + peak lag; responses built on the device
# (sweep delayed by d_p samples, zero-filled front, plus white noise at -40 dB)
pairs, n = 1024, 1 << 20
sweep = G.log_sweep(n)
sw = torch.tensor(sweep, device="cuda")
d = [(p * 2654435761) % 4096 for p in range(pairs)]
A = torch.zeros((pairs, n), device="cuda", dtype=torch.float64)
for p in range(pairs):
    A[p, d[p]:] = sw[: n - d[p]]
gen = torch.Generator(device="cuda"); gen.manual_seed(1000)
A += torch.randn((pairs, n), device="cuda", dtype=torch.float64, generator=gen) * 0.01
B = sw.repeat(pairs, 1)
out = torch.empty((pairs, 2 * n - 1), device="cuda", dtype=torch.float64)
pi = torch.empty(pairs, device="cuda", dtype=torch.int64); pv = torch.empty(pairs, device="cuda", dtype=torch.float64)
lib.adsp_correlate_batch_device(ctx.handle, A.data_ptr(), n, n, B.data_ptr(), n, n, 2, out.data_ptr(), 2 * n - 1, pi.data_ptr(), pv.data_ptr(), 0); ctx.sync()
t0 = time.perf_counter()
stt = lib.adsp_correlate_batch_device(ctx.handle, A.data_ptr(), n, n, B.data_ptr(), n, n, pairs, out.data_ptr(), 2 * n - 1, pi.data_ptr(), pv.data_ptr(), 0)
ctx.sync(); t1 = time.perf_counter()
lags = pi.cpu().numpy() - (n - 1)
print(f"config4 correlate {pairs} pairs 2^20x2^20: status={stt} {1e3*(t1-t0):.1f} ms ({pairs/(t1-t0):.1f} pairs/s, {pairs*(4*n-1)*8/(t1-t0)/1e9:.1f} GB/s algorithmic) "
      f"all {pairs} peak lags exact={np.array_equal(lags, np.array(d))}", flush=True)
from oracle import oracle as O
for p in (3, pairs - 1):
    ref = O.correlate(A[p].cpu().numpy(), sweep)
    print(f"   relL2 pair {p}: {G.rel_l2(out[p].cpu().numpy(), ref):.2e}", flush=True)
