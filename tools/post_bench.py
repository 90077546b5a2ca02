"""FIR block filter and polyphase resampler, device resident, 64 channels x 2^20 samples:
    [ADSP_RESAMPLE_TILED=0] python tools/post_bench.py [up=160] [down=147] [fir_taps=257]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from algo_dsp_b200 import conv, _lib as L
up = int(sys.argv[1]) if len(sys.argv) > 1 else 160
down = int(sys.argv[2]) if len(sys.argv) > 2 else 147
ft = int(sys.argv[3]) if len(sys.argv) > 3 else 257
ctx = conv.default_context(); lib = L.load()
st = torch.cuda.ExternalStream(ctx.stream())
rows, n = 64, 1 << 20
x = torch.rand((rows, n), device="cuda", dtype=torch.float64) * 2 - 1

def timeit(fn, iters=10):
    for _ in range(3): fn()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(iters): fn()
    e1.record(st); ctx.sync(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

fh = C.c_void_p()
taps = np.hanning(ft); taps /= taps.sum()
assert lib.adsp_fir_create(ctx.handle, taps.ctypes.data_as(C.c_void_p), ft, rows, C.byref(fh)) == 0
blk = x.clone()
ms_f = timeit(lambda: lib.adsp_fir_process_block_device(fh, blk.data_ptr(), n, n))
lib.adsp_fir_destroy(fh)
rh = C.c_void_p()
assert lib.adsp_resampler_create(ctx.handle, up, down, 1, 0, C.c_double(0), C.c_double(0), rows, C.byref(rh)) == 0
n_r = int(lib.adsp_resampler_predict_output_len(rh, n))
ro = torch.empty((rows, n_r + 64), device="cuda", dtype=torch.float64)
got = C.c_int64()
def rs():
    lib.adsp_resampler_reset(rh)
    assert lib.adsp_resampler_process_device(rh, x.data_ptr(), n, n, ro.data_ptr(), n_r + 64, n_r + 64, C.byref(got)) == 0
ms_r = timeit(rs)
tpp = int(lib.adsp_resampler_taps_per_phase(rh))
cs = float(ro[:, :n_r].double().abs().sum())
print(f"{os.environ.get('LABEL',''):10s} fir {ft} taps: {ms_f:.3f} ms {rows*n/ms_f/1e6:.1f} Gs/s | resample {up}/{down} ({tpp} taps/phase): {ms_r:.3f} ms "
      f"{rows*n_r/ms_r/1e6:.1f} G out/s checksum {cs!r}", flush=True)
