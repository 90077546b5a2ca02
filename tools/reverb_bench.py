"""Streaming convolution reverb (BASELINE config 3 as a real-time stream): 64 channels through a 288 000-tap IR in
blocks of n samples per call, device resident (adsp_partitioned_process_in_place_batch_device).
    python tools/reverb_bench.py [channels=64] [K=288000] [min_order=7]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from algo_dsp_b200 import conv, siggen as G, _lib as L
import ctypes as C

ch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = int(sys.argv[2]) if len(sys.argv) > 2 else 288000
mn = int(sys.argv[3]) if len(sys.argv) > 3 else 7
ctx = conv.default_context()
stream = torch.cuda.ExternalStream(ctx.stream())
lib = L.load()
r = conv.NewConvolutionReverb(G.decaying_ir(K), mn, channels=ch)
r.SetWetDry(0.3, 0.7)
res = {"channels": ch, "kernel_taps": K, "latency": r.Latency(), "internal_stages": r.internal_stages(), "blocks": []}
for n in (128, 512, 2048, 8192, 65536):
    x = torch.rand((ch, n), device="cuda", dtype=torch.float64) * 2 - 1
    def call():
        st = lib.adsp_partitioned_process_in_place_batch_device(r._h, C.c_void_p(x.data_ptr()), n, n)
        assert st == 0
    for _ in range(3): call()
    ctx.sync()
    iters = max(5, min(200, (1 << 20) // n))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters): call()
    e1.record(stream); ctx.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    res["blocks"].append({"n": n, "ms_per_call": round(ms, 4), "gsamples_s": round(ch * n / ms / 1e6, 3),
                          "x_realtime_48k": round(n / 48000.0 / (ms * 1e-3), 1)})
print(json.dumps(res), flush=True)
