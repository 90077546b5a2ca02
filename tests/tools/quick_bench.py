"""Device-resident timing sweep of the OLS path (development aid; bench.py is the contract)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from algo_dsp_b200 import conv
from algo_dsp_b200 import siggen as G

def run(K, n, channels, dtype=np.float64, iters=5, env=None, check=None):
    env = env or {}
    for k, v in env.items(): os.environ[k] = str(v)
    ctx = conv.default_context()
    h = G.decaying_ir(K)
    plan = conv.NewOverlapSave(h, 0, dtype=dtype)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    x = (torch.rand((channels, n), device="cuda", dtype=tdt) * 2 - 1)
    out_len = n + K - 1
    ostride = (out_len + 31) // 32 * 32
    y = torch.empty((channels, ostride), device="cuda", dtype=tdt)
    st = torch.cuda.ExternalStream(ctx.stream())
    torch.cuda.synchronize()
    def step():
        plan.process_device(x.data_ptr(), n, channels, n, y.data_ptr(), ostride)
    for _ in range(2): step()
    plan.sync()
    evs = []
    for _ in range(iters):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st); step(); e1.record(st); evs.append((e0, e1))
    plan.sync(); torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    med = ms[len(ms) // 2]
    samples = channels * out_len
    esz = 8 if dtype == np.float64 else 4
    res = dict(K=K, n=n, ch=channels, dtype=str(np.dtype(dtype)), geom=plan.internal_geometry(), env=env, ms=med, ms_min=ms[0],
               gsamples_s=samples / med / 1e6, hbm_frac=samples * 2 * esz / (med * 1e-3) / 6555.8e9)
    if check is not None:
        from oracle import oracle as O
        c = check
        ref = O.overlap_save(h, 0, x[c].cpu().numpy().astype(np.float64))
        got = y[c, :out_len].cpu().numpy().astype(np.float64)
        res["relL2"] = float(G.rel_l2(got, ref))
    for k in env: os.environ.pop(k, None)
    plan.Close()
    print(json.dumps(res), flush=True)
    return res

if __name__ == "__main__":
    K, n = 96000, 480000
    run(K, n, 64, check=63)
    for N in (1 << 18, 1 << 19, 1 << 20):
        for N2 in (2048, 4096):
            run(K, n, 256, env=dict(ADSP_FFT_N=N, ADSP_FFT_N2=N2))
    for mb in (24, 48, 96, 200):
        run(K, n, 256, env=dict(ADSP_FFT_N=1 << 19, ADSP_SCRATCH_MB=mb))
    run(K, n, 256, dtype=np.float32, check=5)
    run(1000, 1 << 20, 256, check=3)
    run(64, 1 << 20, 256, check=3)
