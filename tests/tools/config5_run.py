"""BASELINE config 5 at full size on one GPU: a single 2^31-sample signal convolved with a 2^20-tap IR
(overlap-save, device resident), unsharded and as 8 time-block shards with a (K-1) halo -- the per-GPU
work of the 8-GPU decomposition, run back to back on one device.  Parity on sampled
windows against the CPU oracle (first 2^22 outputs, the last 2^22 including the tail, every shard
boundary +- 2^20), fp64 and fp32.

    python tests/tools/config5_run.py [log2_n=31] [log2_K=20] [shards=8] [f64|f32|both]
Prints one JSON line per precision.
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from algo_dsp_b200 import conv, siggen as G
from algo_dsp_b200.shard import time_shards
from oracle import oracle as O

lgn = int(sys.argv[1]) if len(sys.argv) > 1 else 31
lgk = int(sys.argv[2]) if len(sys.argv) > 2 else 20
world = int(sys.argv[3]) if len(sys.argv) > 3 else 8
which = sys.argv[4] if len(sys.argv) > 4 else "both"
n, K = 1 << lgn, 1 << lgk
out_len = n + K - 1
W = min(1 << 22, n // 4)


def run(dtype):
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    tol = 1e-12 if dtype == np.float64 else 1e-5
    ctx = conv.default_context()
    h = G.decaying_ir(K)
    plan = conv.NewOverlapSave(h, 0, dtype=dtype)
    stream = torch.cuda.ExternalStream(ctx.stream())
    gen = torch.Generator(device="cuda"); gen.manual_seed(1)
    x = torch.empty(n, device="cuda", dtype=tdt)
    CH = 1 << 26
    for o in range(0, n, CH):          # white noise (u*2-1), generated on the device in chunks
        m = min(CH, n - o)
        x[o:o + m] = torch.rand(m, device="cuda", dtype=tdt, generator=gen) * 2 - 1
    y = torch.empty(out_len + 32, device="cuda", dtype=tdt)
    torch.cuda.synchronize()

    def timed(fn, iters):
        fn(); plan.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(iters): fn()
        e1.record(stream); plan.sync(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    ms_full = timed(lambda: plan.process_device(x.data_ptr(), n, 1, n, y.data_ptr(), out_len + 32), 3)

    # 8 time-block shards with halo, each the work of one GPU of the 8-GPU run
    shards = time_shards(n, K, world)
    y2 = torch.zeros(out_len, device="cuda", dtype=tdt)
    seg_max = max(s.in_hi - s.in_lo for s in shards)
    tmp = torch.empty(seg_max + K - 1 + 32, device="cuda", dtype=tdt)
    esz = x.element_size()
    shard_ms = []
    for s in shards:
        if s.out_hi <= s.out_lo: continue
        seg = s.in_hi - s.in_lo
        ms = timed(lambda: plan.process_device(x.data_ptr() + s.in_lo * esz, seg, 1, seg, tmp.data_ptr(), seg + K - 1 + 32), 2)
        shard_ms.append(ms)
        y2[s.out_lo:s.out_hi] = tmp[s.skip:s.skip + (s.out_hi - s.out_lo)]
    torch.cuda.synchronize()
    # sharded == unsharded (different block alignment: equal to rounding)
    num = den = 0.0
    for o in range(0, out_len, CH):
        m = min(CH, out_len - o)
        d = (y2[o:o + m].double() - y[o:o + m].double())
        num += float((d * d).sum()); den += float((y[o:o + m].double() ** 2).sum())
    shard_vs_full = (num / den) ** 0.5

    # oracle windows
    wins = [(0, W), (out_len - W, out_len)]
    for s in shards[1:]:
        if s.out_hi > s.out_lo:
            wins.append((max(0, s.out_lo - (1 << 20)), min(out_len, s.out_lo + (1 << 20))))
    worst = 0.0
    t0 = time.perf_counter()
    for a, b in wins:
        lo = max(0, a - (K - 1))
        xs = x[lo:min(b, n)].cpu().numpy().astype(np.float64)
        ref = O.overlap_save(h, 0, xs)[a - lo:a - lo + (b - a)]
        for yy in (y, y2):
            got = yy[a:b].cpu().numpy().astype(np.float64)
            worst = max(worst, float(G.rel_l2(got, ref)))
    oracle_s = time.perf_counter() - t0
    res = {
        "config": "single long signal, time-block sharded", "dtype": str(np.dtype(dtype)), "n": n, "K": K, "out_len": out_len,
        "internal_fft": plan.internal_geometry(),
        "unsharded_ms": ms_full, "unsharded_gsamples_s": out_len / ms_full / 1e6,
        "hbm_frac_unsharded": out_len * 2 * esz / (ms_full * 1e-3) / 6555.8e9,
        "shards": world, "shard_ms": [round(v, 3) for v in shard_ms], "shard_ms_max": max(shard_ms),
        "projected_8gpu_gsamples_s": out_len / max(shard_ms) / 1e6,
        "halo_overhead": (K - 1) / (n / world),
        "sharded_vs_unsharded_rel_l2": shard_vs_full,
        "windows_checked": len(wins), "worst_window_rel_l2_vs_oracle": worst, "tolerance": tol, "parity_ok": bool(worst <= tol),
        "oracle_seconds": round(oracle_s, 1),
    }
    print(json.dumps(res), flush=True)
    plan.Close()
    del x, y, y2, tmp
    torch.cuda.empty_cache()


if which in ("f64", "both"): run(np.float64)
if which in ("f32", "both"): run(np.float32)
