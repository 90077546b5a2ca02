"""One host process, one plan per visible GPU: end-to-end (host buffers) throughput of the bench workload through
adsp_plans_process_batch, and the time-block sharded long-signal call.  python tests/tools/multi_gpu_one_process.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from algo_dsp_b200 import conv, siggen as G
from oracle import oracle as O
K, n, ch_per_gpu = 96000, 480000, 128
ng = torch.cuda.device_count()
h = G.decaying_ir(K)
ctxs = [conv.Context(d) for d in range(ng)]
plans = [conv.OverlapSave(h, 0, ctx=c) for c in ctxs]
res = {"gpus": ng}
for use in sorted({1, ng}):
    ch = ch_per_gpu * use
    x = conv.pinned_empty((ch, n)); y = conv.pinned_empty((ch, n + K - 1))
    x[:] = np.random.default_rng(0).uniform(-1, 1, (ch, n))
    for _ in range(2): conv.ProcessBatchMulti(plans[:use], x, out=y)
    t0 = time.perf_counter()
    for _ in range(5): conv.ProcessBatchMulti(plans[:use], x, out=y)
    dt = (time.perf_counter() - t0) / 5
    err = float(G.rel_l2(y[ch - 1], O.overlap_save(h, 0, x[ch - 1])))
    res[f"e2e_{use}gpu"] = {"channels": ch, "ms": dt * 1e3, "gsamples_s": ch * (n + K - 1) / dt / 1e9, "rel_l2_last_channel": err}
xl = G.white(1 << 24, seed=5)
t0 = time.perf_counter(); yl = conv.ProcessLongMulti(plans, xl); dt = time.perf_counter() - t0
ref = O.overlap_save(h, 0, xl[: 1 << 21])
res["long_signal_time_sharded"] = {"n": len(xl), "ms": dt * 1e3, "rel_l2_first_2M": float(G.rel_l2(yl[: 1 << 21], ref[: 1 << 21]))}
print(json.dumps(res), flush=True)
