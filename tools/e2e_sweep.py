"""End-to-end (host buffers, pinned) timing of the bench workload through adsp_plan_process_batch for several
pipeline chunk sizes: python tools/e2e_sweep.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from algo_dsp_b200 import conv, siggen as G
K, n, ch = 96000, 480000, 256
ol = n + K - 1
xh = conv.pinned_empty((ch, n)); yh = conv.pinned_empty((ch, ol))
xh[:] = np.random.default_rng(0).uniform(-1, 1, (ch, n))
plan = conv.NewOverlapSave(G.decaying_ir(K), 0)
for mb in [int(a) for a in sys.argv[1:]] or [16, 32, 64, 96, 192, 384]:
    os.environ["ADSP_PIPE_CHUNK_MB"] = str(mb)
    for _ in range(2): plan.ProcessBatch(xh, out=yh)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): plan.ProcessBatch(xh, out=yh)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"chunk {mb:4d} MB: {dt*1e3:7.2f} ms/step  {ch*ol/dt/1e9:6.2f} Gs/s  H2D {ch*n*8/dt/1e9:5.1f} GB/s  D2H {ch*ol*8/dt/1e9:5.1f} GB/s", flush=True)
