"""Quick GPU sanity sweep: library vs oracle over many shapes; prints worst errors and timings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from algo_dsp_b200 import conv
from oracle import oracle as O
from algo_dsp_b200 import siggen as G

def rel(y, r): return G.rel_l2(y, r)
rng = np.random.default_rng(0)
worst = 0
os.environ.setdefault("X", "1")
for K, n in [(3, 1000), (4, 500), (64, 4096), (65, 1000), (100, 1000), (257, 5000), (1000, 30000), (1025, 10000), (2048, 100000),
             (5000, 50000), (9000, 9000), (20000, 100000), (50000, 300000), (96000, 480000), (131073, 300000), (300000, 700000), (600000, 700000)]:
    h = G.decaying_ir(K); x = G.white(n, seed=K)
    t0 = time.time(); ref = O.overlap_save(h, 0, x); t1 = time.time()
    c = conv.NewOverlapSave(h, 0); y = c.Process(x); t2 = time.time()
    e = rel(y, ref); worst = max(worst, e)
    print(f"OLS K={K:7d} n={n:7d} geom={c.internal_geometry()} relL2={e:.2e} cpu={t1-t0:.3f}s gpu(host api)={t2-t1:.3f}s", flush=True)
    y2 = conv.Convolve(x, h); e2 = rel(y2, ref)
    print(f"    Convolve relL2={e2:.2e}")
    worst = max(worst, e2)
print("WORST", worst)
