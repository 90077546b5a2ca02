"""Host-path sweep on the GPU box: end-to-end time of adsp_plan_process_batch on PAGEABLE buffers (the Go case) against
copy threads, chunk size and streaming stores, next to the pinned-buffer leg.  One subprocess per setting (the knobs
are read once per process).

    python tools/stage_sweep.py            # prints one JSON line per setting
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    from algo_dsp_b200 import conv, siggen as G
    K, n, ch = 96000, 480000, 256
    ctx = conv.Context(0)
    plan = conv.OverlapSave(G.decaying_ir(K), 0, ctx=ctx)
    pinned = os.environ.get("SWEEP_PINNED") == "1"
    x = conv.pinned_empty((ch, n)) if pinned else np.empty((ch, n))
    y = conv.pinned_empty((ch, n + K - 1)) if pinned else np.empty((ch, n + K - 1))
    x[:] = G.white(n, seed=1)[None, :]
    for _ in range(2):
        plan.ProcessBatch(x, out=y)
    ts = []
    for _ in range(8):
        t0 = time.perf_counter()
        plan.ProcessBatch(x, out=y)
        ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    # config-1 sized mono call, pageable, preallocated output (ProcessTo) and fresh output (Process)
    xm, ym = np.array(x[0]), np.empty(n + K - 1)
    for _ in range(3):
        plan.ProcessTo(ym, xm)
    lat = []
    for _ in range(20):
        t0 = time.perf_counter()
        plan.ProcessTo(ym, xm)
        lat.append((time.perf_counter() - t0) * 1e3)
    lat.sort()
    print(json.dumps({"pinned": pinned, "threads": os.environ.get("ADSP_STAGE_THREADS"), "chunk_mb": os.environ.get("ADSP_STAGE_PIPE_CHUNK_MB"),
                      "nt": os.environ.get("ADSP_STAGE_NT"), "small_chunk_mb": os.environ.get("ADSP_STAGE_CHUNK_MB"),
                      "batch_ms_median": ts[len(ts) // 2], "batch_ms_min": ts[0],
                      "mono_ms_median": lat[len(lat) // 2], "mono_ms_min": lat[0], "mono_ms_max": lat[-1]}), flush=True)


if __name__ == "__main__":
    if os.environ.get("SWEEP_CHILD"):
        child()
        sys.exit(0)
    settings = [{"SWEEP_PINNED": "1"}]
    for th in ("4", "8", "12", "15"):
        settings.append({"ADSP_STAGE_THREADS": th})
    for ck in ("8", "16", "64"):
        settings.append({"ADSP_STAGE_THREADS": "12", "ADSP_STAGE_PIPE_CHUNK_MB": ck})
    settings.append({"ADSP_STAGE_THREADS": "12", "ADSP_STAGE_NT": "0"})
    settings.append({"ADSP_STAGE_THREADS": "12", "ADSP_STAGE_CHUNK_MB": "1"})
    settings.append({"ADSP_STAGE_THREADS": "12", "ADSP_STAGE_CHUNK_MB": "2"})
    for st in settings:
        env = dict(os.environ, SWEEP_CHILD="1", **st)
        r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True)
        sys.stdout.write(r.stdout if r.returncode == 0 else json.dumps({"setting": st, "error": r.stderr[-400:]}) + "\n")
        sys.stdout.flush()
