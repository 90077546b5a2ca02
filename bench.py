#!/usr/bin/env python
"""bench.py -- headline benchmark of the dsp/conv hot path (BASELINE.json metric):
output samples/s for overlap-save convolution with a 96 000-tap IR, 1/2/4/8 B200, % HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload ("ols96k_batch"): config 1's IR and signal shape (96 000-tap decaying IR, 480 000-sample
white-noise signals, fp64) batched to 256 channels per GPU so that the path is throughput- and
not launch-latency-bound (config 1 alone moves 8.4 MB).  One step = one
OverlapSave.Process over the whole batch.  Channels are independent, so N GPUs shard by channel
with no collective (weak scaling: 256 channels per GPU).

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM, CUDA-event timed on the
library's stream.  `e2e`: the same metric through the host-buffer C-ABI call
(adsp_plan_process_batch) with pinned host buffers, H2D and D2H inside the timed region; `e2e.pageable` is the same call
on ordinary pageable numpy memory (what a Go []float64 caller passes: staged through the library's pinned slots), and
`e2e.copy_ceiling` the raw concurrent H2D+D2H cudaMemcpyAsync time of the same byte counts.  `sustained` repeats the
timed loop for >= 2 s with the clock/power sampler on.  `roofline.fp64` / `roofline.lsu` put the step against the FP64
and shared-memory pipe peaks measured in the same run (adsp_ctx_measure_pipes).
`--impl reference` times the CPU restatement of the reference (oracle/, all host threads) on a
bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

K_TAPS = 96000
N_SAMPLES = 480000
CHANNELS_PER_GPU = 256
OUT_LEN = N_SAMPLES + K_TAPS - 1
ALGO_BYTES_PER_SAMPLE = 16  # fp64: one input sample read + one output sample written
METRIC = "ols_conv_96k_tap_output_samples_per_s"
UNIT = "samples/s"


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.power = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            if self._stop.is_set():
                break
            parts = [p.strip() for p in line.split(",")]
            try:
                self.samples.append(float(parts[0]))
                self.max_mhz = float(parts[1])
                for nm, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
                self.power.append(float(parts[6]))
            except Exception:
                pass

    def stop(self):
        self._stop.set()
        if self.proc:
            self.proc.terminate()

    def mark(self):
        """index of the next sample (to summarise a window of the run)"""
        return len(self.samples), len(self.power)

    def summary(self, since=(0, 0)):
        s = sorted(self.samples[since[0]:])
        pw = self.power[since[1]:]
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s),
                "power_w_max": max(pw) if pw else None}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def cpu_baseline(channels, threads, budget_s=12.0, max_passes=64):
    """Times the CPU restatement of the reference (oracle/conv_oracle.c, OverlapSave.Process shape:
    N=262144 complex FFT per block) on `channels` channels of the bench workload, repeated until about
    `budget_s` seconds of wall time have been spent (a bounded sample); returns samples/s over all passes."""
    from oracle import oracle as O
    from algo_dsp_b200 import siggen as G
    O.build()
    h = G.decaying_ir(K_TAPS)
    x = np.stack([G.white(N_SAMPLES, seed=1 + c) for c in range(channels)])
    O.bench_ols(h, x, 0, threads)        # warm-up pass (page faults, thread start)
    passes, t0 = 0, time.perf_counter()
    while passes < max_passes and (passes == 0 or time.perf_counter() - t0 < budget_s):
        O.bench_ols(h, x, 0, threads)
        passes += 1
    dt = time.perf_counter() - t0
    return passes * channels * OUT_LEN / dt, dt, passes


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    threads = O.num_procs()
    channels = CHANNELS_PER_GPU          # the same 256-channel step as the GPU arm (about a second of CPU work per step on 16 cores)
    from algo_dsp_b200 import siggen as G
    h = G.decaying_ir(K_TAPS)
    x = np.stack([G.white(N_SAMPLES, seed=1 + c) for c in range(channels)])
    for _ in range(args.warmup):
        O.bench_ols(h, x, 0, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.bench_ols(h, x, 0, threads)
    dt = (time.perf_counter() - t0) / args.steps
    value = channels * OUT_LEN / dt
    sample = f"{channels} channels x {N_SAMPLES} samples x {K_TAPS} taps per step (the whole {CHANNELS_PER_GPU}-channel step of the GPU arm)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "ols96k_batch", "kernel_taps": K_TAPS, "signal_samples": N_SAMPLES, "channels_per_gpu": channels,
                   "output_samples_per_channel": OUT_LEN,
                   "reference_shape": "OverlapSave.Process, N=262144 complex128 FFT per block (overlap_save.go:126-254)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "C restatement of the Go reference (Go toolchain and algo-fft are absent); one convolver per thread"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def other_configs(torch, conv, G, ctx, stream):
    """Device-resident timings of BASELINE.json configs 1-4 at reduced batch sizes (same kernels, same
    C-ABI entry points).  Auxiliary: the headline metric is unaffected."""
    import ctypes as C
    from algo_dsp_b200 import _lib as L
    lib = L.load()
    peak, _ = measured_peak_gbs()

    def timeit(fn, iters=5):
        for _ in range(2):
            fn()
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(iters):
            fn()
        e1.record(stream)
        ctx.sync()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    def gen_white(rows, n, seed0):
        t = torch.empty((rows, n), device="cuda", dtype=torch.float64)
        G.white_device(ctx, t.data_ptr(), n, rows, n, amp=1.0, seed0=seed0, seed_step=1)
        ctx.sync()                       # the generator runs on the library's stream, torch on its own
        return t

    out = {}
    # config 1: mono 10 s @48 kHz, 96k taps -- latency of one Process (device resident, and through the host API)
    h = G.decaying_ir(K_TAPS)
    plan = conv.OverlapSave(h, 0, ctx=ctx)
    x1 = gen_white(1, N_SAMPLES, 1)
    y1 = torch.empty((1, OUT_LEN + 31), device="cuda", dtype=torch.float64)
    ms = timeit(lambda: plan.process_device(x1.data_ptr(), N_SAMPLES, 1, N_SAMPLES, y1.data_ptr(), OUT_LEN + 31), iters=20)
    xh = x1[0].cpu().numpy()            # pageable, like a Go []float64
    for _ in range(3):
        plan.Process(xh)                 # warm the host path: device buffers, pinned slots, copy threads
    lat = []
    for _ in range(15):
        t0 = time.perf_counter()
        plan.Process(xh)
        lat.append((time.perf_counter() - t0) * 1e3)
    lat.sort()
    ctx.host_profile(True)
    plan.Process(xh)
    breakdown = ctx.host_profile_get()
    ctx.host_profile(False)
    out["config1_mono_ols_96k"] = {"device_latency_ms": ms, "host_api_latency_ms": lat[len(lat) // 2], "host_api_latency_ms_min": lat[0],
                                   "host_api_latency_ms_max": lat[-1], "host_api_calls": len(lat),
                                   "host_api_breakdown_ms": {k: round(v, 4) for k, v in breakdown.items() if k.endswith("_ms")},
                                   "host_api_note": "pageable numpy buffers through adsp_plan_process (Python wrapper included); breakdown phases are serialised by a sync each",
                                   "samples_per_s_device": OUT_LEN / ms * 1e3}
    plan.Close()
    # config 2: direct 64-tap FIR (Convolve auto-select -> direct), 256 of the 1024 channels x 2^20
    ch, n, m = 256, 1 << 20, 64
    x = gen_white(ch, n, 1)
    k = torch.tensor(G.test_kernel(m), device="cuda")
    y = torch.empty((ch, n + m - 1), device="cuda", dtype=torch.float64)
    ms = timeit(lambda: lib.adsp_direct_batch_device(ctx.handle, x.data_ptr(), n, n, k.data_ptr(), m, 0, ch, y.data_ptr(), n + m - 1, 0))
    sps = ch * (n + m - 1) / ms * 1e3
    # the kernel is 64 DFMA per output sample: sample clock and power while it runs back to back (a pure FP64 stream sits at
    # the 1 kW power cap, so the clock it really runs at matters for the FP64-pipe fraction)
    cs = ClockSampler(torch.cuda.current_device())
    cs.start()
    time.sleep(0.2)
    mark = cs.mark()
    t0 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t0 < 1.0:
        for _ in range(20):
            lib.adsp_direct_batch_device(ctx.handle, x.data_ptr(), n, n, k.data_ptr(), m, 0, ch, y.data_ptr(), n + m - 1, 0)
        ctx.sync()
        reps += 20
    sus_ms = (time.perf_counter() - t0) / reps * 1e3
    cl = cs.summary(mark)
    cs.stop()
    dfma_per_s = ch * (n + m - 1) * m / (sus_ms * 1e-3)
    out["config2_direct_64tap"] = {"channels": ch, "samples_per_s": sps, "hbm_frac": sps * 16 / 1e9 / peak, "ms": ms,
                                   "sustained_1s": {"ms": sus_ms, "samples_per_s": ch * (n + m - 1) / sus_ms * 1e3, "sm_mhz_median": cl["sm_mhz"],
                                                    "power_w_max": cl["power_w_max"], "reasons": cl["reasons"], "dfma_per_s": dfma_per_s,
                                                    "dfma_per_clk_per_sm": (dfma_per_s / (cl["sm_mhz"] * 1e6) / 148) if cl["sm_mhz"] else None}}
    del x, y
    time.sleep(1.0)      # the 1 s DFMA burn above ends at the power cap (~1700 MHz): let the clocks recover before the next leg
    # config 3: long-IR reverb shape, 8 of the 64 channels x 14.4 M samples, 288k taps
    ch, n, K = 8, 14_400_000, 288_000
    x = torch.empty((ch, n), device="cuda", dtype=torch.float64)
    G.pink_device(ctx, x.data_ptr(), n, ch, n, amp=1.0, seed0=100, seed_step=1)     # pink, seed 100 + channel
    ctx.sync()
    ol = n + K - 1
    ostr = (ol + 31) // 32 * 32
    y = torch.empty((ch, ostr), device="cuda", dtype=torch.float64)
    plan = conv.OverlapSave(G.decaying_ir(K), 0, ctx=ctx)
    ms = timeit(lambda: plan.process_device(x.data_ptr(), n, ch, n, y.data_ptr(), ostr), iters=5)
    sps = ch * ol / ms * 1e3
    out["config3_reverb_288k"] = {"channels": ch, "samples_per_s": sps, "hbm_frac": sps * 16 / 1e9 / peak, "ms": ms,
                                  "internal_fft": plan.internal_geometry()}
    plan.Close()
    del x, y
    # config 3 as a real-time stream: 64 channels through the same 288k-tap IR in 8192- and 128-sample blocks, wet/dry mixed in
    # place (ConvolutionReverb.ProcessInPlace) on the device-resident frequency-domain delay line
    rv = conv.NewConvolutionReverb(G.decaying_ir(K), 7, ctx=ctx, channels=64)
    rv.SetWetDry(0.3, 0.7)
    stream_res = {"channels": 64, "latency_samples": rv.Latency(), "internal_stages": rv.internal_stages()}
    for nblk in (8192, 128):
        xb = gen_white(64, nblk, 1)
        ms = timeit(lambda: lib.adsp_partitioned_process_in_place_batch_device(rv._h, C.c_void_p(xb.data_ptr()), nblk, nblk), iters=50)
        stream_res[f"block_{nblk}"] = {"ms_per_call": ms, "samples_per_s": 64 * nblk / ms * 1e3, "x_realtime_48k": nblk / 48000.0 / (ms * 1e-3)}
    out["config3_streaming_reverb_288k"] = stream_res
    rv.Close()
    # config 4: sweep/response correlation + peak lag, 256 of the 1024 pairs x 2^20.  SURVEY 8d: every pair correlates against
    # the SAME log sweep b; a_p = b delayed by d_p = hash(p) mod 4096 + white at -40 dB (seed 1000 + p); all built in HBM.
    # (a) b_stride = 0, one b for all pairs: a batched convolution with one cached spectrum; (b) every pair with its own
    # copy of b: packed transform + mirror-bin product per pair.  Wall clock around the (synchronous) call.
    pairs, n = 256, 1 << 20
    B1 = torch.empty((1, n), device="cuda", dtype=torch.float64)
    G.log_sweep_device(ctx, B1.data_ptr(), n)
    A = torch.empty((pairs, n), device="cuda", dtype=torch.float64)
    G.delay_mix_device(ctx, A.data_ptr(), n, pairs, n, B1.data_ptr(), noise_amp=0.01, seed0=1000, delay_seed=0, delay_mod=4096)
    ctx.sync()
    delays = np.array([G.delay_of(p) for p in range(pairs)])
    B = B1.expand(pairs, n).contiguous()
    o = torch.empty((pairs, 2 * n - 1), device="cuda", dtype=torch.float64)
    pi = torch.empty(pairs, device="cuda", dtype=torch.int64)
    pv = torch.empty(pairs, device="cuda", dtype=torch.float64)

    def corr(bptr, bstride, with_out):
        st = lib.adsp_correlate_batch_device(ctx.handle, A.data_ptr(), n, n, bptr, n, bstride, pairs, o.data_ptr() if with_out else None,
                                             2 * n - 1 if with_out else 0, pi.data_ptr(), pv.data_ptr(), 0)
        assert st == 0, L.last_error()
        ctx.sync()

    def wall(fn, iters=3):
        fn()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        return (time.perf_counter() - t0) / iters * 1e3

    c4 = {"pairs": pairs}
    for name, bptr, bstride in (("shared_b", B1.data_ptr(), 0), ("per_pair_b", B.data_ptr(), n)):
        for with_out in (True, False):
            ms = wall(lambda: corr(bptr, bstride, with_out))
            ok = bool(np.array_equal(pi.cpu().numpy() - (n - 1), delays))
            alg = (4 * n - 1) * 8 if with_out else 2 * n * 8
            c4[name + ("" if with_out else "_peaks_only")] = {"pairs_per_s": pairs / ms * 1e3, "ms": ms, "lags_exact": ok,
                                                               "algorithmic_GBps": pairs * alg / ms / 1e6, "hbm_frac": pairs * alg / ms / 1e6 / peak}
    c4["pairs_per_s"] = c4["shared_b"]["pairs_per_s"]
    c4["lags_exact"] = all(v["lags_exact"] for v in c4.values() if isinstance(v, dict))
    out["config4_correlate_peak"] = c4
    del A, B, o
    # deconvolution: 16 problems x 2^20 samples, 4096-tap kernel, regularized spectral division, device resident
    probs, n, m = 16, 1 << 20, 4096
    sig = gen_white(probs, n, 1)
    ker = torch.tensor(G.decaying_ir(m) + (np.arange(m) == 0) * 2.0, device="cuda")
    o = torch.empty((probs, n - m + 1), device="cuda", dtype=torch.float64)
    ms = timeit(lambda: lib.adsp_deconvolve_batch_device(ctx.handle, sig.data_ptr(), n, n, ker.data_ptr(), m, 0, probs, C.c_double(1e-6),
                                                         o.data_ptr(), n - m + 1), iters=3)
    del sig, o
    # the optional fp32 mode on the headline workload (north star: fp32 within 1e-5): same channels, same IR, float inputs
    ch32 = CHANNELS_PER_GPU
    x32 = torch.empty((ch32, N_SAMPLES), device="cuda", dtype=torch.float32)
    G.white_device(ctx, x32.data_ptr(), N_SAMPLES, ch32, N_SAMPLES, amp=1.0, seed0=1, seed_step=1, prec=1)
    ctx.sync()
    os32 = (OUT_LEN + 31) // 32 * 32
    y32 = torch.empty((ch32, os32), device="cuda", dtype=torch.float32)
    h64 = G.decaying_ir(K_TAPS)
    p32 = conv.OverlapSave(h64, 0, ctx=ctx, dtype=np.float32)
    ms32 = timeit(lambda: p32.process_device(x32.data_ptr(), N_SAMPLES, ch32, N_SAMPLES, y32.data_ptr(), os32), iters=10)
    sps32 = ch32 * OUT_LEN / ms32 * 1e3
    out["headline_fp32_mode"] = {"channels": ch32, "samples_per_s": sps32, "ms": ms32, "hbm_frac": sps32 * 8 / 1e9 / peak,
                                 "algorithmic_bytes_per_sample": 8}
    p32.Close()
    del x32, y32
    # the step after the path, device resident: Schroeder integral + onset of 64 IRs of 2^21 samples, a 257-tap
    # block FIR and a 160/147 polyphase resampler over 64 channels x 2^20 samples
    rows_p, n_p = 64, 1 << 21
    irs = gen_white(rows_p, n_p, 500)
    sch = torch.empty((rows_p, n_p), device="cuda", dtype=torch.float64)
    idx = torch.empty(rows_p, device="cuda", dtype=torch.int64)
    ms_s = timeit(lambda: lib.adsp_ir_schroeder_device(ctx.handle, irs.data_ptr(), n_p, rows_p, n_p, sch.data_ptr(), n_p))
    ms_o = timeit(lambda: lib.adsp_ir_find_impulse_start_device(ctx.handle, irs.data_ptr(), n_p, rows_p, n_p, C.c_double(0.1), idx.data_ptr()))
    post = {"schroeder_64x2e21": {"ms": ms_s, "GBps": rows_p * n_p * 16 / ms_s / 1e6, "hbm_frac": rows_p * n_p * 16 / ms_s / 1e6 / peak},
            "impulse_start_64x2e21": {"ms": ms_o, "GBps": rows_p * n_p * 8 / ms_o / 1e6}}
    del sch
    n_f = 1 << 20
    blk = irs[:, :n_f].contiguous()
    fh = C.c_void_p()
    taps = np.hanning(257)
    taps /= taps.sum()
    assert lib.adsp_fir_create(ctx.handle, taps.ctypes.data_as(C.c_void_p), 257, rows_p, C.byref(fh)) == 0
    ms_f = timeit(lambda: lib.adsp_fir_process_block_device(fh, blk.data_ptr(), n_f, n_f))
    post["fir_257tap_64x2e20"] = {"ms": ms_f, "samples_per_s": rows_p * n_f / ms_f * 1e3}
    lib.adsp_fir_destroy(fh)
    rh = C.c_void_p()
    assert lib.adsp_resampler_create(ctx.handle, 160, 147, 1, 0, C.c_double(0), C.c_double(0), rows_p, C.byref(rh)) == 0
    n_r = int(lib.adsp_resampler_predict_output_len(rh, n_f))
    ro = torch.empty((rows_p, n_r + 64), device="cuda", dtype=torch.float64)
    got = C.c_int64()

    def rs():
        lib.adsp_resampler_reset(rh)
        assert lib.adsp_resampler_process_device(rh, blk.data_ptr(), n_f, n_f, ro.data_ptr(), n_r + 64, n_r + 64, C.byref(got)) == 0
    ms_r = timeit(rs)
    post["resample_160_147_64x2e20"] = {"ms": ms_r, "output_samples_per_s": rows_p * n_r / ms_r * 1e3, "taps_per_phase": int(lib.adsp_resampler_taps_per_phase(rh))}
    lib.adsp_resampler_destroy(rh)
    out["post_path"] = post
    del irs, blk, ro
    out["deconvolve_regularized"] = {"problems": probs, "samples": n, "kernel_taps": m, "problems_per_s": probs / ms * 1e3,
                                     "algorithmic_GBps": probs * (2 * n - m + 1 + m) * 8 / ms / 1e6, "ms": ms}
    return out


def run_ours(args):
    import ctypes as C
    import torch
    rank, local_rank, world = dist_env()
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist = None
    if world > 1:
        # N ranks share the host's cores: the library's copy threads (default 3/4 of the cores, for one process) are divided
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
        os.environ.setdefault("ADSP_STAGE_THREADS", str(max(2, (3 * (os.cpu_count() or 8)) // (4 * max(1, local_world)))))
    from algo_dsp_b200 import _lib as L, conv, siggen as G
    lib = L.load()

    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)
    ctx = conv.Context(dev)
    channels = args.channels
    h = G.decaying_ir(K_TAPS)
    plan = conv.OverlapSave(h, 0, ctx=ctx)
    geom = plan.internal_geometry()
    geom["cover"] = plan.describe_cover(N_SAMPLES)   # the transforms one Process() actually runs

    # synthetic white noise (u*2-1), dsp/signal WhiteNoise on the library's hash PRNG, seed = 1 + global channel number,
    # generated straight into HBM by the device generator (adsp_gen_white_device)
    x = torch.empty((channels, N_SAMPLES), device="cuda", dtype=torch.float64)
    G.white_device(ctx, x.data_ptr(), N_SAMPLES, channels, N_SAMPLES, amp=1.0, seed0=1 + rank * channels, seed_step=1)
    ctx.sync()
    ostride = (OUT_LEN + 31) // 32 * 32
    y = torch.empty((channels, ostride), device="cuda", dtype=torch.float64)
    stream = torch.cuda.ExternalStream(ctx.stream())

    def step():
        plan.process_device(x.data_ptr(), N_SAMPLES, channels, N_SAMPLES, y.data_ptr(), ostride)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    plan.sync()

    sampler = ClockSampler(dev) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)

    # ---- timed region: K steps, CUDA events on the library's stream, inputs (0.98 GB) >> L2
    barrier()
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    plan.sync()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    clocks_timed = sampler.summary() if sampler else None

    # ---- sustained leg: the same step back to back for >= sustain_s seconds (clock / power sampler running)
    sustained = None
    if args.sustain_s > 0:
        est = max(ms_total / args.steps, 1e-3)
        chunk = max(int(100.0 / est), 1)               # ~0.1 s of steps between host checks
        mark = sampler.mark() if sampler else (0, 0)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        nsteps, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < args.sustain_s:
            for _ in range(chunk):
                step()
            nsteps += chunk
            plan.sync()
        s1.record(stream)
        plan.sync()
        barrier()
        sus_ms = s0.elapsed_time(s1)
        sustained = {"steps": nsteps, "seconds": sus_ms * 1e-3, "ms_per_step": sus_ms / nsteps}
        if sampler:
            cs = sampler.summary(mark)
            sustained.update({"sm_mhz_median": cs["sm_mhz"], "power_w_max": cs["power_w_max"], "reasons": cs["reasons"]})

    # ---- per-kernel device time of the dominant kernel (same K steps, event pair per launch)
    ctx.kernel_timing(True)
    ctx.kernel_times(reset=True)
    for _ in range(args.steps):
        step()
    plan.sync()
    ktimes = ctx.kernel_times(reset=True)
    # ---- the same kernels timed EXCLUSIVELY: one launch per kernel over the whole batch on one stream, so no
    # launch shares the GPU with another (in the timed schedule four groups are in flight and their per-launch
    # event durations overlap).  Scratch for all pairs (1.2 GB) then lives in HBM, which costs rows ~5 %.
    os.environ["ADSP_STREAMS"] = "1"
    os.environ["ADSP_GROUP_PAIRS"] = "1000000"
    for _ in range(2):
        step()
    plan.sync()
    ctx.kernel_times(reset=True)
    for _ in range(args.steps):
        step()
    plan.sync()
    ktimes_excl = ctx.kernel_times(reset=True)
    os.environ.pop("ADSP_STREAMS")
    os.environ.pop("ADSP_GROUP_PAIRS")
    ctx.kernel_timing(False)

    # ---- pipe peaks of this GPU at this clock (FP64 FMA issue, shared-memory bandwidth)
    dfma, smem = C.c_double(), C.c_double()
    lib.adsp_ctx_measure_pipes(ctx.handle, C.byref(dfma), C.byref(smem))

    # ---- end-to-end through the host-buffer C-ABI call: (a) pinned caller buffers, DMA in place; (b) pageable numpy
    # buffers (a Go []float64), staged through the library's pinned slots by its copy threads
    e2e_channels = channels
    h2d_bytes, d2h_bytes = e2e_channels * N_SAMPLES * 8, e2e_channels * OUT_LEN * 8
    xh = conv.pinned_empty((e2e_channels, N_SAMPLES))
    yh = conv.pinned_empty((e2e_channels, OUT_LEN))
    xh[:] = x[:e2e_channels].cpu().numpy()

    def e2e_leg(xin, yout):
        for _ in range(2):
            plan.ProcessBatch(xin, out=yout)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            plan.ProcessBatch(xin, out=yout)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / args.steps

    e2e_s = e2e_leg(xh, yh)
    xp = np.array(xh)                      # ordinary pageable memory
    yp = np.empty((e2e_channels, OUT_LEN))
    e2e_page_s = e2e_leg(xp, yp)
    page_ok = bool(np.array_equal(yp[0], yh[0]) and np.array_equal(yp[-1], yh[-1]))
    # raw ceiling of the host link for exactly these byte counts (all ranks at the same time)
    barrier()
    c_both, c_in, c_out = C.c_double(), C.c_double(), C.c_double()
    lib.adsp_ctx_copy_ceiling(ctx.handle, h2d_bytes, d2h_bytes, 3, C.byref(c_both), C.byref(c_in), C.byref(c_out))
    barrier()
    # ---- optional collective (reported separately): every rank receives every rank's device-resident output
    # rows through one NCCL all-gather over NVLink / NVSwitch; the data path itself has no collective
    allgather = None
    if dist is not None:
        from algo_dsp_b200 import shard
        ag_rows = min(channels, args.allgather_channels)
        piece = y[:ag_rows]
        sizes = [ag_rows] * world
        for _ in range(2):
            shard.gather_outputs_device(piece, sizes, dist)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(3):
            views = shard.gather_outputs_device(piece, sizes, dist)
        g1.record()
        torch.cuda.synchronize()
        ag_ms = g0.elapsed_time(g1) / 3
        ag_ok = bool(torch.equal(views[rank], piece))
        tag = torch.tensor([ag_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tag, op=dist.ReduceOp.MAX)
        bytes_rank = ag_rows * ostride * 8
        allgather = {"bytes_per_rank": bytes_rank, "ms": float(tag[0]), "own_piece_intact": ag_ok,
                     "bus_GBps": bytes_rank * (world - 1) / (float(tag[0]) * 1e-3) / 1e9,
                     "what": f"torch.distributed all_gather_into_tensor (NCCL) of {ag_rows} output channels per rank, device resident, max over ranks; "
                             "not part of `value` (shards are independent)"}
        del views
    if sampler:
        sampler.stop()

    # ---- the other BASELINE.json configs, reduced batches, device resident (reported as `other_configs`)
    other = {}
    if rank == 0 and not args.skip_other:
        try:
            other = other_configs(torch, conv, G, ctx, stream)
        except Exception as e:   # never lose the headline line to an auxiliary measurement
            other = {"error": repr(e)}

    # spot parity check of the timed output against the host-path output (same inputs)
    chk = float(np.max(np.abs(y[0, :OUT_LEN].cpu().numpy() - yh[0])))

    ms_step = ms_total / args.steps
    vals = [ms_step, e2e_s * 1e3, e2e_page_s * 1e3, c_both.value, c_in.value, c_out.value, sustained["ms_per_step"] if sustained else 0.0]
    t = torch.tensor(vals, device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step, e2e_ms, e2e_page_ms, ceil_ms, ceil_in_ms, ceil_out_ms, sus_ms_step = (float(v) for v in t)
    total_samples = world * channels * OUT_LEN
    value = total_samples / (ms_step * 1e-3)
    e2e_value = world * e2e_channels * OUT_LEN / (e2e_ms * 1e-3)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        # dominant kernel = the one with the largest exclusive device time (agrees with the ncu launch list)
        kinds = ("rows", "cols_fwd", "cols_inv", "full", "fused")
        dom = max(kinds, key=lambda k: ktimes_excl[k][0])
        step_bytes = channels * OUT_LEN * ALGO_BYTES_PER_SAMPLE
        dom_ms, dom_n = ktimes[dom]                      # in the timed schedule (launches of 4 groups overlap)
        samples_per_launch = channels * OUT_LEN * args.steps / max(dom_n, 1)   # output samples attributable to one launch
        dom_avg_ms = dom_ms / max(dom_n, 1)
        achieved = samples_per_launch * ALGO_BYTES_PER_SAMPLE / (dom_avg_ms * 1e-3) / 1e9
        ex_ms, ex_n = ktimes_excl[dom]
        ex_avg_ms = ex_ms / max(ex_n, 1)
        ex_bytes = step_bytes * args.steps / max(ex_n, 1)
        ex_achieved = ex_bytes / (ex_avg_ms * 1e-3) / 1e9
        ex_total = sum(v[0] for v in ktimes_excl.values())
        path_achieved = step_bytes / (ms_step * 1e-3) / 1e9
        kshare = {k: round(v[0], 3) for k, v in ktimes.items() if v[1]}
        # DRAM traffic and instruction counts per step come from committed ncu captures of this same command (profiles/),
        # not from this run: labelled as such
        prof = {}
        for name in ("r02_traffic.json", "r01_traffic.json"):
            try:
                prof = json.load(open(os.path.join(ROOT, "profiles", name)))
                prof["file"] = "profiles/" + name
                break
            except Exception:
                pass
        traffic = (prof.get("fftconv_" + dom) or {}).get("dram_bytes_per_launch")
        per_step_traffic = prof.get("per_step_total_bytes")
        traffic_note = ("static: from the committed ncu capture " + prof.get("file", "") + " (" + str(prof.get("source", "")) + "), not measured in this run") if prof else None
        # The unit of work is a GROUP: one block pair through all three kernels (every output sample passes all of them, and
        # SURVEY 8d's 16 B per output sample -- one input read, one output written -- belongs to the whole path: the row kernel
        # itself touches only L2-resident scratch).  So the roofline line is the pipeline's: algorithmic bytes per step / step
        # time; the dominant member kernel is reported underneath, both exclusively timed and as launched in the schedule.
        groups_per_step = max(ktimes["rows"][1] // max(args.steps, 1), 1)
        # FP64 / LSU pipe rooflines: work per output sample (~150 FP64 instructions and ~3.5 shared-memory/LSU
        # wavefronts of 128 B per complex point; a pair of channels shares N complex points per block) over the peaks
        # measured a moment ago on this GPU
        cover = geom.get("cover") or []
        pts_per_sample = (cover[0]["fft_n"] / 2.0) / OUT_LEN if len(cover) == 1 else None   # one block per channel, two channels per transform
        dp_per_pt = float(prof.get("fp64_instr_per_complex_point", 150.0))
        wf_per_pt = float(prof.get("lsu_wavefronts_per_complex_point", 3.5))
        fp64_roof = lsu_roof = None
        if pts_per_sample and len(geom.get("cover", [])) == 1:
            ips = dp_per_pt * pts_per_sample
            fp64_roof = {"peak_dfma_per_s": dfma.value, "instr_per_sample": ips, "achieved_instr_per_s": ips * value / world,
                         "frac": ips * value / world / dfma.value if dfma.value else None,
                         "samples_per_s_at_peak": dfma.value / ips if ips else None,
                         "source": "peak: adsp_ctx_measure_pipes in this run; instructions per point: " + ("ncu, " + prof["file"] if "fp64_instr_per_complex_point" in prof else "instruction count of the kernels (DESIGN.md 3)")}
            bps = wf_per_pt * 128.0 * pts_per_sample
            lsu_roof = {"peak_smem_bytes_per_s": smem.value, "wavefront_bytes_per_sample": bps, "achieved_bytes_per_s": bps * value / world,
                        "frac": bps * value / world / smem.value if smem.value else None,
                        "samples_per_s_at_peak": smem.value / bps if bps else None,
                        "source": "peak: adsp_ctx_measure_pipes in this run (128-bit shared-memory stores + loads); wavefronts per point: " + ("ncu, " + prof["file"] if "lsu_wavefronts_per_complex_point" in prof else "DESIGN.md 3")}
        roof = {
            "bound": "hbm",
            "kernel": "fftconv four-step pipeline: cols_fwd -> rows -> cols_inv (one launch of each per group of block pairs); dominant member fftconv_" + dom,
            "achieved": path_achieved, "peak": peak, "unit": "GB/s", "frac": path_achieved / peak,
            "traffic": (per_step_traffic / groups_per_step) if per_step_traffic else None,
            "traffic_note": traffic_note, "peak_source": peak_src,
            "avg_launch_ms": ms_step / groups_per_step, "launches": int(groups_per_step * args.steps),
            "algorithmic_bytes_per_launch": step_bytes / groups_per_step,
            "note": "launch = one group (three kernels); groups of four streams overlap, so avg_launch_ms is step time / groups per step",
            "fp64": fp64_roof, "lsu": lsu_roof,
            "dominant_kernel": {
                "name": "fftconv_" + dom,
                "exclusive": {"avg_launch_ms": ex_avg_ms, "launches": ex_n, "algorithmic_bytes_per_launch": ex_bytes,
                              "achieved": ex_achieved, "frac": ex_achieved / peak,
                              "share_of_kernel_time": ex_ms / ex_total if ex_total else None,
                              "kernel_ms_per_step": {k: round(v[0] / args.steps, 4) for k, v in ktimes_excl.items() if v[1]},
                              "schedule": "ADSP_STREAMS=1, one launch per kernel over all block pairs (scratch in HBM): the kernel alone on the GPU"},
                "in_schedule": {"avg_launch_ms": dom_avg_ms, "launches": dom_n, "algorithmic_bytes_per_launch": samples_per_launch * ALGO_BYTES_PER_SAMPLE,
                                "achieved": achieved, "frac": achieved / peak, "dram_bytes_per_launch": traffic,
                                "kernel_device_ms_over_timed_steps": kshare,
                                "note": "event-bracketed launches of four concurrent groups share the GPU, so these durations overlap"},
            },
            "path_achieved": path_achieved, "path_frac": path_achieved / peak,
            "co_bound": "shared-memory (LSU) pipe and fp64 pipe: see roofline.fp64 / roofline.lsu (peaks measured in this run)",
        }
        if sustained:
            sustained["value"] = world * channels * OUT_LEN / (sus_ms_step * 1e-3)
            sustained["ms_per_step"] = sus_ms_step
            sustained["frac_of_hbm_roofline"] = sustained["value"] / world * ALGO_BYTES_PER_SAMPLE / 1e9 / peak
        cpu_threads = os.cpu_count() or 1
        cpu_ch = max(4 * cpu_threads, 8)
        cpu_v = cpu_1 = None
        oracle_err = None
        if world == 1:
            cpu_v, cpu_t, cpu_passes = cpu_baseline(cpu_ch, cpu_threads)
            cpu_1, cpu_1t, cpu_1p = cpu_baseline(2, 1, budget_s=6.0)
            # the timed output against the CPU restatement of the reference (checker only): channel 0 of this rank
            from oracle import oracle as O
            ref0 = O.overlap_save(h, 0, xh[0])
            oracle_err = float(np.linalg.norm(y[0, :OUT_LEN].cpu().numpy() - ref0) / np.linalg.norm(ref0))
        ceil_value = world * e2e_channels * OUT_LEN / (ceil_ms * 1e-3) if ceil_ms > 0 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "ols96k_batch", "kernel_taps": K_TAPS, "signal_samples": N_SAMPLES, "channels_per_gpu": channels,
                       "output_samples_per_channel": OUT_LEN, "sharding": f"channel x{world}, no collective",
                       "internal_fft": geom, "l2_policy": "inputs (0.98 GB/GPU/step) exceed L2; no flush needed",
                       "inputs": "dsp/signal WhiteNoise on the hash PRNG, generated on the device (adsp_gen_white_device), seed 1 + channel",
                       "reference_getters": {"FFTSize": plan.FFTSize(), "StepSize": plan.StepSize()}},
            "roofline": roof,
            "sustained": sustained,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_ms,
                    "api": "adsp_plan_process_batch (pinned host buffers, chunked H2D|compute|D2H pipeline)",
                    "pageable": {"value": world * e2e_channels * OUT_LEN / (e2e_page_ms * 1e-3), "ms_per_step": e2e_page_ms,
                                 "frac_of_pinned": e2e_ms / e2e_page_ms, "outputs_equal_pinned_leg": page_ok, "stage_threads": ctx.stage_threads(),
                                 "api": "same call on pageable numpy buffers (a Go []float64): stage-in | H2D | kernels | D2H | stage-out, "
                                        "pinned slots and copy threads inside the library"},
                    "copy_ceiling": {"ms_per_step": ceil_ms, "value": ceil_value, "h2d_alone_ms": ceil_in_ms, "d2h_alone_ms": ceil_out_ms,
                                     "h2d_GBps_alone": h2d_bytes / ceil_in_ms / 1e6 if ceil_in_ms else None,
                                     "d2h_GBps_alone": d2h_bytes / ceil_out_ms / 1e6 if ceil_out_ms else None,
                                     "e2e_frac_of_ceiling": ceil_ms / e2e_ms if e2e_ms else None,
                                     "what": "one cudaMemcpyAsync H2D + one D2H of the step's byte counts from/to pinned memory, concurrently, "
                                             "every rank at the same time, max over ranks"}},
            "allgather": allgather,
            "gpu_launches": launches,
            "clocks": clocks_timed,
            "other_configs": other,
            "check_max_abs_diff_device_vs_host_path": chk,
            "check_rel_l2_vs_cpu_oracle_channel0": oracle_err,
        }
        if cpu_v is not None:
            line["cpu_baseline"] = {"value": cpu_v, "unit": UNIT, "cores": cpu_threads, "kind": "port",
                                    "sample": f"{cpu_passes} passes over {cpu_ch} channels x {N_SAMPLES} samples x {K_TAPS} taps "
                                              f"({cpu_t:.1f} s wall in total, mean rate)",
                                    "single_thread": {"value": cpu_1, "cores": 1,
                                                      "sample": f"{cpu_1p} passes over 2 channels ({cpu_1t:.1f} s wall): the reference itself never spawns goroutines"},
                                    "note": "C restatement of the Go reference OverlapSave.Process (N=262144 complex FFT per block); "
                                            "the Go toolchain and algo-fft are absent here"}
        print(json.dumps(line), flush=True)
    plan.Close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--channels", type=int, default=CHANNELS_PER_GPU)
    ap.add_argument("--skip-other", action="store_true", help="skip the auxiliary measurements of BASELINE configs 1-4")
    ap.add_argument("--allgather-channels", type=int, default=64, help="N > 1: output channels per rank in the optional all-gather leg")
    ap.add_argument("--sustain-s", type=float, default=2.5, help="seconds of back-to-back steps for the `sustained` leg (0: skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
