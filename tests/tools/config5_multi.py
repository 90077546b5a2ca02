"""BASELINE config 5 on N GPUs: one 2^31-sample signal (x) 2^20-tap IR, time-block sharded with a (K-1) halo, one
process per GPU, no collective on the data path (barrier + max-over-ranks timing only).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/tools/config5_multi.py [log2_n=31] [log2_K=20] [f64|f32]

The signal is defined chunk-wise (white noise, torch generator seeded by the chunk index), so every rank regenerates
exactly the samples of its own shard and halo on its device; nothing is exchanged.  Rank 0 prints one JSON line;
each rank checks a window at the start of its shard (the halo seam) against the CPU oracle.
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import torch.distributed as dist
from algo_dsp_b200 import conv, siggen as G
from algo_dsp_b200.shard import time_shards
from oracle import oracle as O

lgn = int(sys.argv[1]) if len(sys.argv) > 1 else 31
lgk = int(sys.argv[2]) if len(sys.argv) > 2 else 20
prec = sys.argv[3] if len(sys.argv) > 3 else "f64"
n, K = 1 << lgn, 1 << lgk
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dtype, tdt = (np.float64, torch.float64) if prec == "f64" else (np.float32, torch.float32)
CH = 1 << 24


def signal(lo, hi):
    """x[lo:hi] on this device, identical on every rank"""
    out = torch.empty(hi - lo, device="cuda", dtype=tdt)
    for c in range(lo // CH, (hi - 1) // CH + 1):
        g = torch.Generator(device="cuda"); g.manual_seed(1000 + c)
        chunk = torch.rand(min(CH, n - c * CH), device="cuda", dtype=torch.float64, generator=g) * 2 - 1
        a, b = max(lo, c * CH), min(hi, (c + 1) * CH)
        out[a - lo:b - lo] = chunk[a - c * CH:b - c * CH].to(tdt)
    return out


ctx = conv.Context(local)
h = G.decaying_ir(K)
plan = conv.OverlapSave(h, 0, ctx=ctx, dtype=dtype)
s = time_shards(n, K, world)[rank]
x = signal(s.in_lo, s.in_hi)
seg = s.in_hi - s.in_lo
tmp = torch.empty(seg + K - 1 + 32, device="cuda", dtype=tdt)
stream = torch.cuda.ExternalStream(ctx.stream())


def step():
    plan.process_device(x.data_ptr(), seg, 1, seg, tmp.data_ptr(), seg + K - 1 + 32)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(2): step()
plan.sync()
barrier()
iters = 5
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(iters): step()
e1.record(stream)
plan.sync()
barrier()
ms = e0.elapsed_time(e1) / iters
# parity at the seam: the first 2^20 outputs this rank owns depend on the halo
W = min(1 << 20, s.out_hi - s.out_lo)
got = tmp[s.skip:s.skip + W].cpu().numpy().astype(np.float64)
ref = O.overlap_save(h, 0, x[:s.skip + W].cpu().numpy().astype(np.float64))[s.skip:s.skip + W]
err = float(G.rel_l2(got, ref))
t = torch.tensor([ms, err], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    out_len = n + K - 1
    print(json.dumps({"config": "single long signal, time-block sharded with (K-1) halo", "dtype": prec, "n": n, "K": K, "n_gpus": world,
                      "ms_per_pass_max_over_ranks": float(t[0]), "output_samples_per_s": out_len / (float(t[0]) * 1e-3),
                      "worst_seam_rel_l2_vs_oracle": float(t[1]), "halo_overhead": (K - 1) / (n / world),
                      "internal_fft": plan.internal_geometry(), "collective": "none (barrier + max over ranks for timing)"}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
