"""BASELINE config 3 on N GPUs: 64 channels x 14.4 M samples (5 min @48 kHz) through a 288 000-tap IR, channels split
64/N per GPU (strong scaling), one process per GPU, no collective.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/tools/config3_multi.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import torch.distributed as dist
from algo_dsp_b200 import conv, siggen as G
from algo_dsp_b200.shard import channel_range
from oracle import oracle as O
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
CH, n, K = 64, 14_400_000, 288_000
lo, hi = channel_range(CH, rank, world)
ch = hi - lo
ctx = conv.Context(local)
h = G.decaying_ir(K)
plan = conv.OverlapSave(h, 0, ctx=ctx)
gen = torch.Generator(device="cuda"); gen.manual_seed(100 + rank)
x = torch.rand((ch, n), device="cuda", dtype=torch.float64, generator=gen) * 2 - 1
ol = n + K - 1; ostr = (ol + 31) // 32 * 32
y = torch.empty((ch, ostr), device="cuda", dtype=torch.float64)
stream = torch.cuda.ExternalStream(ctx.stream())
step = lambda: plan.process_device(x.data_ptr(), n, ch, n, y.data_ptr(), ostr)
def barrier():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
for _ in range(2): step()
plan.sync(); barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(5): step()
e1.record(stream); plan.sync(); barrier()
ms = e0.elapsed_time(e1) / 5
W = 1_000_000
err = float(G.rel_l2(y[ch - 1, :W].cpu().numpy(), O.overlap_save(h, 0, x[ch - 1, :W].cpu().numpy())[:W]))
t = torch.tensor([ms, err], device="cuda", dtype=torch.float64)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"config": "64 ch x 14.4M samples, 288k-tap IR, channel sharded", "n_gpus": world, "channels_per_gpu": ch,
                      "ms_per_pass_max_over_ranks": float(t[0]), "output_samples_per_s": CH * ol / (float(t[0]) * 1e-3),
                      "worst_rel_l2_first_1M_vs_oracle": float(t[1]), "internal_fft": plan.internal_geometry()}), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
