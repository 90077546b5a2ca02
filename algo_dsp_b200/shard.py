"""Multi-GPU decomposition of the dsp/conv hot path: one process per GPU, shards are
independent, no collective on the data path.

* channel / pair sharding (configs 2, 3, 4): contiguous channel ranges per rank; the IR is
  replicated (each rank builds its own plan).
* time-block sharding with halo (config 5, any single long signal): rank g owns outputs
  [g*S, (g+1)*S) and reads inputs [g*S - (K-1), (g+1)*S) -- overlap-save at shard granularity,
  the same rule as overlap_save.go:205-215; the last rank also emits the K-1 tail (:224-251).
  Halos come from the source, never from a neighbour.

`gather_outputs` (optional) is the only collective: an all-gather for callers that want the whole
output on every rank (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


def channel_range(channels: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced channel range of `rank` (first `channels % world` ranks get one extra)."""
    base, rem = divmod(channels, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


@dataclass(frozen=True)
class TimeShard:
    rank: int
    out_lo: int      # first output sample owned
    out_hi: int      # one past the last output sample owned (includes the K-1 tail on the last rank)
    in_lo: int       # first input sample read (>= 0; the halo is clipped at the signal start)
    in_hi: int       # one past the last input sample read
    skip: int        # outputs of the local convolution to drop (those produced by the halo)


def time_shards(n: int, kernel_len: int, world: int, align: int = 32) -> list[TimeShard]:
    """Cut a length-n signal into `world` time shards with a (kernel_len - 1) halo."""
    out_len = n + kernel_len - 1
    S = -(-n // world)
    S = -(-S // align) * align
    shards = []
    for g in range(world):
        lo = min(g * S, n)
        hi = min((g + 1) * S, n)
        last = g == world - 1 or hi >= n
        out_hi = out_len if last else hi
        in_lo = max(0, lo - (kernel_len - 1))
        shards.append(TimeShard(g, lo, out_hi if lo < n or last else lo, in_lo, hi, lo - in_lo))
        if last:
            for r in range(g + 1, world):   # signal shorter than world*S: trailing ranks own nothing
                shards.append(TimeShard(r, out_len, out_len, n, n, 0))
            break
    return shards


def process_time_shard(convolve, x_segment: np.ndarray, shard: TimeShard) -> np.ndarray:
    """Run `convolve(segment) -> full linear convolution` on this rank's input segment
    (x[in_lo:in_hi]) and keep the outputs the shard owns."""
    if shard.out_hi <= shard.out_lo:
        return np.zeros(0, dtype=x_segment.dtype)
    y = convolve(x_segment)
    return y[shard.skip: shard.skip + (shard.out_hi - shard.out_lo)]


def gather_outputs(local: np.ndarray, shards: list[TimeShard], dist=None) -> np.ndarray:
    """All-gather the per-rank output pieces into the full output (optional collective)."""
    import torch
    if dist is None:
        import torch.distributed as dist  # noqa: PLW0642
    world = dist.get_world_size()
    sizes = [s.out_hi - s.out_lo for s in shards]
    mx = max(sizes)
    pad = torch.zeros(mx, dtype=torch.float64)
    pad[: local.size] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64))
    if dist.get_backend() == "nccl":
        pad = pad.cuda()
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return np.concatenate([b[:sz].cpu().numpy() for b, sz in zip(bufs, sizes)])


def gather_outputs_device(local, sizes, dist=None):
    """Device-resident all-gather ("reported separately"): `local` is this rank's output piece as a torch tensor
    living where the backend wants it (CUDA for NCCL over NVLink / NVSwitch, CPU for gloo); `sizes[r]` the number of valid
    leading elements along dim 0 of rank r's piece.  Every rank receives all pieces in one collective on equally padded
    slots and returns the list of per-rank views (no host round trip, no concatenation copy)."""
    import torch
    if dist is None:
        import torch.distributed as dist  # noqa: PLW0642
    world = dist.get_world_size()
    mx = max(sizes)
    slot = local
    if local.shape[0] != mx:
        slot = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        slot[: local.shape[0]] = local
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    try:
        dist.all_gather_into_tensor(out, slot.contiguous())
    except (RuntimeError, NotImplementedError):          # backends without the single-tensor form
        bufs = [out[r * mx:(r + 1) * mx] for r in range(world)]
        dist.all_gather(bufs, slot.contiguous())
    return [out[r * mx: r * mx + sizes[r]] for r in range(world)]
