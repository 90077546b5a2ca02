"""Host-side multi-GPU logic on CPU: shard planning, and a world_size-2 gloo run of the time-block
(halo) decomposition and the channel decomposition with the oracle standing in for the device."""
import os
import socket

import numpy as np
import pytest

from algo_dsp_b200 import shard, siggen as G


def test_channel_ranges_cover_and_balance():
    for channels, world in [(64, 1), (64, 8), (10, 4), (3, 8), (1024, 8)]:
        rs = [shard.channel_range(channels, r, world) for r in range(world)]
        assert rs[0][0] == 0 and rs[-1][1] == channels
        assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        sizes = [hi - lo for lo, hi in rs]
        assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("n,K,world", [(400000, 5000, 4), (1000, 300, 8), (100, 7, 8), (1 << 20, 1 << 14, 2), (33, 64, 3)])
def test_time_shards_reassemble(n, K, world, oracle):
    x, h = G.white(n, seed=n), G.decaying_ir(K)
    full = oracle.overlap_save(h, 0, x)
    sh = shard.time_shards(n, K, world)
    assert len(sh) == world and sh[0].out_lo == 0 and sh[-1].out_hi == n + K - 1
    parts = [shard.process_time_shard(lambda seg: oracle.overlap_save(h, 0, seg), x[s.in_lo:s.in_hi], s) if s.in_hi > s.in_lo or s.out_hi > s.out_lo
             else np.zeros(0) for s in sh]
    got = np.concatenate(parts)
    assert got.size == full.size and G.rel_l2(got, full) <= 1e-13
    for s in sh:   # halo is exactly K-1 wherever the signal allows it
        if s.out_hi > s.out_lo:
            assert s.skip == min(s.out_lo, K - 1)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, K, channels, q):
    import torch.distributed as dist
    from oracle import oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x, h = G.white(n, seed=3), G.decaying_ir(K)
        sh = shard.time_shards(n, K, world)
        s = sh[rank]
        local = shard.process_time_shard(lambda seg: O.overlap_save(h, 0, seg), x[s.in_lo:s.in_hi], s)
        full = shard.gather_outputs(local, sh, dist)
        # the device-resident form of the same collective (CPU tensors under gloo, CUDA tensors under NCCL)
        import torch
        sizes = [t.out_hi - t.out_lo for t in sh]
        views = shard.gather_outputs_device(torch.from_numpy(np.ascontiguousarray(local)), sizes, dist)
        assert len(views) == world and all(v.shape[0] == sz for v, sz in zip(views, sizes))
        assert np.array_equal(np.concatenate([v.numpy() for v in views]), full)
        # channel sharding: every rank convolves its own channels; gather the per-channel checksums
        lo, hi = shard.channel_range(channels, rank, world)
        sums = torch.zeros(channels, dtype=torch.float64)
        for c in range(lo, hi):
            sums[c] = float(O.overlap_save(h, 0, G.white(2000, seed=100 + c)).sum())
        dist.all_reduce(sums)
        if rank == 0:
            q.put((full, sums.numpy()))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(oracle):
    import torch.multiprocessing as mp
    n, K, channels, world = 50000, 1200, 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, K, channels, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, sums = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x, h = G.white(n, seed=3), G.decaying_ir(K)
    assert G.rel_l2(full, oracle.overlap_save(h, 0, x)) <= 1e-13
    want = np.array([oracle.overlap_save(h, 0, G.white(2000, seed=100 + c)).sum() for c in range(channels)])
    assert np.allclose(sums, want, rtol=1e-12)


def test_c_abi_shard_helpers_match_the_python_planner():
    """adsp_shard_channel_range / adsp_shard_time (pure arithmetic in the C library, no GPU) == algo_dsp_b200/shard.py."""
    import ctypes as C
    from algo_dsp_b200 import _lib as L, shard
    lib = L.load()
    for channels, world in ((64, 8), (7, 3), (5, 8), (1024, 2), (0, 4)):
        for r in range(world):
            lo, hi = C.c_int64(), C.c_int64()
            lib.adsp_shard_channel_range(channels, r, world, C.byref(lo), C.byref(hi))
            assert (lo.value, hi.value) == shard.channel_range(channels, r, world)
    for n, K, world in ((400000, 5000, 4), (1000, 300, 8), (100, 7, 8), (1 << 20, 1 << 14, 2), (33, 64, 3), (1 << 31, 1 << 20, 8), (1, 1, 5)):
        ref = shard.time_shards(n, K, world)
        for r in range(world):
            v = [C.c_int64() for _ in range(5)]
            lib.adsp_shard_time(n, K, r, world, *[C.byref(x) for x in v])
            s = ref[r]
            assert tuple(x.value for x in v) == (s.out_lo, s.out_hi, s.in_lo, s.in_hi, s.skip), (n, K, world, r)
