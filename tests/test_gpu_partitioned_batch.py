"""Multi-channel streaming partitioned convolution on the device-resident frequency-domain delay line
(csrc/fdl.cu) against the CPU restatement of PartitionedConvolutionT (partitioned.go:135-190, 348-396) and of
ConvolutionReverb.ProcessInPlace (dsp/effects/reverb/convolution.go:60-83)."""
import numpy as np
import pytest

from algo_dsp_b200 import siggen as G

pytestmark = pytest.mark.gpu
TOL64, TOL32 = 1e-12, 1e-5


def oracle_stream(oracle, h, mn, mx, x, cuts):
    p = oracle.Partitioned(h, mn, mx)
    out = [p.process_block(x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    return np.concatenate(out)


def cuts_for(n, rng):
    """Irregular block lengths: 1 sample up to several partitions per call."""
    c, pos = [0], 0
    sizes = [1, 3, 17, 64, 127, 128, 129, 500, 1000, 2048, 4097, 9000, 20000]
    while pos < n:
        pos = min(n, pos + int(rng.choice(sizes)))
        c.append(pos)
    return c


@pytest.mark.parametrize("K,mn,mx,channels", [(5000, 6, 10, 5), (300, 3, 5, 2), (40000, 7, 13, 4), (100, 7, 9, 3), (2500, 11, 13, 1),
                                               (1, 4, 6, 2), (9000, 12, 13, 2)])
def test_batch_stream_matches_oracle(conv, oracle, K, mn, mx, channels):
    rng = np.random.default_rng(K + mn)
    n = 60000
    h = G.decaying_ir(K, seed=K)
    x = np.stack([G.white(n, seed=100 + c) for c in range(channels)])
    cuts = cuts_for(n, rng)
    p = conv.NewPartitionedConvolution(h, mn, mx, channels=channels)
    assert p.Latency() == 1 << mn
    st = p.internal_stages()
    assert st[0][0] <= p.Latency() and sum(b * c for b, c, _ in st) >= K
    got = np.concatenate([p.ProcessBlockBatch(x[:, a:b]) for a, b in zip(cuts[:-1], cuts[1:])], axis=1)
    for c in range(channels):
        ref = oracle_stream(oracle, h, mn, mx, x[c], cuts)
        assert G.rel_l2(got[c], ref) <= TOL64
    # output = full linear convolution delayed by the latency (partitioned.go:343-347)
    full = oracle.overlap_save(h, 0, x[0])
    L = p.Latency()
    assert G.rel_l2(got[0][L:], full[: n - L]) <= TOL64 and np.all(got[0][:L] == 0)
    # Reset() restarts the stream (partitioned.go:399-407)
    p.Reset()
    again = p.ProcessBlockBatch(x[:, :5000])
    assert np.array_equal(again, got[:, :5000]) or G.rel_l2(again, got[:, :5000]) <= 1e-15


def test_mono_entry_point_uses_the_same_engine(conv, oracle):
    h, x = G.decaying_ir(3000), G.white(30000, seed=3)
    p = conv.NewPartitionedConvolution(h, 7, 13)
    assert len(p.internal_stages()) >= 1                       # delay-line engine active
    out = np.zeros_like(x)
    p.ProcessBlock(x, out)
    ref = oracle.Partitioned(h, 7, 13).process_block(x)
    assert G.rel_l2(out, ref) <= TOL64
    o = oracle.Partitioned(h, 7, 13)
    assert (p.StageCount(), p.StageInfo(0)) == (o.stage_count(), o.stage_info(0)[:2])


def test_convolution_reverb_wet_dry_in_place(conv, oracle):
    """ProcessInPlace: block = dry*block + wet*reverb(block), block length varies between calls (convolution.go:57-83)."""
    K, channels, n = 20000, 6, 30000
    h = G.decaying_ir(K, seed=2)
    x = np.stack([G.pink(n, seed=100 + c) for c in range(channels)])
    r = conv.NewConvolutionReverb(h, 7, channels=channels)
    r.SetWetDry(0.3, 0.8)
    blk = x.copy()
    pos = 0
    for size in (128, 1000, 77, 8192, 20603):
        part = np.ascontiguousarray(blk[:, pos:pos + size])
        r.ProcessInPlace(part)
        blk[:, pos:pos + size] = part
        pos += size
    assert pos == n
    for c in (0, channels - 1):
        wet = oracle.Partitioned(h, 7, 13).process_block(x[c])
        assert G.rel_l2(blk[c], 0.8 * x[c] + 0.3 * wet) <= TOL64


def test_float32_and_device_pointers(conv, oracle):
    import torch
    K, channels, n = 7000, 3, 20000
    h = G.decaying_ir(K, seed=4)
    x = np.stack([G.white(n, seed=c) for c in range(channels)])
    p32 = conv.NewPartitionedConvolution32(h.astype(np.float32), 6, 13, channels=channels)
    y32 = p32.ProcessBlockBatch(x.astype(np.float32))
    ref = np.stack([oracle.Partitioned(h, 6, 13).process_block(x[c]) for c in range(channels)])
    assert y32.dtype == np.float32 and G.rel_l2(y32.astype(np.float64), ref) <= TOL32
    p = conv.NewPartitionedConvolution(h, 6, 13, channels=channels)
    xd = torch.tensor(x, device="cuda")
    yd = torch.empty_like(xd)
    for a, b in ((0, 4096), (4096, 4100), (4100, n)):
        p.process_block_device(xd.data_ptr() + a * 8, b - a, n, yd.data_ptr() + a * 8, n)
    p.sync()
    assert G.rel_l2(yd.cpu().numpy(), ref) <= TOL64


def test_errors(conv):
    with pytest.raises(Exception):
        conv.NewPartitionedConvolution([1.0, 2.0], 2, 5, channels=4)     # multi-channel needs minBlockOrder >= 3
    p = conv.NewPartitionedConvolution([1.0, 0.5], 5, 8, channels=2)
    with pytest.raises(Exception):
        p.ProcessBlock(np.zeros(8), np.zeros(8))                         # mono entry point on a 2-channel plan
