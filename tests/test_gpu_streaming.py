"""Fixed-block streaming convolvers on the GPU, mirroring the reference's
streaming_overlap_save_test.go / streaming_overlap_add_test.go / streaming_test.go."""
import numpy as np
import pytest

import kat_checks
from algo_dsp_b200 import siggen as G

pytestmark = pytest.mark.gpu


def ctors(conv):
    return [conv.NewStreamingOverlapSave, conv.NewStreamingOverlapAdd]


def test_impulse_kats(conv):
    """TestStreamingOverlapSave / TestStreamingOverlapAdd (:9-56) and the f32 impulse (streaming_test.go:237-268)."""
    k = kat_checks.KATS["streaming_impulse"]
    for new in ctors(conv):
        s = new(k["kernel"], k["block_size"])
        o1 = s.ProcessBlock(k["block1"])
        o2 = s.ProcessBlock(k["block2"])
        assert len(o1) == len(o2) == k["block_size"]
        assert np.max(np.abs(o1 - np.array(k["out1"]))) <= k["tol"]
        assert np.max(np.abs(o2)) <= k["tol"]
    k = kat_checks.KATS["streaming_impulse_f32"]
    for new in (conv.NewStreamingOverlapSave32, conv.NewStreamingOverlapAdd32):
        s = new(np.array(k["kernel"], np.float32), k["block_size"])
        x = np.zeros(k["block_size"], np.float32)
        x[0] = 1
        out = s.ProcessBlock(x)
        assert out.dtype == np.float32 and np.max(np.abs(out - np.array(k["out"], np.float32))) <= k["tol"]


def test_streaming_vs_batch_and_equivalence(conv, oracle):
    """TestStreamingOverlapSaveVsBatch (:51-99, 1e-10) and OLA == OLS (streaming_test.go:122-176, 1e-9)."""
    kernel = [0.5, 1.0, 0.5, 0.2]
    bs, nb = 8, 4
    x = np.sin(np.arange(bs * nb) * 0.1)
    batch = conv.NewOverlapSave(kernel, 0).Process(x)
    outs = []
    for new in ctors(conv):
        s = new(kernel, bs)
        y = np.concatenate([s.ProcessBlock(x[i:i + bs]) for i in range(0, len(x), bs)])
        assert np.max(np.abs(y - batch[: len(x)])) <= 1e-10
        ref = oracle.Streaming(kernel, bs, True)
        yr = np.concatenate([ref.process_block(x[i:i + bs]) for i in range(0, len(x), bs)])
        assert np.max(np.abs(y - yr)) <= 1e-12
        outs.append(y)
    assert np.max(np.abs(outs[0] - outs[1])) <= 1e-9


def test_reset_and_process_block_to(conv):
    """TestStreamingOverlapSaveReset (:101-131), ProcessBlockTo (:133-223)."""
    kernel = [1.0, 0.5, 0.25]
    s = conv.NewStreamingOverlapSave(kernel, 4)
    a = s.ProcessBlock([1, 2, 3, 4])
    s.ProcessBlock([5, 6, 7, 8])
    s.Reset()
    b = s.ProcessBlock([1, 2, 3, 4])
    assert np.max(np.abs(a - b)) <= 1e-12
    out = np.zeros(4)
    s.Reset()
    s.ProcessBlockTo(out, [1, 2, 3, 4])
    assert np.max(np.abs(out - a)) <= 1e-12


def test_errors_and_getters(conv):
    """TestStreamingOverlapSaveErrors (:225-277), Getters (:279-304)."""
    with pytest.raises(conv.ConvError) as ei:
        conv.NewStreamingOverlapSave([], 128)
    assert conv.errors_is(ei.value, conv.ErrEmptyKernel)
    for bad in (0, -1):
        with pytest.raises(conv.ConvError):
            conv.NewStreamingOverlapSave([1.0], bad)
    s = conv.NewStreamingOverlapSave([1.0, 0.5], 4)
    with pytest.raises(conv.ConvError) as ei:
        s.ProcessBlock([1, 2, 3])
    assert conv.errors_is(ei.value, conv.ErrLengthMismatch)
    with pytest.raises(conv.ConvError) as ei:
        s.ProcessBlockTo(np.zeros(3), [1, 2, 3, 4])
    assert conv.errors_is(ei.value, conv.ErrLengthMismatch)
    kernel = [1.0, 0.5, 0.25, 0.1]
    for new in ctors(conv):
        s = new(kernel, 8)
        assert (s.BlockSize(), s.KernelLen(), s.FFTSize()) == (8, 4, conv.nextPowerOf2(8 + 4 - 1))


def test_dirac_long_kernel_and_continuity(conv, oracle):
    """DiracDelta (:306-329), LongKernel (:331-383), Continuity (:385-428, 1e-9)."""
    s = conv.NewStreamingOverlapSave([1.0], 16)
    x = G.white(16, seed=2)
    assert np.max(np.abs(s.ProcessBlock(x) - x)) <= 1e-12
    # kernel longer than the block: history spans several blocks
    K, bs, nb = 300, 64, 20
    h, x = G.decaying_ir(K), G.white(bs * nb, seed=3)
    ref = oracle.overlap_save(h, 0, x)[: len(x)]
    for new in ctors(conv):
        s = new(h, bs)
        y = np.concatenate([s.ProcessBlock(x[i:i + bs]) for i in range(0, len(x), bs)])
        assert G.rel_l2(y, ref) <= 1e-12
    # long IR in big blocks (reverb-style streaming)
    K, bs, nb = 50000, 48000, 5
    h, x = G.decaying_ir(K), G.white(bs * nb, seed=4)
    ref = oracle.overlap_save(h, 0, x)[: len(x)]
    s = conv.NewStreamingOverlapSave(h, bs)
    y = np.concatenate([s.ProcessBlock(x[i:i + bs]) for i in range(0, len(x), bs)])
    assert G.rel_l2(y, ref) <= 1e-12
