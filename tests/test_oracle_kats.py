"""Pins the CPU oracle against every known-answer vector of the reference's own tests for the
dsp/conv hot path (tests/golden/reference_kats.json).  CPU only."""
import numpy as np
import pytest

import kat_checks
import kat_inputs as KI
from adapters import OracleImpl


@pytest.fixture(scope="module")
def impl():
    return OracleImpl()


@pytest.mark.parametrize("check", kat_checks.ALL_CHECKS, ids=lambda f: f.__name__)
def test_reference_kats(impl, check):
    check(impl)


def test_streaming_impulse_kats(impl):
    """TestStreamingOverlapSave/Add (streaming_overlap_*_test.go:9-56), f32 impulse (streaming_test.go:237-268)."""
    O = impl.O
    k = kat_checks.KATS["streaming_impulse"]
    for ols in (True, False):
        s = O.Streaming(k["kernel"], k["block_size"], ols)
        o1 = s.process_block(k["block1"])
        o2 = s.process_block(k["block2"])
        assert len(o1) == len(o2) == k["block_size"]
        assert np.max(np.abs(o1 - np.array(k["out1"]))) <= k["tol"]
    k = kat_checks.KATS["streaming_impulse_f32"]
    for ols in (True, False):
        s = O.Streaming(k["kernel"], k["block_size"], ols, dtype=np.float32)
        x = np.zeros(k["block_size"], np.float32)
        x[0] = 1
        assert np.max(np.abs(s.process_block(x) - np.array(k["out"], np.float32))) <= k["tol"]


def test_partitioned_matches_streaming_ola(impl):
    """TestPartitionedConvolutionMatchesSOA (partitioned_test.go:121-165): the reference's own
    second oracle, StreamingOverlapAdd with blockSize = latency, tol 1e-7."""
    O = impl.O
    for c in kat_checks.KATS["partitioned"]["vs_streaming_ola"]:
        h = 0.99 ** np.arange(c["kernel_len"])
        x = KI.pcg_uniform(c["signal_len"])
        lat = 1 << c["min"]
        s = O.Streaming(h, lat, False)
        soa = np.concatenate([s.process_block(x[i:i + lat]) for i in range(0, len(x), lat)])
        p = O.Partitioned(h, c["min"], c["max"])
        pc = p.process_block(np.concatenate([x, np.zeros(lat)]))[lat:]
        m = min(len(x), len(pc), len(soa))
        assert np.max(np.abs(pc[:m] - soa[:m])) <= c["tol"]


def test_streaming_vs_batch(impl):
    """streaming_overlap_save_test.go:51-99 / streaming_overlap_add_test.go:58-106 (1e-10) and
    OLA == OLS (streaming_test.go:122-176, 1e-9)."""
    O = impl.O
    kernel = [0.5, 1.0, 0.5, 0.2]
    bs, nb = 8, 4
    x = np.sin(np.arange(bs * nb) * 0.1)
    batch = O.overlap_save(kernel, 0, x)
    outs = {}
    for ols in (True, False):
        s = O.Streaming(kernel, bs, ols)
        outs[ols] = np.concatenate([s.process_block(x[i:i + bs]) for i in range(0, len(x), bs)])
        assert np.max(np.abs(outs[ols] - batch[: len(x)])) <= 1e-10
    assert np.max(np.abs(outs[True] - outs[False])) <= 1e-9


@pytest.mark.parametrize("n,K,fft", [(500, 4, 0), (1000, 100, 0), (5000, 300, 1024), (7, 9, 0), (3000, 1, 0), (100, 65, 0), (20000, 2000, 0)])
def test_oracle_vs_numpy(impl, n, K, fft):
    """Independent cross-check of the restatement: every path against numpy.convolve."""
    rng = np.random.default_rng(n + K)
    x, h = rng.standard_normal(n), rng.standard_normal(K)
    ref = np.convolve(x, h)
    for got in (impl.overlap_save(h, fft, x), impl.overlap_add(h, 0, x), impl.convolve(x, h), impl.direct(x, h)):
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-13
    assert np.max(np.abs(impl.correlate_fft(x, h) - impl.correlate(x, h))) < 1e-9


def test_ols_sizing_rules(impl):
    """NewOverlapSave sizing (overlap_save.go:53-76) at the configs of BASELINE.json."""
    assert impl.ols_sizes(96000, 0) == (262144, 166145)
    assert impl.ols_sizes(288000, 0) == (1048576, 760577)
    assert impl.ols_sizes(1 << 20, 0) == (1 << 21, (1 << 20) + 1)
    assert impl.ols_sizes(3, 0) == (256, 254)
    assert impl.ols_sizes(300, 256) == (1024, 725)       # too small -> silently raised (:71-73)
    assert impl.ola_sizes(96000, 0) == (131072, 262144)
