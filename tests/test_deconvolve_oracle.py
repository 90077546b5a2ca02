"""CPU: the oracle's restatement of deconvolve.go against the reference's own checks (conv_test.go:282-330, 563-655,
example_test.go:129-156) and against numpy."""
import math

import numpy as np

from oracle import oracle as O


def test_example_deconvolve_golden_snr():
    """ExampleDeconvolve prints 'Recovery SNR: 39.6 dB', lengths 50 / 50 (example_test.go:129-156)."""
    original = np.sin(2 * np.pi * np.arange(50) / 10)
    kernel = [0.25, 0.5, 0.25]
    recovered = O.deconvolve(O.direct(original, kernel), kernel, O.DECONV_REGULARIZED, 1e-3)
    assert len(recovered) == 50 and f"{O.snr(original, recovered):.1f}" == "39.6"


def test_against_numpy_all_methods():
    rng = np.random.default_rng(0)
    for n, m in ((100, 3), (1000, 31), (5, 9), (4097, 100), (64, 64)):
        x, k = rng.standard_normal(n), rng.standard_normal(m)
        N = 1 << max(0, (n - 1).bit_length())
        if m > N:
            continue
        S, H = np.fft.fft(x, N), np.fft.fft(k, N)
        out_len = n - m + 1 if n - m + 1 > 0 else n
        for method, reg in ((O.DECONV_REGULARIZED, 1e-3), (O.DECONV_WIENER, None)):
            if method == O.DECONV_WIENER:
                got = O.deconvolve(x, k, method, noise_variance=0.02, signal_variance=2.0)
                reg = 0.01
            else:
                got = O.deconvolve(x, k, method, reg)
            ref = np.fft.ifft(S * np.conj(H) / (np.abs(H) ** 2 + reg)).real[:out_len]
            assert len(got) == out_len and np.max(np.abs(got - ref)) <= 1e-10 * max(1.0, np.max(np.abs(ref)))
        got = O.deconvolve(x, k, O.DECONV_NAIVE)
        ref = np.fft.ifft(S / H).real[:out_len]
        assert np.max(np.abs(got - ref)) <= 1e-8 * max(1.0, np.max(np.abs(ref)))
    assert abs(O.variance([1.0, 2.0, 3.0, 4.0]) - 1.25) < 1e-15


def test_reference_checks():
    # TestDeconvolveNaive: identity kernel recovers the signal (conv_test.go:563-583)
    x = np.sin(2 * np.pi * np.arange(50) / 10)
    assert np.max(np.abs(O.deconvolve(x, [1.0], O.DECONV_NAIVE) - x)) < 1e-12
    # division by zero: kernel [1, -1] has H[0] = 0 (deconvolve.go:144-148)
    try:
        O.deconvolve(np.ones(8), [1.0, -1.0], O.DECONV_NAIVE)
        assert False
    except O.OracleError as e:
        assert e.sentinel == "ErrDivisionByZero"
    # TestInverseFilter (conv_test.go:312-340): kernel * inverse has a dominant peak
    inv = O.inverse_filter([0.5, 1.0, 0.5], 64, 1e-3)
    res = O.direct([0.5, 1.0, 0.5], inv)
    idx, val = O.find_peak(res)
    assert len(inv) == 64 and val >= 0.1
    # TestSNR (conv_test.go:633-655)
    assert O.snr([1, 2, 3, 4, 5], [1, 2, 3, 4, 5]) == math.inf
    assert O.snr([1, 2, 3, 4, 5], [1, 2, 3]) == -math.inf and O.snr([], []) == -math.inf
    # errors (conv_test.go:619-631)
    for args, name in ((([], [1, 2]), "ErrEmptyInput"), (([1, 2], []), "ErrEmptyKernel")):
        try:
            O.deconvolve(*args)
            assert False
        except O.OracleError as e:
            assert e.sentinel == name
