"""ctypes binding of oracle/_build/liboracle.so (the plain-C restatement of
/root/reference/dsp/conv).  TEST INFRASTRUCTURE ONLY -- never imported by the product
package algo_dsp_b200.

Function names mirror the reference's Go API (dsp/conv/*.go); each wrapper raises
OracleError carrying the same sentinel the Go code returns.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

OK, EMPTY_INPUT, EMPTY_KERNEL, LENGTH_MISMATCH, INVALID_BLOCK_SIZE, INVALID_BLOCK_ORDER, EMPTY_IR, STAGE_INDEX, INVALID_ARG, DIVISION_BY_ZERO = range(10)
_NAMES = {
    EMPTY_INPUT: "ErrEmptyInput", EMPTY_KERNEL: "ErrEmptyKernel", LENGTH_MISMATCH: "ErrLengthMismatch",
    INVALID_BLOCK_SIZE: "ErrInvalidBlockSize", INVALID_BLOCK_ORDER: "ErrInvalidBlockOrder",
    EMPTY_IR: "ErrEmptyImpulseResponse", STAGE_INDEX: "ErrStageIndexOutOfRange", INVALID_ARG: "ErrInvalidArgument",
    DIVISION_BY_ZERO: "ErrDivisionByZero",
}

MODE_FULL, MODE_SAME, MODE_VALID = 0, 1, 2


class OracleError(Exception):
    def __init__(self, code):
        self.code = code
        self.sentinel = _NAMES.get(code, f"error {code}")
        super().__init__(self.sentinel)


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (no GPU needed)."""
    srcs = [os.path.join(_HERE, f) for f in ("conv_oracle.c", "conv_oracle_impl.h", "Makefile")]
    stale = force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "CC=gcc"], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_next_pow2.restype = C.c_int64
        _lib.orc_next_pow2.argtypes = [C.c_int64]
        _lib.orc_lag_from_index.restype = C.c_int64
        _lib.orc_index_from_lag.restype = C.c_int64
        _lib.orc_lag_from_index.argtypes = [C.c_int64, C.c_int64]
        _lib.orc_index_from_lag.argtypes = [C.c_int64, C.c_int64]
        for sfx in ("_f64", "_f32"):
            for nm in ("orc_stream_create", "orc_part_create"):
                getattr(_lib, nm + sfx).restype = C.c_void_p
            getattr(_lib, "orc_stream_fft_size" + sfx).restype = C.c_int64
    return _lib


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "_f64", C.c_double
    if dtype == np.float32:
        return "_f32", C.c_float
    raise TypeError(dtype)


def _arr(x, dtype):
    return np.ascontiguousarray(x, dtype=dtype).reshape(-1)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _check(st):
    if st != OK:
        raise OracleError(st)


def next_power_of_2(n: int) -> int:
    return int(lib().orc_next_pow2(int(n)))


def ols_sizes(K: int, fft_size: int = 0):
    f, s = C.c_int64(), C.c_int64()
    _check(lib().orc_ols_sizes(C.c_int64(K), C.c_int64(fft_size), C.byref(f), C.byref(s)))
    return f.value, s.value


def ola_sizes(K: int, block_size: int = 0):
    b, f = C.c_int64(), C.c_int64()
    _check(lib().orc_ola_sizes(C.c_int64(K), C.c_int64(block_size), C.byref(b), C.byref(f)))
    return b.value, f.value


def _binary(name, a, b, out_len=None, dtype=np.float64):
    sfx, _ = _sfx(dtype)
    a = _arr(a, dtype)
    b = _arr(b, dtype)
    n, m = a.size, b.size
    if out_len is None:
        out_len = max(n + m - 1, 1)
    out = np.zeros(out_len, dtype=dtype)
    st = getattr(lib(), name + sfx)(_p(a), C.c_int64(n), _p(b), C.c_int64(m), _p(out))
    _check(st)
    return out


def direct(a, b, dtype=np.float64):
    return _binary("orc_direct", a, b, dtype=dtype)


def direct_circular(a, b, dtype=np.float64):
    return _binary("orc_direct_circular", a, b, out_len=max(len(a), 1), dtype=dtype)


def convolve(a, b, dtype=np.float64):
    return _binary("orc_convolve", a, b, dtype=dtype)


def trim_to_mode(full, len_a, len_b, mode):
    s, l = C.c_int64(), C.c_int64()
    lib().orc_trim_mode(C.c_int64(len_a), C.c_int64(len_b), C.c_int(mode), C.byref(s), C.byref(l))
    return full[s.value:s.value + l.value]


def convolve_mode(a, b, mode, dtype=np.float64):
    return trim_to_mode(convolve(a, b, dtype), len(a), len(b), mode)


def overlap_add(kernel, block_size, signal, dtype=np.float64):
    """NewOverlapAdd(kernel, block_size).Process(signal)."""
    sfx, _ = _sfx(dtype)
    k = _arr(kernel, dtype)
    x = _arr(signal, dtype)
    out = np.zeros(max(x.size + k.size - 1, 1), dtype=dtype)
    _check(getattr(lib(), "orc_ola_process" + sfx)(_p(k), C.c_int64(k.size), C.c_int64(block_size), _p(x), C.c_int64(x.size), _p(out)))
    return out


def overlap_save(kernel, fft_size, signal, dtype=np.float64):
    """NewOverlapSave(kernel, fft_size).Process(signal)."""
    sfx, _ = _sfx(dtype)
    k = _arr(kernel, dtype)
    x = _arr(signal, dtype)
    out = np.zeros(max(x.size + k.size - 1, 1), dtype=dtype)
    _check(getattr(lib(), "orc_ols_process" + sfx)(_p(k), C.c_int64(k.size), C.c_int64(fft_size), _p(x), C.c_int64(x.size), _p(out)))
    return out


def overlap_add_convolve(signal, kernel, dtype=np.float64):
    return overlap_add(kernel, 0, signal, dtype)


def overlap_save_convolve(signal, kernel, dtype=np.float64):
    return overlap_save(kernel, 0, signal, dtype)


def correlate(a, b, dtype=np.float64):
    return _binary("orc_correlate", a, b, dtype=dtype)


def correlate_direct(a, b, dtype=np.float64):
    return _binary("orc_correlate_direct", a, b, dtype=dtype)


def correlate_fft(a, b, dtype=np.float64):
    return _binary("orc_correlate_fft", a, b, dtype=dtype)


def correlate_normalized(a, b, dtype=np.float64):
    return _binary("orc_correlate_normalized", a, b, dtype=dtype)


def correlate_mode(a, b, mode, dtype=np.float64):
    return trim_to_mode(correlate(a, b, dtype), len(a), len(b), mode)


def auto_correlate(a, dtype=np.float64):
    return correlate(a, a, dtype)


def auto_correlate_normalized(a, dtype=np.float64):
    sfx, _ = _sfx(dtype)
    a = _arr(a, dtype)
    out = np.zeros(max(2 * a.size - 1, 1), dtype=dtype)
    _check(getattr(lib(), "orc_autocorrelate_normalized" + sfx)(_p(a), C.c_int64(a.size), _p(out)))
    return out


def find_peak(corr, dtype=np.float64):
    sfx, ct = _sfx(dtype)
    c = _arr(corr, dtype)
    idx, val = C.c_int64(), ct()
    getattr(lib(), "orc_find_peak" + sfx)(_p(c), C.c_int64(c.size), C.byref(idx), C.byref(val))
    return idx.value, val.value


def lag_from_index(index, len_b):
    return int(lib().orc_lag_from_index(index, len_b))


def index_from_lag(lag, len_b):
    return int(lib().orc_index_from_lag(lag, len_b))


class Streaming:
    """StreamingOverlapAddT / StreamingOverlapSaveT (streaming_overlap_*.go)."""

    def __init__(self, kernel, block_size, ols: bool, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        self.sfx, _ = _sfx(dtype)
        k = _arr(kernel, dtype)
        st = C.c_int()
        self.block_size = block_size
        self.h = getattr(lib(), "orc_stream_create" + self.sfx)(_p(k), C.c_int64(k.size), C.c_int64(block_size), C.c_int(1 if ols else 0), C.byref(st))
        _check(st.value)

    def fft_size(self):
        return int(getattr(lib(), "orc_stream_fft_size" + self.sfx)(C.c_void_p(self.h)))

    def process_block(self, x):
        x = _arr(x, self.dtype)
        out = np.zeros(max(x.size, 1), dtype=self.dtype)
        _check(getattr(lib(), "orc_stream_process_block" + self.sfx)(C.c_void_p(self.h), _p(x), C.c_int64(x.size), _p(out)))
        return out[:x.size]

    def reset(self):
        getattr(lib(), "orc_stream_reset" + self.sfx)(C.c_void_p(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            getattr(lib(), "orc_stream_destroy" + self.sfx)(C.c_void_p(self.h))
            self.h = None


class Partitioned:
    """PartitionedConvolutionT (partitioned.go)."""

    def __init__(self, kernel, min_order, max_order, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        self.sfx, _ = _sfx(dtype)
        k = _arr(kernel, dtype)
        st = C.c_int()
        self.h = getattr(lib(), "orc_part_create" + self.sfx)(_p(k), C.c_int64(k.size), C.c_int(min_order), C.c_int(max_order), C.byref(st))
        _check(st.value)

    def _i(self, name):
        return int(getattr(lib(), name + self.sfx)(C.c_void_p(self.h)))

    def latency(self):
        return self._i("orc_part_latency")

    def kernel_len(self):
        return self._i("orc_part_kernel_len")

    def stage_count(self):
        return self._i("orc_part_stage_count")

    def stage_info(self, index):
        ps, bc, sp = C.c_int(), C.c_int(), C.c_int()
        _check(getattr(lib(), "orc_part_stage_info" + self.sfx)(C.c_void_p(self.h), C.c_int(index), C.byref(ps), C.byref(bc), C.byref(sp)))
        return ps.value, bc.value, sp.value

    def process_block(self, x, out_len=None):
        x = _arr(x, self.dtype)
        n_out = x.size if out_len is None else out_len
        out = np.zeros(max(n_out, 1), dtype=self.dtype)
        _check(getattr(lib(), "orc_part_process_block" + self.sfx)(C.c_void_p(self.h), _p(x), C.c_int64(x.size), _p(out), C.c_int64(n_out)))
        return out[:n_out]

    def reset(self):
        getattr(lib(), "orc_part_reset" + self.sfx)(C.c_void_p(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            getattr(lib(), "orc_part_destroy" + self.sfx)(C.c_void_p(self.h))
            self.h = None


DECONV_NAIVE, DECONV_REGULARIZED, DECONV_WIENER = 0, 1, 2


def deconvolve(signal, kernel, method=DECONV_REGULARIZED, epsilon=1e-6, noise_variance=0.0, signal_variance=0.0):
    """Deconvolve(signal, kernel, opts) -- deconvolve.go:72 (float64 only, like the reference)."""
    x, k = _arr(signal, np.float64), _arr(kernel, np.float64)
    L = lib()
    L.orc_deconv_out_len.restype = C.c_int64
    out = np.zeros(max(int(L.orc_deconv_out_len(C.c_int64(x.size), C.c_int64(k.size))), 1))
    bad = C.c_int64(-1)
    _check(L.orc_deconvolve(_p(x), C.c_int64(x.size), _p(k), C.c_int64(k.size), C.c_int(method), C.c_double(epsilon),
                            C.c_double(noise_variance), C.c_double(signal_variance), _p(out), C.byref(bad)))
    return out[: int(L.orc_deconv_out_len(C.c_int64(x.size), C.c_int64(k.size)))]


def inverse_filter(kernel, length, epsilon):
    """InverseFilter(kernel, length, epsilon) -- deconvolve.go:359."""
    k = _arr(kernel, np.float64)
    out = np.zeros(max(int(length), 1))
    _check(lib().orc_inverse_filter(_p(k), C.c_int64(k.size), C.c_int64(length), C.c_double(epsilon), _p(out)))
    return out[: int(length)]


def snr(original, recovered):
    """SNR(original, recovered) -- deconvolve.go:417."""
    a, b = _arr(original, np.float64), _arr(recovered, np.float64)
    f = lib().orc_snr
    f.restype = C.c_double
    return float(f(_p(a), C.c_int64(0 if len(original) == 0 else a.size), _p(b), C.c_int64(0 if len(recovered) == 0 else b.size)))


def variance(x):
    a = _arr(x, np.float64)
    f = lib().orc_variance
    f.restype = C.c_double
    return float(f(_p(a), C.c_int64(a.size)))


def bench_ols(kernel, signal2d, fft_size=0, threads=1):
    """Timed-baseline helper: OverlapSave built once, Process per channel (rows of signal2d)."""
    k = _arr(kernel, np.float64)
    x = np.ascontiguousarray(signal2d, dtype=np.float64)
    if x.ndim == 1:
        x = x[None, :]
    ch, n = x.shape
    out = np.zeros((ch, n + k.size - 1), dtype=np.float64)
    _check(lib().orc_bench_ols_f64(_p(k), C.c_int64(k.size), C.c_int64(fft_size), _p(x), C.c_int64(n), C.c_int64(ch), C.c_int64(n),
                                   _p(out), C.c_int64(out.shape[1]), C.c_int(threads)))
    return out


def num_procs():
    return int(lib().orc_num_procs())
