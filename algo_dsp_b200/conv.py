"""Python mirror of the reference's Go package `conv` (dsp/conv/*.go) on top of the C ABI.

Same names, argument meaning and error behaviour as the Go API, so the parity tests read like
the reference's own tests:

    result = conv.Convolve(signal, kernel)            # conv.go:194
    c = conv.NewOverlapSave(kernel, 0); y = c.Process(x)   # overlap_save.go:53,126
    idx, val = conv.FindPeak(conv.Correlate(a, b))    # correlate.go:16,200

Go returns (value, error); here errors are raised as ConvError whose `.sentinel` is one of the
Err* objects below (`errors_is(err, ErrEmptyInput)` mirrors errors.Is).  All arithmetic runs in
libalgodsp_cuda on the GPU -- there is no CPU path in this module.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib as L

# ---------------------------------------------------------------- errors (conv.go:41-46, partitioned.go:11-15)


class _Sentinel:
    def __init__(self, name, text):
        self.name, self.text = name, text

    def __repr__(self):
        return self.name


ErrEmptyInput = _Sentinel("ErrEmptyInput", "conv: empty input")
ErrEmptyKernel = _Sentinel("ErrEmptyKernel", "conv: empty kernel")
ErrLengthMismatch = _Sentinel("ErrLengthMismatch", "conv: buffer length mismatch")
ErrInvalidBlockSize = _Sentinel("ErrInvalidBlockSize", "conv: invalid block size")
ErrInvalidBlockOrder = _Sentinel("ErrInvalidBlockOrder", "conv: invalid block order")
ErrEmptyImpulseResponse = _Sentinel("ErrEmptyImpulseResponse", "conv: empty impulse response")
ErrStageIndexOutOfRange = _Sentinel("ErrStageIndexOutOfRange", "conv: stage index out of range")
ErrDivisionByZero = _Sentinel("ErrDivisionByZero", "conv: division by zero in deconvolution")      # deconvolve.go:14
ErrInvalidEpsilon = _Sentinel("ErrInvalidEpsilon", "conv: epsilon must be positive")                # deconvolve.go:15 (never returned by the reference)
ErrInvalidNoiseVar = _Sentinel("ErrInvalidNoiseVar", "conv: noise variance must be positive")       # deconvolve.go:16 (never returned)
ErrInvalidArgument = _Sentinel("ErrInvalidArgument", "algodsp: invalid argument")
ErrCUDA = _Sentinel("ErrCUDA", "algodsp: CUDA error")
ErrOutOfMemory = _Sentinel("ErrOutOfMemory", "algodsp: out of memory")

_SENTINELS = {
    L.ERR_EMPTY_INPUT: ErrEmptyInput, L.ERR_EMPTY_KERNEL: ErrEmptyKernel, L.ERR_LENGTH_MISMATCH: ErrLengthMismatch,
    L.ERR_INVALID_BLOCK_SIZE: ErrInvalidBlockSize, L.ERR_INVALID_BLOCK_ORDER: ErrInvalidBlockOrder,
    L.ERR_EMPTY_IR: ErrEmptyImpulseResponse, L.ERR_STAGE_INDEX: ErrStageIndexOutOfRange,
    L.ERR_INVALID_ARG: ErrInvalidArgument, L.ERR_CUDA: ErrCUDA, L.ERR_OOM: ErrOutOfMemory,
    L.ERR_DIVISION_BY_ZERO: ErrDivisionByZero,
}


class ConvError(Exception):
    def __init__(self, status, detail=""):
        self.status = status
        self.sentinel = _SENTINELS.get(status, ErrInvalidArgument)
        msg = self.sentinel.text + (f": {detail}" if detail else "")
        super().__init__(msg)


def errors_is(err, sentinel) -> bool:
    """errors.Is(err, sentinel)."""
    return isinstance(err, ConvError) and err.sentinel is sentinel


def _check(st):
    if st != L.OK:
        detail = L.last_error() if st in (L.ERR_CUDA, L.ERR_OOM, L.ERR_INVALID_ARG, L.ERR_DIVISION_BY_ZERO) else ""
        raise ConvError(st, detail)


# ---------------------------------------------------------------- Mode (conv.go:57-69)
ModeFull, ModeSame, ModeValid = 0, 1, 2

# ---------------------------------------------------------------- context


class Context:
    """One GPU (adsp_ctx): streams, twiddle tables, L2-resident scratch, staging buffers."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        _check(L.load().adsp_ctx_create(int(device), C.byref(self._h)))
        self.device = device

    @property
    def handle(self):
        return self._h

    def sync(self):
        _check(L.load().adsp_ctx_sync(self._h))

    def launch_count(self) -> int:
        return int(L.load().adsp_ctx_launch_count(self._h))

    def stream(self) -> int:
        return int(L.load().adsp_ctx_stream(self._h) or 0)

    KERNEL_KINDS = ("cols_fwd", "rows", "cols_inv", "full", "direct", "other", "fused")

    def kernel_timing(self, enable: bool):
        L.load().adsp_ctx_kernel_timing(self._h, 1 if enable else 0)

    def kernel_times(self, reset=True):
        """{kind: (total_ms, launches)} accumulated while kernel_timing was on."""
        out = {}
        for i, name in enumerate(self.KERNEL_KINDS):
            ms, n = C.c_double(), C.c_uint64()
            _check(L.load().adsp_ctx_kernel_time(self._h, i, C.byref(ms), C.byref(n), 0))
            out[name] = (ms.value, n.value)
        if reset:
            _check(L.load().adsp_ctx_kernel_time(self._h, 0, None, None, 1))
        return out

    def host_profile(self, enable: bool):
        """Record a phase breakdown of the next small host-buffer call (adsp_ctx_host_profile)."""
        L.load().adsp_ctx_host_profile(self._h, 1 if enable else 0)

    def host_profile_get(self):
        """{alloc_ms, upload_ms, kernels_ms, download_ms, total_ms, staged_in_bytes, staged_out_bytes} of the last profiled call."""
        ms = (C.c_double * 12)()
        bi, bo = C.c_uint64(), C.c_uint64()
        _check(L.load().adsp_ctx_host_profile_get(self._h, ms, 12, C.byref(bi), C.byref(bo)))
        return {"alloc_ms": ms[0], "upload_ms": ms[1], "kernels_ms": ms[2], "download_ms": ms[3], "total_ms": ms[5],
                "stage_in_copy_ms": ms[6], "d2h_wait_ms": ms[7], "stage_out_copy_ms": ms[8], "pointer_query_ms": ms[9],
                "staged_in_bytes": int(bi.value), "staged_out_bytes": int(bo.value)}

    def stage_threads(self) -> int:
        return int(L.load().adsp_ctx_stage_threads(self._h))

    def close(self):
        if self._h:
            L.load().adsp_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float64):
    """numpy array backed by page-locked host memory (adsp_host_alloc_pinned); keep a reference to
    the returned array -- the memory is released when it is garbage collected."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    ptr = C.c_void_p()
    _check(L.load().adsp_host_alloc_pinned(n * dtype.itemsize, C.byref(ptr)))
    buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    import weakref
    weakref.finalize(buf, L.load().adsp_host_free_pinned, ptr)
    return arr


_default_ctx: dict[int, Context] = {}
_default_ctx_lock = threading.Lock()


def default_context(device: int = 0) -> Context:
    """The process-wide context of a device (sync.Once in the Go shim).  Created under a lock: two threads racing here would
    each build a context, and the loser's would be destroyed while its handle is already on its way into a call."""
    ctx = _default_ctx.get(device)
    if ctx is None:
        with _default_ctx_lock:
            ctx = _default_ctx.get(device)
            if ctx is None:
                ctx = _default_ctx[device] = Context(device)
    return ctx


def _ctx(ctx):
    return (ctx or default_context()).handle


def _f64(x):
    return np.ascontiguousarray(x, dtype=np.float64).reshape(-1)


def _as(x, dtype):
    return np.ascontiguousarray(x, dtype=dtype).reshape(-1)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a.size else C.c_void_p(0)


def _binary(name, a, b, ctx=None, dtype=np.float64):
    a, b = _as(a, dtype), _as(b, dtype)
    out = np.empty(max(a.size + b.size - 1, 1), dtype=dtype)
    _check(getattr(L.load(), name)(_ctx(ctx), _p(a), a.size, _p(b), b.size, _p(out)))
    return out[: a.size + b.size - 1]


# ---------------------------------------------------------------- package-level functions


def Direct(a, b, ctx=None):
    """Direct(a, b) -- conv.go:76."""
    return _binary("adsp_direct", a, b, ctx)


def DirectTo(dst, a, b, ctx=None):
    """DirectTo(dst, a, b) -- conv.go:97 (dst must have len(a)+len(b)-1 elements)."""
    dst[:] = Direct(a, b, ctx)


def DirectCircular(a, b, ctx=None):
    """DirectCircular(a, b) -- conv.go:158."""
    a, b = _f64(a), _f64(b)
    out = np.empty(max(a.size, 1), dtype=np.float64)
    _check(L.load().adsp_direct_circular(_ctx(ctx), _p(a), a.size, _p(b), b.size, _p(out)))
    return out[: a.size]


def Convolve(a, b, ctx=None):
    """Convolve(a, b): direct iff the shorter operand has <= 64 taps, else FFT -- conv.go:194."""
    return _binary("adsp_convolve", a, b, ctx)


def trimToMode(full, lenA, lenB, mode):
    """trimToMode -- conv.go:229."""
    s, l = C.c_int64(), C.c_int64()
    L.load().adsp_trim_mode(lenA, lenB, int(mode), C.byref(s), C.byref(l))
    return full[s.value: s.value + l.value]


def ConvolveMode(a, b, mode, ctx=None):
    """ConvolveMode -- conv.go:219."""
    return trimToMode(Convolve(a, b, ctx), len(a), len(b), mode)


def OverlapAddConvolve(signal, kernel, ctx=None):
    """OverlapAddConvolve -- overlap_add.go:221."""
    return _binary("adsp_overlap_add_convolve", signal, kernel, ctx)


def OverlapSaveConvolve(signal, kernel, ctx=None):
    """OverlapSaveConvolve -- overlap_save.go:313."""
    return _binary("adsp_overlap_save_convolve", signal, kernel, ctx)


def OverlapAddConvolveTo(output, signal, kernel, ctx=None):
    """OverlapAddConvolveTo -- overlap_add.go:297."""
    oa = NewOverlapAdd(kernel, 0, ctx=ctx)
    oa.ProcessTo(output, signal)


def Correlate(a, b, ctx=None):
    """Correlate(a, b) = Convolve(a, reverse(b)) -- correlate.go:16."""
    return _binary("adsp_correlate", a, b, ctx)


def CorrelateDirect(a, b, ctx=None):
    """CorrelateDirect -- correlate.go:31."""
    return _binary("adsp_correlate_direct", a, b, ctx)


def CorrelateFFT(a, b, ctx=None):
    """CorrelateFFT -- correlate.go:111."""
    return _binary("adsp_correlate_fft", a, b, ctx)


def CorrelateMode(a, b, mode, ctx=None):
    """CorrelateMode -- correlate.go:45."""
    return trimToMode(Correlate(a, b, ctx), len(a), len(b), mode)


def AutoCorrelate(a, ctx=None):
    """AutoCorrelate -- correlate.go:57."""
    return Correlate(a, a, ctx)


def AutoCorrelateNormalized(a, ctx=None):
    """AutoCorrelateNormalized -- correlate.go:63."""
    a = _f64(a)
    out = np.empty(max(2 * a.size - 1, 1), dtype=np.float64)
    _check(L.load().adsp_autocorrelate_normalized(_ctx(ctx), _p(a), a.size, _p(out)))
    return out[: 2 * a.size - 1]


def CorrelateNormalized(a, b, ctx=None):
    """CorrelateNormalized -- correlate.go:86."""
    return _binary("adsp_correlate_normalized", a, b, ctx)


# ---------------------------------------------------------------- deconvolution (deconvolve.go)
DeconvNaive, DeconvRegularized, DeconvWiener = 0, 1, 2        # DeconvMethod, deconvolve.go:20-35


class DeconvOptions:
    """DeconvOptions -- deconvolve.go:37-54."""

    def __init__(self, Method=DeconvRegularized, Epsilon=0.0, NoiseVariance=0.0, SignalVariance=0.0):
        self.Method, self.Epsilon, self.NoiseVariance, self.SignalVariance = Method, Epsilon, NoiseVariance, SignalVariance


def DefaultDeconvOptions():
    """DefaultDeconvOptions -- deconvolve.go:56."""
    return DeconvOptions(DeconvRegularized, 1e-6)


def Deconvolve(signal, kernel, opts=None, ctx=None):
    """Deconvolve -- deconvolve.go:72: len(signal)-len(kernel)+1 samples (len(signal) if that is not positive)."""
    opts = opts or DefaultDeconvOptions()
    x, k = _f64(signal), _f64(kernel)
    if len(signal) == 0:
        raise ConvError(L.ERR_EMPTY_INPUT)
    if len(kernel) == 0:
        raise ConvError(L.ERR_EMPTY_KERNEL)
    n_out = int(L.load().adsp_deconv_out_len(x.size, k.size))
    out = np.empty(n_out)
    _check(L.load().adsp_deconvolve(_ctx(ctx), _p(x), x.size, _p(k), k.size, int(opts.Method), float(opts.Epsilon),
                                    float(opts.NoiseVariance), float(opts.SignalVariance), _p(out), n_out))
    return out


def InverseFilter(kernel, length, epsilon, ctx=None):
    """InverseFilter -- deconvolve.go:359."""
    k = _f64(kernel)
    if len(kernel) == 0:
        raise ConvError(L.ERR_EMPTY_KERNEL)
    out = np.empty(max(int(length), 0))
    if length > 0:
        _check(L.load().adsp_inverse_filter(_ctx(ctx), _p(k), k.size, int(length), float(epsilon), _p(out)))
    return out


def SNR(original, recovered):
    """SNR -- deconvolve.go:417 (dB; +Inf for identical signals, -Inf for mismatched or empty ones)."""
    a, b = _f64(original), _f64(recovered)
    return float(L.load().adsp_snr(_p(a), len(original), _p(b), len(recovered)))


def FindPeak(corr, ctx=None):
    """FindPeak -- correlate.go:200: (index, value) of the signed maximum, (-1, 0) if empty."""
    c = _f64(corr)
    idx, val = C.c_int64(), C.c_double()
    _check(L.load().adsp_find_peak(_ctx(ctx), _p(c), c.size, C.byref(idx), C.byref(val)))
    return idx.value, val.value


def LagFromIndex(index, lenB):
    """LagFromIndex -- correlate.go:221."""
    return int(L.load().adsp_lag_from_index(index, lenB))


def IndexFromLag(lag, lenB):
    """IndexFromLag -- correlate.go:227."""
    return int(L.load().adsp_index_from_lag(lag, lenB))


def nextPowerOf2(n):
    """nextPowerOf2 -- conv.go:250."""
    return int(L.load().adsp_next_pow2(n))


def isPowerOf2(n):
    """isPowerOf2 -- conv.go:264."""
    return bool(L.load().adsp_is_pow2(n))


# f32 twins (optional fp32 mode)
def Direct32(a, b, ctx=None):
    return _binary("adsp_direct_f32", a, b, ctx, np.float32)


def Convolve32(a, b, ctx=None):
    return _binary("adsp_convolve_f32", a, b, ctx, np.float32)


def Correlate32(a, b, ctx=None):
    return _binary("adsp_correlate_f32", a, b, ctx, np.float32)


# batched forms (configs 2 and 4)
def DirectBatch(a2d, b, ctx=None):
    """Direct over the rows of a2d; b is one kernel (1-D) or one per row (2-D)."""
    a = np.ascontiguousarray(a2d, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    batch, n = a.shape
    m = b.shape[-1]
    out = np.empty((batch, n + m - 1), dtype=np.float64)
    bs = m if b.ndim == 2 else 0
    _check(L.load().adsp_direct_batch(_ctx(ctx), _p(a), n, n, _p(b), m, bs, batch, _p(out), n + m - 1))
    return out


def CorrelateBatch(a2d, b2d, want_output=True, ctx=None):
    """Correlate row p of a2d with row p of b2d; returns (out or None, peak_index, peak_value)."""
    a = np.ascontiguousarray(a2d, dtype=np.float64)
    b = np.ascontiguousarray(b2d, dtype=np.float64)
    pairs, n = a.shape
    m = b.shape[1]
    out = np.empty((pairs, n + m - 1), dtype=np.float64) if want_output else None
    pi = np.empty(pairs, dtype=np.int64)
    pv = np.empty(pairs, dtype=np.float64)
    _check(L.load().adsp_correlate_batch(_ctx(ctx), _p(a), n, n, _p(b), m, m, pairs,
                                         _p(out) if out is not None else C.c_void_p(0), n + m - 1, _p(pi), _p(pv)))
    return out, pi, pv


# ---------------------------------------------------------------- reusable convolvers


class _Plan:
    _dtype = np.float64

    def __init__(self):
        self._h = C.c_void_p()
        self._ctx_obj = None

    def _np(self, x):
        return _as(x, self._dtype)

    def KernelLen(self):
        return int(L.load().adsp_plan_kernel_len(self._h))

    def FFTSize(self):
        return int(L.load().adsp_plan_fft_size(self._h))

    def Reset(self):
        L.load().adsp_plan_reset(self._h)

    def internal_geometry(self):
        v = [C.c_int64() for _ in range(5)]
        L.load().adsp_plan_internal_geometry(self._h, *[C.byref(x) for x in v])
        return dict(zip(("fft_n", "n1", "n2", "step", "partitions"), (x.value for x in v)))

    def describe_cover(self, n):
        """Transforms the GPU runs for Process() on n samples: [{fft_n, out_offset, out_len, single_block}] (diagnostic)."""
        buf = (C.c_int64 * 32)()
        k = L.load().adsp_plan_describe_cover(self._h, int(n), buf, 8)
        return [dict(fft_n=buf[4 * i], out_offset=buf[4 * i + 1], out_len=buf[4 * i + 2], single_block=bool(buf[4 * i + 3]))
                for i in range(min(k, 8))]

    def Process(self, input):
        """Process(input) -> full linear convolution, len(input)+KernelLen()-1 samples."""
        x = self._np(input)
        out = np.empty(max(x.size + self.KernelLen() - 1, 1), dtype=self._dtype)
        n_out = x.size + self.KernelLen() - 1
        _check(L.load().adsp_plan_process(self._h, _p(x), x.size, _p(out), n_out))
        return out[:n_out]

    def ProcessTo(self, output, input):
        """ProcessTo(output, input): ErrLengthMismatch unless len(output) == len(input)+K-1."""
        x = self._np(input)
        if not (isinstance(output, np.ndarray) and output.dtype == self._dtype and output.flags.c_contiguous):
            raise TypeError("output must be a contiguous numpy array of the plan's dtype")
        _check(L.load().adsp_plan_process(self._h, _p(x), x.size, _p(output), output.size))

    def ProcessBatch(self, x2d, out=None):
        """All rows of x2d (channels x n) in one call (host buffers; pinned ones DMA directly)."""
        x = x2d if (isinstance(x2d, np.ndarray) and x2d.dtype == self._dtype and x2d.flags.c_contiguous) else \
            np.ascontiguousarray(x2d, dtype=self._dtype)
        ch, n = x.shape
        if out is None:
            out = np.empty((ch, n + self.KernelLen() - 1), dtype=self._dtype)
        _check(L.load().adsp_plan_process_batch(self._h, _p(x), n, ch, n, _p(out), out.shape[1]))
        return out

    def process_device(self, in_ptr, n, channels, in_stride, out_ptr, out_stride):
        """Device-resident call (raw device pointers); asynchronous on the context stream."""
        _check(L.load().adsp_plan_process_device(self._h, C.c_void_p(in_ptr), n, channels, in_stride, C.c_void_p(out_ptr), out_stride))

    def sync(self):
        _check(L.load().adsp_plan_sync(self._h))

    def Close(self):
        if self._h:
            L.load().adsp_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.Close()
        except Exception:
            pass


class OverlapSave(_Plan):
    """OverlapSave -- overlap_save.go:32."""

    def __init__(self, kernel, fftSize=0, ctx=None, dtype=np.float64):
        super().__init__()
        self._dtype = np.dtype(dtype).type
        self._ctx_obj = ctx or default_context()
        k = self._np(kernel)
        prec = L.F64 if self._dtype == np.float64 else L.F32
        _check(L.load().adsp_overlap_save_create(self._ctx_obj.handle, _p(k), k.size, int(fftSize), prec, C.byref(self._h)))

    def StepSize(self):
        return int(L.load().adsp_plan_step_size(self._h))


class OverlapAdd(_Plan):
    """OverlapAdd -- overlap_add.go:24."""

    def __init__(self, kernel, blockSize=0, ctx=None, dtype=np.float64):
        super().__init__()
        self._dtype = np.dtype(dtype).type
        self._ctx_obj = ctx or default_context()
        k = self._np(kernel)
        prec = L.F64 if self._dtype == np.float64 else L.F32
        _check(L.load().adsp_overlap_add_create(self._ctx_obj.handle, _p(k), k.size, int(blockSize), prec, C.byref(self._h)))

    def BlockSize(self):
        return int(L.load().adsp_plan_block_size(self._h))


def NewOverlapSave(kernel, fftSize=0, ctx=None, dtype=np.float64):
    """NewOverlapSave(kernel, fftSize) -- overlap_save.go:53."""
    return OverlapSave(kernel, fftSize, ctx, dtype)


def NewOverlapAdd(kernel, blockSize=0, ctx=None, dtype=np.float64):
    """NewOverlapAdd(kernel, blockSize) -- overlap_add.go:44."""
    return OverlapAdd(kernel, blockSize, ctx, dtype)


def ProcessBatchMulti(plans, x2d, out=None):
    """One host call over several GPUs: `plans` are OverlapSave/OverlapAdd plans of the same kernel, one per Context
    (device); channels (rows) are split contiguously across them, no collective (adsp_plans_process_batch)."""
    p0 = plans[0]
    x = np.ascontiguousarray(x2d, dtype=p0._dtype)
    ol = x.shape[1] + p0.KernelLen() - 1
    y = np.empty((x.shape[0], ol), dtype=p0._dtype) if out is None else out
    arr = (C.c_void_p * len(plans))(*[p._h for p in plans])
    _check(L.load().adsp_plans_process_batch(arr, len(plans), _p(x), x.shape[1], x.shape[0], x.shape[1], _p(y), y.shape[1]))
    return y


def ProcessLongMulti(plans, x):
    """One long signal over several GPUs by time block with a (K-1) halo (adsp_plans_process_long)."""
    p0 = plans[0]
    xx = np.ascontiguousarray(x, dtype=p0._dtype)
    y = np.empty(xx.size + p0.KernelLen() - 1, dtype=p0._dtype)
    arr = (C.c_void_p * len(plans))(*[p._h for p in plans])
    _check(L.load().adsp_plans_process_long(arr, len(plans), _p(xx), xx.size, _p(y), y.size))
    return y


class PartitionedConvolution(_Plan):
    """PartitionedConvolutionT -- partitioned.go:27 (f64: PartitionedConvolution, f32: PartitionedConvolution32).
    channels > 1: that many independent streams through one IR, one launch set per call (rows of 2-D blocks)."""

    def __init__(self, kernel, minBlockOrder, maxBlockOrder, ctx=None, dtype=np.float64, channels=1):
        super().__init__()
        self._dtype = np.dtype(dtype).type
        self._ctx_obj = ctx or default_context()
        self._channels = int(channels)
        k = self._np(kernel)
        prec = L.F64 if self._dtype == np.float64 else L.F32
        if self._channels == 1:
            _check(L.load().adsp_partitioned_create(self._ctx_obj.handle, _p(k), k.size, int(minBlockOrder), int(maxBlockOrder),
                                                    prec, C.byref(self._h)))
        else:
            _check(L.load().adsp_partitioned_create_batch(self._ctx_obj.handle, _p(k), k.size, int(minBlockOrder), int(maxBlockOrder),
                                                          self._channels, prec, C.byref(self._h)))

    def ProcessBlock(self, input, output):
        """ProcessBlock(input, output): equal lengths; output delayed by Latency() samples."""
        x = self._np(input)
        if not (isinstance(output, np.ndarray) and output.dtype == self._dtype and output.flags.c_contiguous):
            raise TypeError("output must be a contiguous numpy array of the plan's dtype")
        _check(L.load().adsp_partitioned_process_block(self._h, _p(x), x.size, _p(output), output.size))

    def ProcessBlockBatch(self, input2d, output2d=None):
        """Rows = channels: every row is ProcessBlock'ed as its own stream (same call for all channels)."""
        x = np.ascontiguousarray(input2d, dtype=self._dtype)
        if x.ndim != 2 or x.shape[0] != self._channels:
            raise ValueError("input must be [channels, n]")
        out = np.empty_like(x) if output2d is None else output2d
        _check(L.load().adsp_partitioned_process_block_batch(self._h, _p(x), x.shape[1], x.shape[1], _p(out), out.shape[1]))
        return out

    def process_block_device(self, in_ptr, n, in_stride, out_ptr, out_stride):
        _check(L.load().adsp_partitioned_process_block_batch_device(self._h, C.c_void_p(in_ptr), int(n), int(in_stride),
                                                                    C.c_void_p(out_ptr), int(out_stride)))

    def SetWetDry(self, wet, dry):
        """ConvolutionReverb.SetWetDry -- dsp/effects/reverb/convolution.go:51."""
        _check(L.load().adsp_partitioned_set_wet_dry(self._h, float(wet), float(dry)))

    def ProcessInPlace(self, block):
        """ConvolutionReverb.ProcessInPlace -- convolution.go:60: block = dry*block + wet*reverb(block); 1-D (mono plan) or
        [channels, n]."""
        if not (isinstance(block, np.ndarray) and block.dtype == self._dtype and block.flags.c_contiguous):
            raise TypeError("block must be a contiguous numpy array of the plan's dtype")
        n = block.shape[-1]
        rows = 1 if block.ndim == 1 else block.shape[0]
        if rows != self._channels:
            raise ValueError("block must have one row per channel")
        _check(L.load().adsp_partitioned_process_in_place_batch(self._h, _p(block), n, n))

    def process_in_place_device(self, ptr, n, stride):
        """ProcessInPlace on device rows ([channels] rows of n samples, `stride` elements apart); asynchronous."""
        _check(L.load().adsp_partitioned_process_in_place_batch_device(self._h, C.c_void_p(ptr), int(n), int(stride)))

    def Latency(self):
        return int(L.load().adsp_partitioned_latency(self._h))

    def StageCount(self):
        return int(L.load().adsp_partitioned_stage_count(self._h))

    def StageInfo(self, index):
        ps, bc = C.c_int(), C.c_int()
        _check(L.load().adsp_partitioned_stage_info(self._h, int(index), C.byref(ps), C.byref(bc)))
        return ps.value, bc.value

    def internal_stages(self):
        """[(partition size, partitions, IR offset)] of the delay-line engine (diagnostic)."""
        out = []
        for i in range(int(L.load().adsp_partitioned_internal_stage_count(self._h))):
            ps, cnt, off = C.c_int(), C.c_int(), C.c_int64()
            _check(L.load().adsp_partitioned_internal_stage_info(self._h, i, C.byref(ps), C.byref(cnt), C.byref(off)))
            out.append((ps.value, cnt.value, off.value))
        return out


def NewPartitionedConvolution(kernel, minBlockOrder, maxBlockOrder, ctx=None, channels=1):
    """NewPartitionedConvolution -- partitioned.go:335."""
    return PartitionedConvolution(kernel, minBlockOrder, maxBlockOrder, ctx, np.float64, channels)


def NewPartitionedConvolution32(kernel, minBlockOrder, maxBlockOrder, ctx=None, channels=1):
    """NewPartitionedConvolution32 -- partitioned.go:340."""
    return PartitionedConvolution(kernel, minBlockOrder, maxBlockOrder, ctx, np.float32, channels)


def NewConvolutionReverb(kernel, minBlockOrder, ctx=None, channels=1, dtype=np.float64):
    """NewConvolutionReverb -- dsp/effects/reverb/convolution.go:27 (maxBlockOrder 13, wet = dry = 1): use SetWetDry and
    ProcessInPlace on the returned engine."""
    return PartitionedConvolution(kernel, minBlockOrder, 13, ctx, dtype, channels)


class _Streaming(_Plan):
    """StreamingOverlapAddT / StreamingOverlapSaveT -- streaming_overlap_add.go:20, streaming_overlap_save.go:20."""
    _ols = 1

    def __init__(self, kernel, blockSize, ctx=None, dtype=np.float64):
        super().__init__()
        self._dtype = np.dtype(dtype).type
        self._ctx_obj = ctx or default_context()
        k = self._np(kernel)
        prec = L.F64 if self._dtype == np.float64 else L.F32
        _check(L.load().adsp_streaming_create(self._ctx_obj.handle, _p(k), k.size, int(blockSize), self._ols, prec, C.byref(self._h)))

    def BlockSize(self):
        return int(L.load().adsp_plan_block_size(self._h))

    def ProcessBlock(self, input):
        """ProcessBlock(input) -> output block (both BlockSize() samples); state persists across calls."""
        x = self._np(input)
        out = np.empty(max(self.BlockSize(), 1), dtype=self._dtype)
        _check(L.load().adsp_streaming_process_block(self._h, _p(x), x.size, _p(out), self.BlockSize()))
        return out[: self.BlockSize()]

    def ProcessBlockTo(self, output, input):
        x = self._np(input)
        if not (isinstance(output, np.ndarray) and output.dtype == self._dtype and output.flags.c_contiguous):
            raise TypeError("output must be a contiguous numpy array of the plan's dtype")
        _check(L.load().adsp_streaming_process_block(self._h, _p(x), x.size, _p(output), output.size))


class StreamingOverlapSave(_Streaming):
    _ols = 1


class StreamingOverlapAdd(_Streaming):
    _ols = 0


def NewStreamingOverlapSave(kernel, blockSize, ctx=None):
    """NewStreamingOverlapSave -- streaming_overlap_save.go:88."""
    return StreamingOverlapSave(kernel, blockSize, ctx, np.float64)


def NewStreamingOverlapSave32(kernel, blockSize, ctx=None):
    return StreamingOverlapSave(kernel, blockSize, ctx, np.float32)


def NewStreamingOverlapAdd(kernel, blockSize, ctx=None):
    """NewStreamingOverlapAdd -- streaming_overlap_add.go:88."""
    return StreamingOverlapAdd(kernel, blockSize, ctx, np.float64)


def NewStreamingOverlapAdd32(kernel, blockSize, ctx=None):
    return StreamingOverlapAdd(kernel, blockSize, ctx, np.float32)
