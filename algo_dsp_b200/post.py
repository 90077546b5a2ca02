"""Python mirrors of the reference packages right after the dsp/conv path, on top of the C ABI
(csrc/post.cu): measure/ir (SchroederIntegral, FindImpulseStart), measure/sweep (LogSweep), dsp/filter/fir (Filter),
dsp/resample (Resampler).  Same names and argument meaning as the Go API; all arithmetic runs on the GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .conv import ConvError, _check, _ctx, _f64, _p, default_context


# ---------------------------------------------------------------- measure/ir
class Analyzer:
    """ir.Analyzer (measure/ir/ir.go): the parts that consume a deconvolved impulse response."""

    def __init__(self, sampleRate=48000.0, ctx=None):
        self.SampleRate = float(sampleRate)
        self._ctx = ctx

    def SchroederIntegral(self, ir):
        """SchroederIntegral -- ir.go:94: backward-integrated energy decay in dB."""
        x = _f64(ir)
        if x.size == 0:
            raise ConvError(L.ERR_EMPTY_IR)
        out = np.empty(x.size)
        _check(L.load().adsp_ir_schroeder(_ctx(self._ctx), _p(x), x.size, _p(out)))
        return out

    def FindImpulseStart(self, ir, thresholdRatio=0.1):
        """FindImpulseStart -- ir.go:381: first sample at or above -20 dB re peak."""
        x = _f64(ir)
        if x.size == 0:
            raise ConvError(L.ERR_EMPTY_IR)
        idx = C.c_int64()
        _check(L.load().adsp_ir_find_impulse_start(_ctx(self._ctx), _p(x), x.size, float(thresholdRatio), C.byref(idx)))
        return int(idx.value)

    def findPeak(self, ir):
        """findPeak -- ir.go:406: index of the absolute maximum (first one wins)."""
        return self.FindImpulseStart(ir, 1.0)


# ---------------------------------------------------------------- measure/sweep
class LogSweep:
    """sweep.LogSweep -- measure/sweep/sweep.go:28."""

    def __init__(self, StartFreq, EndFreq, Duration, SampleRate, ctx=None):
        self.StartFreq, self.EndFreq, self.Duration, self.SampleRate = float(StartFreq), float(EndFreq), float(Duration), float(SampleRate)
        self._ctx = ctx

    def _args(self):
        return self.StartFreq, self.EndFreq, self.Duration, self.SampleRate

    def samples(self):
        return int(L.load().adsp_logsweep_samples(self.Duration, self.SampleRate))

    def Generate(self):
        """Generate -- sweep.go:73 (host twin of the device generator, bit-identical to it)."""
        out = np.empty(max(self.samples(), 1))
        _check(L.load().adsp_logsweep_generate_host(_p(out), *self._args()))
        return out[: self.samples()]

    def InverseFilter(self):
        """InverseFilter -- sweep.go:104."""
        out = np.empty(max(self.samples(), 1))
        _check(L.load().adsp_logsweep_inverse_filter_host(_p(out), *self._args()))
        return out[: self.samples()]

    def Deconvolve(self, response):
        """Deconvolve -- sweep.go:164: response (*) inverse filter, len(response) + samples - 1 values."""
        x = _f64(response)
        if x.size == 0:
            raise ConvError(L.ERR_EMPTY_INPUT)                    # sweep.ErrEmptyResponse
        n_out = x.size + self.samples() - 1
        out = np.empty(max(n_out, 1))
        _check(L.load().adsp_logsweep_deconvolve(_ctx(self._ctx), _p(x), x.size, *self._args(), _p(out), n_out))
        return out[:n_out]


# ---------------------------------------------------------------- dsp/filter/fir
class Filter:
    """fir.Filter -- dsp/filter/fir/filter.go:11: stateful block FIR, `channels` rows filtered independently."""

    def __init__(self, coeffs, ctx=None, channels=1):
        self._h = C.c_void_p()
        self._ctx_obj = ctx or default_context()
        c = _f64(coeffs)
        self._n = c.size
        self._channels = int(channels)
        _check(L.load().adsp_fir_create(self._ctx_obj.handle, _p(c), c.size, self._channels, C.byref(self._h)))

    def Order(self):
        return int(L.load().adsp_fir_order(self._h))

    def ProcessBlock(self, buf):
        """ProcessBlock(buf) -- filter.go:61: in place; buf is [n] (one channel) or [channels, n]."""
        if not (isinstance(buf, np.ndarray) and buf.dtype == np.float64 and buf.flags.c_contiguous):
            raise TypeError("buf must be a contiguous float64 numpy array")
        n = buf.shape[-1]
        _check(L.load().adsp_fir_process_block(self._h, _p(buf), n, n))

    def process_block_device(self, ptr, n, stride):
        _check(L.load().adsp_fir_process_block_device(self._h, C.c_void_p(ptr), int(n), int(stride)))

    def Reset(self):
        L.load().adsp_fir_reset(self._h)

    def Close(self):
        if self._h:
            L.load().adsp_fir_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.Close()
        except Exception:
            pass


def New(coeffs, ctx=None, channels=1):
    """fir.New(coeffs) -- filter.go:18."""
    return Filter(coeffs, ctx, channels)


# ---------------------------------------------------------------- dsp/resample
QualityFast, QualityBalanced, QualityBest = 0, 1, 2


def approximateRatio(v, maxDen=4096):
    """approximateRatio -- resample_design.go:74."""
    num, den = C.c_int(), C.c_int()
    L.load().adsp_resample_approximate_ratio(float(v), int(maxDen), C.byref(num), C.byref(den))
    return num.value, den.value


class Resampler:
    """resample.Resampler -- dsp/resample/resample.go:138."""

    def __init__(self, handle, ctx_obj, channels):
        self._h, self._ctx_obj, self._channels = handle, ctx_obj, channels

    def Ratio(self):
        up, down = C.c_int(), C.c_int()
        L.load().adsp_resampler_ratio(self._h, C.byref(up), C.byref(down))
        return up.value, down.value

    def TapsPerPhase(self):
        return int(L.load().adsp_resampler_taps_per_phase(self._h))

    def Prototype(self):
        n = int(L.load().adsp_resampler_prototype(self._h, None, 0))
        out = np.empty(n)
        L.load().adsp_resampler_prototype(self._h, _p(out), n)
        return out

    def PredictOutputLen(self, inputLen):
        return int(L.load().adsp_resampler_predict_output_len(self._h, int(inputLen)))

    def Process(self, input):
        """Process(input) -- resample.go:249; [n] for one channel or [channels, n]."""
        x = np.ascontiguousarray(input, dtype=np.float64)
        if x.size == 0:
            return np.empty(0)
        one = x.ndim == 1
        x2 = x.reshape(1, -1) if one else x
        n = x2.shape[1]
        nout = self.PredictOutputLen(n)
        out = np.empty((x2.shape[0], max(nout, 1)))
        got = C.c_int64()
        _check(L.load().adsp_resampler_process(self._h, _p(x2), n, n, _p(out), out.shape[1], out.shape[1], C.byref(got)))
        out = out[:, : got.value]
        return out[0] if one else out

    def process_device(self, in_ptr, n, in_stride, out_ptr, out_cap, out_stride):
        """Device rows in, device rows out (adsp_resampler_process_device); returns the output samples per channel."""
        got = C.c_int64()
        _check(L.load().adsp_resampler_process_device(self._h, C.c_void_p(in_ptr), int(n), int(in_stride), C.c_void_p(out_ptr), int(out_cap), int(out_stride),
                                                      C.byref(got)))
        return got.value

    def Reset(self):
        L.load().adsp_resampler_reset(self._h)

    def Close(self):
        if self._h:
            L.load().adsp_resampler_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.Close()
        except Exception:
            pass


def NewRational(up, down, quality=QualityBalanced, tapsPerPhase=0, cutoffScale=0.0, kaiserBeta=0.0, ctx=None, channels=1):
    """NewRational(up, down, opts...) -- resample.go:153."""
    ctx_obj = ctx or default_context()
    h = C.c_void_p()
    _check(L.load().adsp_resampler_create(ctx_obj.handle, int(up), int(down), int(quality), int(tapsPerPhase), float(cutoffScale), float(kaiserBeta),
                                          int(channels), C.byref(h)))
    return Resampler(h, ctx_obj, channels)


def NewForRates(inRate, outRate, quality=QualityBalanced, maxDen=4096, ctx=None, channels=1):
    """NewForRates(inRate, outRate, opts...) -- resample.go:194."""
    ctx_obj = ctx or default_context()
    h = C.c_void_p()
    _check(L.load().adsp_resampler_create_for_rates(ctx_obj.handle, float(inRate), float(outRate), int(quality), int(maxDen), int(channels), C.byref(h)))
    return Resampler(h, ctx_obj, channels)


def Resample(input, up, down, **kw):
    """Resample -- resample.go:232 (one-shot helper)."""
    r = NewRational(up, down, **kw)
    try:
        return r.Process(input)
    finally:
        r.Close()
