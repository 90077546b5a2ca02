"""dsp/signal generators of the library: host twins and device versions.

Formulas of the reference's dsp/signal/generate.go (white :188, pink :210, linear sweep :134, log sweep :157,
Normalize :253, RemoveDC :306) on a stateless hash PRNG of (seed, stream, index) -- Go's math/rand table is not in the
tree, so "identical inputs" means: one generator feeds both the oracle and the GPU.  The host functions below call the
library's adsp_gen_*_host twins, which run the very same arithmetic as the CUDA kernels (csrc/siggen_core.h) and are
bit-identical to the fp64 device output; the *_device functions fill device memory directly.

The independent numpy restatement used by the tests as the checker lives in oracle/siggen_oracle.py."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def _out(n):
    return np.empty(int(n), dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def uniform(n, seed=1, index0=0):
    """u_i in [0, 1): hash(seed, stream 0, index0 + i)."""
    out = _out(n)
    L.load().adsp_gen_uniform_host(_p(out), int(n), int(seed), int(index0))
    return out


def white(n, seed=1, amp=1.0, index0=0):
    """(u*2-1)*amp -- generate.go:199-202."""
    out = _out(n)
    L.load().adsp_gen_white_host(_p(out), int(n), float(amp), int(seed), int(index0))
    return out


def pink(n, seed=1, amp=1.0, index0=0):
    """Voss-McCartney, 5 bands -- generate.go:210-245."""
    out = _out(n)
    L.load().adsp_gen_pink_host(_p(out), int(n), float(amp), int(seed), int(index0))
    return out


def linear_sweep(n, f0=20.0, f1=20000.0, fs=48000.0, amp=1.0, index0=0, total=None):
    """sin(2*pi*(f0 t + k t^2/2)), k = (f1-f0)/T -- generate.go:134-154."""
    out = _out(n)
    L.load().adsp_gen_linear_sweep_host(_p(out), int(n), int(index0), int(total or n), float(f0), float(f1), float(amp), float(fs))
    return out


def log_sweep(n, f0=20.0, f1=20000.0, fs=48000.0, amp=1.0, index0=0, total=None):
    """sin(2*pi*f0*(exp(k t)-1)/k), k = ln(f1/f0)/T -- generate.go:157-185."""
    out = _out(n)
    L.load().adsp_gen_log_sweep_host(_p(out), int(n), int(index0), int(total or n), float(f0), float(f1), float(amp), float(fs))
    return out


def decaying_ir(K, seed=7, decades=3.0):
    """h[i] = (u_i*2-1) * 10^(-3 i / K): -60 dB at the last tap."""
    out = _out(K)
    L.load().adsp_gen_decaying_ir_host(_p(out), int(K), float(decades), int(seed))
    return out


def delay_of(row, delay_seed=0, delay_mod=4096):
    """d_p = hash(p) mod 4096 of config 4."""
    return int(L.load().adsp_gen_delay_host(int(delay_seed), int(row), int(delay_mod)))


def exp_kernel(K, r=0.99):
    """makeImpulseKernel -- partitioned_test.go:11-20."""
    return r ** np.arange(K)


def test_kernel(n):
    """makeTestKernel -- conv_bench_test.go:296-312 (Hann-windowed sinc)."""
    i = np.arange(n, dtype=np.float64)
    x = i - (n - 1) / 2
    with np.errstate(divide="ignore", invalid="ignore"):
        k = np.where(x == 0, 1.0, np.sin(np.pi * x / 4) / (np.pi * x / 4))
    return k * 0.5 * (1 - np.cos(2 * np.pi * i / (n - 1)))


test_kernel.__test__ = False   # not a pytest test


def rel_l2(y, ref):
    y = np.asarray(y, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    d = np.linalg.norm(y - ref)
    r = np.linalg.norm(ref)
    return d / r if r > 0 else d


# ---------------------------------------------------------------- device side
class DeviceArray:
    """`rows` x `n` elements of device memory owned by the library (adsp_device_alloc), row stride `stride` elements."""

    def __init__(self, ctx, rows, n, dtype=np.float64, stride=None):
        self.ctx, self.rows, self.n = ctx, int(rows), int(n)
        self.dtype = np.dtype(dtype)
        self.stride = int(stride) if stride else (self.n + 31) // 32 * 32     # rows start 256-byte aligned (fp64)
        self.prec = L.F64 if self.dtype == np.float64 else L.F32
        ptr = C.c_void_p()
        st = L.load().adsp_device_alloc(ctx.handle, self.rows * self.stride * self.dtype.itemsize, C.byref(ptr))
        if st != L.OK:
            raise MemoryError(L.last_error())
        self.ptr = ptr.value

    def row_ptr(self, r, offset=0):
        return self.ptr + (int(r) * self.stride + int(offset)) * self.dtype.itemsize

    def get(self, r0=0, r1=None, c0=0, c1=None):
        """rows [r0, r1) x columns [c0, c1) as a numpy array (synchronous)."""
        r1 = self.rows if r1 is None else r1
        c1 = self.n if c1 is None else c1
        out = np.empty((r1 - r0, c1 - c0), dtype=self.dtype)
        lib = L.load()
        for r in range(r0, r1):
            row = out[r - r0]
            if lib.adsp_memcpy_d2h(self.ctx.handle, row.ctypes.data_as(C.c_void_p), C.c_void_p(self.row_ptr(r, c0)), row.nbytes) != L.OK:
                raise RuntimeError(L.last_error())
        return out

    def free(self):
        if self.ptr:
            L.load().adsp_device_free(self.ctx.handle, C.c_void_p(self.ptr))
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _chk(st):
    if st != L.OK:
        raise RuntimeError(f"generator failed ({st}): {L.last_error()}")


def white_device(ctx, ptr, n, rows=1, stride=None, amp=1.0, seed0=1, seed_step=1, index0=0, prec=L.F64):
    _chk(L.load().adsp_gen_white_device(ctx.handle, C.c_void_p(ptr), int(n), int(rows), int(stride or n), float(amp), int(seed0), int(seed_step),
                                        int(index0), prec))


def uniform_device(ctx, ptr, n, rows=1, stride=None, seed0=1, seed_step=1, index0=0, prec=L.F64):
    _chk(L.load().adsp_gen_uniform_device(ctx.handle, C.c_void_p(ptr), int(n), int(rows), int(stride or n), int(seed0), int(seed_step), int(index0), prec))


def pink_device(ctx, ptr, n, rows=1, stride=None, amp=1.0, seed0=1, seed_step=1, index0=0, prec=L.F64):
    _chk(L.load().adsp_gen_pink_device(ctx.handle, C.c_void_p(ptr), int(n), int(rows), int(stride or n), float(amp), int(seed0), int(seed_step),
                                       int(index0), prec))


def decaying_ir_device(ctx, ptr, K, rows=1, stride=None, decades=3.0, seed0=7, seed_step=1, prec=L.F64):
    _chk(L.load().adsp_gen_decaying_ir_device(ctx.handle, C.c_void_p(ptr), int(K), int(rows), int(stride or K), float(decades), int(seed0),
                                              int(seed_step), prec))


def linear_sweep_device(ctx, ptr, n, f0=20.0, f1=20000.0, fs=48000.0, amp=1.0, index0=0, total=None, prec=L.F64):
    _chk(L.load().adsp_gen_linear_sweep_device(ctx.handle, C.c_void_p(ptr), int(n), int(index0), int(total or n), float(f0), float(f1), float(amp),
                                               float(fs), prec))


def log_sweep_device(ctx, ptr, n, f0=20.0, f1=20000.0, fs=48000.0, amp=1.0, index0=0, total=None, prec=L.F64):
    _chk(L.load().adsp_gen_log_sweep_device(ctx.handle, C.c_void_p(ptr), int(n), int(index0), int(total or n), float(f0), float(f1), float(amp),
                                            float(fs), prec))


def delay_mix_device(ctx, ptr, n, rows, stride, src_ptr, noise_amp=0.01, seed0=1000, seed_step=1, delay_seed=0, delay_mod=4096, delays_ptr=0,
                     prec=L.F64):
    _chk(L.load().adsp_gen_delay_mix_device(ctx.handle, C.c_void_p(ptr), int(n), int(rows), int(stride), C.c_void_p(src_ptr), float(noise_amp),
                                            int(seed0), int(seed_step), int(delay_seed), int(delay_mod), C.c_void_p(delays_ptr), prec))


def normalize_device(ctx, in_ptr, n, rows, in_stride, target_peak, out_ptr, out_stride, prec=L.F64):
    _chk(L.load().adsp_normalize_device(ctx.handle, C.c_void_p(in_ptr), int(n), int(rows), int(in_stride), float(target_peak), C.c_void_p(out_ptr),
                                        int(out_stride), prec))


def remove_dc_device(ctx, in_ptr, n, rows, in_stride, out_ptr, out_stride, prec=L.F64):
    _chk(L.load().adsp_remove_dc_device(ctx.handle, C.c_void_p(in_ptr), int(n), int(rows), int(in_stride), C.c_void_p(out_ptr), int(out_stride), prec))
