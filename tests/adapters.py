"""Two implementations behind one vocabulary so the same known-answer checks run against the
CPU oracle (-m "not gpu") and the CUDA product path (-m gpu).  Errors are normalised to the
reference's sentinel names."""
import numpy as np


class ImplError(Exception):
    def __init__(self, sentinel):
        self.sentinel = sentinel
        super().__init__(sentinel)


class OracleImpl:
    name = "oracle"

    def __init__(self):
        from oracle import oracle as O
        O.build()
        self.O = O

    def _wrap(self, fn, *a, **k):
        try:
            return fn(*a, **k)
        except self.O.OracleError as e:
            raise ImplError(e.sentinel) from None

    def direct(self, a, b): return self._wrap(self.O.direct, a, b)
    def direct_circular(self, a, b): return self._wrap(self.O.direct_circular, a, b)
    def convolve(self, a, b): return self._wrap(self.O.convolve, a, b)
    def convolve_mode(self, a, b, m): return self._wrap(self.O.convolve_mode, a, b, m)
    def overlap_add_convolve(self, s, k): return self._wrap(self.O.overlap_add_convolve, s, k)
    def overlap_save_convolve(self, s, k): return self._wrap(self.O.overlap_save_convolve, s, k)
    def correlate(self, a, b): return self._wrap(self.O.correlate, a, b)
    def correlate_direct(self, a, b): return self._wrap(self.O.correlate_direct, a, b)
    def correlate_fft(self, a, b): return self._wrap(self.O.correlate_fft, a, b)
    def correlate_mode(self, a, b, m): return self._wrap(self.O.correlate_mode, a, b, m)
    def correlate_normalized(self, a, b): return self._wrap(self.O.correlate_normalized, a, b)
    def auto_correlate(self, a): return self._wrap(self.O.auto_correlate, a)
    def auto_correlate_normalized(self, a): return self._wrap(self.O.auto_correlate_normalized, a)
    def find_peak(self, c): return self.O.find_peak(c)
    def lag_from_index(self, i, lb): return self.O.lag_from_index(i, lb)
    def index_from_lag(self, l, lb): return self.O.index_from_lag(l, lb)
    def next_power_of_2(self, n): return self.O.next_power_of_2(n)

    def ols_sizes(self, K, f): return self._wrap(self.O.ols_sizes, K, f)
    def ola_sizes(self, K, b): return self._wrap(self.O.ola_sizes, K, b)
    def overlap_save(self, kernel, fft_size, x): return self._wrap(self.O.overlap_save, kernel, fft_size, x)
    def overlap_add(self, kernel, block, x): return self._wrap(self.O.overlap_add, kernel, block, x)

    def overlap_save_to(self, kernel, fft_size, x, out_len):
        # ProcessTo, overlap_save.go:258-272: length check first
        if out_len != len(x) + len(kernel) - 1:
            raise ImplError("ErrLengthMismatch")
        return self.overlap_save(kernel, fft_size, x)

    def partitioned(self, kernel, mn, mx, dtype=np.float64):
        return self._wrap(self.O.Partitioned, kernel, mn, mx, dtype)

    def part_process(self, p, x, out_len=None):
        return self._wrap(p.process_block, x, out_len)

    def part_info(self, p):
        return dict(latency=p.latency(), kernel_len=p.kernel_len(), stages=[p.stage_info(i)[:2] for i in range(p.stage_count())])

    def part_stage_info(self, p, i):
        return self._wrap(p.stage_info, i)[:2]

    def part_reset(self, p): p.reset()


class CudaImpl:
    name = "cuda"

    def __init__(self):
        from algo_dsp_b200 import conv
        self.c = conv

    def _wrap(self, fn, *a, **k):
        try:
            return fn(*a, **k)
        except self.c.ConvError as e:
            raise ImplError(e.sentinel.name) from None

    def direct(self, a, b): return self._wrap(self.c.Direct, a, b)
    def direct_circular(self, a, b): return self._wrap(self.c.DirectCircular, a, b)
    def convolve(self, a, b): return self._wrap(self.c.Convolve, a, b)
    def convolve_mode(self, a, b, m): return self._wrap(self.c.ConvolveMode, a, b, m)
    def overlap_add_convolve(self, s, k): return self._wrap(self.c.OverlapAddConvolve, s, k)
    def overlap_save_convolve(self, s, k): return self._wrap(self.c.OverlapSaveConvolve, s, k)
    def correlate(self, a, b): return self._wrap(self.c.Correlate, a, b)
    def correlate_direct(self, a, b): return self._wrap(self.c.CorrelateDirect, a, b)
    def correlate_fft(self, a, b): return self._wrap(self.c.CorrelateFFT, a, b)
    def correlate_mode(self, a, b, m): return self._wrap(self.c.CorrelateMode, a, b, m)
    def correlate_normalized(self, a, b): return self._wrap(self.c.CorrelateNormalized, a, b)
    def auto_correlate(self, a): return self._wrap(self.c.AutoCorrelate, a)
    def auto_correlate_normalized(self, a): return self._wrap(self.c.AutoCorrelateNormalized, a)
    def find_peak(self, c): return self._wrap(self.c.FindPeak, c)
    def lag_from_index(self, i, lb): return self.c.LagFromIndex(i, lb)
    def index_from_lag(self, l, lb): return self.c.IndexFromLag(l, lb)
    def next_power_of_2(self, n): return self.c.nextPowerOf2(n)

    def ols_sizes(self, K, f):
        p = self._wrap(self.c.NewOverlapSave, np.ones(K), f)
        return p.FFTSize(), p.StepSize()

    def ola_sizes(self, K, b):
        p = self._wrap(self.c.NewOverlapAdd, np.ones(K), b)
        return p.BlockSize(), p.FFTSize()

    def overlap_save(self, kernel, fft_size, x):
        return self._wrap(self._wrap(self.c.NewOverlapSave, kernel, fft_size).Process, x)

    def overlap_add(self, kernel, block, x):
        return self._wrap(self._wrap(self.c.NewOverlapAdd, kernel, block).Process, x)

    def overlap_save_to(self, kernel, fft_size, x, out_len):
        p = self._wrap(self.c.NewOverlapSave, kernel, fft_size)
        out = np.zeros(out_len)
        self._wrap(p.ProcessTo, out, x)
        return out

    def partitioned(self, kernel, mn, mx, dtype=np.float64):
        return self._wrap(self.c.PartitionedConvolution, kernel, mn, mx, None, dtype)

    def part_process(self, p, x, out_len=None):
        x = np.asarray(x)
        out = np.zeros(len(x) if out_len is None else out_len, dtype=p._dtype)
        self._wrap(p.ProcessBlock, x, out)
        return out

    def part_info(self, p):
        return dict(latency=p.Latency(), kernel_len=p.KernelLen(), stages=[p.StageInfo(i) for i in range(p.StageCount())])

    def part_stage_info(self, p, i):
        return self._wrap(p.StageInfo, i)

    def part_reset(self, p): p.Reset()
