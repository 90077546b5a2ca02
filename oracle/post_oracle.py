"""ORACLE (test infrastructure, never imported by the product): numpy / pure-Python restatement of the reference code right
after the dsp/conv path, each function citing the Go lines it follows:
  measure/ir/ir.go            schroederIntegral :103-130, findImpulseStart :386-403, findPeak :406-424
  measure/sweep/sweep.go      LogSweep.Generate :73-94, InverseFilter :104-155, deconvolveWithInverse :182-239
  dsp/filter/fir/filter.go    ProcessSample :36-59, ProcessBlock :61-103 (both branches, as written)
  dsp/resample/resample*.go   designPolyphaseFIR :9-72, approximateRatio :74-113, Process :249-292, PredictOutputLen :295-314
Pinned by the reference's own known-answer tests where it has them (tests/test_post.py)."""
import math

import numpy as np


# ---------------------------------------------------------------- measure/ir
def schroeder_integral(ir):
    ir = np.asarray(ir, dtype=np.float64)
    n = len(ir)
    result = np.empty(n)
    cum = 0.0
    for i in range(n - 1, -1, -1):                      # ir.go:108-112
        cum += ir[i] * ir[i]
        result[i] = cum
    total = result[0]
    if total <= 0:                                      # :115-118
        return result
    for i in range(n):                                  # :120-127
        ratio = result[i] / total
        result[i] = -200.0 if ratio <= 0 else 10 * math.log10(ratio)
    return result


def find_impulse_start(ir, threshold_ratio=0.1):
    peak = 0.0
    for v in ir:                                        # ir.go:387-392
        av = abs(v)
        if av > peak:
            peak = av
    thr = peak * threshold_ratio
    for i, v in enumerate(ir):                          # :394-399
        if abs(v) >= thr:
            return i
    return 0


def find_peak(ir):
    idx, val = 0, 0.0
    for i, v in enumerate(ir):                          # ir.go:410-417
        av = abs(v)
        if av > val:
            val, idx = av, i
    return idx


# ---------------------------------------------------------------- measure/sweep
def logsweep_samples(duration, sr):
    return int(round(duration * sr))                    # sweep.go:58-60 (math.Round: half away from zero; inputs here are exact)


def logsweep_generate(f1, f2, duration, sr, dtype=np.float64):
    n = logsweep_samples(duration, sr)
    T, ln_ratio = dtype(duration), np.log(dtype(f2) / dtype(f1))
    t = np.arange(n, dtype=dtype) / dtype(sr)
    pi = dtype(np.pi) if dtype is np.float64 else np.longdouble(3.14159265358979323846264338327950288)
    phase = 2 * pi * dtype(f1) * T / ln_ratio * (np.exp(t / T * ln_ratio) - 1)     # :88-90
    return np.sin(phase).astype(np.float64)


def logsweep_inverse_filter(f1, f2, duration, sr, dtype=np.float64):
    n = logsweep_samples(duration, sr)
    sweep = logsweep_generate(f1, f2, duration, sr, dtype).astype(dtype)
    T, ln_ratio = dtype(duration), np.log(dtype(f2) / dtype(f1))
    j = n - 1 - np.arange(n)                            # :128-130
    t = j.astype(dtype) / dtype(sr)
    finst = dtype(f1) * np.exp(t / T * ln_ratio)        # :137
    inv = sweep[j] * (dtype(f1) / finst)                # :141-143
    norm = T * dtype(f1) / ln_ratio * dtype(sr)         # :148
    if norm > 0:
        inv = inv * (1.0 / norm)
    return np.asarray(inv, dtype=np.float64)


def deconvolve_with_inverse(response, inv):
    """sweep.go:182-239: zero-padded FFT product = full linear convolution (numpy's FFT stands in for algo-fft)."""
    n = len(response) + len(inv) - 1
    size = 1 << max(0, (n - 1).bit_length())
    return np.fft.irfft(np.fft.rfft(response, size) * np.fft.rfft(inv, size), size)[:n]


# ---------------------------------------------------------------- dsp/filter/fir
class Fir:
    def __init__(self, coeffs):
        self.coeffs = [float(c) for c in coeffs]
        n = len(self.coeffs)
        self.delay = [0.0] * n
        self.linear = [0.0] * (2 * n)
        self.pos = 0

    def process_sample(self, x):                        # filter.go:36-59
        n = len(self.coeffs)
        self.delay[self.pos] = x
        y = 0.0
        p = self.pos
        for k in range(n):
            y += self.coeffs[k] * self.delay[p]
            p -= 1
            if p < 0:
                p = n - 1
        self.pos += 1
        if self.pos >= n:
            self.pos = 0
        return y

    def process_block(self, buf):                       # filter.go:61-103
        n = len(self.coeffs)
        if n == 0:
            return
        if n < 32:
            for i in range(len(buf)):
                buf[i] = self.process_sample(buf[i])
            return
        for i in range(len(buf)):
            x = buf[i]
            self.linear[self.pos] = x
            self.linear[self.pos + n] = x
            self.delay[self.pos] = x
            start = self.pos + 1
            buf[i] = float(np.dot(self.coeffs, self.linear[start:start + n]))     # vecmath.DotProduct
            self.pos += 1
            if self.pos >= n:
                self.pos = 0


# ---------------------------------------------------------------- dsp/resample
PROFILES = {0: (16, 0.88, 5.0), 1: (32, 0.92, 7.5), 2: (64, 0.96, 9.0)}           # resample.go:35-44


def _gcd(a, b):
    a, b = abs(a), abs(b)
    while b:
        a, b = b, a % b
    return a or 1


def approximate_ratio(v, max_den=4096):                 # resample_design.go:74-113
    if max_den <= 0:
        max_den = 4096
    if v <= 0 or math.isnan(v) or math.isinf(v):
        return 1, 1
    a0 = math.floor(v)
    p0, q0, p1, q1, x = 1.0, 0.0, a0, 1.0, v
    while True:
        frac = x - math.floor(x)
        if frac == 0:
            break
        x = 1 / frac
        a = math.floor(x)
        p2, q2 = a * p1 + p0, a * q1 + q0
        if q2 > max_den:
            break
        p0, q0, p1, q1 = p1, q1, p2, q2
    num, den = int(round(p1)), int(round(q1))
    if den <= 0:
        return 1, 1
    g = _gcd(num, den)
    return num // g, den // g


def _i0(x):                                             # resample_design.go:161-173
    s, term = 1.0, 1.0
    x2 = (x * x) / 4
    for k in range(1, 64):
        term *= x2 / float(k * k)
        s += term
        if term < 1e-16 * s:
            break
    return s


def _kaiser(i, n, beta):                                # :150-158
    if n <= 1 or beta == 0:
        return 1.0
    t = 2 * float(i) / float(n - 1) - 1
    return _i0(beta * math.sqrt(max(0.0, 1 - t * t))) / _i0(beta)


def _sinc(x):                                           # :139-147
    if abs(x) < 1e-12:
        return 1.0
    pix = math.pi * x
    return math.sin(pix) / pix


class Resampler:
    def __init__(self, up, down, quality=1):
        if up <= 0 or down <= 0:
            raise ValueError("resample: invalid ratio")
        g = _gcd(up, down)
        self.up, self.down = up // g, down // g
        tpp, cs, kb = PROFILES[quality]
        n_taps = tpp * self.up                          # resample_design.go:22
        fc = (0.5 / float(max(self.up, self.down))) * cs
        center = 0.5 * float(n_taps - 1)
        taps = [2 * fc * _sinc(2 * fc * (float(n) - center)) * _kaiser(n, n_taps, kb) for n in range(n_taps)]
        s = 0.0
        for v in taps:
            s += v
        scale = float(self.up) / s
        self.taps = [v * scale for v in taps]
        self.phases = [self.taps[p::self.up] for p in range(self.up)]             # :54-67
        self.max_phase_len = max(len(p) for p in self.phases)
        self.phase = self.input_index = self.total_in = 0
        self.history = []

    def predict_output_len(self, input_len):            # resample.go:295-314
        if input_len <= 0:
            return 0
        last = self.total_in + input_len - 1
        i, phase, count = self.input_index, self.phase, 0
        while i <= last:
            count += 1
            phase += self.down
            i += phase // self.up
            phase %= self.up
        return count

    def process(self, x):                               # resample.go:249-292
        if len(x) == 0:
            return np.empty(0)
        work = list(self.history) + [float(v) for v in x]
        base = self.total_in - len(self.history)
        last = self.total_in + len(x) - 1
        out = []
        while self.input_index <= last:
            taps = self.phases[self.phase]
            y = 0.0
            for k, c in enumerate(taps):
                idx = self.input_index - k
                if idx < base or idx > last:
                    continue
                y += c * work[idx - base]
            out.append(y)
            self.phase += self.down
            self.input_index += self.phase // self.up
            self.phase %= self.up
        self.total_in += len(x)
        keep = min(max(0, self.max_phase_len - 1), len(work))
        self.history = work[len(work) - keep:] if keep else []
        return np.array(out)
