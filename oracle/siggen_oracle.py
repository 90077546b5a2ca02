"""ORACLE (test infrastructure, never imported by the product): numpy restatement of the reference's dsp/signal generators
(/root/reference/dsp/signal/generate.go) with the uniform stream replaced by the library's documented stateless hash
(include/algodsp_cuda.h, "dsp/signal generators"): u(seed, stream, index) = top 53 bits of
mix64(key + (index+1)*0x9E3779B97F4A7C15), key = mix64(mix64(seed + 0x9E37...) ^ (stream*0xD1B5... + 0x2545...)), mix64 = the
splitmix64 finaliser.  Go's math/rand lagged-Fibonacci table is not in the tree, so the reference's own KAT for
the generators (dsp/signal/example_test.go:29-46) cannot be reproduced: the uniform stream is pinned by the known-answer
vector in tests/test_siggen.py instead, the formulas on top of it are the reference's.

Sweeps use numpy's libm (and a long-double variant as ground truth): agreement with the library's deterministic exp/sin is
limited by the conditioning of the formula itself -- the phase reaches ~4e5 rad at 2^20 samples, so one rounding of it is
already ~5e-11 -- not by either implementation."""
import numpy as np

M64 = (1 << 64) - 1
GOLD = 0x9E3779B97F4A7C15


def _mix64_int(z):
    z &= M64
    z ^= z >> 30
    z = (z * 0xBF58476D1CE4E5B9) & M64
    z ^= z >> 27
    z = (z * 0x94D049BB133111EB) & M64
    z ^= z >> 31
    return z


def hash_key(seed, stream=0):
    return _mix64_int(_mix64_int((seed + GOLD) & M64) ^ ((stream * 0xD1B54A32D192ED03 + 0x2545F4914F6CDD1D) & M64))


def _mix64(z):
    with np.errstate(over="ignore"):
        z = z ^ (z >> np.uint64(30))
        z = z * np.uint64(0xBF58476D1CE4E5B9)
        z = z ^ (z >> np.uint64(27))
        z = z * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def hash_u64(key, index):
    idx = np.asarray(index, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return _mix64(np.uint64(key) + (idx + np.uint64(1)) * np.uint64(GOLD))


def uniform(n, seed=1, index0=0, stream=0):
    """u_i in [0,1), a multiple of 2^-53 (rng.Float64() of the reference has the same range)."""
    i = np.arange(index0, index0 + n, dtype=np.uint64)
    return (hash_u64(hash_key(seed, stream), i) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53


def white(n, seed=1, amp=1.0, index0=0):
    """WhiteNoise generate.go:199-202: out[i] = (rng.Float64()*2 - 1) * amplitude."""
    return (uniform(n, seed, index0) * 2.0 - 1.0) * amp


PA = np.array([0.23980, 0.18727, 0.16380, 0.194685, 0.214463])      # generate.go:220
PSUM = np.array([0.00198, 0.01478, 0.06378, 0.23378, 0.91578])      # generate.go:221


def pink(n, seed=1, amp=1.0):
    """PinkNoise generate.go:210-250, the loop as written (two uniforms per sample, at most one band rewritten)."""
    key = hash_key(seed, 0)
    i = np.arange(n, dtype=np.uint64)
    ur1 = (hash_u64(key, i * np.uint64(2)) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    ur2 = (hash_u64(key, i * np.uint64(2) + np.uint64(1)) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    contributions = [0.0] * 5
    out = np.empty(n)
    for s in range(n):
        val = ur2[s] * 2 - 1
        for b in range(5):
            if ur1[s] <= PSUM[b]:
                contributions[b] = val * PA[b]
                break
        total = 0.0
        for c in contributions:
            total += c
        out[s] = total * amp
    return out


def linear_sweep(n, f0=20.0, f1=20000.0, fs=48000.0, amp=1.0, dtype=np.float64):
    """LinearSweep generate.go:144-151."""
    T = dtype(n) / dtype(fs)
    k = (dtype(f1) - dtype(f0)) / T
    t = np.arange(n, dtype=dtype) / dtype(fs)
    return (dtype(amp) * np.sin(2 * dtype(np.pi if dtype is np.float64 else np.longdouble(3.14159265358979323846264338327950288)) * (dtype(f0) * t + dtype(0.5) * k * t * t))).astype(np.float64)


def log_sweep(n, f0=20.0, f1=20000.0, fs=48000.0, amp=1.0, dtype=np.float64):
    """LogSweep generate.go:170-181."""
    T = dtype(n) / dtype(fs)
    k = np.log(dtype(f1) / dtype(f0)) / T
    t = np.arange(n, dtype=dtype) / dtype(fs)
    pi = dtype(np.pi) if dtype is np.float64 else np.longdouble(3.14159265358979323846264338327950288)
    return (dtype(amp) * np.sin(2 * pi * dtype(f0) * ((np.exp(k * t) - 1) / k))).astype(np.float64)


def decaying_ir(K, seed=7, decades=3.0):
    """h[i] = (u_i*2-1) * 10^(-3 i/K)."""
    return (uniform(K, seed) * 2.0 - 1.0) * 10.0 ** (-decades * np.arange(K) / K)


def delay_of(row, delay_seed=0, delay_mod=4096):
    return int(hash_u64(hash_key(delay_seed, 1), np.uint64(row))) % delay_mod


def normalize(data, target_peak):
    """Normalize generate.go:253-283."""
    data = np.asarray(data, dtype=np.float64)
    m = 0.0
    for v in np.abs(data):
        if v > m:
            m = v
    if m == 0 or target_peak == 0:
        return np.zeros_like(data)
    return data * (target_peak / m)


def remove_dc(data):
    """RemoveDC generate.go:306-324 (sequential sum)."""
    data = np.asarray(data, dtype=np.float64)
    s = 0.0
    for v in data:
        s += v
    return data - s / len(data)
