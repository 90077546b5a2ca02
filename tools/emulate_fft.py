"""numpy emulation of the index math in fft_core.cuh / conv_kernels.cuh (design check, no GPU)."""
import numpy as np

def shape(L):
    lg = L.bit_length() - 1
    P = (lg - 1) // 4
    R0 = L >> (4 * P)
    return R0, P, L // 16

def cta_fft(x, inv=False):
    """x: length-L vector. Emulates per-thread ownership e[q] <-> x[j + q*TPF]."""
    L = len(x); R0, P, TPF = shape(L); S0 = 16 // R0
    sgn = 1 if inv else -1
    e = np.array([[x[j + q * TPF] for q in range(16)] for j in range(TPF)], dtype=complex)
    # pass 0
    for j in range(TPF):
        for u in range(S0):
            v = e[j, u::S0].copy()          # e[u + r*S0]
            w = np.exp(sgn * 2j * np.pi * np.outer(np.arange(R0), np.arange(R0)) / R0)
            e[j, u::S0] = w @ v
    if P == 0:
        out = np.zeros(L, complex)
        for j in range(TPF):
            for q in range(16): out[j + q * TPF] = e[j, q]
        return out
    buf = np.zeros(L, complex)
    for j in range(TPF):
        for u in range(S0):
            b = j + u * TPF
            for r in range(R0): buf[R0 * b + r] = e[j, u + r * S0]
    ns = R0
    for t in range(1, P + 1):
        for j in range(TPF):
            for q in range(16): e[j, q] = buf[j + q * TPF]
        nb = np.zeros(L, complex)
        for j in range(TPF):
            k = j & (ns - 1)
            v = e[j].copy()
            for r in range(1, 16): v[r] *= np.exp(sgn * 2j * np.pi * r * k / (16 * ns))
            w = np.exp(sgn * 2j * np.pi * np.outer(np.arange(16), np.arange(16)) / 16)
            e[j] = w @ v
            j0 = (j - k) * 16 + k
            for r in range(16): nb[j0 + r * ns] = e[j, r]
        buf = nb
        ns *= 16
    # ownership claim: after last pass e[j][r] == X[j + r*TPF]
    out = np.zeros(L, complex)
    for j in range(TPF):
        for q in range(16): out[j + q * TPF] = e[j, q]
    return out

rng = np.random.default_rng(0)
for L in (16, 32, 64, 128, 256, 512, 1024, 2048, 4096):
    x = rng.standard_normal(L) + 1j * rng.standard_normal(L)
    f = cta_fft(x); g = cta_fft(x, True)
    print(L, shape(L), np.abs(f - np.fft.fft(x)).max(), np.abs(g - np.fft.ifft(x) * L).max())

# four-step conv check
def fourstep_conv(xa, xb, h, N1, N2):
    N = N1 * N2
    z = (xa + 1j * xb).reshape(N1, N2)           # [n1][n2]
    A = np.fft.fft(z, axis=0)                      # column FFTs -> [k1][n2]
    k1 = np.arange(N1)[:, None]; n2 = np.arange(N2)[None, :]
    A = A * np.exp(-2j * np.pi * k1 * n2 / N)
    X = np.fft.fft(A, axis=1)                      # rows -> [k1][k2] holds X[k1 + N1*k2]
    H = np.fft.fft(np.concatenate([h, np.zeros(N - len(h))]))
    Hperm = H[(k1 + N1 * np.arange(N2)[None, :])] / N
    Y = X * Hperm
    B = np.fft.ifft(Y, axis=1) * N2               # unnormalised inverse rows
    B = B * np.exp(+2j * np.pi * k1 * n2 / N)
    y = np.fft.ifft(B, axis=0) * N1               # unnormalised inverse cols -> [n1][n2]
    return y.reshape(N)

N1, N2 = 16, 256
N = N1 * N2
xa = rng.standard_normal(N); xb = rng.standard_normal(N); h = rng.standard_normal(100)
y = fourstep_conv(xa, xb, h, N1, N2)
ra = np.fft.ifft(np.fft.fft(xa) * np.fft.fft(h, N)).real
rb = np.fft.ifft(np.fft.fft(xb) * np.fft.fft(h, N)).real
print("fourstep", np.abs(y.real - ra).max(), np.abs(y.imag - rb).max())
