"""Shared-memory bank-conflict simulator for the Stockham exchange layouts (design aid).
Element = complex (16 B fp64 / 8 B fp32).  phys(i) = i ^ ((i >> s) & mask)."""
import itertools, sys

def conflicts(addrs_bytes, esize):
    # phase = lanes per wavefront
    lanes = 128 // esize
    worst = 0
    for p in range(0, 32, lanes):
        banks = {}
        for a in addrs_bytes[p:p+lanes]:
            for w in range(esize // 4):
                banks.setdefault(((a // 4) + w) % 32, set()).add(a)
        worst = max(worst, max(len(v) for v in banks.values()))
    return worst

def swz(i, s, mask):
    return i ^ ((i >> s) & mask) if s >= 0 else i

def row_store_idx(j, u, r, R, NS, TPR):
    b = j + u * TPR
    k = b % NS
    return (b - k) * R + k + r * NS

def check(L, radices, esize):
    TPR = L // 16
    mask = (128 // esize) - 1
    NS = 1
    res = []
    for p, R in enumerate(radices[:-1]):
        S = 16 // R
        best = None
        for s in [-1] + list(range(2, 9)):
            w = 0
            for warp0 in range(0, max(TPR, 32), 32):
                lanes = [warp0 + l for l in range(32)]
                js = [l % TPR for l in lanes]   # (rows beyond handled identically)
                for u in range(S):
                    for r in range(R):
                        a = [swz(row_store_idx(j, u, r, R, NS, TPR), s, mask) * esize for j in js]
                        w = max(w, conflicts(a, esize))
                for q in range(16):
                    a = [swz(j + q * TPR, s, mask) * esize for j in js]
                    w = max(w, conflicts(a, esize))
            if best is None or w < best[1]:
                best = (s, w)
        res.append(best)
        NS *= R
    return res

if __name__ == "__main__":
    for esize in (16, 8):
        for L, rad in [(256, (16, 16)), (512, (2, 16, 16)), (1024, (4, 16, 16)), (2048, (8, 16, 16)), (4096, (16, 16, 16)), (8192,(2,16,16,16))]:
            print(esize, L, rad, check(L, rad, esize))

def all_ok(L, radices, esize):
    TPR = L // 16
    mask = (128 // esize) - 1
    NS = 1
    out = []
    for p, R in enumerate(radices[:-1]):
        S = 16 // R
        oks = []
        for s in [-1] + list(range(2, 9)):
            w = 0
            for warp0 in range(0, max(TPR, 32), 32):
                js = [(warp0 + l) % TPR for l in range(32)]
                for u in range(S):
                    for r in range(R):
                        w = max(w, conflicts([swz(row_store_idx(j, u, r, R, NS, TPR), s, mask) * esize for j in js], esize))
                for q in range(16):
                    w = max(w, conflicts([swz(j + q * TPR, s, mask) * esize for j in js], esize))
            if w == 1: oks.append(s)
        out.append(oks)
        NS *= R
    return out
