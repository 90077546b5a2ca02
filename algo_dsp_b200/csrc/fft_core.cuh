// fft_core.cuh -- register/shared-memory Stockham FFT building blocks for sm_100a.
//
// Replaces the third-party CPU FFT the reference calls (algo-fft Plan.Forward/Inverse,
// call sites dsp/conv/overlap_save.go:166,177; overlap_add.go:138,149; correlate.go:143-164;
// partitioned.go:145,154,170).  Contract kept: forward unnormalised, inverse 1/N (the 1/N is
// folded into the cached IR spectrum by the callers).
//
// Design: every thread owns E=16 complex points of one transform in registers.  A transform of
// length L = R0 * 16^P (R0 in {2,4,8,16}) is P+1 passes: one twiddle-free radix-R0 pass, then P
// radix-16 passes, exchanging through shared memory between passes (Stockham autosort, so the
// result is in natural order and thread j again owns points j + q*L/16 -- the same ownership as
// on entry, which lets forward -> spectral multiply -> inverse chain in registers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace adsp {

// Timing-attribution switches (development only, results are WRONG when non-zero): skip a memory
// phase of a tile to see what hiding it would be worth.  bit0 rows load, bit1 rows store, bit2 rows
// H prefetch+multiply, bit3 cols_fwd load, bit4 cols_fwd store, bit5 cols_inv load, bit6 cols_inv store,
// bit7 all butterfly/twiddle arithmetic, bit8 all shared-memory exchange traffic (barriers kept).
#ifdef ADSP_PHASE_DEBUG
static __constant__ int g_phase_skip = 0;
#define ADSP_SKIP(bit) ((g_phase_skip >> (bit)) & 1)
// steady-state timing aid: map every block pair's scratch onto the first g_scratch_alias slots (results WRONG)
static __constant__ int g_scratch_alias = 0;
#define ADSP_ALIAS(pl) (g_scratch_alias > 0 ? (pl) % g_scratch_alias : (pl))
#else
#define ADSP_SKIP(bit) 0
#define ADSP_ALIAS(pl) (pl)
#endif

// 1: form the 15 twiddle powers of EVERY twiddled pass from the first power (14 complex products) instead of
// reading all 15 from the shared-memory table in the short passes: trades 15 LDS.128 for 56 FP64 instructions
#ifndef ADSP_TW_CHAIN
#define ADSP_TW_CHAIN 0
#endif
#ifndef ADSP_TW_TREE_ALL
#define ADSP_TW_TREE_ALL 0
#endif

template <typename T> struct cpx_of;
template <> struct cpx_of<double> { using type = double2; };
template <> struct cpx_of<float> { using type = float2; };
template <typename T> using cpx = typename cpx_of<T>::type;

template <typename C> __device__ __forceinline__ C cadd(C a, C b) { C r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { C r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
    C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r;
}
// a * w (forward) or a * conj(w) (inverse)
template <bool INV, typename C> __device__ __forceinline__ C cmul_tw(C a, C w) {
    C r;
    if (INV) { r.x = a.x * w.x + a.y * w.y; r.y = a.y * w.x - a.x * w.y; }
    else     { r.x = a.x * w.x - a.y * w.y; r.y = a.x * w.y + a.y * w.x; }
    return r;
}
// multiply by -i (forward) / +i (inverse)
template <bool INV, typename C> __device__ __forceinline__ C mul_mi(C a) {
    C r;
    if (INV) { r.x = -a.y; r.y = a.x; }
    else     { r.x = a.y;  r.y = -a.x; }
    return r;
}

// ---------------------------------------------------------------- small DFTs in registers
template <bool INV, typename C> __device__ __forceinline__ void dft2(C &a0, C &a1) {
    C t = csub(a0, a1); a0 = cadd(a0, a1); a1 = t;
}
template <bool INV, typename C> __device__ __forceinline__ void dft4(C &a0, C &a1, C &a2, C &a3) {
    C t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_mi<INV>(csub(a1, a3));
    a0 = cadd(t0, t2); a1 = cadd(t1, t3); a2 = csub(t0, t2); a3 = csub(t1, t3);
}

template <typename C> struct real_of;
template <> struct real_of<double2> { using type = double; };
template <> struct real_of<float2> { using type = float; };

template <bool INV, typename C> __device__ __forceinline__ C mkc(double re, double im_fwd) {
    using T = typename real_of<C>::type;
    C r; r.x = (T)re; r.y = (T)(INV ? -im_fwd : im_fwd); return r;
}

// v: R elements at stride S (compile time), natural order in and out.
template <int R, int S, bool INV, typename C> struct Dft;

template <int S, bool INV, typename C> struct Dft<2, S, INV, C> {
    static __device__ __forceinline__ void run(C *v) { dft2<INV>(v[0], v[S]); }
};
template <int S, bool INV, typename C> struct Dft<4, S, INV, C> {
    static __device__ __forceinline__ void run(C *v) { dft4<INV>(v[0], v[S], v[2 * S], v[3 * S]); }
};
template <int S, bool INV, typename C> struct Dft<8, S, INV, C> {
    static __device__ __forceinline__ void run(C *v) {
        constexpr double h = 0.70710678118654752440;
        C y0[4] = { v[0], v[2 * S], v[4 * S], v[6 * S] };
        C y1[4] = { v[S], v[3 * S], v[5 * S], v[7 * S] };
        dft4<INV>(y0[0], y0[1], y0[2], y0[3]);
        dft4<INV>(y1[0], y1[1], y1[2], y1[3]);
        y1[1] = cmul(y1[1], mkc<INV, C>(h, -h));
        y1[2] = mul_mi<INV>(y1[2]);
        y1[3] = cmul(y1[3], mkc<INV, C>(-h, -h));
#pragma unroll
        for (int k = 0; k < 4; k++) {
            v[k * S] = cadd(y0[k], y1[k]);
            v[(k + 4) * S] = csub(y0[k], y1[k]);
        }
    }
};
template <int S, bool INV, typename C> struct Dft<16, S, INV, C> {
    static __device__ __forceinline__ void run(C *v) {
        constexpr double h = 0.70710678118654752440;
        constexpr double c1 = 0.92387953251128675613;  // cos(pi/8)
        constexpr double s1 = 0.38268343236508977173;  // sin(pi/8)
        C y[4][4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            y[q][0] = v[q * S]; y[q][1] = v[(q + 4) * S]; y[q][2] = v[(q + 8) * S]; y[q][3] = v[(q + 12) * S];
            dft4<INV>(y[q][0], y[q][1], y[q][2], y[q][3]);
        }
        // y[q][k] *= W16^(q*k)
        y[1][1] = cmul(y[1][1], mkc<INV, C>(c1, -s1));
        y[1][2] = cmul(y[1][2], mkc<INV, C>(h, -h));
        y[1][3] = cmul(y[1][3], mkc<INV, C>(s1, -c1));
        y[2][1] = cmul(y[2][1], mkc<INV, C>(h, -h));
        y[2][2] = mul_mi<INV>(y[2][2]);
        y[2][3] = cmul(y[2][3], mkc<INV, C>(-h, -h));
        y[3][1] = cmul(y[3][1], mkc<INV, C>(s1, -c1));
        y[3][2] = cmul(y[3][2], mkc<INV, C>(-h, -h));
        y[3][3] = cmul(y[3][3], mkc<INV, C>(-c1, s1));
#pragma unroll
        for (int k = 0; k < 4; k++) {
            dft4<INV>(y[0][k], y[1][k], y[2][k], y[3][k]);
            v[k * S] = y[0][k]; v[(k + 4) * S] = y[1][k]; v[(k + 8) * S] = y[2][k]; v[(k + 12) * S] = y[3][k];
        }
    }
};

// ---------------------------------------------------------------- small odd DFTs (mixed-radix column passes)
// cos / sin of 2*pi*m/P for m = 0..(P-1)/2 (long-double values rounded once); ternary chains fold to
// immediates once the caller's loops are unrolled.
template <int P> struct OddRoots;
template <> struct OddRoots<3> {
    static __host__ __device__ constexpr double c(int m) { return m == 0 ? 1.0 : -0.5; }
    static __host__ __device__ constexpr double s(int m) { return m == 0 ? 0.0 : 0.8660254037844386; }
};
template <> struct OddRoots<5> {
    static __host__ __device__ constexpr double c(int m) { return m == 0 ? 1.0 : m == 1 ? 0.30901699437494745 : -0.8090169943749475; }
    static __host__ __device__ constexpr double s(int m) { return m == 0 ? 0.0 : m == 1 ? 0.9510565162951535 : 0.5877852522924731; }
};
template <> struct OddRoots<7> {
    static __host__ __device__ constexpr double c(int m) {
        return m == 0 ? 1.0 : m == 1 ? 0.6234898018587335 : m == 2 ? -0.2225209339563144 : -0.9009688679024191;
    }
    static __host__ __device__ constexpr double s(int m) {
        return m == 0 ? 0.0 : m == 1 ? 0.7818314824680298 : m == 2 ? 0.9749279121818236 : 0.4338837391175581;
    }
};
template <> struct OddRoots<9> {
    static __host__ __device__ constexpr double c(int m) {
        return m == 0 ? 1.0 : m == 1 ? 0.766044443118978 : m == 2 ? 0.17364817766693036 : m == 3 ? -0.5 : -0.9396926207859084;
    }
    static __host__ __device__ constexpr double s(int m) {
        return m == 0 ? 0.0 : m == 1 ? 0.6427876096865394 : m == 2 ? 0.984807753012208 : m == 3 ? 0.8660254037844386 : 0.3420201433256687;
    }
};

// In-place P-point DFT (P odd), natural order in and out: X[k] = x0 + sum_m (a_m cos(t k m) -/+ i b_m sin(t k m)),
// a_m = x[m] + x[P-m], b_m = x[m] - x[P-m], t = 2*pi/P; X[P-k] uses the opposite sign.
template <int P, bool INV, typename C> __device__ __forceinline__ void odd_dft(C (&e)[P]) {
    using T = typename real_of<C>::type;
    constexpr int H = (P - 1) / 2;
    C a[H + 1], b[H + 1];
#pragma unroll
    for (int m = 1; m <= H; m++) { a[m] = cadd(e[m], e[P - m]); b[m] = csub(e[m], e[P - m]); }
    const C x0 = e[0];
    C sum = x0;
#pragma unroll
    for (int m = 1; m <= H; m++) sum = cadd(sum, a[m]);
    e[0] = sum;
#pragma unroll
    for (int k = 1; k <= H; k++) {
        C R = x0, I;
        I.x = (T)0; I.y = (T)0;
#pragma unroll
        for (int m = 1; m <= H; m++) {
            const int r0 = (k * m) % P;
            const int r = r0 > H ? P - r0 : r0;
            const T cs = (T)OddRoots<P>::c(r);
            const T sn = (T)(r0 > H ? -OddRoots<P>::s(r) : OddRoots<P>::s(r));
            R.x += a[m].x * cs; R.y += a[m].y * cs;
            I.x += b[m].x * sn; I.y += b[m].y * sn;
        }
        // forward: X[k] = R - i*I, X[P-k] = R + i*I ; inverse: the other way round
        C lo, hi;
        lo.x = R.x + I.y; lo.y = R.y - I.x;     // R - i*I
        hi.x = R.x - I.y; hi.y = R.y + I.x;     // R + i*I
        e[k] = INV ? hi : lo;
        e[P - k] = INV ? lo : hi;
    }
}

// In-place M-point DFT for the mixed-radix column passes: M = P (odd) or M = 2P, the latter by the prime-factor
// (Good-Thomas) map i = (P*i1 + 2*i2) mod 2P, k = (P*k1 + (P+1)*k2) mod 2P -- two P-point DFTs and P two-point
// butterflies, no twiddles between them (2 and P are coprime).
template <int M, bool INV, typename C> __device__ __forceinline__ void small_dft(C (&e)[M]) {
    if constexpr (M % 2 == 1) {
        odd_dft<M, INV>(e);
    } else {
        constexpr int P = M / 2;
        C g0[P], g1[P];
#pragma unroll
        for (int i2 = 0; i2 < P; i2++) { g0[i2] = e[(2 * i2) % M]; g1[i2] = e[(P + 2 * i2) % M]; }
        odd_dft<P, INV>(g0);
        odd_dft<P, INV>(g1);
#pragma unroll
        for (int k2 = 0; k2 < P; k2++) {
            e[((P + 1) * k2) % M] = cadd(g0[k2], g1[k2]);
            e[(P + (P + 1) * k2) % M] = csub(g0[k2], g1[k2]);
        }
    }
}

// ---------------------------------------------------------------- transform geometry
constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }

// Twiddle storage for a twiddled radix-16 pass with NS sub-transform length:
//   NS <= 16 : all 15 powers, [15][NS]  (direct shared-memory lookups)
//   NS  > 16 : first power only, [NS]; powers 2..15 are formed in registers by a short product
//              tree (depth <= 6).  Keeps the per-CTA table small enough for shared memory so no
//              twiddle ever costs a global-memory round trip.
constexpr int tw_pass_entries(int ns) { return ns <= 16 ? 15 * ns : ns; }
constexpr int tw_total_entries(int r0, int p) {
    int n = 0, ns = r0;
    for (int t = 0; t < p; t++) { n += tw_pass_entries(ns); ns *= 16; }
    return n;
}

template <int L> struct FftShape {
    static_assert(L >= 16 && (L & (L - 1)) == 0, "L must be a power of two >= 16");
    static constexpr int LG = ilog2(L);
    static constexpr int P = (LG - 1) / 4;        // number of twiddled radix-16 passes
    static constexpr int R0 = L >> (4 * P);       // first (twiddle-free) radix: 2,4,8,16
    static constexpr int TPF = L / 16;            // threads per transform
    static constexpr int TW_ENTRIES = tw_total_entries(R0, P);  // compact table size (elements)
};

// Address policies ---------------------------------------------------------------------------
// Row layout: one transform contiguous in shared memory; per-exchange XOR swizzle on the
// 16-byte (fp64) / 8-byte (fp32) element index keeps every 128-bit access conflict free
// (validated by tools/bank_sim.py).
template <typename T, int R0> struct RowAddr {
    static constexpr int MASK = (sizeof(T) == 8) ? 7 : 15;
    static constexpr int S1 = (R0 == 16) ? 4 : (R0 == 8) ? 3 : (R0 == 4) ? 2 : ((sizeof(T) == 8) ? 3 : 2);
    int base;  // element offset of this transform's row inside the CTA buffer
    __device__ __forceinline__ int at(int idx, int exch) const {
        const int s = (exch == 0) ? S1 : 4;
        return base + (idx ^ ((idx >> s) & MASK));
    }
};
// Column layout: TC transforms interleaved (column index fastest): conflict free as is.
template <int TC> struct ColAddr {
    int c;
    __device__ __forceinline__ int at(int idx, int /*exch*/) const { return idx * TC + c; }
};

struct NoHook {
    template <typename C> __device__ __forceinline__ void operator()(C * /*buf*/) const {}
};
struct NoHook0 {
    __device__ __forceinline__ void operator()() const {}
};

// 16-byte / 8-byte asynchronous global -> shared copy (LDGSTS), used to prefetch the spectrum
template <typename C> __device__ __forceinline__ void cp_async_elem(C *smem_dst, const C *gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    if (sizeof(C) == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int PENDING> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;\n" ::"n"(PENDING) : "memory"); }
// real element (8 / 4 bytes) global -> shared; !valid writes a zero without touching global memory
template <typename T> __device__ __forceinline__ void cp_async_real_zfill(T *smem_dst, const T *gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int n = valid ? (int)sizeof(T) : 0;
    if (sizeof(T) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}


// ---------------------------------------------------------------- L2 residency hints
// The four-step intermediates ("scratch") and the cached IR spectrum must stay L2 resident while the input and
// output signals stream through the same cache.  Scratch/spectrum accesses carry an evict_last policy, the
// final read of a scratch line (inverse columns) an evict_first one; signals use ld.cs / st.cs already.
#ifndef ADSP_L2_HINTS
#define ADSP_L2_HINTS 1
#endif
__device__ __forceinline__ uint64_t l2_policy_keep() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_drop() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ double2 ld_scratch(const double2 *p, uint64_t pol) {
#if ADSP_L2_HINTS
    double2 r;
    asm volatile("ld.global.cg.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(pol));
    return r;
#else
    return __ldcg(p);
#endif
}
__device__ __forceinline__ float2 ld_scratch(const float2 *p, uint64_t pol) {
#if ADSP_L2_HINTS
    float2 r;
    asm volatile("ld.global.cg.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(r.x), "=f"(r.y) : "l"(p), "l"(pol));
    return r;
#else
    return __ldcg(p);
#endif
}
__device__ __forceinline__ void st_scratch(double2 *p, double2 v, uint64_t pol) {
#if ADSP_L2_HINTS
    asm volatile("st.global.cg.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
#else
    __stcg(p, v);
#endif
}
__device__ __forceinline__ void st_scratch(float2 *p, float2 v, uint64_t pol) {
#if ADSP_L2_HINTS
    asm volatile("st.global.cg.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
#else
    __stcg(p, v);
#endif
}
// asynchronous global -> shared copy of one complex element with an L2 policy (IR spectrum prefetch)
template <typename C> __device__ __forceinline__ void cp_async_elem_keep(C *smem_dst, const C *gsrc, uint64_t pol) {
#if ADSP_L2_HINTS
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    if (sizeof(C) == 16) asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "l"(pol) : "memory");
    else asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "l"(pol) : "memory");
#else
    cp_async_elem(smem_dst, gsrc);
#endif
}

// cooperative ASYNCHRONOUS copy of the compact twiddle table (global -> shared, cp.async: no register
// dependency, so the tile's own global loads issue right behind it instead of waiting a full memory
// round trip).  cta_fft waits for it (cp.async.wait_all + barrier) just before the first twiddled pass.
template <typename T, int L>
__device__ __forceinline__ void load_tw_smem(cpx<T> *stw, const cpx<T> *__restrict__ gtw, int tid, int nthreads) {
    for (int i = tid; i < FftShape<L>::TW_ENTRIES; i += nthreads) cp_async_elem(&stw[i], &gtw[i]);
}

// ---------------------------------------------------------------- phase gates
// A transform alternates FP64-pipe phases (butterflies, "D") and shared-memory phases (Stockham
// exchanges, "L").  CtaGate: one tile per CTA, whole-CTA barriers, no gating.  (The gate is a template parameter of cta_fft:
// round 1 measured a ping-pong gate that phase-locks two tiles per CTA with named barriers -- slower)
struct CtaGate {
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ void d_begin() {}
    __device__ __forceinline__ void d_end() {}
    __device__ __forceinline__ void l_begin() {}
    __device__ __forceinline__ void l_end() {}
};

// ---------------------------------------------------------------- the in-CTA transform
// e[q] holds point (j + q*TPF) of the transform on entry and on exit (natural order).
// stw: compact twiddle table for this L in SHARED memory (see tw_pass_entries; entry for power r,
//      position k of a pass with NS: exp(-2*pi*i * r*k / (16*NS))).
// hook(buf): called once, right after the last pass has read its inputs from `buf` (the calling
//      thread may then reuse exactly the slots it read: addr.at(j + q*TPF, P-1)).
// gate: barrier/phase policy (see above).  D_OPEN_IN: the caller already holds the D phase;
//      D_OPEN_OUT: leave the final D phase open for the caller to close (gate.d_end()).
// hook0(): called once, right after the pass-0 butterflies (every input register has been consumed).
// WAIT_TW: wait for the asynchronous twiddle-table copy before the first twiddled pass (kernels that
//      keep other cp.async groups in flight wait for the table themselves, once, and pass false).
// All threads of the CTA (or ping-pong group) must call this together.
template <typename T, int L, bool INV, bool D_OPEN_IN = false, bool D_OPEN_OUT = false, bool WAIT_TW = true, typename Addr,
          typename Gate, typename Hook = NoHook, typename Hook0 = NoHook0>
__device__ __forceinline__ void cta_fft(cpx<T> (&e)[16], cpx<T> *buf, const Addr &addr, const cpx<T> *stw, int j,
                                        Gate &gate, const Hook &hook = Hook(), const Hook0 &hook0 = Hook0()) {
    using C = cpx<T>;
    using Sh = FftShape<L>;
    constexpr int R0 = Sh::R0, P = Sh::P, TPF = Sh::TPF, S0 = 16 / R0;

    // pass 0: radix R0, no twiddles
    if (!D_OPEN_IN) gate.d_begin();
    if (!ADSP_SKIP(7)) {
#pragma unroll
        for (int u = 0; u < S0; u++) Dft<R0, S0, INV, C>::run(&e[u]);
    }
    hook0();
    if (P == 0) {
        if (!D_OPEN_OUT) gate.d_end();
        return;
    }
    gate.d_end();

    gate.l_begin();
    gate.sync();  // buffer may still be read by the previous user
    if (!ADSP_SKIP(8)) {
#pragma unroll
        for (int u = 0; u < S0; u++) {
            const int b = j + u * TPF;
#pragma unroll
            for (int r = 0; r < R0; r++) buf[addr.at(R0 * b + r, 0)] = e[u + r * S0];
        }
    }
    if (WAIT_TW) cp_async_wait_all();   // twiddle table copy (load_tw_smem) issued by this thread has landed
    gate.sync();

    int ns = R0;
    int off = 0;
#pragma unroll
    for (int t = 1; t <= P; t++) {
        if (!ADSP_SKIP(8)) {
#pragma unroll
            for (int q = 0; q < 16; q++) e[q] = buf[addr.at(j + q * TPF, t - 1)];
        }
        if (t == P) hook(buf);
        gate.l_end();

        gate.d_begin();
        const int k = j & (ns - 1);
        if (ADSP_SKIP(7)) {
        } else if (ns <= 16 && !ADSP_TW_TREE_ALL) {
            const C *twp = stw + off + k;
#pragma unroll
            for (int r = 1; r < 16; r++) e[r] = cmul_tw<INV>(e[r], twp[(r - 1) * ns]);
        } else if (ADSP_TW_CHAIN) {
            // powers by a running product: two twiddle values live instead of eight (register pressure), 15 dependent products
            const C w1 = stw[off + k];
            C w = w1;
            e[1] = cmul_tw<INV>(e[1], w);
#pragma unroll
            for (int r = 2; r < 16; r++) { w = cmul(w, w1); e[r] = cmul_tw<INV>(e[r], w); }
        } else {
            const C w1 = stw[off + k];
            const C w2 = cmul(w1, w1), w4 = cmul(w2, w2), w8 = cmul(w4, w4);
            e[1] = cmul_tw<INV>(e[1], w1);
            e[2] = cmul_tw<INV>(e[2], w2);
            e[3] = cmul_tw<INV>(e[3], cmul(w2, w1));
            e[4] = cmul_tw<INV>(e[4], w4);
            e[5] = cmul_tw<INV>(e[5], cmul(w4, w1));
            const C w6 = cmul(w4, w2);
            e[6] = cmul_tw<INV>(e[6], w6);
            e[7] = cmul_tw<INV>(e[7], cmul(w6, w1));
            e[8] = cmul_tw<INV>(e[8], w8);
            e[9] = cmul_tw<INV>(e[9], cmul(w8, w1));
            const C w10 = cmul(w8, w2);
            e[10] = cmul_tw<INV>(e[10], w10);
            e[11] = cmul_tw<INV>(e[11], cmul(w10, w1));
            const C w12 = cmul(w8, w4);
            e[12] = cmul_tw<INV>(e[12], w12);
            e[13] = cmul_tw<INV>(e[13], cmul(w12, w1));
            const C w14 = cmul(w12, w2);
            e[14] = cmul_tw<INV>(e[14], w14);
            e[15] = cmul_tw<INV>(e[15], cmul(w14, w1));
        }
        if (!ADSP_SKIP(7)) Dft<16, 1, INV, C>::run(&e[0]);
        if (t < P) {
            gate.d_end();
            gate.l_begin();
            gate.sync();
            if (!ADSP_SKIP(8)) {
                const int j0 = (j - k) * 16 + k;
#pragma unroll
                for (int r = 0; r < 16; r++) buf[addr.at(j0 + r * ns, t)] = e[r];
            }
            gate.sync();
        }
        off += tw_pass_entries(ns);
        ns *= 16;
    }
    if (!D_OPEN_OUT) gate.d_end();
}

}  // namespace adsp
