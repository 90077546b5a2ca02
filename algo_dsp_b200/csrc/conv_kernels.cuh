// conv_kernels.cuh -- the overlap-save / overlap-add FFT-convolution kernels (sm_100a).
//
// Replaces the per-block loop of the reference (dsp/conv/overlap_save.go:144-251,
// overlap_add.go:120-161): zero-fill, history copy, complex FFT, multiply by the kernel
// spectrum, inverse FFT, discard/extract.  Here all of that bookkeeping is index arithmetic in
// the load and store of the transform kernels; two real blocks ride in one complex transform
// (re = block 2p, im = block 2p+1), valid because the IR is real.
//
// Two shapes:
//   * N <= 4096: one kernel (fftconv_full) does load -> FFT -> *H -> IFFT -> store in shared memory.
//   * N = N1*N2 larger: four-step.  cols_fwd (N1-point transforms down the columns of the
//     N1 x N2 matrix + four-step twiddle) -> rows (N2-point FFT, *H, N2-point IFFT, in place in
//     an L2-resident scratch) -> cols_inv (conj twiddle, N1-point inverse, discard, store).
#pragma once
#include "fft_core.cuh"

namespace adsp {

// Block structure of one batched convolution call.  Output block b of a channel covers output
// samples [b*S, b*S+S) and is produced from input samples [b*S - D, b*S - D + N) (zero outside
// [0, n)); positions i >= D of the circular result are valid (D >= K-1), same rule as
// overlap_save.go:151-186 with the history/zero-fill folded into bounds checks.
struct ConvGeom {
    long long n;            // input samples per channel
    long long out_len;      // output samples per channel to produce (n + K - 1)
    long long in_stride;    // elements between channels (input)
    long long out_stride;   // elements between channels (output)
    long long S;            // new output samples per block
    long long D;            // discarded leading positions per block (N - S)
    long long total_blocks; // channels * nblk
    long long in_shift;     // added to every block's input start (K-1 for "valid" streaming calls)
    long long out_shift;    // added to every block's output start (IR partition p: p * part_len)
    int nblk;               // blocks per channel
    int accumulate;         // 0: store, 1: add into the output (IR-partition sums)
};

template <typename T> struct BlockIO {
    const T *in;   // channel base + (b*S - D): may point before the channel start
    long long lo, hi;  // valid i-range: lo <= i < hi  maps to in-range samples
    T *out;        // channel base + b*S
    long long cnt; // valid outputs
};

template <typename T>
__device__ __forceinline__ BlockIO<T> block_io(const ConvGeom &g, const T *x, T *y, long long bid) {
    BlockIO<T> io;
    if (bid >= g.total_blocks) { io.in = x; io.lo = 0; io.hi = 0; io.out = y; io.cnt = 0; return io; }
    const long long ch = bid / g.nblk;
    const long long b = bid - ch * g.nblk;
    const long long start = b * g.S - g.D + g.in_shift;
    io.in = x + ch * g.in_stride + start;
    io.lo = start < 0 ? -start : 0;
    io.hi = g.n - start;  // i < hi  <=> start + i < n
    io.out = y + ch * g.out_stride + g.out_shift + b * g.S;
    long long c = g.out_len - b * g.S;
    io.cnt = c < g.S ? (c < 0 ? 0 : c) : g.S;
    return io;
}

template <typename T> __device__ __forceinline__ T ld_stream(const T *p) { return __ldcs(p); }

// ------------------------------------------------------------------------------------------
// Single-kernel path, N = L <= 4096.  256 threads; 4096/L block-pairs per CTA.
// SPECTRUM mode: forward transform of x only, scaled, written to `spec` (used once per plan to
// build the cached IR spectrum, replacing overlap_save.go:96-101).
template <typename T, int L, bool SPECTRUM>
__global__ void __launch_bounds__(256, 2)
fftconv_full(ConvGeom g, const T *__restrict__ x, T *__restrict__ y,
             const cpx<T> *__restrict__ H, cpx<T> *__restrict__ spec, T scale,
             const cpx<T> *__restrict__ tw, long long npairs) {
    using C = cpx<T>;
    using Sh = FftShape<L>;
    constexpr int TPF = Sh::TPF;
    constexpr int ROWS = 256 / TPF;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *buf = reinterpret_cast<C *>(smem_raw);
    C *stw = buf + ROWS * L;
    load_tw_smem<T, L>(stw, tw, threadIdx.x, 256);  // visible after the first __syncthreads in cta_fft

    const int row = threadIdx.x / TPF;
    const int j = threadIdx.x % TPF;
    const long long pair = (long long)blockIdx.x * ROWS + row;
    RowAddr<T, Sh::R0> addr{row * L};

    const BlockIO<T> a = block_io<T>(g, x, y, pair < npairs ? 2 * pair : g.total_blocks);
    const BlockIO<T> b = block_io<T>(g, x, y, pair < npairs ? 2 * pair + 1 : g.total_blocks);

    C e[16];
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const long long i = j + q * TPF;
        e[q].x = (i >= a.lo && i < a.hi) ? ld_stream(a.in + i) : (T)0;
        e[q].y = (i >= b.lo && i < b.hi) ? ld_stream(b.in + i) : (T)0;
    }
    // prefetch this thread's 16 spectrum values into the slots it has just emptied (no barrier
    // needed: same thread writes and reads them); overlaps with the last forward pass
    auto prefetch_h = [&](C *b) {
        if (!SPECTRUM) {
#pragma unroll
            for (int q = 0; q < 16; q++) cp_async_elem(&b[addr.at(j + q * TPF, Sh::P - 1)], &H[j + q * TPF]);
        }
    };
    cta_fft<T, L, false>(e, buf, addr, stw, j, prefetch_h);
    if (SPECTRUM) {
        if (pair < npairs) {
#pragma unroll
            for (int q = 0; q < 16; q++) {
                C v; v.x = e[q].x * scale; v.y = e[q].y * scale;
                spec[pair * L + j + q * TPF] = v;
            }
        }
        return;
    }
    cp_async_wait_all();
#pragma unroll
    for (int q = 0; q < 16; q++) e[q] = cmul(e[q], buf[addr.at(j + q * TPF, Sh::P - 1)]);
    cta_fft<T, L, true>(e, buf, addr, stw, j);
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const long long o = (long long)(j + q * TPF) - g.D;
        if (o >= 0) {
            if (g.accumulate) {
                if (o < a.cnt) a.out[o] += e[q].x;
                if (o < b.cnt) b.out[o] += e[q].y;
            } else {
                if (o < a.cnt) __stcs(a.out + o, e[q].x);
                if (o < b.cnt) __stcs(b.out + o, e[q].y);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Four-step twiddle W_N^m from two small tables: W_N^m = hi[m >> 10] * lo[m & 1023].
template <typename T>
__device__ __forceinline__ cpx<T> twiddle_n(const cpx<T> *__restrict__ tw_hi, const cpx<T> *__restrict__ tw_lo,
                                            unsigned m) {
    return cmul(__ldg(&tw_hi[m >> 10]), __ldg(&tw_lo[m & 1023u]));
}

// g_r = base * rho^r, r = 0..15, short dependency chains (depth <= 6 products).
template <typename C> __device__ __forceinline__ void geometric16(C base, C rho, C (&g)[16]) {
    const C rho2 = cmul(rho, rho);
    const C rho4 = cmul(rho2, rho2);
    g[0] = base;
    g[4] = cmul(g[0], rho4);
    g[8] = cmul(g[4], rho4);
    g[12] = cmul(g[8], rho4);
#pragma unroll
    for (int a = 0; a < 16; a += 4) {
        g[a + 1] = cmul(g[a], rho);
        g[a + 2] = cmul(g[a], rho2);
        g[a + 3] = cmul(g[a + 2], rho);
    }
}

template <int N1> struct ColShape {
    static constexpr int TPF = N1 / 16;                         // threads per column transform
    static constexpr int TC = (N1 <= 256) ? (256 / TPF) : ((N1 == 512) ? 16 : 8);  // columns per tile
    static constexpr int THREADS = TPF * TC;
    static constexpr int SMEM_ELEMS = N1 * TC;
    static constexpr int MIN_CTAS = (THREADS <= 256) ? 2 : 1;
};

// Forward column pass.  grid = (N2/TC, pairs in this group).  Writes A[k1][n2] * W_N^(k1*n2)
// to scratch (row-major N1 x N2 per pair).
template <typename T, int N1>
__global__ void __launch_bounds__(ColShape<N1>::THREADS, ColShape<N1>::MIN_CTAS)
fftconv_cols_fwd(ConvGeom g, const T *__restrict__ x, cpx<T> *__restrict__ scratch, int N2, int lgN,
                 const cpx<T> *__restrict__ tw, const cpx<T> *__restrict__ tw_hi,
                 const cpx<T> *__restrict__ tw_lo, long long pair0) {
    using C = cpx<T>;
    using CS = ColShape<N1>;
    constexpr int TPF = CS::TPF, TC = CS::TC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *buf = reinterpret_cast<C *>(smem_raw);

    C *stw = buf + ((FftShape<N1>::P > 0) ? CS::SMEM_ELEMS : 0);
    load_tw_smem<T, N1>(stw, tw, threadIdx.x, CS::THREADS);

    const int c = threadIdx.x % TC;
    const int j = threadIdx.x / TC;
    const int n2 = blockIdx.x * TC + c;
    const long long pair = pair0 + blockIdx.y;
    ColAddr<TC> addr{c};

    // four-step twiddle seeds, fetched first so their latency hides behind the data loads
    const unsigned maskN = (1u << lgN) - 1u;
    const C tw_base = twiddle_n<T>(tw_hi, tw_lo, ((unsigned)n2 * (unsigned)j) & maskN);
    const C tw_rho = twiddle_n<T>(tw_hi, tw_lo, ((unsigned)n2 * (unsigned)TPF) & maskN);

    const BlockIO<T> a = block_io<T>(g, x, (T *)nullptr, 2 * pair);
    const BlockIO<T> b = block_io<T>(g, x, (T *)nullptr, 2 * pair + 1);

    C e[16];
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const long long i = (long long)(j + q * TPF) * N2 + n2;
        e[q].x = (i >= a.lo && i < a.hi) ? ld_stream(a.in + i) : (T)0;
        e[q].y = (i >= b.lo && i < b.hi) ? ld_stream(b.in + i) : (T)0;
    }
    cta_fft<T, N1, false>(e, buf, addr, stw, j);

    // four-step twiddle: k1 = j + r*TPF  ->  W_N^(n2*j) * (W_N^(n2*TPF))^r
    C gtw[16];
    geometric16(tw_base, tw_rho, gtw);
    C *dst = scratch + (size_t)blockIdx.y * ((size_t)N1 * N2) + n2;
#pragma unroll
    for (int r = 0; r < 16; r++) dst[(size_t)(j + r * TPF) * N2] = cmul(e[r], gtw[r]);
}

// Row pass, in place on scratch: N2-point FFT, multiply by the cached IR spectrum (already
// permuted to the four-step order and scaled by 1/N), N2-point inverse FFT.
// grid = (N1 / ROWS, pairs in this group); 256 threads; ROWS = 4096/L rows per CTA.
// SPECTRUM mode: forward only, scaled, written to `spec` (IR spectrum construction).
template <typename T, int L, bool SPECTRUM>
__global__ void __launch_bounds__(256, 2)
fftconv_rows(cpx<T> *__restrict__ scratch, const cpx<T> *__restrict__ H, cpx<T> *__restrict__ spec, T scale,
             int N1, const cpx<T> *__restrict__ tw) {
    using C = cpx<T>;
    using Sh = FftShape<L>;
    constexpr int TPF = Sh::TPF;
    constexpr int ROWS = 256 / TPF;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *buf = reinterpret_cast<C *>(smem_raw);
    C *stw = buf + ROWS * L;
    load_tw_smem<T, L>(stw, tw, threadIdx.x, 256);

    const int row = threadIdx.x / TPF;
    const int j = threadIdx.x % TPF;
    const size_t k1 = (size_t)blockIdx.x * ROWS + row;
    RowAddr<T, Sh::R0> addr{row * L};
    C *p = scratch + (size_t)blockIdx.y * ((size_t)N1 * L) + k1 * L + j;
    const size_t hoff = k1 * L + j;

    C e[16];
#pragma unroll
    for (int q = 0; q < 16; q++) e[q] = p[q * TPF];
    auto prefetch_h = [&](C *b) {
        if (!SPECTRUM) {
#pragma unroll
            for (int q = 0; q < 16; q++) cp_async_elem(&b[addr.at(j + q * TPF, Sh::P - 1)], &H[hoff + q * TPF]);
        }
    };
    cta_fft<T, L, false>(e, buf, addr, stw, j, prefetch_h);
    if (SPECTRUM) {
#pragma unroll
        for (int q = 0; q < 16; q++) {
            C v; v.x = e[q].x * scale; v.y = e[q].y * scale;
            spec[(size_t)blockIdx.y * ((size_t)N1 * L) + hoff + q * TPF] = v;
        }
        return;
    }
    cp_async_wait_all();
#pragma unroll
    for (int q = 0; q < 16; q++) e[q] = cmul(e[q], buf[addr.at(j + q * TPF, Sh::P - 1)]);
    cta_fft<T, L, true>(e, buf, addr, stw, j);
#pragma unroll
    for (int q = 0; q < 16; q++) p[q * TPF] = e[q];
}

// Inverse column pass: conj four-step twiddle, N1-point inverse, keep positions >= D, split
// re/im to the two real output blocks.
template <typename T, int N1>
__global__ void __launch_bounds__(ColShape<N1>::THREADS, ColShape<N1>::MIN_CTAS)
fftconv_cols_inv(ConvGeom g, const cpx<T> *__restrict__ scratch, const T *__restrict__ x, T *__restrict__ y,
                 int N2, int lgN, const cpx<T> *__restrict__ tw, const cpx<T> *__restrict__ tw_hi,
                 const cpx<T> *__restrict__ tw_lo, long long pair0) {
    using C = cpx<T>;
    using CS = ColShape<N1>;
    constexpr int TPF = CS::TPF, TC = CS::TC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *buf = reinterpret_cast<C *>(smem_raw);

    C *stw = buf + ((FftShape<N1>::P > 0) ? CS::SMEM_ELEMS : 0);
    load_tw_smem<T, N1>(stw, tw, threadIdx.x, CS::THREADS);

    const int c = threadIdx.x % TC;
    const int j = threadIdx.x / TC;
    const int n2 = blockIdx.x * TC + c;
    const long long pair = pair0 + blockIdx.y;
    ColAddr<TC> addr{c};

    const unsigned maskN = (1u << lgN) - 1u;
    C gtw[16];
    geometric16(twiddle_n<T>(tw_hi, tw_lo, ((unsigned)n2 * (unsigned)j) & maskN),
                twiddle_n<T>(tw_hi, tw_lo, ((unsigned)n2 * (unsigned)TPF) & maskN), gtw);
    const C *src = scratch + (size_t)blockIdx.y * ((size_t)N1 * N2) + n2;
    C e[16];
#pragma unroll
    for (int q = 0; q < 16; q++) e[q] = cmul_tw<true>(src[(size_t)(j + q * TPF) * N2], gtw[q]);
    cta_fft<T, N1, true>(e, buf, addr, stw, j);

    const BlockIO<T> a = block_io<T>(g, x, y, 2 * pair);
    const BlockIO<T> b = block_io<T>(g, x, y, 2 * pair + 1);
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const long long o = (long long)(j + r * TPF) * N2 + n2 - g.D;
        if (o >= 0) {
            if (g.accumulate) {
                if (o < a.cnt) a.out[o] += e[r].x;
                if (o < b.cnt) b.out[o] += e[r].y;
            } else {
                if (o < a.cnt) __stcs(a.out + o, e[r].x);
                if (o < b.cnt) __stcs(b.out + o, e[r].y);
            }
        }
    }
}

}  // namespace adsp
