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
#include "aux_kernels.cuh"
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
    // total_blocks < 2^31 always (it is a grid dimension): 32-bit division is ~5x cheaper than 64-bit
    const long long ch = (long long)((unsigned)bid / (unsigned)g.nblk);
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

// The element index inside a block is < N <= 2^22, so the per-element bounds tests of a tile run in 32 bits:
// inputs  valid  <=>  (unsigned)(i - lo) < span   with lo, span clamped into [0, N];
// outputs valid  <=>  (unsigned)o < cnt           with o = i - D.
template <typename T> struct TileIn {
    const T *p;          // block base (element 0 of the block; may point before the channel start, only valid indices are read)
    int lo;
    unsigned span;
};
template <typename T> __device__ __forceinline__ TileIn<T> tile_in(const BlockIO<T> &b, long long N) {
    TileIn<T> t;
    long long lo = b.lo < 0 ? 0 : (b.lo > N ? N : b.lo);
    long long hi = b.hi < lo ? lo : (b.hi > N ? N : b.hi);
    t.p = b.in; t.lo = (int)lo; t.span = (unsigned)(hi - lo);
    return t;
}
template <typename T> struct TileOut {
    T *p;                // block output base (output sample 0 of the block)
    unsigned cnt;
};
template <typename T> __device__ __forceinline__ TileOut<T> tile_out(const BlockIO<T> &b, long long N) {
    TileOut<T> t;
    t.p = b.out; t.cnt = (unsigned)(b.cnt < 0 ? 0 : (b.cnt > N ? N : b.cnt));
    return t;
}



// CTA shapes (compile-time knobs; tools/ builds variants with -D to A/B them on the GPU).
// A row CTA owns 16 points per thread: L/16 threads per row, rows_cta_threads(L)/(L/16) rows per CTA.
// (Round 1 also carried ping-pong, persistent fused, stage-merged, prefetching and interleaved variants of these kernels;
// all were measured slower and have been removed -- results and artefacts in DESIGN.md section 7.)
#ifndef ADSP_H_DIRECT
#define ADSP_H_DIRECT 0
#endif
#ifndef ADSP_ROWS_SMALL_CTA
#define ADSP_ROWS_SMALL_CTA 1
#endif
#ifndef ADSP_COLS_CTA_THREADS
#define ADSP_COLS_CTA_THREADS 128
#endif
#ifndef ADSP_MIN_CTAS_128
#define ADSP_MIN_CTAS_128 4      // resident 128-thread CTAs per SM the register allocator must allow (4 -> 128 regs, 5 -> 102)
#endif
constexpr int rows_cta_threads(int L) { return (ADSP_ROWS_SMALL_CTA && L <= 2048) ? 128 : 256; }
constexpr int rows_min_ctas(int L) { return rows_cta_threads(L) == 128 ? ADSP_MIN_CTAS_128 : 2; }
// fp32 needs half the registers for the same 16 points: 6 resident 128-thread CTAs (85 registers) measured +21 % (141.8 -> 172.2 Gsamples/s)
#ifndef ADSP_MIN_CTAS_128_F32
#define ADSP_MIN_CTAS_128_F32 6
#endif
template <typename T> constexpr int min_ctas_for(int threads, int fp64_ctas) {
    return (sizeof(T) == 4 && threads <= 128) ? ((fp64_ctas + 2 > ADSP_MIN_CTAS_128_F32) ? fp64_ctas + 2 : ADSP_MIN_CTAS_128_F32) : fp64_ctas;
}

// ------------------------------------------------------------------------------------------
// Single-kernel path, N = L <= 4096.  256 threads; 4096/L block-pairs per CTA.
// SPECTRUM mode: forward transform of x only, scaled, written to `spec` (used once per plan to
// build the cached IR spectrum, replacing overlap_save.go:96-101).
template <typename T, int L, bool SPECTRUM>
__global__ void __launch_bounds__(rows_cta_threads(L), min_ctas_for<T>(rows_cta_threads(L), rows_min_ctas(L)))
fftconv_full(ConvGeom g, const T *__restrict__ x, T *__restrict__ y,
             const cpx<T> *__restrict__ H, cpx<T> *__restrict__ spec, T scale,
             const cpx<T> *__restrict__ tw, long long npairs) {
    using C = cpx<T>;
    using Sh = FftShape<L>;
    constexpr int TPF = Sh::TPF;
    constexpr int ROWS = rows_cta_threads(L) / TPF;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *buf = reinterpret_cast<C *>(smem_raw);
    C *stw = buf + ROWS * L;
    load_tw_smem<T, L>(stw, tw, threadIdx.x, rows_cta_threads(L));  // visible after the first __syncthreads in cta_fft

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
    CtaGate gate;
    cta_fft<T, L, false>(e, buf, addr, stw, j, gate, prefetch_h);
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
    cta_fft<T, L, true>(e, buf, addr, stw, j, gate);
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

// e[r] *= base * rho^r (forward) or its conjugate (INV), r = 0..15, forming the factors on the fly so
// that only base/rho/rho^2/rho^4 stay live (the 16-entry table variant below costs 64 registers).
template <bool INV, typename C> __device__ __forceinline__ void apply_geometric16(C (&e)[16], C base, C rho) {
    const C rho2 = cmul(rho, rho);
    const C rho4 = cmul(rho2, rho2);
    C ga = base;
#pragma unroll
    for (int a = 0; a < 16; a += 4) {
        if (a) ga = cmul(ga, rho4);
        e[a] = cmul_tw<INV>(e[a], ga);
        e[a + 1] = cmul_tw<INV>(e[a + 1], cmul(ga, rho));
        const C g2 = cmul(ga, rho2);
        e[a + 2] = cmul_tw<INV>(e[a + 2], g2);
        e[a + 3] = cmul_tw<INV>(e[a + 3], cmul(g2, rho));
    }
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

#ifndef ADSP_COLS_MIN_CTAS_128
#define ADSP_COLS_MIN_CTAS_128 5     // column kernels: 5 resident CTAs (102 registers) measured +2 % on 2^20-point transforms
#endif
#ifndef ADSP_COLS_MIN_CTAS_256
#define ADSP_COLS_MIN_CTAS_256 2
#endif
#ifndef ADSP_COLS_TC_512
#define ADSP_COLS_TC_512 8      // 256-thread CTAs (3 per SM); 16 columns / 512 threads measured 9 % slower
#endif
#ifndef ADSP_COLS_TC_1024
#define ADSP_COLS_TC_1024 4     // 256-thread CTAs; 8 columns / 512 threads measured 6 % slower
#endif
template <int N1> struct ColShape {
    static constexpr int TPF = N1 / 16;                         // threads per column transform
    static constexpr int TC = (N1 <= 256) ? (ADSP_COLS_CTA_THREADS / TPF) : ((N1 == 512) ? ADSP_COLS_TC_512 : ADSP_COLS_TC_1024);  // columns per tile
    static constexpr int THREADS = TPF * TC;
    static constexpr int SMEM_ELEMS = N1 * TC;
    static constexpr int MIN_CTAS = (THREADS <= 128) ? ADSP_COLS_MIN_CTAS_128 : ((THREADS <= 256) ? ADSP_COLS_MIN_CTAS_256 : 1);
};

// ------------------------------------------------------------------------------------------
// Tile bodies (device functions) of the stand-alone kernels.  `tid` is the thread's index inside its tile, `gate` the
// barrier/phase policy (fft_core.cuh), `active` false for padding tiles (they run the same barrier
// sequence on zeros and store nothing).  Scratch is read with ld.global.cg (L2 only): another CTA
// produced it and L1 could hold stale lines from an earlier use of the same scratch slot.

// Forward column tile: N1-point transforms down TC columns of the N1 x N2 view of block pair `pair`,
// times the four-step twiddle W_N^(k1*n2), written to scratch (row-major N1 x N2).
template <typename T, int N1, typename Gate>
__device__ __forceinline__ void cols_fwd_tile(const ConvGeom &g, const T *__restrict__ x, cpx<T> *__restrict__ scratch_pair,
                                              int N2, int lgN, const cpx<T> *stw, const cpx<T> *__restrict__ tw_hi,
                                              const cpx<T> *__restrict__ tw_lo, long long pair, int tile, cpx<T> *buf,
                                              int tid, Gate &gate, bool active) {
    using C = cpx<T>;
    using CS = ColShape<N1>;
    constexpr int TPF = CS::TPF, TC = CS::TC;
    const int c = tid % TC;
    const int j = tid / TC;
    const int n2 = tile * TC + c;
    ColAddr<TC> addr{c};

    // four-step twiddle seeds, fetched first so their latency hides behind the data loads
    const unsigned maskN = (1u << lgN) - 1u;
    const C tw_base = twiddle_n<T>(tw_hi, tw_lo, ((unsigned)n2 * (unsigned)j) & maskN);
    const C tw_rho = twiddle_n<T>(tw_hi, tw_lo, ((unsigned)n2 * (unsigned)TPF) & maskN);

    const BlockIO<T> a = block_io<T>(g, x, (T *)nullptr, active ? 2 * pair : g.total_blocks);
    const BlockIO<T> b = block_io<T>(g, x, (T *)nullptr, active ? 2 * pair + 1 : g.total_blocks);

    C e[16];
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const long long i = (long long)(j + q * TPF) * N2 + n2;
        e[q].x = (i >= a.lo && i < a.hi && !ADSP_SKIP(3)) ? ld_stream(a.in + i) : (T)0;
        e[q].y = (i >= b.lo && i < b.hi && !ADSP_SKIP(3)) ? ld_stream(b.in + i) : (T)0;
    }
    cta_fft<T, N1, false, false, true>(e, buf, addr, stw, j, gate);   // leaves the last D phase open

    // k1 = j + r*TPF  ->  W_N^(n2*j) * (W_N^(n2*TPF))^r
    apply_geometric16<false>(e, tw_base, tw_rho);
    gate.d_end();
    if (active && !ADSP_SKIP(4)) {
        C *dst = scratch_pair + n2;
        const uint64_t keep = l2_policy_keep();
#pragma unroll
        for (int r = 0; r < 16; r++) st_scratch(&dst[(size_t)(j + r * TPF) * N2], e[r], keep);
    }
}

// Row tile, in place on scratch: ROWS rows; N2-point FFT, multiply by the cached IR spectrum
// (four-step order, pre-scaled by 1/N, prefetched by cp.async), N2-point inverse FFT.
template <typename T, int L, typename Gate>
__device__ __forceinline__ void rows_tile(cpx<T> *__restrict__ scratch_pair, const cpx<T> *__restrict__ H, int rowtile,
                                          cpx<T> *buf, const cpx<T> *stw, int tid, Gate &gate, bool active) {
    using C = cpx<T>;
    using Sh = FftShape<L>;
    constexpr int TPF = Sh::TPF;
    constexpr int ROWS = rows_cta_threads(L) / TPF;
    const int row = tid / TPF;
    const int j = tid % TPF;
    const size_t k1 = (size_t)rowtile * ROWS + row;
    RowAddr<T, Sh::R0> addr{row * L};
    const size_t hoff = k1 * L + j;
    C *p = scratch_pair + hoff;

    C e[16];
    const uint64_t keep = l2_policy_keep();
    if (active && !ADSP_SKIP(0)) {
#pragma unroll
        for (int q = 0; q < 16; q++) e[q] = ld_scratch(&p[q * TPF], keep);
    } else {
#pragma unroll
        for (int q = 0; q < 16; q++) { e[q].x = (T)0; e[q].y = (T)0; }
    }
#if ADSP_H_DIRECT
    // spectrum straight from L2 into registers after the last forward pass (no shared-memory staging: two LSU passes less,
    // one exposed L2 round trip more)
    cta_fft<T, L, false, false, true>(e, buf, addr, stw, j, gate);
    if (active && !ADSP_SKIP(2)) {
        C hv[16];
#pragma unroll
        for (int q = 0; q < 16; q++) hv[q] = ld_scratch(&H[hoff + q * TPF], keep);
#pragma unroll
        for (int q = 0; q < 16; q++) e[q] = cmul(e[q], hv[q]);
    }
#else
    auto prefetch_h = [&](C *b) {
        if (active && !ADSP_SKIP(2)) {
#pragma unroll
            for (int q = 0; q < 16; q++) cp_async_elem_keep(&b[addr.at(j + q * TPF, Sh::P - 1)], &H[hoff + q * TPF], keep);
        }
    };
    cta_fft<T, L, false, false, true>(e, buf, addr, stw, j, gate, prefetch_h);   // D phase stays open ...
    cp_async_wait_all();
    if (active && !ADSP_SKIP(2)) {
#pragma unroll
        for (int q = 0; q < 16; q++) e[q] = cmul(e[q], buf[addr.at(j + q * TPF, Sh::P - 1)]);
    }
#endif
    cta_fft<T, L, true, true, false>(e, buf, addr, stw, j, gate);                // ... through the inverse's first pass
    if (active && !ADSP_SKIP(1)) {
#pragma unroll
        for (int q = 0; q < 16; q++) st_scratch(&p[q * TPF], e[q], keep);
    }
}

// Inverse column tile: conj four-step twiddle, N1-point inverse, keep positions >= D, split re/im to
// the two real output blocks.
template <typename T, int N1, typename Gate>
__device__ __forceinline__ void cols_inv_tile(const ConvGeom &g, const cpx<T> *__restrict__ scratch_pair, const T *x,
                                              T *__restrict__ y, int N2, int lgN, const cpx<T> *stw,
                                              const cpx<T> *__restrict__ tw_hi, const cpx<T> *__restrict__ tw_lo,
                                              long long pair, int tile, cpx<T> *buf, int tid, Gate &gate, bool active) {
    using C = cpx<T>;
    using CS = ColShape<N1>;
    constexpr int TPF = CS::TPF, TC = CS::TC;
    const int c = tid % TC;
    const int j = tid / TC;
    const int n2 = tile * TC + c;
    ColAddr<TC> addr{c};

    const unsigned maskN = (1u << lgN) - 1u;
    const C tw_base = twiddle_n<T>(tw_hi, tw_lo, ((unsigned)n2 * (unsigned)j) & maskN);
    const C tw_rho = twiddle_n<T>(tw_hi, tw_lo, ((unsigned)n2 * (unsigned)TPF) & maskN);
    const C *src = scratch_pair + n2;
    C e[16];
    if (active && !ADSP_SKIP(5)) {
        const uint64_t drop = l2_policy_drop();   // last use of these scratch lines
#pragma unroll
        for (int q = 0; q < 16; q++) e[q] = ld_scratch(&src[(size_t)(j + q * TPF) * N2], drop);
    } else {
#pragma unroll
        for (int q = 0; q < 16; q++) { e[q].x = (T)0; e[q].y = (T)0; }
    }
    gate.d_begin();
    apply_geometric16<true>(e, tw_base, tw_rho);
    cta_fft<T, N1, true, true, false>(e, buf, addr, stw, j, gate);

    const BlockIO<T> a = block_io<T>(g, x, y, active ? 2 * pair : g.total_blocks);
    const BlockIO<T> b = block_io<T>(g, x, y, active ? 2 * pair + 1 : g.total_blocks);
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const long long o = (long long)(j + r * TPF) * N2 + n2 - g.D;
        if (o >= 0 && !ADSP_SKIP(6)) {
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

// ------------------------------------------------------------------------------------------
// Stand-alone kernels (one tile per CTA).  grid = (tiles, pairs in this group).
template <typename T, int N1>
__global__ void __launch_bounds__(ColShape<N1>::THREADS, min_ctas_for<T>(ColShape<N1>::THREADS, ColShape<N1>::MIN_CTAS))
fftconv_cols_fwd(ConvGeom g, const T *__restrict__ x, cpx<T> *__restrict__ scratch, int N2, int lgN,
                 const cpx<T> *__restrict__ tw, const cpx<T> *__restrict__ tw_hi,
                 const cpx<T> *__restrict__ tw_lo, long long pair0, int ntiles) {
    using C = cpx<T>;
    using CS = ColShape<N1>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *buf = reinterpret_cast<C *>(smem_raw);
    C *stw = buf + ((FftShape<N1>::P > 0) ? CS::SMEM_ELEMS : 0);
    load_tw_smem<T, N1>(stw, tw, threadIdx.x, CS::THREADS);
    CtaGate gate;
    // 1-D grid; a CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... (table copy and prologue amortised)
    const int tiles_per_pair = N2 / CS::TC;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int pl = t / tiles_per_pair, tile = t - pl * tiles_per_pair;
        cols_fwd_tile<T, N1>(g, x, scratch + (size_t)ADSP_ALIAS(pl) * ((size_t)N1 * N2), N2, lgN, stw, tw_hi, tw_lo, pair0 + pl, tile, buf,
                             threadIdx.x, gate, true);
    }
}

// MODE 0: convolution (FFT, *H, IFFT, in place).  MODE 1: forward only, scaled, written to `spec`
// (IR spectrum construction once per plan; correlation forward pass with spec == scratch).
// MODE 2: inverse only, written to `spec` (correlation inverse pass, in place).
template <typename T, int L, int MODE>
__global__ void __launch_bounds__(rows_cta_threads(L), min_ctas_for<T>(rows_cta_threads(L), rows_min_ctas(L)))
fftconv_rows(cpx<T> *__restrict__ scratch, const cpx<T> *__restrict__ H, cpx<T> *spec, T scale,
             int N1, const cpx<T> *__restrict__ tw, int ntiles) {
    using C = cpx<T>;
    using Sh = FftShape<L>;
    constexpr int TPF = Sh::TPF;
    constexpr int ROWS = rows_cta_threads(L) / TPF;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *buf = reinterpret_cast<C *>(smem_raw);
    C *stw = buf + ROWS * L;
    load_tw_smem<T, L>(stw, tw, threadIdx.x, rows_cta_threads(L));
    CtaGate gate;
    if (MODE == 0) {   // 1-D grid over flattened (pair, row tile); a CTA walks tiles with stride gridDim.x
        const int tiles_per_pair = N1 / ROWS;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int pl = t / tiles_per_pair, tile = t - pl * tiles_per_pair;
            rows_tile<T, L>(scratch + (size_t)ADSP_ALIAS(pl) * ((size_t)N1 * L), H, tile, buf, stw, threadIdx.x, gate, true);
        }
        return;
    }
    C *pairbase = scratch + (size_t)blockIdx.y * ((size_t)N1 * L);
    const int row = threadIdx.x / TPF;
    const int j = threadIdx.x % TPF;
    const size_t k1 = (size_t)blockIdx.x * ROWS + row;
    RowAddr<T, Sh::R0> addr{row * L};
    const size_t hoff = k1 * L + j;
    C e[16];
#pragma unroll
    for (int q = 0; q < 16; q++) e[q] = __ldcg(&pairbase[hoff + q * TPF]);
    if (MODE == 1) cta_fft<T, L, false>(e, buf, addr, stw, j, gate);
    else cta_fft<T, L, true>(e, buf, addr, stw, j, gate);
    C *dst = spec + (size_t)blockIdx.y * ((size_t)N1 * L) + hoff;
#pragma unroll
    for (int q = 0; q < 16; q++) {
        C v; v.x = e[q].x * scale; v.y = e[q].y * scale;
        __stcg(&dst[q * TPF], v);
    }
}

template <typename T, int N1>
__global__ void __launch_bounds__(ColShape<N1>::THREADS, min_ctas_for<T>(ColShape<N1>::THREADS, ColShape<N1>::MIN_CTAS))
fftconv_cols_inv(ConvGeom g, const cpx<T> *__restrict__ scratch, const T *__restrict__ x, T *__restrict__ y,
                 int N2, int lgN, const cpx<T> *__restrict__ tw, const cpx<T> *__restrict__ tw_hi,
                 const cpx<T> *__restrict__ tw_lo, long long pair0, int ntiles) {
    using C = cpx<T>;
    using CS = ColShape<N1>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *buf = reinterpret_cast<C *>(smem_raw);
    C *stw = buf + ((FftShape<N1>::P > 0) ? CS::SMEM_ELEMS : 0);
    load_tw_smem<T, N1>(stw, tw, threadIdx.x, CS::THREADS);
    CtaGate gate;
    const int tiles_per_pair = N2 / CS::TC;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int pl = t / tiles_per_pair, tile = t - pl * tiles_per_pair;
        cols_inv_tile<T, N1>(g, scratch + (size_t)ADSP_ALIAS(pl) * ((size_t)N1 * N2), x, y, N2, lgN, stw, tw_hi, tw_lo, pair0 + pl, tile, buf,
                             threadIdx.x, gate, true);
    }
}

// ------------------------------------------------------------------------------------------
// Pairwise FFT correlation (correlate.go:16-28 evaluated as one transform per pair instead of a
// generic long-kernel convolution): z = a + i*reverse(b) -> Z; for real a, b
//     FFT(a)*FFT(rev b) = (Z[k]^2 - conj(Z[N-k])^2) / (4i),
// and two pairs share one inverse transform (Q = P_A + i*P_B, outputs in re / im).
// corr_cols_fwd: forward column pass of one pair; grid = (N2/TC, pairs).
template <typename T, int N1>
__global__ void __launch_bounds__(ColShape<N1>::THREADS, min_ctas_for<T>(ColShape<N1>::THREADS, ColShape<N1>::MIN_CTAS))
corr_cols_fwd(const T *__restrict__ a, long long n, long long a_stride, const T *__restrict__ b, long long m,
              long long b_stride, cpx<T> *__restrict__ scratch, int N2, int lgN, const cpx<T> *__restrict__ tw,
              const cpx<T> *__restrict__ tw_hi, const cpx<T> *__restrict__ tw_lo, long long pair0, int reverse_b) {
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
    ColAddr<TC> addr{c};
    const unsigned maskN = (1u << lgN) - 1u;
    const C tw_base = twiddle_n<T>(tw_hi, tw_lo, ((unsigned)n2 * (unsigned)j) & maskN);
    const C tw_rho = twiddle_n<T>(tw_hi, tw_lo, ((unsigned)n2 * (unsigned)TPF) & maskN);
    const long long pair = pair0 + blockIdx.y;
    const T *ap = a + pair * a_stride;
    const T *bp = b + pair * b_stride;
    C e[16];
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const long long i = (long long)(j + q * TPF) * N2 + n2;
        e[q].x = (i < n) ? ld_stream(ap + i) : (T)0;
        e[q].y = (i < m) ? ld_stream(bp + (reverse_b ? (m - 1 - i) : i)) : (T)0;   // reverse(b): correlate.go:22-25; plain b: deconvolve.go:121-123
    }
    CtaGate gate;
    cta_fft<T, N1, false, false, true>(e, buf, addr, stw, j, gate);
    apply_geometric16<false>(e, tw_base, tw_rho);
    C *dst = scratch + (size_t)blockIdx.y * ((size_t)N1 * N2) + n2;
#pragma unroll
    for (int r = 0; r < 16; r++) __stcg(&dst[(size_t)(j + r * TPF) * N2], e[r]);
}

// ------------------------------------------------------------------------------------------
// Fused correlation rows: forward row transforms of the packed spectrum Z = FFT(a + i*reverse(b)), the spectral product
//     P[k] = (Z[k]^2 - conj(Z[N-k])^2) / (4i)            (= FFT(a)[k] * FFT(reverse b)[k], both real sequences)
// of TWO pairs (Q = P_A + i*P_B: one inverse transform serves both, results in re / im) and the inverse row transforms,
// in one launch -- what corr rows-forward + corr_pointwise + rows-inverse did in three with the spectrum making two extra
// round trips through L2 / HBM.  The mirror bin N-k of row k1 lives in row N1-k1 at position N2-1-k2 (row 0: same row,
// position (N2-k2) mod N2), so a CTA transforms the two partner rows side by side (2 * L/16 threads) and exchanges the
// spectra through the shared-memory buffers the transforms have just finished with.  Q is written over pair A's rows
// (the CTA owns rows k1 and N1-k1 of both pairs).  grid = (N1/2, pairs of pairs).
template <typename T, int L>
__global__ void __launch_bounds__(2 * (L / 16), (2 * (L / 16) <= 128) ? 4 : ((2 * (L / 16) <= 256) ? 2 : 1))
corr_rows_fused(cpx<T> *__restrict__ Z, int npairs, int N1, T scale, const cpx<T> *__restrict__ tw) {
    using C = cpx<T>;
    using Sh = FftShape<L>;
    constexpr int TPF = Sh::TPF;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *buf = reinterpret_cast<C *>(smem_raw);
    C *stw = buf + 2 * L;
    load_tw_smem<T, L>(stw, tw, threadIdx.x, 2 * TPF);
    const int grp = threadIdx.x / TPF;            // 0: row k1, 1: its partner row
    const int j = threadIdx.x % TPF;
    const int task = blockIdx.x;                  // 0: rows 0 and N1/2 (each its own partner), t >= 1: rows t and N1 - t
    const int k1 = (task == 0) ? (grp == 0 ? 0 : N1 / 2) : (grp == 0 ? task : N1 - task);
    const bool self = task == 0;                  // partner spectrum sits in this thread's own row buffer
    const bool rule_a = self && grp == 0;         // row 0: mirror of k2 is (L - k2) mod L; every other row: L - 1 - k2
    const size_t Nel = (size_t)N1 * L;
    const int pa_idx = 2 * blockIdx.y, pb_idx = 2 * blockIdx.y + 1;
    RowAddr<T, Sh::R0> addr{grp * L};
    RowAddr<T, Sh::R0> paddr{(self ? grp : 1 - grp) * L};
    CtaGate gate;
    C *rowA = Z + (size_t)pa_idx * Nel + (size_t)k1 * L + j;
    C e[16];
    // one pair: forward transform of this thread's row, product with the mirror bins -> e[] = P[k1 + N1*k2], k2 = j + q*TPF
    auto spectrum_product = [&](const C *row) {
#pragma unroll
        for (int q = 0; q < 16; q++) e[q] = __ldcg(&row[q * TPF]);
        cta_fft<T, L, false>(e, buf, addr, stw, j, gate);
        // every thread parks its spectrum in the slots it alone read in the last pass, then reads the mirror bins
#pragma unroll
        for (int q = 0; q < 16; q++) buf[addr.at(j + q * TPF, Sh::P - 1)] = e[q];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int k2 = j + q * TPF;
            const int m2 = rule_a ? ((L - k2) & (L - 1)) : (L - 1 - k2);
            const C zm = buf[paddr.at(m2, Sh::P - 1)];
            const C zk = e[q];
            const T ar = zk.x * zk.x - zk.y * zk.y, ai = 2 * zk.x * zk.y;      // zk^2
            const T br = zm.x * zm.x - zm.y * zm.y, bi = -2 * zm.x * zm.y;     // conj(zm)^2
            const T dr = ar - br, di = ai - bi;                                  // (dr + i di) / (4i) = (di - i dr) / 4
            e[q].x = di * (scale * (T)0.25);
            e[q].y = -dr * (scale * (T)0.25);
        }
    };
    spectrum_product(rowA);
    if (pb_idx < npairs) {
        // park P_A in its own rows (L2), transform pair B, then Q = P_A + i*P_B
#pragma unroll
        for (int q = 0; q < 16; q++) __stcg(&rowA[q * TPF], e[q]);
        __syncthreads();                          // everyone has read the mirror bins of pair A: the buffers are free again
        spectrum_product(Z + (size_t)pb_idx * Nel + (size_t)k1 * L + j);
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const C pa = __ldcg(&rowA[q * TPF]);
            const C pb = e[q];
            e[q].x = pa.x - pb.y;
            e[q].y = pa.y + pb.x;
        }
    }
    cta_fft<T, L, true>(e, buf, addr, stw, j, gate);     // (its first barrier also covers the mirror reads above)
#pragma unroll
    for (int q = 0; q < 16; q++) __stcg(&rowA[q * TPF], e[q]);
}

// Inverse column pass of the correlation with the peak search folded into its epilogue (FindPeak, correlate.go:200-216:
// signed maximum, first index wins, NaN never wins): every CTA leaves one (value, index) candidate per real output block
// in part_v / part_i [pair][tile]; peak_final_kernel reduces them.  `slot_mult`: Q of pairs (2q, 2q+1) sits in the scratch
// slot of pair 2q.  STORE = false skips the output stores (peak lags only: the correlation itself never reaches HBM).
template <typename T, int N1, bool STORE>
__global__ void __launch_bounds__(ColShape<N1>::THREADS, min_ctas_for<T>(ColShape<N1>::THREADS, ColShape<N1>::MIN_CTAS > 2 ? ColShape<N1>::MIN_CTAS - 1 : ColShape<N1>::MIN_CTAS))
corr_cols_inv_peak(ConvGeom g, const cpx<T> *__restrict__ scratch, int slot_mult, T *__restrict__ y, int N2, int lgN, const cpx<T> *__restrict__ tw,
                   const cpx<T> *__restrict__ tw_hi, const cpx<T> *__restrict__ tw_lo, long long pair0, T *__restrict__ part_v,
                   long long *__restrict__ part_i, T *__restrict__ first_v) {
    using C = cpx<T>;
    using CS = ColShape<N1>;
    constexpr int TPF = CS::TPF, TC = CS::TC, NW = CS::THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *buf = reinterpret_cast<C *>(smem_raw);
    C *stw = buf + ((FftShape<N1>::P > 0) ? CS::SMEM_ELEMS : 0);
    load_tw_smem<T, N1>(stw, tw, threadIdx.x, CS::THREADS);
    __shared__ T sv[2][NW > 0 ? NW : 1];
    __shared__ long long si[2][NW > 0 ? NW : 1];
    const int tid = threadIdx.x;
    const int c = tid % TC, j = tid / TC;
    const int tile = blockIdx.x, pl = blockIdx.y;
    const int n2 = tile * TC + c;
    ColAddr<TC> addr{c};
    const unsigned maskN = (1u << lgN) - 1u;
    const C tw_base = twiddle_n<T>(tw_hi, tw_lo, ((unsigned)n2 * (unsigned)j) & maskN);
    const C tw_rho = twiddle_n<T>(tw_hi, tw_lo, ((unsigned)n2 * (unsigned)TPF) & maskN);
    const C *src = scratch + (size_t)pl * (size_t)slot_mult * ((size_t)N1 * N2) + n2;
    C e[16];
    const uint64_t drop = l2_policy_drop();
#pragma unroll
    for (int q = 0; q < 16; q++) e[q] = ld_scratch(&src[(size_t)(j + q * TPF) * N2], drop);
    CtaGate gate;
    apply_geometric16<true>(e, tw_base, tw_rho);
    cta_fft<T, N1, true, true, false>(e, buf, addr, stw, j, gate);
    const long long pair = pair0 + pl;
    const BlockIO<T> a = block_io<T>(g, (const T *)nullptr, y, 2 * pair);
    const BlockIO<T> b = block_io<T>(g, (const T *)nullptr, y, 2 * pair + 1);
    T va = (T)0, vb = (T)0;
    long long ia = -1, ib = -1;
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const long long o = (long long)(j + r * TPF) * N2 + n2;
        if (o < a.cnt) {
            const T x = e[r].x;
            if (STORE) __stcs(a.out + o, x);
            if (x == x && (ia < 0 || x > va)) { va = x; ia = o; }
            if (o == 0) first_v[2 * pair] = x;
        }
        if (o < b.cnt) {
            const T x = e[r].y;
            if (STORE) __stcs(b.out + o, x);
            if (x == x && (ib < 0 || x > vb)) { vb = x; ib = o; }
            if (o == 0) first_v[2 * pair + 1] = x;
        }
    }
    for (int o = 16; o; o >>= 1) {
        peak_combine(va, ia, __shfl_xor_sync(0xffffffffu, va, o), __shfl_xor_sync(0xffffffffu, ia, o));
        peak_combine(vb, ib, __shfl_xor_sync(0xffffffffu, vb, o), __shfl_xor_sync(0xffffffffu, ib, o));
    }
    if ((tid & 31) == 0) { sv[0][tid >> 5] = va; si[0][tid >> 5] = ia; sv[1][tid >> 5] = vb; si[1][tid >> 5] = ib; }
    __syncthreads();
    if (tid < 2) {
        T v = sv[tid][0];
        long long i = si[tid][0];
        for (int w = 1; w < NW; w++) peak_combine(v, i, sv[tid][w], si[tid][w]);
        const long long blk = 2 * pair + tid;
        if (blk < g.total_blocks) {
            part_v[blk * gridDim.x + tile] = v;
            part_i[blk * gridDim.x + tile] = i;
        }
    }
}

// Regularised / naive spectral division for deconvolution (deconvolve.go:143-151, 216-220, 304-308), in four-step order,
// in place.  Z is the transform of z = signal + i*kernel; for real signal and kernel
//     S[k] = (Z[k] + conj(Z[N-k])) / 2,   H[k] = (Z[k] - conj(Z[N-k])) / (2i),
//     R[k] = S[k] * conj(H[k]) / (|H[k]|^2 + reg)        (reg < 0: naive S/H, bins with |H| < 1e-15 reported in *bad_bin)
// and R[N-k] = conj(R[k]) because the result is real.  One thread handles k and N-k.
template <typename T> __device__ __forceinline__ cpx<T> deconv_bin(cpx<T> zk, cpx<T> zm, T scale, T reg, long long k, long long *bad_bin) {
    const T sr = (zk.x + zm.x) * (T)0.5, si = (zk.y - zm.y) * (T)0.5;          // S = (zk + conj(zm))/2
    const T hr = (zk.y + zm.y) * (T)0.5, hi = (zm.x - zk.x) * (T)0.5;          // H = (zk - conj(zm))/(2i)
    T den = hr * hr + hi * hi;
    if (reg < (T)0) {
        if (sqrt((double)den) < 1e-15) { atomicMin((unsigned long long *)bad_bin, (unsigned long long)k); den = (T)1; }
    } else den += reg;
    cpx<T> r;
    r.x = (sr * hr + si * hi) / den * scale;                                     // S * conj(H) / den
    r.y = (si * hr - sr * hi) / den * scale;
    return r;
}

// grid.y = q: problems 2q (spectrum Z[2q]) and 2q+1 (Z[2q+1], absent when nprob is odd) share one inverse transform:
// Q[q] = R_A + i*R_B; both R are Hermitian (real results), so Q[N-k] = conj(R_A[k]) + i*conj(R_B[k]).
template <typename T>
__global__ void deconv_pointwise(const cpx<T> *Z, cpx<T> *Q, int nprob, int N1, int N2, T scale, T reg,
                                 long long *bad_bin) {
    using C = cpx<T>;
    const long long N = (long long)N1 * N2;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N) return;
    const long long k1 = idx / N2, k2 = idx - k1 * N2;
    const long long k = k1 + (long long)N1 * k2;
    const long long km = (N - k) & (N - 1);
    if (k > km) return;
    const long long m1 = km & (N1 - 1), m2 = km / N1;
    const long long midx = m1 * N2 + m2;
    const int q = blockIdx.y;
    const C *ZA = Z + (size_t)(2 * q) * N;
    const C ra = deconv_bin<T>(__ldcg(&ZA[idx]), __ldcg(&ZA[midx]), scale, reg, k, bad_bin);
    C rb; rb.x = (T)0; rb.y = (T)0;
    if (2 * q + 1 < nprob) {
        const C *ZB = ZA + N;
        rb = deconv_bin<T>(__ldcg(&ZB[idx]), __ldcg(&ZB[midx]), scale, reg, k, bad_bin);
    }
    C *Qq = Q + (size_t)q * N;
    C qk, qm;
    qk.x = ra.x - rb.y; qk.y = ra.y + rb.x;
    qm.x = ra.x + rb.y; qm.y = -ra.y + rb.x;
    __stcg(&Qq[idx], qk);
    if (midx != idx) __stcg(&Qq[midx], qm);
}

// Deconvolution with a transform that fits one CTA (N = L <= 4096): load z = signal + i*kernel, FFT, division through a
// mirrored read of the spectrum in shared memory, inverse FFT, keep the first out_len samples.  grid = problems.
template <typename T, int L>
__global__ void __launch_bounds__(FftShape<L>::TPF)
deconv_small(const T *__restrict__ sig, long long n, long long s_stride, const T *__restrict__ ker, long long m, long long k_stride,
             T *__restrict__ out, long long out_stride, long long out_len, T reg, const cpx<T> *__restrict__ tw, long long *bad_bin) {
    using C = cpx<T>;
    using Sh = FftShape<L>;
    constexpr int TPF = Sh::TPF;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *buf = reinterpret_cast<C *>(smem_raw);
    C *stw = buf + L;
    load_tw_smem<T, L>(stw, tw, threadIdx.x, TPF);
    const int j = threadIdx.x;
    const T *sp = sig + (long long)blockIdx.x * s_stride;
    const T *kp = ker + (long long)blockIdx.x * k_stride;
    C e[16];
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const long long i = j + q * TPF;
        e[q].x = (i < n) ? sp[i] : (T)0;
        e[q].y = (i < m) ? kp[i] : (T)0;
    }
    RowAddr<T, Sh::R0> addr{0};
    CtaGate gate;
    cta_fft<T, L, false>(e, buf, addr, stw, j, gate);
    __syncthreads();                                   // everyone is done reading the exchange buffer
#pragma unroll
    for (int q = 0; q < 16; q++) buf[j + q * TPF] = e[q];
    __syncthreads();
    const T scale = (T)1 / (T)L;
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const int k = j + q * TPF;
        const C zk = e[q], zm = buf[(L - k) & (L - 1)];
        const T sr = (zk.x + zm.x) * (T)0.5, si = (zk.y - zm.y) * (T)0.5;
        const T hr = (zk.y + zm.y) * (T)0.5, hi = (zm.x - zk.x) * (T)0.5;
        T den = hr * hr + hi * hi;
        if (reg < (T)0) {
            if (sqrt((double)den) < 1e-15) { atomicMin((unsigned long long *)bad_bin, (unsigned long long)k); den = (T)1; }
        } else den += reg;
        e[q].x = (sr * hr + si * hi) / den * scale;
        e[q].y = (si * hr - sr * hi) / den * scale;
    }
    __syncthreads();                                   // mirrored reads done before the inverse reuses the buffer
    cta_fft<T, L, true>(e, buf, addr, stw, j, gate);
    T *op = out + (long long)blockIdx.x * out_stride;
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const long long i = j + q * TPF;
        if (i < out_len) op[i] = e[q].x;
    }
}

// N = 1, 2, 4, 8: direct DFT by one thread per problem (same formulas); grid = problems, one thread each
template <typename T>
__global__ void deconv_tiny(const T *sig, long long n, long long s_stride, const T *ker, long long m, long long k_stride, T *out,
                            long long out_stride, long long out_len, int N, T reg, long long *bad_bin) {
    const T *sp = sig + (long long)blockIdx.x * s_stride;
    const T *kp = ker + (long long)blockIdx.x * k_stride;
    double rr[8], ri[8];
    for (int k = 0; k < N; k++) {
        double sr = 0, si = 0, hr = 0, hi = 0;
        for (int i = 0; i < N; i++) {
            const double c = cospi(2.0 * (double)((k * i) % N) / N), s = -sinpi(2.0 * (double)((k * i) % N) / N);
            const double xs = i < n ? (double)sp[i] : 0.0, xk = i < m ? (double)kp[i] : 0.0;
            sr += xs * c; si += xs * s; hr += xk * c; hi += xk * s;
        }
        double den = hr * hr + hi * hi;
        if (reg < (T)0) { if (sqrt(den) < 1e-15) { atomicMin((unsigned long long *)bad_bin, (unsigned long long)k); den = 1; } }
        else den += (double)reg;
        rr[k] = (sr * hr + si * hi) / den; ri[k] = (si * hr - sr * hi) / den;
    }
    T *op = out + (long long)blockIdx.x * out_stride;
    for (int i = 0; i < N && i < out_len; i++) {
        double acc = 0;
        for (int k = 0; k < N; k++) acc += rr[k] * cospi(2.0 * (double)((k * i) % N) / N) - ri[k] * sinpi(2.0 * (double)((k * i) % N) / N);
        op[i] = (T)(acc / N);
    }
}

}  // namespace adsp
