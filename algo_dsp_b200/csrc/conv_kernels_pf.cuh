// conv_kernels_pf.cuh -- software-pipelined ("prefetching") four-step kernels (sm_100a).
//
// Same three transforms as conv_kernels.cuh (forward columns -> rows with the spectral multiply ->
// inverse columns; reference loop dsp/conv/overlap_save.go:144-251), restructured so that no thread
// ever waits on a global-memory round trip inside a tile:
//   * CTAs are persistent (grid = resident CTAs) and walk tiles t = blockIdx.x, +gridDim.x, ...
//   * every thread owns a private 16-element landing zone in shared memory; the inputs of tile t+1 are
//     fetched into it with cp.async (LDGSTS, no destination registers) while tile t is being
//     transformed, so a tile starts with 16 conflict-free shared-memory reads instead of 16 strided
//     global loads.  The zone is private (the thread that issued a copy is the one that reads it), so
//     cp.async.wait_group is the only synchronisation it needs -- no extra barrier.
//   * the zero history / zero tail of overlap-save are zero-filling copies (src-size 0).
//   * the per-tile four-step twiddle seeds are loaded a tile ahead into registers.
// Steady-state measurements that motivated this: DESIGN.md section 7 (tools/steady_sweep.sh).
#pragma once
#include "conv_kernels.cuh"

namespace adsp {

#ifndef ADSP_PF_CTAS
#define ADSP_PF_CTAS 3          // resident 128-thread CTAs per SM (exchange buffer + landing zone = 64 KB each in fp64)
#endif

template <typename T, int N1> struct ColShapePF {
    using CS = ColShape<N1>;
    static constexpr int STAGE_ELEMS = CS::THREADS * 16;
    static constexpr int BUF_ELEMS = (FftShape<N1>::P > 0) ? CS::SMEM_ELEMS : 0;
    static constexpr size_t SMEM = ((size_t)BUF_ELEMS + STAGE_ELEMS + FftShape<N1>::TW_ENTRIES) * sizeof(cpx<T>);
    static constexpr int MIN_CTAS = (CS::THREADS <= 128) ? ADSP_PF_CTAS : 1;
};
template <typename T, int L> struct RowShapePF {
    static constexpr int THREADS = rows_cta_threads(L);
    static constexpr int ROWS = THREADS / FftShape<L>::TPF;
    static constexpr int STAGE_ELEMS = THREADS * 16;
    static constexpr size_t SMEM = ((size_t)ROWS * L + STAGE_ELEMS + FftShape<L>::TW_ENTRIES) * sizeof(cpx<T>);
    static constexpr int MIN_CTAS = (THREADS <= 128) ? ADSP_PF_CTAS : 1;
};

// four-step twiddle seeds of a column tile: W_N^(n2*j) and W_N^(n2*TPF), each as a (hi, lo) table pair
template <typename T>
__device__ __forceinline__ void fetch_tw4(cpx<T> (&w)[4], const cpx<T> *__restrict__ tw_hi, const cpx<T> *__restrict__ tw_lo,
                                          unsigned n2, unsigned j, unsigned tpf, unsigned maskN) {
    const unsigned mb = (n2 * j) & maskN, mr = (n2 * tpf) & maskN;
    w[0] = __ldg(&tw_hi[mb >> 10]); w[1] = __ldg(&tw_lo[mb & 1023u]);
    w[2] = __ldg(&tw_hi[mr >> 10]); w[3] = __ldg(&tw_lo[mr & 1023u]);
}

// ------------------------------------------------------------------------------------------
// Forward column tiles: x (two real blocks per pair) -> N1-point column FFTs -> four-step twiddle -> scratch
template <typename T, int N1>
__global__ void __launch_bounds__(ColShape<N1>::THREADS, ColShapePF<T, N1>::MIN_CTAS)
fftconv_cols_fwd_pf(ConvGeom g, const T *__restrict__ x, cpx<T> *__restrict__ scratch, int N2, int lgN,
                    const cpx<T> *__restrict__ tw, const cpx<T> *__restrict__ tw_hi, const cpx<T> *__restrict__ tw_lo,
                    long long pair0, int ntiles) {
    using C = cpx<T>;
    using CS = ColShape<N1>;
    using PF = ColShapePF<T, N1>;
    constexpr int TPF = CS::TPF, TC = CS::TC, THREADS = CS::THREADS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *buf = reinterpret_cast<C *>(smem_raw);
    C *stage = buf + PF::BUF_ELEMS + threadIdx.x;          // this thread's landing zone: stage[q * THREADS]
    C *stw = buf + PF::BUF_ELEMS + PF::STAGE_ELEMS;
    load_tw_smem<T, N1>(stw, tw, threadIdx.x, THREADS);
    const int c = threadIdx.x % TC;
    const int j = threadIdx.x / TC;
    ColAddr<TC> addr{c};
    const int tiles_per_pair = N2 / TC;
    const unsigned maskN = (1u << lgN) - 1u;
    const size_t pair_elems = (size_t)N1 * N2;

    auto issue = [&](int t) {
        const int pl = t / tiles_per_pair, tile = t - pl * tiles_per_pair;
        const int n2 = tile * TC + c;
        const BlockIO<T> a = block_io<T>(g, x, (T *)nullptr, 2 * (pair0 + pl));
        const BlockIO<T> b = block_io<T>(g, x, (T *)nullptr, 2 * (pair0 + pl) + 1);
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const long long i = (long long)(j + q * TPF) * N2 + n2;
            const bool va = (i >= a.lo && i < a.hi) && !ADSP_SKIP(3);
            const bool vb = (i >= b.lo && i < b.hi) && !ADSP_SKIP(3);
            T *dst = reinterpret_cast<T *>(&stage[q * THREADS]);
            cp_async_real_zfill<T>(dst, va ? a.in + i : x, va);
            cp_async_real_zfill<T>(dst + 1, vb ? b.in + i : x, vb);
        }
    };

    int t = blockIdx.x;
    C w4[4];
    if (t < ntiles) {
        issue(t);
        const int pl = t / tiles_per_pair, tile = t - pl * tiles_per_pair;
        fetch_tw4<T>(w4, tw_hi, tw_lo, (unsigned)(tile * TC + c), (unsigned)j, (unsigned)TPF, maskN);
    }
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();                                        // twiddle table visible to everyone
    CtaGate gate;
    for (; t < ntiles; t += gridDim.x) {
        const int tn = t + gridDim.x;
        cp_async_wait_all();
        C e[16];
#pragma unroll
        for (int q = 0; q < 16; q++) e[q] = stage[q * THREADS];
        auto prefetch_next = [&]() {                        // after pass 0: the landing zone has been consumed
            if (tn < ntiles) issue(tn);
            cp_async_commit();
        };
        cta_fft<T, N1, false, false, true, false>(e, buf, addr, stw, j, gate, NoHook(), prefetch_next);
        const C tw_base = cmul(w4[0], w4[1]), tw_rho = cmul(w4[2], w4[3]);
        const int pl = t / tiles_per_pair, tile = t - pl * tiles_per_pair;
        const int n2 = tile * TC + c;
        if (tn < ntiles) {                                  // next tile's seeds: a whole tile of latency cover
            const int pln = tn / tiles_per_pair, tilen = tn - pln * tiles_per_pair;
            fetch_tw4<T>(w4, tw_hi, tw_lo, (unsigned)(tilen * TC + c), (unsigned)j, (unsigned)TPF, maskN);
        }
        apply_geometric16<false>(e, tw_base, tw_rho);
        gate.d_end();
        if (!ADSP_SKIP(4)) {
            C *dst = scratch + (size_t)ADSP_ALIAS(pl) * pair_elems + n2;
#pragma unroll
            for (int r = 0; r < 16; r++) __stcg(&dst[(size_t)(j + r * TPF) * N2], e[r]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Row tiles, in place on scratch: N2-point FFT, multiply by the cached IR spectrum, N2-point inverse FFT.
template <typename T, int L>
__global__ void __launch_bounds__(rows_cta_threads(L), RowShapePF<T, L>::MIN_CTAS)
fftconv_rows_pf(cpx<T> *__restrict__ scratch, const cpx<T> *__restrict__ H, int N1, const cpx<T> *__restrict__ tw, int ntiles) {
    using C = cpx<T>;
    using Sh = FftShape<L>;
    using PF = RowShapePF<T, L>;
    constexpr int TPF = Sh::TPF, THREADS = PF::THREADS, ROWS = PF::ROWS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *buf = reinterpret_cast<C *>(smem_raw);
    C *stage = buf + ROWS * L + threadIdx.x;
    C *stw = buf + ROWS * L + PF::STAGE_ELEMS;
    load_tw_smem<T, L>(stw, tw, threadIdx.x, THREADS);
    const int row = threadIdx.x / TPF;
    const int j = threadIdx.x % TPF;
    RowAddr<T, Sh::R0> addr{row * L};
    const int tiles_per_pair = N1 / ROWS;
    const size_t pair_elems = (size_t)N1 * L;

    // element offset (inside scratch / inside H) of this thread's first point of tile t
    auto offsets = [&](int t, size_t &soff, size_t &hoff) {
        const int pl = t / tiles_per_pair, tile = t - pl * tiles_per_pair;
        hoff = ((size_t)tile * ROWS + row) * L + j;
        soff = (size_t)ADSP_ALIAS(pl) * pair_elems + hoff;
    };
    auto issue = [&](int t) {
        size_t soff, hoff;
        offsets(t, soff, hoff);
        if (!ADSP_SKIP(0)) {
#pragma unroll
            for (int q = 0; q < 16; q++) cp_async_elem(&stage[q * THREADS], &scratch[soff + q * TPF]);
        }
    };

    int t = blockIdx.x;
    if (t < ntiles) issue(t);
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    CtaGate gate;
    for (; t < ntiles; t += gridDim.x) {
        const int tn = t + gridDim.x;
        size_t soff, hoff;
        offsets(t, soff, hoff);
        cp_async_wait_all();
        C e[16];
#pragma unroll
        for (int q = 0; q < 16; q++) e[q] = stage[q * THREADS];
        // at the last forward pass: spectrum into the exchange slots this thread has just emptied (group 1),
        // then the next tile into the landing zone (group 2)
        auto prefetch = [&](C *b) {
            if (!ADSP_SKIP(2)) {
#pragma unroll
                for (int q = 0; q < 16; q++) cp_async_elem(&b[addr.at(j + q * TPF, Sh::P - 1)], &H[hoff + q * TPF]);
            }
            cp_async_commit();
            if (tn < ntiles) issue(tn);
            cp_async_commit();
        };
        cta_fft<T, L, false, false, true, false>(e, buf, addr, stw, j, gate, prefetch);
        cp_async_wait_group<1>();                           // spectrum landed; the next tile may still be in flight
        if (!ADSP_SKIP(2)) {
#pragma unroll
            for (int q = 0; q < 16; q++) e[q] = cmul(e[q], buf[addr.at(j + q * TPF, Sh::P - 1)]);
        }
        cta_fft<T, L, true, true, false, false>(e, buf, addr, stw, j, gate);
        if (!ADSP_SKIP(1)) {
#pragma unroll
            for (int q = 0; q < 16; q++) __stcg(&scratch[soff + q * TPF], e[q]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Inverse column tiles: scratch -> conj four-step twiddle -> N1-point inverse FFTs -> discard D -> two real blocks
template <typename T, int N1>
__global__ void __launch_bounds__(ColShape<N1>::THREADS, ColShapePF<T, N1>::MIN_CTAS)
fftconv_cols_inv_pf(ConvGeom g, const cpx<T> *__restrict__ scratch, const T *__restrict__ x, T *__restrict__ y, int N2, int lgN,
                    const cpx<T> *__restrict__ tw, const cpx<T> *__restrict__ tw_hi, const cpx<T> *__restrict__ tw_lo,
                    long long pair0, int ntiles) {
    using C = cpx<T>;
    using CS = ColShape<N1>;
    using PF = ColShapePF<T, N1>;
    constexpr int TPF = CS::TPF, TC = CS::TC, THREADS = CS::THREADS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *buf = reinterpret_cast<C *>(smem_raw);
    C *stage = buf + PF::BUF_ELEMS + threadIdx.x;
    C *stw = buf + PF::BUF_ELEMS + PF::STAGE_ELEMS;
    load_tw_smem<T, N1>(stw, tw, threadIdx.x, THREADS);
    const int c = threadIdx.x % TC;
    const int j = threadIdx.x / TC;
    ColAddr<TC> addr{c};
    const int tiles_per_pair = N2 / TC;
    const unsigned maskN = (1u << lgN) - 1u;
    const size_t pair_elems = (size_t)N1 * N2;

    auto issue = [&](int t) {
        const int pl = t / tiles_per_pair, tile = t - pl * tiles_per_pair;
        const C *src = scratch + (size_t)ADSP_ALIAS(pl) * pair_elems + (tile * TC + c);
        if (!ADSP_SKIP(5)) {
#pragma unroll
            for (int q = 0; q < 16; q++) cp_async_elem(&stage[q * THREADS], &src[(size_t)(j + q * TPF) * N2]);
        }
    };

    int t = blockIdx.x;
    C w4[4];
    if (t < ntiles) {
        issue(t);
        const int pl = t / tiles_per_pair, tile = t - pl * tiles_per_pair;
        fetch_tw4<T>(w4, tw_hi, tw_lo, (unsigned)(tile * TC + c), (unsigned)j, (unsigned)TPF, maskN);
    }
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    CtaGate gate;
    for (; t < ntiles; t += gridDim.x) {
        const int tn = t + gridDim.x;
        cp_async_wait_all();
        C e[16];
#pragma unroll
        for (int q = 0; q < 16; q++) e[q] = stage[q * THREADS];
        const C tw_base = cmul(w4[0], w4[1]), tw_rho = cmul(w4[2], w4[3]);
        apply_geometric16<true>(e, tw_base, tw_rho);       // consumes every landing-zone value
        if (tn < ntiles) {
            issue(tn);
            const int pln = tn / tiles_per_pair, tilen = tn - pln * tiles_per_pair;
            fetch_tw4<T>(w4, tw_hi, tw_lo, (unsigned)(tilen * TC + c), (unsigned)j, (unsigned)TPF, maskN);
        }
        cp_async_commit();
        cta_fft<T, N1, true, true, false, false>(e, buf, addr, stw, j, gate);

        const int pl = t / tiles_per_pair, tile = t - pl * tiles_per_pair;
        const int n2 = tile * TC + c;
        const BlockIO<T> a = block_io<T>(g, x, y, 2 * (pair0 + pl));
        const BlockIO<T> b = block_io<T>(g, x, y, 2 * (pair0 + pl) + 1);
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
}

}  // namespace adsp
