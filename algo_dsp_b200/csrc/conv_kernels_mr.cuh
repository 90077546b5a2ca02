// conv_kernels_mr.cuh -- mixed-radix column kernels: transform lengths N = (16*M) * N2, M = P or 2P, P in {3,5,7,9}.
//
// Why: the reference sizes every block as a power of two (overlap_save.go:58-74, overlap_add.go:52-59), and a
// power-of-two-only engine wastes up to half of each transform.  With N = P * 2^k available, a one-shot
// convolution whose full result (n + K - 1 samples) fits one transform runs as a single zero-padded block
// with nothing discarded -- a 96 000-tap IR on 480 000 samples needs 9 * 2^16 = 589 824 points per channel
// instead of 2^19 + 2^18 = 786 432.  Results are unchanged (same linear convolution, overlap_save.go:146-251).
//
// Only the column transforms (length N1 = 16*M) are new; the rows stay power-of-two (fftconv_rows).  M = 2P
// (prime-factor 2 x P DFT in registers) exists so that P * 2^16 keeps 2048-point rows, the fastest row shape.
// Column transform, n1 = 16*i + j, k1 = km + M*k16:
//     X[km + M*k16] = sum_j W16^(j*k16) * [ W_N1^(j*km) * sum_i x[16*i + j] * W_M^(i*km) ]
// forward: M-point DFTs in registers (thread j < 16 of a column owns rows j, j+16, ...), twiddle, exchange through
// shared memory, then threads 0..M-1 of the column each run one radix-16 butterfly (outputs rows t + M*r);
// the inverse runs the same graph backwards.  A column is served by max(16, M) threads.
#pragma once
#include "conv_kernels.cuh"

namespace adsp {

#ifndef ADSP_MR_TC
#define ADSP_MR_TC 8
#endif
#ifndef ADSP_MR_MIN_CTAS
#define ADSP_MR_MIN_CTAS (ADSP_MR_TC == 8 ? 5 : 2)   // 5 x 128 threads per SM: 102 registers, measured +3.7 % over 4
#endif
// 1: read the column twiddles W_N1^(j*km) straight from global memory through L1 (4.6 KB, hot on every SM) instead of
// staging them in shared memory behind a barrier at the start of every CTA (measured: 99.0 vs 101.2 Gsamples/s, so off)
#ifndef ADSP_MR_TW_LDG
#define ADSP_MR_TW_LDG 0
#endif
#ifndef ADSP_MR_WIDE_CTAS
#define ADSP_MR_WIDE_CTAS 3
#endif
template <int M> struct ColShapeMR {
    static constexpr int N1 = 16 * M;
    static constexpr int TC = ADSP_MR_TC;        // columns per tile
    static constexpr int TPC = M > 16 ? M : 16;  // threads per column
    static constexpr int THREADS = TPC * TC;
    static constexpr int SMEM_ELEMS = N1 * TC;
    static constexpr int TW_ENTRIES = 16 * M;    // W_N1^(j*km) at [km*16 + j]
    static constexpr int MIN_CTAS = M > 16 ? (ADSP_MR_TC <= 8 ? ADSP_MR_WIDE_CTAS : 2) : ADSP_MR_MIN_CTAS;
};

// W_N^m for any N that is a multiple of 1024: hi[m >> 10] * lo[m & 1023]
template <typename T>
__device__ __forceinline__ cpx<T> twiddle_any(const cpx<T> *__restrict__ tw_hi, const cpx<T> *__restrict__ tw_lo, unsigned m) {
    return cmul(__ldg(&tw_hi[m >> 10]), __ldg(&tw_lo[m & 1023u]));
}

template <typename T, int M>
__global__ void __launch_bounds__(ColShapeMR<M>::THREADS, min_ctas_for<T>(ColShapeMR<M>::THREADS > 128 ? 128 : ColShapeMR<M>::THREADS, ColShapeMR<M>::MIN_CTAS))
fftconv_cols_fwd_mr(ConvGeom g, const T *__restrict__ x, cpx<T> *__restrict__ scratch, int N2, unsigned N,
                    const cpx<T> *__restrict__ tw, const cpx<T> *__restrict__ tw_hi, const cpx<T> *__restrict__ tw_lo,
                    long long pair0, int ntiles) {
    using C = cpx<T>;
    using CS = ColShapeMR<M>;
    constexpr int TC = CS::TC, N1 = CS::N1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *buf = reinterpret_cast<C *>(smem_raw);
    C *stw = buf + CS::SMEM_ELEMS;
    if (!ADSP_MR_TW_LDG) { for (int i = threadIdx.x; i < CS::TW_ENTRIES; i += CS::THREADS) stw[i] = tw[i]; }
    const int c = threadIdx.x % TC;
    const int j = threadIdx.x / TC;
    const int tiles_per_pair = N2 / TC;
    const size_t pair_elems = (size_t)N1 * N2;
    const uint64_t keep = l2_policy_keep();
    __syncthreads();
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int pl = t / tiles_per_pair, tile = t - pl * tiles_per_pair;
        const int n2 = tile * TC + c;
        C tw_base, tw_rho;
        if (j < M) {   // seeds of the four-step twiddle W_N^(n2*k1), k1 = j + M*r; fetched first, used last
            tw_base = twiddle_any<T>(tw_hi, tw_lo, ((unsigned)n2 * (unsigned)j) % N);
            tw_rho = twiddle_any<T>(tw_hi, tw_lo, ((unsigned)n2 * (unsigned)M) % N);
        }
        if (j < 16) {
            const TileIn<T> a = tile_in<T>(block_io<T>(g, x, (T *)nullptr, 2 * (pair0 + pl)), (long long)N);
            const TileIn<T> b = tile_in<T>(block_io<T>(g, x, (T *)nullptr, 2 * (pair0 + pl) + 1), (long long)N);
            C e[M];
            const int idx0 = j * N2 + n2;
#pragma unroll
            for (int i = 0; i < M; i++) {
                const int idx = idx0 + i * 16 * N2;
                e[i].x = ((unsigned)(idx - a.lo) < a.span) ? ld_stream(a.p + idx) : (T)0;
                e[i].y = ((unsigned)(idx - b.lo) < b.span) ? ld_stream(b.p + idx) : (T)0;
            }
            small_dft<M, false>(e);
#pragma unroll
            for (int km = 1; km < M; km++) e[km] = cmul_tw<false>(e[km], ADSP_MR_TW_LDG ? __ldg(&tw[km * 16 + j]) : stw[km * 16 + j]);
            // (the barrier at the end of the previous tile guarantees its readers are done with buf)
#pragma unroll
            for (int km = 0; km < M; km++) buf[(km * 16 + j) * TC + c] = e[km];
        }
        __syncthreads();
        if (j < M) {
            C f[16];
#pragma unroll
            for (int jj = 0; jj < 16; jj++) f[jj] = buf[(j * 16 + jj) * TC + c];
            Dft<16, 1, false, C>::run(&f[0]);
            apply_geometric16<false>(f, tw_base, tw_rho);
            C *dst = scratch + (size_t)ADSP_ALIAS(pl) * pair_elems + n2;
#pragma unroll
            for (int r = 0; r < 16; r++) st_scratch(&dst[(size_t)(j + M * r) * N2], f[r], keep);
        }
        __syncthreads();                                   // buf free for the next tile
    }
}

template <typename T, int M>
__global__ void __launch_bounds__(ColShapeMR<M>::THREADS, min_ctas_for<T>(ColShapeMR<M>::THREADS > 128 ? 128 : ColShapeMR<M>::THREADS, ColShapeMR<M>::MIN_CTAS))
fftconv_cols_inv_mr(ConvGeom g, const cpx<T> *__restrict__ scratch, const T *__restrict__ x, T *__restrict__ y, int N2, unsigned N,
                    const cpx<T> *__restrict__ tw, const cpx<T> *__restrict__ tw_hi, const cpx<T> *__restrict__ tw_lo,
                    long long pair0, int ntiles) {
    using C = cpx<T>;
    using CS = ColShapeMR<M>;
    constexpr int TC = CS::TC, N1 = CS::N1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *buf = reinterpret_cast<C *>(smem_raw);
    C *stw = buf + CS::SMEM_ELEMS;
    if (!ADSP_MR_TW_LDG) { for (int i = threadIdx.x; i < CS::TW_ENTRIES; i += CS::THREADS) stw[i] = tw[i]; }
    const int c = threadIdx.x % TC;
    const int j = threadIdx.x / TC;
    const int tiles_per_pair = N2 / TC;
    const size_t pair_elems = (size_t)N1 * N2;
    const uint64_t drop = l2_policy_drop();   // last use of these scratch lines
    __syncthreads();
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int pl = t / tiles_per_pair, tile = t - pl * tiles_per_pair;
        const int n2 = tile * TC + c;
        __syncthreads();                                   // previous tile's readers are done with buf
        if (j < M) {
            const C tw_base = twiddle_any<T>(tw_hi, tw_lo, ((unsigned)n2 * (unsigned)j) % N);
            const C tw_rho = twiddle_any<T>(tw_hi, tw_lo, ((unsigned)n2 * (unsigned)M) % N);
            const C *src = scratch + (size_t)ADSP_ALIAS(pl) * pair_elems + n2;
            C f[16];
#pragma unroll
            for (int r = 0; r < 16; r++) f[r] = ld_scratch(&src[(size_t)(j + M * r) * N2], drop);
            apply_geometric16<true>(f, tw_base, tw_rho);
            Dft<16, 1, true, C>::run(&f[0]);
#pragma unroll
            for (int jj = 0; jj < 16; jj++) buf[(j * 16 + jj) * TC + c] = f[jj];
        }
        __syncthreads();
        if (j >= 16) continue;                             // helper threads of wide columns (M > 16): both barriers of this tile passed
        C e[M];
#pragma unroll
        for (int km = 0; km < M; km++) e[km] = buf[(km * 16 + j) * TC + c];
#pragma unroll
        for (int km = 1; km < M; km++) e[km] = cmul_tw<true>(e[km], ADSP_MR_TW_LDG ? __ldg(&tw[km * 16 + j]) : stw[km * 16 + j]);
        small_dft<M, true>(e);
        const TileOut<T> a = tile_out<T>(block_io<T>(g, x, y, 2 * (pair0 + pl)), (long long)N);
        const TileOut<T> b = tile_out<T>(block_io<T>(g, x, y, 2 * (pair0 + pl) + 1), (long long)N);
        const int o0 = j * N2 + n2 - (int)g.D;
        const bool acc = g.accumulate != 0;
#pragma unroll
        for (int i = 0; i < M; i++) {
            const int o = o0 + i * 16 * N2;                 // o < 0 wraps to a huge unsigned: one compare does both bounds
            if ((unsigned)o < a.cnt) { if (acc) a.p[o] += e[i].x; else __stcs(a.p + o, e[i].x); }
            if ((unsigned)o < b.cnt) { if (acc) b.p[o] += e[i].y; else __stcs(b.p + o, e[i].y); }
        }
    }
}

}  // namespace adsp
