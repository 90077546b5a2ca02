// conv_kernels_mr.cuh -- mixed-radix column kernels: transform lengths N = (16*P) * N2 with P in {3,5,7,9}.
//
// Why: the reference sizes every block as a power of two (overlap_save.go:58-74, overlap_add.go:52-59), and a
// power-of-two-only engine wastes up to half of each transform.  With N = P * 2^k available, a one-shot
// convolution whose full result (n + K - 1 samples) fits one transform runs as a single zero-padded block
// with nothing discarded -- a 96 000-tap IR on 480 000 samples needs 9 * 2^16 = 589 824 points per channel
// instead of 2^19 + 2^18 = 786 432.  Results are unchanged (same linear convolution, overlap_save.go:146-251).
//
// Only the column transforms (length N1 = 16*P) are new; the rows stay power-of-two (fftconv_rows).
// Column transform, n1 = 16*i + j, k1 = kp + P*k16:
//     X[kp + P*k16] = sum_j W16^(j*k16) * [ W_N1^(j*kp) * sum_i x[16*i + j] * W_P^(i*kp) ]
// forward: P-point DFTs in registers (thread j of a column owns rows j, j+16, ...), twiddle, exchange through
// shared memory, then threads 0..P-1 of the column each run one radix-16 butterfly (outputs rows t + P*r);
// the inverse runs the same graph backwards.
#pragma once
#include "conv_kernels.cuh"

namespace adsp {

#ifndef ADSP_MR_TC
#define ADSP_MR_TC 8
#endif
#ifndef ADSP_MR_MIN_CTAS
#define ADSP_MR_MIN_CTAS (ADSP_MR_TC == 8 ? 5 : 2)   // 5 x 128 threads per SM: 102 registers, measured +3.7 % over 4
#endif
template <int P> struct ColShapeMR {
    static constexpr int N1 = 16 * P;
    static constexpr int TC = ADSP_MR_TC;        // columns per tile
    static constexpr int THREADS = 16 * TC;      // 16 threads per column
    static constexpr int SMEM_ELEMS = N1 * TC;
    static constexpr int TW_ENTRIES = 16 * P;    // W_N1^(j*kp) at [kp*16 + j]
};

// W_N^m for any N that is a multiple of 1024: hi[m >> 10] * lo[m & 1023]
template <typename T>
__device__ __forceinline__ cpx<T> twiddle_any(const cpx<T> *__restrict__ tw_hi, const cpx<T> *__restrict__ tw_lo, unsigned m) {
    return cmul(__ldg(&tw_hi[m >> 10]), __ldg(&tw_lo[m & 1023u]));
}

template <typename T, int P>
__global__ void __launch_bounds__(ColShapeMR<P>::THREADS, ADSP_MR_MIN_CTAS)
fftconv_cols_fwd_mr(ConvGeom g, const T *__restrict__ x, cpx<T> *__restrict__ scratch, int N2, unsigned N,
                    const cpx<T> *__restrict__ tw, const cpx<T> *__restrict__ tw_hi, const cpx<T> *__restrict__ tw_lo,
                    long long pair0, int ntiles) {
    using C = cpx<T>;
    using CS = ColShapeMR<P>;
    constexpr int TC = CS::TC, N1 = CS::N1;
    __shared__ __align__(16) C buf[CS::SMEM_ELEMS];
    __shared__ __align__(16) C stw[CS::TW_ENTRIES];
    for (int i = threadIdx.x; i < CS::TW_ENTRIES; i += CS::THREADS) stw[i] = tw[i];
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
        if (j < P) {   // seeds of the four-step twiddle W_N^(n2*k1), k1 = j + P*r; fetched first, used last
            tw_base = twiddle_any<T>(tw_hi, tw_lo, ((unsigned)n2 * (unsigned)j) % N);
            tw_rho = twiddle_any<T>(tw_hi, tw_lo, ((unsigned)n2 * (unsigned)P) % N);
        }
        const BlockIO<T> a = block_io<T>(g, x, (T *)nullptr, 2 * (pair0 + pl));
        const BlockIO<T> b = block_io<T>(g, x, (T *)nullptr, 2 * (pair0 + pl) + 1);
        C e[P];
#pragma unroll
        for (int i = 0; i < P; i++) {
            const long long idx = (long long)(j + 16 * i) * N2 + n2;
            e[i].x = (idx >= a.lo && idx < a.hi) ? ld_stream(a.in + idx) : (T)0;
            e[i].y = (idx >= b.lo && idx < b.hi) ? ld_stream(b.in + idx) : (T)0;
        }
        odd_dft<P, false>(e);
#pragma unroll
        for (int kp = 1; kp < P; kp++) e[kp] = cmul_tw<false>(e[kp], stw[kp * 16 + j]);
        __syncthreads();                                   // previous tile's readers are done with buf
#pragma unroll
        for (int kp = 0; kp < P; kp++) buf[(kp * 16 + j) * TC + c] = e[kp];
        __syncthreads();
        if (j < P) {
            C f[16];
#pragma unroll
            for (int jj = 0; jj < 16; jj++) f[jj] = buf[(j * 16 + jj) * TC + c];
            Dft<16, 1, false, C>::run(&f[0]);
            apply_geometric16<false>(f, tw_base, tw_rho);
            C *dst = scratch + (size_t)ADSP_ALIAS(pl) * pair_elems + n2;
#pragma unroll
            for (int r = 0; r < 16; r++) st_scratch(&dst[(size_t)(j + P * r) * N2], f[r], keep);
        }
    }
}

template <typename T, int P>
__global__ void __launch_bounds__(ColShapeMR<P>::THREADS, ADSP_MR_MIN_CTAS)
fftconv_cols_inv_mr(ConvGeom g, const cpx<T> *__restrict__ scratch, const T *__restrict__ x, T *__restrict__ y, int N2, unsigned N,
                    const cpx<T> *__restrict__ tw, const cpx<T> *__restrict__ tw_hi, const cpx<T> *__restrict__ tw_lo,
                    long long pair0, int ntiles) {
    using C = cpx<T>;
    using CS = ColShapeMR<P>;
    constexpr int TC = CS::TC, N1 = CS::N1;
    __shared__ __align__(16) C buf[CS::SMEM_ELEMS];
    __shared__ __align__(16) C stw[CS::TW_ENTRIES];
    for (int i = threadIdx.x; i < CS::TW_ENTRIES; i += CS::THREADS) stw[i] = tw[i];
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
        if (j < P) {
            const C tw_base = twiddle_any<T>(tw_hi, tw_lo, ((unsigned)n2 * (unsigned)j) % N);
            const C tw_rho = twiddle_any<T>(tw_hi, tw_lo, ((unsigned)n2 * (unsigned)P) % N);
            const C *src = scratch + (size_t)ADSP_ALIAS(pl) * pair_elems + n2;
            C f[16];
#pragma unroll
            for (int r = 0; r < 16; r++) f[r] = ld_scratch(&src[(size_t)(j + P * r) * N2], drop);
            apply_geometric16<true>(f, tw_base, tw_rho);
            Dft<16, 1, true, C>::run(&f[0]);
#pragma unroll
            for (int jj = 0; jj < 16; jj++) buf[(j * 16 + jj) * TC + c] = f[jj];
        }
        __syncthreads();
        C e[P];
#pragma unroll
        for (int kp = 0; kp < P; kp++) e[kp] = buf[(kp * 16 + j) * TC + c];
#pragma unroll
        for (int kp = 1; kp < P; kp++) e[kp] = cmul_tw<true>(e[kp], stw[kp * 16 + j]);
        odd_dft<P, true>(e);
        const BlockIO<T> a = block_io<T>(g, x, y, 2 * (pair0 + pl));
        const BlockIO<T> b = block_io<T>(g, x, y, 2 * (pair0 + pl) + 1);
#pragma unroll
        for (int i = 0; i < P; i++) {
            const long long o = (long long)(j + 16 * i) * N2 + n2 - g.D;
            if (o >= 0) {
                if (g.accumulate) {
                    if (o < a.cnt) a.out[o] += e[i].x;
                    if (o < b.cnt) b.out[o] += e[i].y;
                } else {
                    if (o < a.cnt) __stcs(a.out + o, e[i].x);
                    if (o < b.cnt) __stcs(b.out + o, e[i].y);
                }
            }
        }
    }
}

}  // namespace adsp
