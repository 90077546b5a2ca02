// aux_kernels.cuh -- direct (time-domain) convolution, peak search and small helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace adsp {

// ------------------------------------------------------------------------------------------
// Sliding-window direct convolution.  Replaces the n*m MAC loop of the reference
// (dsp/conv/conv.go:117-154: dst[i+j] += a[i]*b[j]).  For each output k the products are
// accumulated in the reference's order (input index i ascending, i.e. tap index descending).
// CTA = 128 threads, each thread produces RO consecutive outputs from a register window;
// the signal tile and the tap chunk are staged in shared memory.
constexpr int DIRECT_THREADS = 128;
#ifndef ADSP_DIRECT_RO
#define ADSP_DIRECT_RO 8
#endif
constexpr int DIRECT_RO = ADSP_DIRECT_RO;
constexpr int DIRECT_PAD_SHIFT = (DIRECT_RO == 16) ? 4 : 3;   // one pad word per RO samples: per-thread windows hit distinct banks
constexpr int DIRECT_TILE = DIRECT_THREADS * DIRECT_RO;  // outputs per CTA
constexpr int DIRECT_MC = 64;                            // taps per staged chunk

// taps of one staged chunk, highest first; `sa` is the padded signal tile, `sb` the taps
template <typename T, bool FUSED, bool GUARD>
__device__ __forceinline__ void direct_chunk(T (&acc)[DIRECT_RO], const T *sa, const T *sb, int base, int lim) {
#define ADSP_SA(i) sa[(i) + ((i) >> DIRECT_PAD_SHIFT)]
    T w[DIRECT_RO];
#pragma unroll
    for (int r = 0; r < DIRECT_RO - 1; r++) w[r + 1] = ADSP_SA(base + r);
#pragma unroll
    for (int jj = DIRECT_MC - 1; jj >= 0; jj--) {
#pragma unroll
        for (int r = 0; r < DIRECT_RO - 1; r++) w[r] = w[r + 1];
        w[DIRECT_RO - 1] = ADSP_SA(base + DIRECT_RO - 1 + (DIRECT_MC - 1) - jj);
        if (GUARD && jj >= lim) continue;
        const T bj = sb[jj];
#pragma unroll
        for (int r = 0; r < DIRECT_RO; r++) {
            if (FUSED) acc[r] = fma(w[r], bj, acc[r]);
            else if (sizeof(T) == 8) acc[r] = __dadd_rn((double)acc[r], __dmul_rn((double)w[r], (double)bj));   // two roundings, as Go on amd64
            else acc[r] = __fadd_rn((float)acc[r], __fmul_rn((float)w[r], (float)bj));
        }
    }
#undef ADSP_SA
}

// Output epilogue of a tile: a thread owns RO CONSECUTIVE outputs, so storing them directly makes every warp store touch
// 32 segments of RO*sizeof(T) bytes (16 LSU wavefronts per fp64 instruction; half of the kernel's wavefronts, ncu round 2).
// The outputs are transposed through the (now idle) signal tile instead: padded writes and unit-stride reads are both
// conflict free, and every warp store covers 32 consecutive samples (2 wavefronts).  No alignment requirement.
template <typename T>
__device__ __forceinline__ void direct_store_tile(const T (&acc)[DIRECT_RO], T *sa, T *__restrict__ oc, long long k0, long long out_len, int t) {
#define ADSP_SA(i) sa[(i) + ((i) >> DIRECT_PAD_SHIFT)]
    __syncthreads();                                   // every thread has finished reading its window
#pragma unroll
    for (int r = 0; r < DIRECT_RO; r++) ADSP_SA(t * DIRECT_RO + r) = acc[r];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < DIRECT_RO; r++) {
        const int i = r * DIRECT_THREADS + t;
        const long long k = k0 + i;
        if (k < out_len) __stcs(oc + k, ADSP_SA(i));
    }
#undef ADSP_SA
}

template <typename T, bool FUSED>
__global__ void __launch_bounds__(DIRECT_THREADS)
direct_conv_kernel(const T *__restrict__ a, long long n, long long a_stride,
                   const T *__restrict__ b, long long m, long long b_stride,
                   T *__restrict__ out, long long out_stride, long long tiles_per_ch) {
    // padded by one word per 8 so that the per-thread windows (stride RO=8) hit distinct banks
    __shared__ T sa[(DIRECT_TILE + DIRECT_MC) + ((DIRECT_TILE + DIRECT_MC) >> DIRECT_PAD_SHIFT) + 1];
    __shared__ T sb[DIRECT_MC];
#define ADSP_SA(i) sa[(i) + ((i) >> DIRECT_PAD_SHIFT)]
    const long long ch = blockIdx.x / tiles_per_ch;
    const long long tile = blockIdx.x - ch * tiles_per_ch;
    const long long k0 = tile * DIRECT_TILE;
    const T *ac = a + ch * a_stride;
    const T *bc = b + ch * b_stride;
    const long long out_len = n + m - 1;
    const int t = threadIdx.x;

    T acc[DIRECT_RO];
#pragma unroll
    for (int r = 0; r < DIRECT_RO; r++) acc[r] = (T)0;

    // chunks of taps, highest taps first so that input index ascends across the whole sum
    const long long nchunks = (m + DIRECT_MC - 1) / DIRECT_MC;
    for (long long cidx = nchunks - 1; cidx >= 0; cidx--) {
        const long long j0 = cidx * DIRECT_MC;
        // taps j0 .. j0+MC-1 (zero beyond m); signal a[k0 - (j0+MC-1) .. k0 + TILE - 1 - j0]
        const long long abase = k0 - (j0 + DIRECT_MC - 1);
        __syncthreads();
        for (int i = t; i < DIRECT_TILE + DIRECT_MC - 1; i += DIRECT_THREADS) {
            const long long ai = abase + i;
            ADSP_SA(i) = (ai >= 0 && ai < n) ? ac[ai] : (T)0;
        }
        if (t < DIRECT_MC) sb[t] = (j0 + t < m) ? bc[j0 + t] : (T)0;
        __syncthreads();
        // thread's outputs k = k0 + t*RO + r ; for tap j = j0 + jj the sample is
        // a[k - j] = sa[(k - j) - abase] = sa[t*RO + r + (MC-1) - jj]
        const int base = t * DIRECT_RO;
        const int lim = (int)((m - j0 < DIRECT_MC) ? (m - j0) : DIRECT_MC);   // taps of this chunk that exist
        // Full chunks run branch free; a partial chunk skips its padded taps (not "multiplies by zero"):
        // the reference never forms those products, so a NaN/Inf sample must only reach the outputs
        // its real taps touch.
        if (lim == DIRECT_MC) direct_chunk<T, FUSED, false>(acc, sa, sb, base, lim);
        else direct_chunk<T, FUSED, true>(acc, sa, sb, base, lim);
    }
    direct_store_tile<T>(acc, sa, out + ch * out_stride, k0, out_len, t);
#undef ADSP_SA
}

// Same kernel for the strategy auto-select case (conv.go:209-211: at most 64 taps, one kernel for every channel):
// the taps travel as a kernel PARAMETER, so each tap is a constant-bank operand of the DFMA instead of a
// shared-memory broadcast load -- one LSU wavefront less per 8 DFMA in a kernel whose LSU and FP64 pipes are
// otherwise co-limited (2 + 1 wavefronts vs 4 FP64 issue cycles per tap).
template <typename T> struct DirectTaps { T v[DIRECT_MC]; };

template <typename T, bool FUSED, bool GUARD>
__global__ void __launch_bounds__(DIRECT_THREADS)
direct_conv_ctaps_kernel(const T *__restrict__ a, long long n, long long a_stride, const __grid_constant__ DirectTaps<T> taps, int m,
                         T *__restrict__ out, long long out_stride, long long tiles_per_ch) {
    __shared__ T sa[(DIRECT_TILE + DIRECT_MC) + ((DIRECT_TILE + DIRECT_MC) >> DIRECT_PAD_SHIFT) + 1];
#define ADSP_SA(i) sa[(i) + ((i) >> DIRECT_PAD_SHIFT)]
    const long long ch = blockIdx.x / tiles_per_ch;
    const long long tile = blockIdx.x - ch * tiles_per_ch;
    const long long k0 = tile * DIRECT_TILE;
    const T *ac = a + ch * a_stride;
    const long long out_len = n + m - 1;
    const int t = threadIdx.x;
    const long long abase = k0 - (DIRECT_MC - 1);
    for (int i = t; i < DIRECT_TILE + DIRECT_MC - 1; i += DIRECT_THREADS) {
        const long long ai = abase + i;
        ADSP_SA(i) = (ai >= 0 && ai < n) ? ac[ai] : (T)0;
    }
    __syncthreads();
    const int base = t * DIRECT_RO;
    T acc[DIRECT_RO], w[DIRECT_RO];
#pragma unroll
    for (int r = 0; r < DIRECT_RO; r++) acc[r] = (T)0;
#pragma unroll
    for (int r = 0; r < DIRECT_RO - 1; r++) w[r + 1] = ADSP_SA(base + r);
#pragma unroll
    for (int jj = DIRECT_MC - 1; jj >= 0; jj--) {            // highest tap first: input index ascends (reference order)
#pragma unroll
        for (int r = 0; r < DIRECT_RO - 1; r++) w[r] = w[r + 1];
        w[DIRECT_RO - 1] = ADSP_SA(base + DIRECT_RO - 1 + (DIRECT_MC - 1) - jj);
        if (GUARD && jj >= m) continue;                        // padded taps are skipped, not multiplied by zero
        const T bj = taps.v[jj];
#pragma unroll
        for (int r = 0; r < DIRECT_RO; r++) {
            if (FUSED) acc[r] = fma(w[r], bj, acc[r]);
            else if (sizeof(T) == 8) acc[r] = __dadd_rn((double)acc[r], __dmul_rn((double)w[r], (double)bj));
            else acc[r] = __fadd_rn((float)acc[r], __fmul_rn((float)w[r], (float)bj));
        }
    }
    direct_store_tile<T>(acc, sa, out + ch * out_stride, k0, out_len, t);
#undef ADSP_SA
}

// DirectCircular (dsp/conv/conv.go:176-189): dst[(i+j)%n] += a[i]*b[j]; per output the terms
// arrive in order of ascending i.
template <typename T>
__global__ void direct_circular_kernel(const T *__restrict__ a, const T *__restrict__ b, T *__restrict__ out, long long n) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    T acc = (T)0;
    for (long long i = 0; i < n; i++) {
        long long j = k - i;
        if (j < 0) j += n;
        acc = fma(a[i], b[j], acc);
    }
    out[k] = acc;
}

// reversed copy: dst[i] = src[m-1-i] (Correlate: correlate.go:22-25), batched.
template <typename T>
__global__ void reverse_kernel(const T *__restrict__ src, long long m, long long src_stride,
                               T *__restrict__ dst, long long dst_stride, long long batch) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * batch) return;
    const long long bch = idx / m, i = idx - bch * m;
    dst[bch * dst_stride + i] = src[bch * src_stride + (m - 1 - i)];
}

// out[i] /= denom[0]-derived scalar handled by caller: scale by a device scalar unless zero.
template <typename T>
__global__ void scale_by_inverse_kernel(T *__restrict__ x, long long len, const T *__restrict__ denom) {
    const T d = *denom;
    if (d == (T)0) return;  // "if zeroLag == 0 return result" correlate.go:72-74, :97-99
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) x[i] = x[i] / d;
}

// sum of squares (l2Norm, correlate.go:189-196), single block per vector; result sqrt'ed by caller kernel
template <typename T>
__global__ void norm_product_kernel(const T *__restrict__ a, long long n, const T *__restrict__ b, long long m,
                                    T *__restrict__ out) {
    __shared__ double red[2][32];
    double sa = 0, sb = 0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) sa += (double)a[i] * (double)a[i];
    for (long long i = threadIdx.x; i < m; i += blockDim.x) sb += (double)b[i] * (double)b[i];
    for (int o = 16; o; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sa; red[1][threadIdx.x >> 5] = sb; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0, tb = 0;
        for (int w = 0; w < (int)(blockDim.x + 31) / 32; w++) { ta += red[0][w]; tb += red[1][w]; }
        *out = (T)(sqrt(ta) * sqrt(tb));
    }
}

// ------------------------------------------------------------------------------------------
// FindPeak (dsp/conv/correlate.go:200-216): signed maximum, first index wins.
template <typename T> struct PeakPair { T v; long long i; };

template <typename T>
__device__ __forceinline__ void peak_combine(T &v, long long &i, T ov, long long oi) {
    if (oi >= 0 && (i < 0 || ov > v || (ov == v && oi < i))) { v = ov; i = oi; }
}

// stage 1: grid = (blocks_per_vec, batch); partial results to part_v/part_i
template <typename T>
__global__ void __launch_bounds__(256)
peak_partial_kernel(const T *__restrict__ x, long long len, long long stride, T *__restrict__ part_v,
                    long long *__restrict__ part_i) {
    const T *xv = x + (long long)blockIdx.y * stride;
    T v = (T)0;
    long long idx = -1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (long long)gridDim.x * blockDim.x) {
        const T xi = xv[i];
        if (xi == xi && (idx < 0 || xi > v)) { v = xi; idx = i; }  // NaN never compares greater
    }
    for (int o = 16; o; o >>= 1) {
        const T ov = __shfl_xor_sync(0xffffffffu, v, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, idx, o);
        peak_combine(v, idx, ov, oi);
    }
    __shared__ T sv[8];
    __shared__ long long si[8];
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = v; si[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) peak_combine(v, idx, sv[w], si[w]);
        part_v[(long long)blockIdx.y * gridDim.x + blockIdx.x] = v;
        part_i[(long long)blockIdx.y * gridDim.x + blockIdx.x] = idx;
    }
}

// stage 2: one warp per vector
template <typename T>
__global__ void peak_final_kernel(const T *__restrict__ x, long long stride, const T *__restrict__ part_v,
                                  const long long *__restrict__ part_i, int nparts, long long batch,
                                  T *__restrict__ out_v, long long *__restrict__ out_i) {
    const long long vec = blockIdx.x;
    if (vec >= batch) return;
    T v = (T)0;
    long long idx = -1;
    for (int p = threadIdx.x; p < nparts; p += 32) peak_combine(v, idx, part_v[vec * nparts + p], part_i[vec * nparts + p]);
    for (int o = 16; o; o >>= 1) {
        const T ov = __shfl_xor_sync(0xffffffffu, v, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, idx, o);
        peak_combine(v, idx, ov, oi);
    }
    if (threadIdx.x == 0) {
        // reference semantics when corr[0] is NaN: nothing compares greater, (0, NaN) is returned
        const T x0 = x[vec * stride];
        if (x0 != x0) { v = x0; idx = 0; }
        out_v[vec] = v;
        out_i[vec] = idx;
    }
}

}  // namespace adsp
