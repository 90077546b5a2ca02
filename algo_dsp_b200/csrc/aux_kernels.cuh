// aux_kernels.cuh -- direct (time-domain) convolution, peak search and small helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace adsp {

// ------------------------------------------------------------------------------------------
// Sliding-window direct convolution.  Replaces the n*m MAC loop of the reference
// (dsp/conv/conv.go:117-154: dst[i+j] += a[i]*b[j]).  For each output k the products are
// accumulated in the reference's order (input index i ascending, i.e. tap index descending).
// CTA = 128 threads, each thread produces RO consecutive outputs from a register window over a signal tile staged in
// shared memory; the outputs leave through the same tile (direct_store_tile).  Three tap sources: kernel parameters for up
// to 64 taps (direct_conv_ctaps_kernel) and up to 1024 taps (direct_conv_ctapsn_kernel) shared by all channels, shared memory
// for per-channel or longer kernels (direct_conv_kernel).  fir_inplace_kernel is the second one walking a row in place.
constexpr int DIRECT_THREADS = 128;
#ifndef ADSP_DIRECT_RO
#define ADSP_DIRECT_RO 8
#endif
constexpr int DIRECT_RO = ADSP_DIRECT_RO;
constexpr int DIRECT_PAD_SHIFT = (DIRECT_RO == 16) ? 4 : 3;   // one pad word per RO samples: per-thread windows hit distinct banks
constexpr int DIRECT_TILE = DIRECT_THREADS * DIRECT_RO;  // outputs per CTA
constexpr int DIRECT_MC = 64;                            // taps per unrolled chunk
constexpr int DIRECT_GC = 4;                             // chunks per staged signal window (direct_conv_kernel)

// taps of one staged chunk, highest first; `sa` is the padded signal tile, `sb` the taps
template <typename T, bool FUSED, bool GUARD>
__device__ __forceinline__ void direct_chunk(T (&acc)[DIRECT_RO], const T *sa, const T *sb, int base, int lim) {
    const T *wp = sa + base + (base >> DIRECT_PAD_SHIFT);   // base is a multiple of the padding period: offsets pad alone
#define ADSP_SA(c) wp[(c) + ((c) >> DIRECT_PAD_SHIFT)]
    T w[DIRECT_RO];
#pragma unroll
    for (int r = 0; r < DIRECT_RO - 1; r++) w[r + 1] = ADSP_SA(r);
#pragma unroll
    for (int jj = DIRECT_MC - 1; jj >= 0; jj--) {
#pragma unroll
        for (int r = 0; r < DIRECT_RO - 1; r++) w[r] = w[r + 1];
        w[DIRECT_RO - 1] = ADSP_SA(DIRECT_RO - 1 + (DIRECT_MC - 1) - jj);
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
#if defined(ADSP_DIRECT_STORE_STRIDED)                  // the round-1 store, kept for the ncu A/B in profiles/: RO outputs per lane, 64 bytes apart
#pragma unroll
    for (int r = 0; r < DIRECT_RO; r++) {
        const long long k = k0 + (long long)t * DIRECT_RO + r;
        if (k < out_len) oc[k] = acc[r];
    }
    return;
#endif
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
    constexpr int SPAN = DIRECT_GC * DIRECT_MC;         // taps per staged signal window
    __shared__ T sa[(DIRECT_TILE + SPAN) + ((DIRECT_TILE + SPAN) >> DIRECT_PAD_SHIFT) + 1];
    __shared__ T sb[SPAN];
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

    // spans of GC chunks of taps, highest taps first so that the input index ascends across the whole sum; one staged
    // window of TILE + SPAN - 1 samples serves all chunks of a span (a window per chunk re-read 17 samples per output and
    // chunk, and took two barriers each)
    const long long nspans = (m + SPAN - 1) / SPAN;
    for (long long sidx = nspans - 1; sidx >= 0; sidx--) {
        const long long J0 = sidx * SPAN;
        // taps J0 .. J0+SPAN-1 (zero beyond m); signal a[k0 - (J0+SPAN-1) .. k0 + TILE - 1 - J0]
        const long long abase = k0 - (J0 + SPAN - 1);
        __syncthreads();
        for (int i = t; i < DIRECT_TILE + SPAN - 1; i += DIRECT_THREADS) {
            const long long ai = abase + i;
            ADSP_SA(i) = (ai >= 0 && ai < n) ? ac[ai] : (T)0;
        }
        for (int i = t; i < SPAN; i += DIRECT_THREADS) sb[i] = (J0 + i < m) ? bc[J0 + i] : (T)0;
        __syncthreads();
        // thread's outputs k = k0 + t*RO + r ; for tap j = J0 + sub*MC + jj the sample is
        // a[k - j] = sa[(k - j) - abase] = sa[t*RO + (GC-1-sub)*MC + r + (MC-1) - jj]
#pragma unroll 1
        for (int sub = DIRECT_GC - 1; sub >= 0; sub--) {
            const long long left = m - (J0 + (long long)sub * DIRECT_MC);   // taps of this chunk that exist
            if (left <= 0) continue;
            const int base = t * DIRECT_RO + (DIRECT_GC - 1 - sub) * DIRECT_MC;
            // Full chunks run branch free; a partial chunk skips its padded taps (not "multiplies by zero"):
            // the reference never forms those products, so a NaN/Inf sample must only reach the outputs
            // its real taps touch.
            if (left >= DIRECT_MC) direct_chunk<T, FUSED, false>(acc, sa, sb + sub * DIRECT_MC, base, DIRECT_MC);
            else direct_chunk<T, FUSED, true>(acc, sa, sb + sub * DIRECT_MC, base, (int)left);
        }
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

// Longer shared kernels (65 .. NCH*64 taps; FIR blocks, explicit Direct calls): the same constant-bank taps, NCH chunks of
// 64 per launch (kernel parameters may be up to 32 KB), one staged window of TILE + NCH*64 - 1 samples for all of them.
// Against the shared-memory-tap kernel above this drops the broadcast tap load, a third of its LSU wavefronts.
template <typename T, int NCH> struct DirectTapsN { T v[NCH * DIRECT_MC]; };

// all taps of one tile from its staged window `sa` (TILE + nch*MC - 1 samples, padded), taps from the kernel parameter
template <typename T, bool FUSED, int NCH>
__device__ __forceinline__ void direct_ctapsn_tile(T (&acc)[DIRECT_RO], const T *sa, const DirectTapsN<T, NCH> &taps, int m, int nch, int t) {
#pragma unroll
    for (int r = 0; r < DIRECT_RO; r++) acc[r] = (T)0;
    // tap j = sub*MC + jj of output k = k0 + t*RO + r reads sa[t*RO + (nch-1-sub)*MC + r + (MC-1) - jj]; highest tap first
#pragma unroll 1
    for (int sub = nch - 1; sub >= 0; sub--) {
        const int base = t * DIRECT_RO + (nch - 1 - sub) * DIRECT_MC;   // a multiple of the padding period: offsets pad alone
        const T *wp = sa + base + (base >> DIRECT_PAD_SHIFT);
#define ADSP_WP(c) wp[(c) + ((c) >> DIRECT_PAD_SHIFT)]
        const int left = m - sub * DIRECT_MC;                    // < MC only in the highest chunk
        const T *tv = taps.v + sub * DIRECT_MC;
        T w[DIRECT_RO];
#pragma unroll
        for (int r = 0; r < DIRECT_RO - 1; r++) w[r + 1] = ADSP_WP(r);
        if (left >= DIRECT_MC) {
#pragma unroll
            for (int jj = DIRECT_MC - 1; jj >= 0; jj--) {
#pragma unroll
                for (int r = 0; r < DIRECT_RO - 1; r++) w[r] = w[r + 1];
                w[DIRECT_RO - 1] = ADSP_WP(DIRECT_RO - 1 + (DIRECT_MC - 1) - jj);
                const T bj = tv[jj];
#pragma unroll
                for (int r = 0; r < DIRECT_RO; r++) {
                    if (FUSED) acc[r] = fma(w[r], bj, acc[r]);
                    else if (sizeof(T) == 8) acc[r] = __dadd_rn((double)acc[r], __dmul_rn((double)w[r], (double)bj));
                    else acc[r] = __fadd_rn((float)acc[r], __fmul_rn((float)w[r], (float)bj));
                }
            }
        } else {
            // partial chunk: its padded taps are skipped, not multiplied by zero (see direct_conv_kernel)
#pragma unroll 4
            for (int jj = DIRECT_MC - 1; jj >= 0; jj--) {
#pragma unroll
                for (int r = 0; r < DIRECT_RO - 1; r++) w[r] = w[r + 1];
                w[DIRECT_RO - 1] = ADSP_WP(DIRECT_RO - 1 + (DIRECT_MC - 1) - jj);
                if (jj >= left) continue;
                const T bj = tv[jj];
#pragma unroll
                for (int r = 0; r < DIRECT_RO; r++) {
                    if (FUSED) acc[r] = fma(w[r], bj, acc[r]);
                    else if (sizeof(T) == 8) acc[r] = __dadd_rn((double)acc[r], __dmul_rn((double)w[r], (double)bj));
                    else acc[r] = __fadd_rn((float)acc[r], __fmul_rn((float)w[r], (float)bj));
                }
            }
        }
    }
#undef ADSP_WP
}

template <typename T, bool FUSED, int NCH>
__global__ void __launch_bounds__(DIRECT_THREADS)
direct_conv_ctapsn_kernel(const T *__restrict__ a, long long n, long long a_stride, const __grid_constant__ DirectTapsN<T, NCH> taps, int m,
                          T *__restrict__ out, long long out_stride, long long tiles_per_ch) {
    constexpr int SPAN = NCH * DIRECT_MC;
    __shared__ T sa[(DIRECT_TILE + SPAN) + ((DIRECT_TILE + SPAN) >> DIRECT_PAD_SHIFT) + 1];
#define ADSP_SA(i) sa[(i) + ((i) >> DIRECT_PAD_SHIFT)]
    const long long ch = blockIdx.x / tiles_per_ch;
    const long long tile = blockIdx.x - ch * tiles_per_ch;
    const long long k0 = tile * DIRECT_TILE;
    const T *ac = a + ch * a_stride;
    const long long out_len = n + m - 1;
    const int t = threadIdx.x;
    const int nch = (m + DIRECT_MC - 1) / DIRECT_MC;             // chunks that hold taps
    const int span = nch * DIRECT_MC;
    const long long abase = k0 - (span - 1);
    for (int i = t; i < DIRECT_TILE + span - 1; i += DIRECT_THREADS) {
        const long long ai = abase + i;
        ADSP_SA(i) = (ai >= 0 && ai < n) ? ac[ai] : (T)0;
    }
    __syncthreads();
    T acc[DIRECT_RO];
    direct_ctapsn_tile<T, FUSED, NCH>(acc, sa, taps, m, nch, t);
    direct_store_tile<T>(acc, sa, out + ch * out_stride, k0, out_len, t);
#undef ADSP_SA
}

// ------------------------------------------------------------------------------------------
// Block FIR IN PLACE (dsp/filter/fir Filter.ProcessBlock, filter.go:61-103: the block is input and output).  y[t] needs
// x[t-H .. t] (H = taps - 1), so a CTA that walks the tiles of its segment from the TOP down and stages a tile's window in
// shared memory before storing that tile's outputs never overwrites a sample it still needs; the only samples another CTA
// could destroy are the H in front of each segment, and those (plus the next call's history) are saved by fir_halo_kernel
// before the main kernel starts.  Traffic: the block is read once and written once -- no work-row copy in, no result copy out.
constexpr int FIR_SEG_TILES = 8;                       // tiles per segment (8192 samples per CTA)

template <typename T>
__global__ void __launch_bounds__(256) fir_halo_kernel(const T *__restrict__ buf, long long stride, long long n, const T *__restrict__ hist_old,
                                                       T *__restrict__ hist_new, T *__restrict__ halo, int H, long long seg_len, int nseg) {
    const int s = blockIdx.x;
    const long long ch = blockIdx.y;
    const T *x = buf + ch * stride;
    const T *ho = hist_old + ch * H;                   // x[-H .. -1]
    T *d = (s < nseg) ? halo + (ch * nseg + s) * H : hist_new + ch * H;
    const long long start = (s < nseg) ? (long long)s * seg_len - H : n - H;   // >= -H (H <= seg_len)
    for (int i = threadIdx.x; i < H; i += blockDim.x) {
        const long long idx = start + i;
        d[i] = idx < 0 ? ho[H + idx] : x[idx];
    }
}

template <typename T, bool FUSED, int NCH>
__global__ void __launch_bounds__(DIRECT_THREADS)
fir_inplace_kernel(T *__restrict__ buf, long long stride, long long n, const T *__restrict__ halo, int H, const __grid_constant__ DirectTapsN<T, NCH> taps, int m,
                   long long seg_len, int nseg) {
    constexpr int SPAN = NCH * DIRECT_MC;
    __shared__ T sa[(DIRECT_TILE + SPAN) + ((DIRECT_TILE + SPAN) >> DIRECT_PAD_SHIFT) + 1];
#define ADSP_SA(i) sa[(i) + ((i) >> DIRECT_PAD_SHIFT)]
    const long long ch = blockIdx.x / nseg;
    const int s = (int)(blockIdx.x - ch * nseg);
    T *x = buf + ch * stride;
    const T *hl = halo + (ch * nseg + s) * H;          // x[seg_lo - H .. seg_lo - 1]
    const long long seg_lo = (long long)s * seg_len;
    const long long seg_hi = (n < seg_lo + seg_len) ? n : seg_lo + seg_len;
    const int t = threadIdx.x;
    const int nch = (m + DIRECT_MC - 1) / DIRECT_MC;
    const int span = nch * DIRECT_MC;
    for (long long k0 = seg_lo + ((seg_hi - seg_lo - 1) / DIRECT_TILE) * DIRECT_TILE; k0 >= seg_lo; k0 -= DIRECT_TILE) {
        const long long abase = k0 - (span - 1);
        __syncthreads();                               // the previous tile's outputs have left the buffer
        for (int i = t; i < DIRECT_TILE + span - 1; i += DIRECT_THREADS) {
            const long long ai = abase + i;
            T v = (T)0;                                // slots no stored output reads: beyond the segment, below its halo
            if (ai >= seg_lo) { if (ai < seg_hi) v = x[ai]; }
            else if (ai >= seg_lo - H) v = hl[ai - (seg_lo - H)];
            ADSP_SA(i) = v;
        }
        __syncthreads();
        T acc[DIRECT_RO];
        direct_ctapsn_tile<T, FUSED, NCH>(acc, sa, taps, m, nch, t);
        direct_store_tile<T>(acc, sa, x, k0, seg_hi, t);
    }
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
