// fdl.cu -- multi-channel streaming partitioned convolution on a device-resident FREQUENCY-DOMAIN DELAY LINE.
//
// Replaces the streaming hot loop of the reference's PartitionedConvolutionT.ProcessBlock and its stages
// (dsp/conv/partitioned.go:135-190 `process`, :348-396 `ProcessBlock`) and the wet/dry mix of
// ConvolutionReverb.ProcessInPlace (dsp/effects/reverb/convolution.go:60-83) for many channels per launch.
// Contract kept: every output sample equals the full linear convolution delayed by Latency() = 2^minBlockOrder
// samples (zero before), for arbitrary block lengths per call.  The partition layout is internal and free
// (StageCount/StageInfo keep reporting the reference's layout, api.cu partition_layout):
//
//   stage s: partition size B_s = L*2^s (L = latency), one partition per stage, IR offset off_s = L*(2^s - 1);
//   the last stage keeps B_max (<= 2048 and <= 2^maxBlockOrder) and holds all remaining partitions.
//   B_s <= off_s + L + 1 is what lets a stage wait for whole blocks and still meet the latency.
//
// Where the reference runs one inverse FFT per IR partition and accumulates in the time domain
// (partitioned.go:165-186), each stage here keeps the spectra of its last `count` input blocks (the delay
// line) and evaluates  Y = sum_p X[m-p] * H_p  in the frequency domain: one forward and ONE inverse transform
// per block and stage, the multiply-accumulate fused with the inverse transform in one kernel.
// Two channels share a complex transform (re = channel 2c, im = channel 2c+1; valid because the IR is real).
// All firings of a stage inside one call are independent once their forward spectra exist, so a call costs
// two launches per stage whatever its length.
#include <algorithm>
#include <vector>

#include "engine.cuh"

namespace adsp {

namespace {

constexpr int FDL_MAX_B = 2048;          // largest partition: 4096-point transforms fit one CTA (fft_core.cuh)
constexpr long long FDL_CHUNK = 16384;   // a call is processed in chunks of at most this many samples
constexpr int FDL_SPLIT_MIN_COUNT = 8;   // stages with at least this many partitions use the bin-tiled MAC kernel

template <int L2> struct FdlShape {
    static constexpr int TPF = L2 / 16;                                   // threads per transform
    static constexpr int THREADS = (TPF < 128) ? 128 : TPF;               // 128-thread CTAs, 256 for 4096 points
    static constexpr int ROWS = THREADS / TPF;                            // transforms per CTA
    static constexpr int MIN_CTAS = (THREADS == 128) ? 4 : 2;
};

struct AttrFlags {
    bool done[64] = {};
    bool need(int dev) { if (dev < 0 || dev >= 64) return true; if (done[dev]) return false; done[dev] = true; return true; }
};

struct StageGeom {
    int B;            // partition size
    int count;        // partitions (delay-line taps)
    int ring;         // delay-line slots per channel pair
    long long off;    // IR offset of the stage = output delay of its blocks
};

// Window of firing m of a stage with partition size B: x[(m-1)B, (m+1)B), zero before the stream start.  The input rows are
// RINGS: the sample of stream time t sits at index t & xmask (xmask = -1: a plain row that starts at time 0), so a call
// only appends its samples -- no history slide.
template <typename T, int L2>
__device__ __forceinline__ void fdl_load_window(cpx<T> (&e)[16], const T *__restrict__ xbuf, long long xstride, long long xmask, int channels, int pair,
                                                long long m, int j, bool active) {
    constexpr int TPF = FftShape<L2>::TPF, B = L2 / 2;
    const long long w0 = (m - 1) * B;
    const int ca = 2 * pair, cb = 2 * pair + 1;
    const T *xa = xbuf + (long long)ca * xstride;
    const T *xb = xbuf + (long long)cb * xstride;
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const long long t = w0 + j + q * TPF;
        e[q].x = (active && t >= 0) ? xa[t & xmask] : (T)0;
        e[q].y = (active && t >= 0 && cb < channels) ? xb[t & xmask] : (T)0;
    }
}

// overlap-save extraction of an inverse transform: position i >= B of the circular result is stream time (m-1)B + off + i
template <typename T, int L2>
__device__ __forceinline__ void fdl_accumulate(const cpx<T> (&e)[16], T *__restrict__ acc, long long acc_stride, long long acc_mask, int channels, int pair,
                                               long long m, long long off, int j) {
    constexpr int TPF = FftShape<L2>::TPF, B = L2 / 2;
    const int ca = 2 * pair, cb = 2 * pair + 1;
    T *aa = acc + (long long)ca * acc_stride;
    T *ab = acc + (long long)cb * acc_stride;
    const long long tbase = m * B + off - B;
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const int i = j + q * TPF;
        if (i >= B) {
            const long long t = (tbase + i) & acc_mask;
            aa[t] += e[q].x;
            if (cb < channels) ab[t] += e[q].y;
        }
    }
}

// Forward transforms of the blocks of one stage that complete inside this chunk.
template <typename T, int L2>
__global__ void __launch_bounds__(FdlShape<L2>::THREADS, FdlShape<L2>::MIN_CTAS)
fdl_forward(const T *__restrict__ xbuf, long long xstride, long long xmask, int channels, long long m0, int nfire, int ring,
            cpx<T> *__restrict__ fdl, const cpx<T> *__restrict__ tw) {
    using C = cpx<T>;
    using Sh = FftShape<L2>;
    using FS = FdlShape<L2>;
    constexpr int TPF = Sh::TPF, B = L2 / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *buf = reinterpret_cast<C *>(smem_raw);
    C *stw = buf + FS::ROWS * L2;
    load_tw_smem<T, L2>(stw, tw, threadIdx.x, FS::THREADS);
    const int row = threadIdx.x / TPF, j = threadIdx.x % TPF;
    const int pair = blockIdx.y;
    const int f = blockIdx.x * FS::ROWS + row;          // firing inside this launch
    const bool active = f < nfire;
    const long long m = m0 + (active ? f : 0);
    C e[16];
    fdl_load_window<T, L2>(e, xbuf, xstride, xmask, channels, pair, m, j, active);
    RowAddr<T, Sh::R0> addr{row * L2};
    CtaGate gate;
    cta_fft<T, L2, false>(e, buf, addr, stw, j, gate);
    if (active) {
        C *dst = fdl + ((size_t)pair * ring + (size_t)(m % ring)) * L2 + j;
#pragma unroll
        for (int q = 0; q < 16; q++) dst[q * TPF] = e[q];
    }
}

// A stage with ONE partition needs no delay line: window -> forward transform -> x H -> inverse transform -> accumulate, in
// one launch (the short stages of a low-latency layout: a 128-sample real-time block fires stage 0 on every call).
// The forward transform leaves every thread with the 16 bins the inverse starts from, so the product stays in registers.
template <typename T, int L2>
__global__ void __launch_bounds__(FdlShape<L2>::THREADS, FdlShape<L2>::MIN_CTAS)
fdl_fused_single(const T *__restrict__ xbuf, long long xstride, long long xmask, int channels, long long m0, int nfire, const cpx<T> *__restrict__ H,
                 long long off, T *__restrict__ acc, long long acc_stride, long long acc_mask, const cpx<T> *__restrict__ tw) {
    using C = cpx<T>;
    using Sh = FftShape<L2>;
    using FS = FdlShape<L2>;
    constexpr int TPF = Sh::TPF;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *buf = reinterpret_cast<C *>(smem_raw);
    C *stw = buf + FS::ROWS * L2;
    load_tw_smem<T, L2>(stw, tw, threadIdx.x, FS::THREADS);
    const int row = threadIdx.x / TPF, j = threadIdx.x % TPF;
    const int pair = blockIdx.y;
    const int f = blockIdx.x * FS::ROWS + row;
    const bool active = f < nfire;
    const long long m = m0 + (active ? f : 0);
    C e[16];
    fdl_load_window<T, L2>(e, xbuf, xstride, xmask, channels, pair, m, j, active);
    RowAddr<T, Sh::R0> addr{row * L2};
    CtaGate gate;
    cta_fft<T, L2, false>(e, buf, addr, stw, j, gate);
#pragma unroll
    for (int q0 = 0; q0 < 16; q0 += 4) {              // four spectrum values in flight: the kernel stays inside 128 registers
        C h[4];
#pragma unroll
        for (int q = 0; q < 4; q++) h[q] = __ldg(&H[j + (q0 + q) * TPF]);
#pragma unroll
        for (int q = 0; q < 4; q++) e[q0 + q] = cmul(e[q0 + q], h[q]);
    }
    cta_fft<T, L2, true>(e, buf, addr, stw, j, gate);
    if (active) fdl_accumulate<T, L2>(e, acc, acc_stride, acc_mask, channels, pair, m, off, j);
}

// Fused spectral multiply-accumulate over the stage's partitions + inverse transform + overlap-save extraction:
//   Y[k] = sum_{p < count} FDL[(m-p) mod ring][k] * H_p[k];  y = IFFT(Y)[B .. 2B)  ->  acc[t = m*B + off + r] += y[r]
// A thread owns the 16 bins the inverse transform wants it to own, so the sum lands in the registers the
// butterflies start from.  Blocks before the stream start (m - p < 0) contribute nothing.
template <typename T, int L2>
__global__ void __launch_bounds__(FdlShape<L2>::THREADS, FdlShape<L2>::MIN_CTAS)
fdl_mac_inverse(const cpx<T> *__restrict__ fdl, const cpx<T> *__restrict__ H, int count, int ring, long long m0, int nfire,
                long long off, T *__restrict__ acc, long long acc_stride, long long acc_mask, int channels,
                const cpx<T> *__restrict__ tw) {
    using C = cpx<T>;
    using Sh = FftShape<L2>;
    using FS = FdlShape<L2>;
    constexpr int TPF = Sh::TPF, B = L2 / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *buf = reinterpret_cast<C *>(smem_raw);
    C *stw = buf + FS::ROWS * L2;
    load_tw_smem<T, L2>(stw, tw, threadIdx.x, FS::THREADS);
    const int row = threadIdx.x / TPF, j = threadIdx.x % TPF;
    const int pair = blockIdx.y;
    const int f = blockIdx.x * FS::ROWS + row;
    const bool active = f < nfire;
    const long long m = m0 + (active ? f : 0);
    const C *ring_base = fdl + (size_t)pair * ring * L2 + j;
    C e[16];
#pragma unroll
    for (int q = 0; q < 16; q++) { e[q].x = (T)0; e[q].y = (T)0; }
    const int np = active ? (int)((m + 1 < count) ? m + 1 : count) : 0;     // taps with m - p >= 0
    int slot = (int)(m % ring);
    for (int p = 0; p < np; p++) {
        const C *xr = ring_base + (size_t)slot * L2;
        const C *hr = H + (size_t)p * L2 + j;
#pragma unroll
        for (int q0 = 0; q0 < 16; q0 += 4) {          // 4 + 4 loads in flight per step keeps the kernel inside 128 registers
            C xv[4], hv[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                xv[q] = __ldcg(&xr[(q0 + q) * TPF]);
                if (H) hv[q] = __ldg(&hr[(q0 + q) * TPF]); else { hv[q].x = (T)1; hv[q].y = (T)0; }   // H == null: Y is already summed
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                e[q0 + q].x += xv[q].x * hv[q].x - xv[q].y * hv[q].y;
                e[q0 + q].y += xv[q].x * hv[q].y + xv[q].y * hv[q].x;
            }
        }
        slot = (slot == 0) ? ring - 1 : slot - 1;
    }
    RowAddr<T, Sh::R0> addr{row * L2};
    CtaGate gate;
    cta_fft<T, L2, true>(e, buf, addr, stw, j, gate);
    if (active) fdl_accumulate<T, L2>(e, acc, acc_stride, acc_mask, channels, pair, m, off, j);
}

// Spectral multiply-accumulate of a long stage, tiled over BINS instead of firings: one thread owns one bin of FT
// consecutive firings of a channel pair,
//     Y_m[k] = sum_{p < count} X_{m-p}[k] * H_p[k],   m = mA .. mA + FT - 1,
// walks the delay line once (count + FT - 1 slots, newest first) and keeps the FT spectrum values each slot meets
// in a shift register -- every delay-line value is loaded once per FT firings instead of once per firing, and a
// single firing still spreads over pairs * L2/128 CTAs (a real-time block is one firing: 32 pairs alone would
// leave most SMs idle).  Y goes to a small scratch; fdl_mac_inverse (H == null) then inverts and accumulates it.
template <typename T, int FT>
__global__ void __launch_bounds__(128)
fdl_mac_bins(const cpx<T> *__restrict__ fdl, const cpx<T> *__restrict__ H, int count, int ring, long long m0, int nfire, int L2,
             cpx<T> *__restrict__ Y, int yring) {
    using C = cpx<T>;
    const int bin = blockIdx.x * 128 + threadIdx.x;
    const int pair = blockIdx.y;
    const long long mA = m0 + (long long)blockIdx.z * FT;
    const C *ring_base = fdl + (size_t)pair * ring * L2 + bin;
    const C *hb = H + bin;
    C acc[FT], hw[FT];
#pragma unroll
    for (int f = 0; f < FT; f++) { acc[f].x = acc[f].y = (T)0; hw[f].x = hw[f].y = (T)0; }
    const long long s_hi = mA + FT - 1;
    long long s_lo = mA - count + 1;
    if (s_lo < 0) s_lo = 0;                                  // blocks before the stream start are zero
    int slot = (int)(s_hi % ring);
    int ptop = 0;                                            // spectrum index met by the newest firing at this slot
#pragma unroll 4
    for (long long sidx = s_hi; sidx >= s_lo; sidx--) {
#pragma unroll
        for (int f = 0; f < FT - 1; f++) hw[f] = hw[f + 1];
        if (ptop < count) hw[FT - 1] = __ldg(&hb[(size_t)ptop * L2]); else { hw[FT - 1].x = (T)0; hw[FT - 1].y = (T)0; }
        const C x = __ldcg(&ring_base[(size_t)slot * L2]);
        // firing f meets partition ptop - (FT-1-f) at this slot; outside [0, count) the slot does not belong to that
        // firing (newer than it, or older than the IR) and is skipped rather than multiplied by a zero tap, so a
        // non-finite or stale delay-line value cannot leak into outputs the reference leaves untouched (0 * NaN)
#pragma unroll
        for (int f = 0; f < FT; f++) {
            if ((unsigned)(ptop - (FT - 1 - f)) < (unsigned)count) {
                acc[f].x += x.x * hw[f].x - x.y * hw[f].y;
                acc[f].y += x.x * hw[f].y + x.y * hw[f].x;
            }
        }
        ptop++;
        slot = (slot == 0) ? ring - 1 : slot - 1;
    }
#pragma unroll
    for (int f = 0; f < FT; f++) {
        const long long m = mA + f;
        if (m < m0 + nfire) Y[((size_t)pair * yring + (size_t)(m % yring)) * L2 + bin] = acc[f];
    }
}

// Emit the call's outputs (stream times [tout, tout + n)), clear them in the accumulator, optionally mix with
// the dry input:  out = dry * x + wet * y   (ConvolutionReverb.ProcessInPlace, convolution.go:76-80)
template <typename T>
__global__ void fdl_emit(T *__restrict__ acc, long long acc_stride, long long acc_mask, long long tout, long long n,
                         const T *__restrict__ xring, long long xstride, long long xmask, long long tin, T *__restrict__ out, long long out_stride, int mix,
                         T wet, T dry) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (i >= n) return;
    const long long t = tout + i;
    T y = (T)0;
    if (t >= 0) {
        T *a = acc + (long long)c * acc_stride + (t & acc_mask);
        y = *a;
        *a = (T)0;
    }
    if (mix) y = dry * xring[(long long)c * xstride + ((tin + i) & xmask)] + wet * y;   // the call's own input, from the ring
    out[(long long)c * out_stride + i] = y;
}

template <typename T> __global__ void fdl_scale(cpx<T> *p, long long n, T s) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { p[i].x *= s; p[i].y *= s; }
}

}  // namespace

struct FdlEngine {
    adsp_ctx *ctx = nullptr;
    adsp_precision prec = ADSP_F64;
    long long K = 0;
    int latency = 0, channels = 0, pairs = 0;
    std::vector<StageGeom> stages;
    std::vector<void *> d_H, d_fdl;           // per stage: spectra [count][2B], delay line [pairs][ring][2B]
    std::vector<const void *> d_tw;           // per stage: twiddle table of the 2B-point transform
    long long HX = 0;                         // input history a firing may reach back to (2*B_max)
    long long XL = 0;                         // ring length of an input row: power of two >= HX + FDL_CHUNK
    DevBuf xbuf, acc, d_io_out, yscratch;     // xbuf: input rows as rings, time t at index t & (XL - 1)
    long long acc_len = 0;
    long long pos = 0;                        // samples consumed so far
    double wet = 1.0, dry = 1.0;
};

namespace {

struct FwdArgs {
    const void *xbuf; long long xstride, xmask; int channels, pairs; long long m0; int nfire, ring; void *fdl; const void *tw;
};
struct FusedArgs {
    const void *xbuf; long long xstride, xmask; int channels, pairs; long long m0; int nfire; const void *H; long long off; void *acc; long long acc_len;
    const void *tw;
};
struct MacArgs {
    const void *fdl, *H; int count, ring; long long m0; int nfire; long long off; void *acc; long long acc_len; int channels, pairs;
    const void *tw;
};

template <typename T, int L2> adsp_status fdl_forward_launch(adsp_ctx *ctx, const FwdArgs &a) {
    using FS = FdlShape<L2>;
    const size_t smem = ((size_t)FS::ROWS * L2 + FftShape<L2>::TW_ENTRIES) * sizeof(cpx<T>);
    static AttrFlags once;
    if (once.need(ctx->device) && smem > 48 * 1024)
        ADSP_CUDA(cudaFuncSetAttribute(fdl_forward<T, L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((a.nfire + FS::ROWS - 1) / FS::ROWS), (unsigned)a.pairs);
    LaunchTimer lt(ctx, ctx->main, KK_OTHER);
    fdl_forward<T, L2><<<grid, FS::THREADS, smem, ctx->main>>>((const T *)a.xbuf, a.xstride, a.xmask, a.channels, a.m0, a.nfire, a.ring,
                                                             (cpx<T> *)a.fdl, (const cpx<T> *)a.tw);
    count_launch(ctx);
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}

template <typename T, int L2> adsp_status fdl_mac_launch(adsp_ctx *ctx, const MacArgs &a) {
    using FS = FdlShape<L2>;
    const size_t smem = ((size_t)FS::ROWS * L2 + FftShape<L2>::TW_ENTRIES) * sizeof(cpx<T>);
    static AttrFlags once;
    if (once.need(ctx->device) && smem > 48 * 1024)
        ADSP_CUDA(cudaFuncSetAttribute(fdl_mac_inverse<T, L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((a.nfire + FS::ROWS - 1) / FS::ROWS), (unsigned)a.pairs);
    LaunchTimer lt(ctx, ctx->main, KK_OTHER);
    fdl_mac_inverse<T, L2><<<grid, FS::THREADS, smem, ctx->main>>>((const cpx<T> *)a.fdl, (const cpx<T> *)a.H, a.count, a.ring, a.m0, a.nfire,
                                                                 a.off, (T *)a.acc, a.acc_len, a.acc_len - 1, a.channels,
                                                                 (const cpx<T> *)a.tw);
    count_launch(ctx);
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}

template <typename T, int L2> adsp_status fdl_fused_launch(adsp_ctx *ctx, const FusedArgs &a) {
    using FS = FdlShape<L2>;
    const size_t smem = ((size_t)FS::ROWS * L2 + FftShape<L2>::TW_ENTRIES) * sizeof(cpx<T>);
    static AttrFlags once;
    if (once.need(ctx->device) && smem > 48 * 1024)
        ADSP_CUDA(cudaFuncSetAttribute(fdl_fused_single<T, L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((a.nfire + FS::ROWS - 1) / FS::ROWS), (unsigned)a.pairs);
    LaunchTimer lt(ctx, ctx->main, KK_OTHER);
    fdl_fused_single<T, L2><<<grid, FS::THREADS, smem, ctx->main>>>((const T *)a.xbuf, a.xstride, a.xmask, a.channels, a.m0, a.nfire, (const cpx<T> *)a.H, a.off,
                                                                  (T *)a.acc, a.acc_len, a.acc_len - 1, (const cpx<T> *)a.tw);
    count_launch(ctx);
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}

#define ADSP_FDL_DISPATCH(FN, L2v, ...)                                                      \
    switch (L2v) {                                                                           \
    case 16:   return FN<T, 16>(__VA_ARGS__);                                                \
    case 32:   return FN<T, 32>(__VA_ARGS__);                                                \
    case 64:   return FN<T, 64>(__VA_ARGS__);                                                \
    case 128:  return FN<T, 128>(__VA_ARGS__);                                               \
    case 256:  return FN<T, 256>(__VA_ARGS__);                                               \
    case 512:  return FN<T, 512>(__VA_ARGS__);                                               \
    case 1024: return FN<T, 1024>(__VA_ARGS__);                                              \
    case 2048: return FN<T, 2048>(__VA_ARGS__);                                              \
    case 4096: return FN<T, 4096>(__VA_ARGS__);                                              \
    default: set_error("fdl: unsupported partition size"); return ADSP_ERR_INVALID_ARG;      \
    }

template <typename T> adsp_status fdl_forward_any(adsp_ctx *ctx, int L2, const FwdArgs &a) { ADSP_FDL_DISPATCH(fdl_forward_launch, L2, ctx, a) }
template <typename T> adsp_status fdl_mac_any(adsp_ctx *ctx, int L2, const MacArgs &a) { ADSP_FDL_DISPATCH(fdl_mac_launch, L2, ctx, a) }
template <typename T> adsp_status fdl_fused_any(adsp_ctx *ctx, int L2, const FusedArgs &a) { ADSP_FDL_DISPATCH(fdl_fused_launch, L2, ctx, a) }

template <typename T> adsp_status fdl_build(FdlEngine *e, const T *d_kernel) {
    adsp_ctx *ctx = e->ctx;
    const size_t ns = e->stages.size();
    e->d_H.assign(ns, nullptr); e->d_fdl.assign(ns, nullptr); e->d_tw.assign(ns, nullptr);
    ADSP_TRY(e->xbuf.reserve((size_t)e->XL * (size_t)(2 * e->pairs) * sizeof(T)));
    ADSP_TRY(e->acc.reserve((size_t)e->acc_len * (size_t)(2 * e->pairs) * sizeof(T)));
    for (size_t s = 0; s < ns; s++) {
        const StageGeom &g = e->stages[s];
        const int L2 = 2 * g.B;
        const cpx<T> *tw = nullptr;
        ADSP_TRY(get_tw_table<T>(ctx, L2, &tw));
        e->d_tw[s] = tw;
        ADSP_CUDA(cudaMalloc(&e->d_H[s], (size_t)g.count * L2 * sizeof(cpx<T>)));
        if (g.count > 1) ADSP_CUDA(cudaMalloc(&e->d_fdl[s], (size_t)e->pairs * g.ring * L2 * sizeof(cpx<T>)));   // one partition: no delay line
        // IR spectra through the same forward kernel: a staging row holds partition p at times [2pB, 2pB + B) and
        // zeros at [2pB + B, 2pB + 2B), so the window of "firing" m = 2p + 1 is exactly [h_p | 0] (partition in the
        // first half, as overlap-save with outputs taken from [B, 2B) needs).
        const long long row_len = 2LL * g.count * g.B + 2 * g.B;
        const int ring_tmp = 2 * g.count + 2;
        struct Scoped { DevBuf b; ~Scoped() { b.release(); } } row_s, spec_s;   // freed on every exit path
        DevBuf &row = row_s.b, &spec = spec_s.b;
        ADSP_TRY(row.reserve((size_t)row_len * sizeof(T)));
        ADSP_TRY(spec.reserve((size_t)ring_tmp * L2 * sizeof(cpx<T>)));
        ADSP_CUDA(cudaMemsetAsync(row.p, 0, (size_t)row_len * sizeof(T), ctx->main));
        for (int p = 0; p < g.count; p++) {
            const long long k0 = g.off + (long long)p * g.B;
            const long long len = std::min<long long>(g.B, e->K - k0);
            if (len > 0)
                ADSP_CUDA(cudaMemcpyAsync((T *)row.p + 2LL * p * g.B, d_kernel + k0, (size_t)len * sizeof(T), cudaMemcpyDeviceToDevice, ctx->main));
        }
        FwdArgs fa{row.p, row_len, -1LL, 1, 1, 1, 2 * g.count - 1, ring_tmp, spec.p, tw};    // xmask -1: a plain row from time 0
        adsp_status st = fdl_forward_any<T>(ctx, L2, fa);
        for (int p = 0; p < g.count && st == ADSP_OK; p++)
            if (cudaMemcpyAsync((cpx<T> *)e->d_H[s] + (size_t)p * L2, (cpx<T> *)spec.p + (size_t)(2 * p + 1) * L2, (size_t)L2 * sizeof(cpx<T>),
                                cudaMemcpyDeviceToDevice, ctx->main) != cudaSuccess) st = ADSP_ERR_CUDA;
        const long long tot = (long long)g.count * L2;
        if (st == ADSP_OK) fdl_scale<T><<<(unsigned)((tot + 255) / 256), 256, 0, ctx->main>>>((cpx<T> *)e->d_H[s], tot, (T)(1.0L / (long double)L2));
        cudaStreamSynchronize(ctx->main);
        ADSP_TRY(st);
    }
    return ADSP_OK;
}

template <typename T> adsp_status fdl_reset_t(FdlEngine *e) {
    adsp_ctx *ctx = e->ctx;
    for (size_t s = 0; s < e->stages.size(); s++)
        if (e->d_fdl[s]) ADSP_CUDA(cudaMemsetAsync(e->d_fdl[s], 0, (size_t)e->pairs * e->stages[s].ring * 2 * e->stages[s].B * sizeof(cpx<T>), ctx->main));
    ADSP_CUDA(cudaMemsetAsync(e->xbuf.p, 0, e->xbuf.cap, ctx->main));
    ADSP_CUDA(cudaMemsetAsync(e->acc.p, 0, e->acc.cap, ctx->main));
    e->pos = 0;
    ADSP_CUDA(cudaStreamSynchronize(ctx->main));
    return ADSP_OK;
}

// one chunk (n <= FDL_CHUNK) whose input already sits in the rings at times [pos, pos + n)
template <typename T>
adsp_status fdl_chunk(FdlEngine *e, long long n, T *d_out, long long out_stride, bool mix) {
    adsp_ctx *ctx = e->ctx;
    const long long xstride = e->XL, xmask = e->XL - 1;
    for (size_t s = 0; s < e->stages.size(); s++) {
        const StageGeom &g = e->stages[s];
        const long long m_first = e->pos / g.B;                  // blocks with (m+1)*B in (pos, pos+n]
        const long long m_last = (e->pos + n) / g.B - 1;
        const long long nf = m_last - m_first + 1;
        if (nf <= 0) continue;
        if (g.count == 1) {
            FusedArgs ua{e->xbuf.p, xstride, xmask, e->channels, e->pairs, m_first, (int)nf, e->d_H[s], g.off, e->acc.p, e->acc_len, e->d_tw[s]};
            ADSP_TRY(fdl_fused_any<T>(ctx, 2 * g.B, ua));
            continue;
        }
        FwdArgs fa{e->xbuf.p, xstride, xmask, e->channels, e->pairs, m_first, (int)nf, g.ring, e->d_fdl[s], e->d_tw[s]};
        ADSP_TRY(fdl_forward_any<T>(ctx, 2 * g.B, fa));
        if (g.count >= FDL_SPLIT_MIN_COUNT && 2 * g.B >= 128) {
            // long stage: bin-tiled MAC into a scratch spectrum, then the inverse kernel on Y (H == null, one "tap")
            const int L2 = 2 * g.B, yring = (int)nf;
            ADSP_TRY(e->yscratch.reserve((size_t)e->pairs * (size_t)yring * L2 * sizeof(cpx<T>)));
            const int FT = nf >= 4 ? 4 : (nf >= 2 ? 2 : 1);
            dim3 grid((unsigned)(L2 / 128), (unsigned)e->pairs, (unsigned)((nf + FT - 1) / FT));
            {
                LaunchTimer lt(ctx, ctx->main, KK_OTHER);
                if (FT == 4) fdl_mac_bins<T, 4><<<grid, 128, 0, ctx->main>>>((const cpx<T> *)e->d_fdl[s], (const cpx<T> *)e->d_H[s], g.count, g.ring, m_first, (int)nf, L2, (cpx<T> *)e->yscratch.p, yring);
                else if (FT == 2) fdl_mac_bins<T, 2><<<grid, 128, 0, ctx->main>>>((const cpx<T> *)e->d_fdl[s], (const cpx<T> *)e->d_H[s], g.count, g.ring, m_first, (int)nf, L2, (cpx<T> *)e->yscratch.p, yring);
                else fdl_mac_bins<T, 1><<<grid, 128, 0, ctx->main>>>((const cpx<T> *)e->d_fdl[s], (const cpx<T> *)e->d_H[s], g.count, g.ring, m_first, (int)nf, L2, (cpx<T> *)e->yscratch.p, yring);
                count_launch(ctx);
            }
            MacArgs ma{e->yscratch.p, nullptr, 1, yring, m_first, (int)nf, g.off, e->acc.p, e->acc_len, e->channels, e->pairs, e->d_tw[s]};
            ADSP_TRY(fdl_mac_any<T>(ctx, L2, ma));
        } else {
            MacArgs ma{e->d_fdl[s], e->d_H[s], g.count, g.ring, m_first, (int)nf, g.off, e->acc.p, e->acc_len, e->channels, e->pairs, e->d_tw[s]};
            ADSP_TRY(fdl_mac_any<T>(ctx, 2 * g.B, ma));
        }
    }
    for (int c0 = 0; c0 < e->channels; c0 += 65535) {   // grid.y is limited to 65535 rows per launch
        const int nc = std::min(65535, e->channels - c0);
        dim3 grid((unsigned)((n + 255) / 256), (unsigned)nc);
        fdl_emit<T><<<grid, 256, 0, ctx->main>>>((T *)e->acc.p + (long long)c0 * e->acc_len, e->acc_len, e->acc_len - 1, e->pos - e->latency, n,
                                                 (const T *)e->xbuf.p + (long long)c0 * xstride, xstride, xmask, e->pos, d_out + (long long)c0 * out_stride,
                                                 out_stride, mix ? 1 : 0, (T)e->wet, (T)e->dry);
        count_launch(ctx);
    }
    e->pos += n;
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}

template <typename T>
adsp_status fdl_process_t(FdlEngine *e, const T *in, long long n, long long in_stride, T *out, long long out_stride, bool host, bool mix) {
    adsp_ctx *ctx = e->ctx;
    const long long xstride = e->XL;
    for (long long o = 0; o < n; o += FDL_CHUNK) {
        T *xb = (T *)e->xbuf.p;
        const long long len = std::min(FDL_CHUNK, n - o);
        const cudaMemcpyKind kin = host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
        // append the chunk to the rings at time pos (two pieces when it crosses the end of the ring)
        const long long at = e->pos & (e->XL - 1), first = std::min(len, e->XL - at);
        ADSP_CUDA(cudaMemcpy2DAsync(xb + at, (size_t)xstride * sizeof(T), in + o, (size_t)in_stride * sizeof(T), (size_t)first * sizeof(T),
                                    (size_t)e->channels, kin, ctx->main));
        if (first < len)
            ADSP_CUDA(cudaMemcpy2DAsync(xb, (size_t)xstride * sizeof(T), in + o + first, (size_t)in_stride * sizeof(T), (size_t)(len - first) * sizeof(T),
                                        (size_t)e->channels, kin, ctx->main));
        if (host) {
            ADSP_TRY(e->d_io_out.reserve((size_t)FDL_CHUNK * (size_t)e->channels * sizeof(T)));
            T *dout = (T *)e->d_io_out.p;
            ADSP_TRY(fdl_chunk<T>(e, len, dout, FDL_CHUNK, mix));
            ADSP_CUDA(cudaMemcpy2DAsync(out + o, (size_t)out_stride * sizeof(T), dout, (size_t)FDL_CHUNK * sizeof(T), (size_t)len * sizeof(T),
                                        (size_t)e->channels, cudaMemcpyDeviceToHost, ctx->main));
        } else {
            ADSP_TRY(fdl_chunk<T>(e, len, out + o, out_stride, mix));
        }
    }
    if (host) ADSP_CUDA(cudaStreamSynchronize(ctx->main));
    return ADSP_OK;
}

}  // namespace

bool fdl_supported(int min_order) { return min_order >= 3 && min_order <= 24; }

// Internal stage layout (pure host arithmetic): sizes L*2^s with one partition each, capped at min(2048, 2^maxOrder);
// the last stage holds all remaining partitions.  Invariant: part_size <= ir_offset + L + 1 for every stage.
std::vector<FdlStage> fdl_layout(long long K, int min_order, int max_order) {
    std::vector<FdlStage> out;
    const long long L = 1LL << min_order;
    long long bmax = std::min<long long>(FDL_MAX_B, 1LL << std::min(max_order, 30));
    if (bmax < 8) bmax = 8;
    long long off = 0, B = std::min(L, bmax);
    while (off < K) {
        const bool last = (B >= bmax) || (off + B >= K);
        const int count = last ? (int)((K - off + B - 1) / B) : 1;
        out.push_back({(int)B, count, off});
        off += (long long)count * B;
        if (last) break;
        B *= 2;
    }
    return out;
}

adsp_status fdl_create(adsp_ctx *ctx, const void *d_kernel, long long K, int min_order, int max_order, int channels,
                       adsp_precision prec, FdlEngine **out) {
    *out = nullptr;
    if (!fdl_supported(min_order) || channels <= 0) return ADSP_ERR_INVALID_ARG;
    FdlEngine *e = new FdlEngine();
    e->ctx = ctx; e->prec = prec; e->K = K; e->latency = 1 << min_order; e->channels = channels; e->pairs = (channels + 1) / 2;
    for (const FdlStage &st : fdl_layout(K, min_order, max_order)) {
        StageGeom g{};
        g.B = st.part_size; g.count = st.count; g.off = st.ir_offset;
        g.ring = g.count + (int)(FDL_CHUNK / g.B) + 2;
        e->stages.push_back(g);
    }
    const long long L = e->latency;
    const long long bl = e->stages.back().B;
    e->HX = 2 * bl;
    e->XL = 1;
    while (e->XL < e->HX + FDL_CHUNK) e->XL *= 2;
    long long need = FDL_CHUNK + e->stages.back().off + 2 * bl + L + 64;
    long long al = 1;
    while (al < need) al *= 2;
    e->acc_len = al;
    adsp_status st = (prec == ADSP_F64) ? fdl_build<double>(e, (const double *)d_kernel) : fdl_build<float>(e, (const float *)d_kernel);
    if (st == ADSP_OK) st = (prec == ADSP_F64) ? fdl_reset_t<double>(e) : fdl_reset_t<float>(e);
    if (st != ADSP_OK) { fdl_destroy(e); return st; }
    *out = e;
    return ADSP_OK;
}

void fdl_destroy(FdlEngine *e) {
    if (!e) return;
    for (void *p : e->d_H) if (p) cudaFree(p);
    for (void *p : e->d_fdl) if (p) cudaFree(p);
    e->xbuf.release(); e->acc.release(); e->d_io_out.release(); e->yscratch.release();
    delete e;
}

adsp_status fdl_reset(FdlEngine *e) { return e->prec == ADSP_F64 ? fdl_reset_t<double>(e) : fdl_reset_t<float>(e); }
void fdl_set_wet_dry(FdlEngine *e, double wet, double dry) { e->wet = wet; e->dry = dry; }
int fdl_channels(const FdlEngine *e) { return e->channels; }
int fdl_stage_count(const FdlEngine *e) { return (int)e->stages.size(); }
void fdl_stage_info(const FdlEngine *e, int i, int *part, int *count, long long *off) {
    const StageGeom &g = e->stages[(size_t)i];
    if (part) *part = g.B;
    if (count) *count = g.count;
    if (off) *off = g.off;
}

adsp_status fdl_process(FdlEngine *e, const void *in, long long n, long long in_stride, void *out, long long out_stride, bool host_ptrs,
                        bool mix) {
    if (n <= 0) return ADSP_OK;
    if (e->prec == ADSP_F64) return fdl_process_t<double>(e, (const double *)in, n, in_stride, (double *)out, out_stride, host_ptrs, mix);
    return fdl_process_t<float>(e, (const float *)in, n, in_stride, (float *)out, out_stride, host_ptrs, mix);
}

}  // namespace adsp
