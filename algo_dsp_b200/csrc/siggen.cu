// siggen.cu -- device-side dsp/signal generators and their bit-identical host twins.
//
// Step BEFORE the hot path: every benchmark and full-size parity test can build its inputs in HBM instead of pushing
// gigabytes through PCIe.  Formulas: /root/reference/dsp/signal/generate.go -- WhiteNoise :188-205, PinkNoise :210-250
// (Voss-McCartney, 5 bands), LinearSweep :134-154, LogSweep :157-185, Normalize :253-283, RemoveDC :306-324.  The
// reference draws from Go's math/rand (a lagged-Fibonacci generator whose seeding table is not in the tree); here the
// uniforms come from a stateless hash of (seed, stream, index) (siggen_core.h), so a sample depends only on its index:
// any shard of any stream can be generated anywhere (config 5: each GPU generates its own time block plus halo).
#include <algorithm>

#include "engine.cuh"
#include "siggen_core.h"

namespace adsp {
namespace {

template <typename T> __device__ __forceinline__ void put(T *p, double v) { *p = (T)v; }

// ---------------------------------------------------------------- white noise / decaying IR / sweeps: one thread per sample
enum GenKind { GEN_WHITE = 0, GEN_DECAY_IR = 1, GEN_LIN_SWEEP = 2, GEN_LOG_SWEEP = 3, GEN_UNIFORM = 4 };

struct GenArgs {
    long long n, rows, stride, index0;
    long long seed0, seed_step;
    double amplitude;   // white, sweeps
    double p0, p1, p2;  // decaying IR: K, decades; sweeps: f0, k, sample rate
};

template <typename T, int KIND>
__global__ void __launch_bounds__(256) gen_pointwise(T *__restrict__ out, GenArgs a) {
    const long long row = blockIdx.y;
    const uint64_t key = adsp_hash_key((uint64_t)(a.seed0 + row * a.seed_step), 0);
    T *o = out + row * a.stride;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < a.n; j += (long long)gridDim.x * blockDim.x) {
        const uint64_t i = (uint64_t)(a.index0 + j);
        double v;
        if (KIND == GEN_WHITE) v = adsp_white_sample(key, i, a.amplitude);
        else if (KIND == GEN_UNIFORM) v = adsp_hash_uniform(key, i);
        else if (KIND == GEN_DECAY_IR) v = adsp_decaying_ir_sample(key, i, a.p0, a.p1);
        else if (KIND == GEN_LIN_SWEEP) v = adsp_lin_sweep_sample(i, a.p0, a.p1, a.p2, a.amplitude);
        else v = adsp_log_sweep_sample(i, a.p0, a.p1, a.p2, a.amplitude);
        put(o + j, v);
    }
}

// ---------------------------------------------------------------- pink noise
// The reference keeps five "contributions" and rewrites at most one per sample (generate.go:228-245): the output is the
// sum of the most recent value of every band.  With a stateless PRNG that is a segmented "last write wins" scan:
//   1. every thread walks its PINK_R samples and notes the last write per band,
//   2. an exclusive scan over the CTA's threads (operator: the right operand wins where it has a value),
//   3. bands nobody in front of a thread has written yet take the last write BEFORE the CTA's first sample, found by
//      the whole CTA scanning backwards 256 samples at a time (expected distance 1/p_band: 505 samples for the rarest),
//   4. every thread replays its samples from that state and stores sum * amplitude (bands added in index order, :240-243).
constexpr int PINK_R = 8, PINK_THREADS = 256, PINK_TILE = PINK_R * PINK_THREADS;

struct PinkState { double v[5]; unsigned mask; };

__device__ __forceinline__ void pink_merge(PinkState &left, const PinkState &right) {   // left = left then right
#pragma unroll
    for (int b = 0; b < 5; b++) if (right.mask & (1u << b)) left.v[b] = right.v[b];
    left.mask |= right.mask;
}

template <typename T>
__global__ void __launch_bounds__(PINK_THREADS) gen_pink(T *__restrict__ out, GenArgs a) {
    __shared__ long long best[5];
    __shared__ PinkState warp_tot[PINK_THREADS / 32];
    __shared__ int done_flag;
    const long long row = blockIdx.y;
    const uint64_t key = adsp_hash_key((uint64_t)(a.seed0 + row * a.seed_step), 0);
    const long long tile0 = (long long)blockIdx.x * PINK_TILE;              // first sample of the CTA inside this call
    const long long base = a.index0 + tile0;                               // its stream index
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 5) best[tid] = -1;
    // 1. own samples
    int band[PINK_R];
    PinkState mine;
    mine.mask = 0;
#pragma unroll
    for (int b = 0; b < 5; b++) mine.v[b] = 0.0;
    const long long my0 = base + (long long)tid * PINK_R;
#pragma unroll
    for (int r = 0; r < PINK_R; r++) {
        const uint64_t i = (uint64_t)(my0 + r);
        band[r] = adsp_pink_band(adsp_hash_uniform(key, 2 * i));
        if (band[r] >= 0) {
            const double val = adsp_pink_value(key, i, band[r]);
#pragma unroll
            for (int b = 0; b < 5; b++) if (band[r] == b) mine.v[b] = val;
            mine.mask |= 1u << band[r];
        }
    }
    // 2. exclusive scan: warp level by shuffles, then across the 8 warps
    PinkState incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        PinkState other;
        other.mask = __shfl_up_sync(0xffffffffu, incl.mask, o);
#pragma unroll
        for (int b = 0; b < 5; b++) other.v[b] = __shfl_up_sync(0xffffffffu, incl.v[b], o);
        if (lane >= o) { PinkState t = other; pink_merge(t, incl); incl = t; }
    }
    if (lane == 31) warp_tot[warp] = incl;
    PinkState excl;                                                         // everything in front of this thread inside its warp
    excl.mask = __shfl_up_sync(0xffffffffu, incl.mask, 1);
#pragma unroll
    for (int b = 0; b < 5; b++) excl.v[b] = __shfl_up_sync(0xffffffffu, incl.v[b], 1);
    if (lane == 0) { excl.mask = 0; for (int b = 0; b < 5; b++) excl.v[b] = 0.0; }
    __syncthreads();
    PinkState front;
    front.mask = 0;
#pragma unroll
    for (int b = 0; b < 5; b++) front.v[b] = 0.0;
    for (int w = 0; w < warp; w++) pink_merge(front, warp_tot[w]);
    pink_merge(front, excl);
    // 3. carry-in from before the CTA
    for (long long lo = base - PINK_THREADS;; lo -= PINK_THREADS) {
        const long long i = lo + tid;
        if (i >= 0 && i < base) {
            const int b = adsp_pink_band(adsp_hash_uniform(key, 2 * (uint64_t)i));
            if (b >= 0) atomicMax((long long *)&best[b], i);
        }
        __syncthreads();
        if (tid == 0) done_flag = (lo <= 0) || (best[0] >= 0 && best[1] >= 0 && best[2] >= 0 && best[3] >= 0 && best[4] >= 0);
        __syncthreads();
        if (done_flag) break;
    }
    double c[5];
#pragma unroll
    for (int b = 0; b < 5; b++) {
        if (front.mask & (1u << b)) c[b] = front.v[b];
        else { const long long bi = best[b]; c[b] = bi >= 0 ? adsp_pink_value(key, (uint64_t)bi, b) : 0.0; }
    }
    // 4. replay
    T *o = out + row * a.stride + tile0 + (long long)tid * PINK_R;
#pragma unroll
    for (int r = 0; r < PINK_R; r++) {
        const long long j = tile0 + (long long)tid * PINK_R + r;
        if (band[r] >= 0) {
            const double val = adsp_pink_value(key, (uint64_t)(my0 + r), band[r]);
#pragma unroll
            for (int b = 0; b < 5; b++) if (band[r] == b) c[b] = val;
        }
        double sum = 0.0;
#pragma unroll
        for (int b = 0; b < 5; b++) sum = ADSP_ADD(sum, c[b]);
        if (j < a.n) put(o + r, ADSP_MUL(sum, a.amplitude));
    }
}

// ---------------------------------------------------------------- delayed copy + noise (config 4 responses)
// out[p][i] = (i >= d_p ? src[i - d_p] : 0) + white(seed0 + p*seed_step)[i] * noise_amp,  d_p = hash(delay_seed, p) mod delay_mod
template <typename T>
__global__ void __launch_bounds__(256) gen_delay_mix(T *__restrict__ out, const T *__restrict__ src, GenArgs a, long long delay_seed,
                                                     long long delay_mod, long long *__restrict__ delays) {
    const long long row = blockIdx.y;
    const uint64_t key = adsp_hash_key((uint64_t)(a.seed0 + row * a.seed_step), 0);
    const long long d = delay_mod > 0 ? (long long)(adsp_hash_u64(adsp_hash_key((uint64_t)delay_seed, 1), (uint64_t)row) % (uint64_t)delay_mod) : 0;
    if (delays && blockIdx.x == 0 && threadIdx.x == 0) delays[row] = d;
    T *o = out + row * a.stride;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < a.n; j += (long long)gridDim.x * blockDim.x) {
        const double s = j >= d ? (double)src[j - d] : 0.0;
        put(o + j, ADSP_ADD(s, adsp_white_sample(key, (uint64_t)j, a.amplitude)));
    }
}

// ---------------------------------------------------------------- Normalize / RemoveDC
template <typename T>
__global__ void __launch_bounds__(256) absmax_kernel(const T *__restrict__ x, long long n, long long stride, unsigned long long *__restrict__ bits) {
    const T *r = x + (long long)blockIdx.y * stride;
    double m = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double av = fabs((double)r[i]);
        if (av > m) m = av;                       // NaN never compares greater (generate.go:266-269)
    }
    for (int o = 16; o; o >>= 1) { const double ov = __shfl_xor_sync(0xffffffffu, m, o); if (ov > m) m = ov; }
    // non-negative doubles order like their bit patterns
    if ((threadIdx.x & 31) == 0) atomicMax(&bits[blockIdx.y], (unsigned long long)__double_as_longlong(m));
}
template <typename T>
__global__ void __launch_bounds__(256) normalize_kernel(const T *__restrict__ x, T *__restrict__ out, long long n, long long in_stride,
                                                        long long out_stride, const unsigned long long *__restrict__ bits, double target) {
    const double maxabs = __longlong_as_double((long long)bits[blockIdx.y]);
    const bool zero = (maxabs == 0.0) || (target == 0.0);                   // :272-274: all zeros
    const double scale = zero ? 0.0 : target / maxabs;
    const T *r = x + (long long)blockIdx.y * in_stride;
    T *o = out + (long long)blockIdx.y * out_stride;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        o[i] = zero ? (T)0 : (T)ADSP_MUL((double)r[i], scale);
}
// fixed-shape tree sum: 1024 partials per row (thread t of 1024 sums i = t, t + 1024, ... in index order), then pairwise
template <typename T>
__global__ void __launch_bounds__(1024) row_sum_kernel(const T *__restrict__ x, long long n, long long stride, double *__restrict__ sums) {
    __shared__ double sh[1024];
    const T *r = x + (long long)blockIdx.x * stride;
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += 1024) s = ADSP_ADD(s, (double)r[i]);
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int w = 512; w; w >>= 1) {
        if ((int)threadIdx.x < w) sh[threadIdx.x] = ADSP_ADD(sh[threadIdx.x], sh[threadIdx.x + w]);
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[blockIdx.x] = sh[0];
}
template <typename T>
__global__ void __launch_bounds__(256) remove_dc_kernel(const T *__restrict__ x, T *__restrict__ out, long long n, long long in_stride,
                                                        long long out_stride, const double *__restrict__ sums) {
    const double mean = sums[blockIdx.y] / (double)n;                       // :316
    const T *r = x + (long long)blockIdx.y * in_stride;
    T *o = out + (long long)blockIdx.y * out_stride;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        o[i] = (T)ADSP_ADD((double)r[i], -mean);
}

unsigned grid_x(long long n, int per_cta) { return (unsigned)std::max<long long>(1, std::min<long long>((n + per_cta - 1) / per_cta, 1 << 20)); }

template <typename T, int KIND> adsp_status launch_pointwise(adsp_ctx *ctx, void *out, const GenArgs &a) {
    for (long long r0 = 0; r0 < a.rows; r0 += 65535) {
        GenArgs b = a;
        b.rows = std::min<long long>(65535, a.rows - r0);
        b.seed0 = a.seed0 + r0 * a.seed_step;
        dim3 grid(grid_x(a.n, 256 * 4), (unsigned)b.rows);
        gen_pointwise<T, KIND><<<grid, 256, 0, ctx->main>>>((T *)out + r0 * a.stride, b);
        count_launch(ctx);
    }
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}

adsp_status check_gen(adsp_ctx *ctx, const void *out, long long n, long long rows, long long stride, long long index0) {
    if (!ctx || !out) return ADSP_ERR_INVALID_ARG;
    if (n <= 0) { set_error("generator: samples must be > 0"); return ADSP_ERR_INVALID_ARG; }   // generate.go:135,158,189,211
    if (rows <= 0 || index0 < 0 || (rows > 1 && stride < n)) { set_error("generator: bad rows / stride / index"); return ADSP_ERR_INVALID_ARG; }
    return ADSP_OK;
}

template <int KIND>
adsp_status gen_any(adsp_ctx *ctx, void *out, const GenArgs &a, adsp_precision prec) {
    ADSP_TRY(check_gen(ctx, out, a.n, a.rows, a.stride, a.index0));
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    return prec == ADSP_F64 ? launch_pointwise<double, KIND>(ctx, out, a) : launch_pointwise<float, KIND>(ctx, out, a);
}

}  // namespace
}  // namespace adsp

using namespace adsp;

extern "C" {

adsp_status adsp_gen_uniform_device(adsp_ctx *ctx, void *out, int64_t n, int64_t rows, int64_t stride, int64_t seed0, int64_t seed_step,
                                    int64_t index0, adsp_precision prec) {
    GenArgs a{n, rows, stride, index0, seed0, seed_step, 1.0, 0, 0, 0};
    return gen_any<GEN_UNIFORM>(ctx, out, a, prec);
}

adsp_status adsp_gen_white_device(adsp_ctx *ctx, void *out, int64_t n, int64_t rows, int64_t stride, double amplitude, int64_t seed0,
                                  int64_t seed_step, int64_t index0, adsp_precision prec) {
    if (amplitude < 0) { set_error("noise amplitude must be >= 0"); return ADSP_ERR_INVALID_ARG; }   // generate.go:193-195
    GenArgs a{n, rows, stride, index0, seed0, seed_step, amplitude, 0, 0, 0};
    return gen_any<GEN_WHITE>(ctx, out, a, prec);
}

adsp_status adsp_gen_pink_device(adsp_ctx *ctx, void *out, int64_t n, int64_t rows, int64_t stride, double amplitude, int64_t seed0,
                                 int64_t seed_step, int64_t index0, adsp_precision prec) {
    if (amplitude < 0) { set_error("noise amplitude must be >= 0"); return ADSP_ERR_INVALID_ARG; }   // generate.go:215-217
    ADSP_TRY(check_gen(ctx, out, n, rows, stride, index0));
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    const long long tiles = (n + PINK_TILE - 1) / PINK_TILE;
    if (tiles > 0x7fffffffLL) { set_error("generator: row too long for one call"); return ADSP_ERR_INVALID_ARG; }
    for (long long r0 = 0; r0 < rows; r0 += 65535) {
        GenArgs a{n, std::min<long long>(65535, rows - r0), stride, index0, seed0 + r0 * seed_step, seed_step, amplitude, 0, 0, 0};
        dim3 grid((unsigned)tiles, (unsigned)a.rows);
        if (prec == ADSP_F64) gen_pink<double><<<grid, PINK_THREADS, 0, ctx->main>>>((double *)out + r0 * stride, a);
        else gen_pink<float><<<grid, PINK_THREADS, 0, ctx->main>>>((float *)out + r0 * stride, a);
        count_launch(ctx);
    }
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}

adsp_status adsp_gen_decaying_ir_device(adsp_ctx *ctx, void *out, int64_t taps, int64_t rows, int64_t stride, double decades, int64_t seed0,
                                        int64_t seed_step, adsp_precision prec) {
    GenArgs a{taps, rows, stride, 0, seed0, seed_step, 1.0, (double)taps, decades, 0};
    return gen_any<GEN_DECAY_IR>(ctx, out, a, prec);
}

// samples [index0, index0 + n) of a sweep that is `total_samples` long (k depends on the total duration)
adsp_status adsp_gen_linear_sweep_device(adsp_ctx *ctx, void *out, int64_t n, int64_t index0, int64_t total_samples, double start_hz,
                                         double end_hz, double amplitude, double sample_rate, adsp_precision prec) {
    if (total_samples <= 0 || !(sample_rate > 0)) { set_error("linear sweep: samples and sample rate must be > 0"); return ADSP_ERR_INVALID_ARG; }
    const double duration = (double)total_samples / sample_rate;                         // generate.go:144-145
    GenArgs a{n, 1, n, index0, 0, 0, amplitude, start_hz, (end_hz - start_hz) / duration, sample_rate};
    return gen_any<GEN_LIN_SWEEP>(ctx, out, a, prec);
}

adsp_status adsp_gen_log_sweep_device(adsp_ctx *ctx, void *out, int64_t n, int64_t index0, int64_t total_samples, double start_hz, double end_hz,
                                      double amplitude, double sample_rate, adsp_precision prec) {
    if (total_samples <= 0 || !(sample_rate > 0)) { set_error("log sweep: samples and sample rate must be > 0"); return ADSP_ERR_INVALID_ARG; }
    if (!(start_hz > 0) || !(end_hz > 0)) { set_error("log sweep frequencies must be > 0"); return ADSP_ERR_INVALID_ARG; }   // :166-168
    const double duration = (double)total_samples / sample_rate;
    const double k = log(end_hz / start_hz) / duration;                                  // :170-171 (host libm, once per call)
    if (k == 0) {                                                                          // :174-176: a plain sine = linear sweep with k = 0
        GenArgs a{n, 1, n, index0, 0, 0, amplitude, start_hz, 0.0, sample_rate};
        return gen_any<GEN_LIN_SWEEP>(ctx, out, a, prec);
    }
    GenArgs a{n, 1, n, index0, 0, 0, amplitude, start_hz, k, sample_rate};
    return gen_any<GEN_LOG_SWEEP>(ctx, out, a, prec);
}

adsp_status adsp_gen_delay_mix_device(adsp_ctx *ctx, void *out, int64_t n, int64_t rows, int64_t stride, const void *src, double noise_amplitude,
                                      int64_t seed0, int64_t seed_step, int64_t delay_seed, int64_t delay_mod, int64_t *delays_dev,
                                      adsp_precision prec) {
    ADSP_TRY(check_gen(ctx, out, n, rows, stride, 0));
    if (!src || delay_mod < 0) return ADSP_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    for (long long r0 = 0; r0 < rows; r0 += 65535) {
        GenArgs a{n, std::min<long long>(65535, rows - r0), stride, 0, seed0 + r0 * seed_step, seed_step, noise_amplitude, 0, 0, 0};
        dim3 grid(grid_x(n, 256 * 4), (unsigned)a.rows);
        // the delay of row r is hash(delay_seed, r): rows of later launches must keep their global row number
        if (r0 != 0) { set_error("delay mix: at most 65535 rows per call"); return ADSP_ERR_INVALID_ARG; }
        if (prec == ADSP_F64) gen_delay_mix<double><<<grid, 256, 0, ctx->main>>>((double *)out, (const double *)src, a, delay_seed, delay_mod, (long long *)delays_dev);
        else gen_delay_mix<float><<<grid, 256, 0, ctx->main>>>((float *)out, (const float *)src, a, delay_seed, delay_mod, (long long *)delays_dev);
        count_launch(ctx);
    }
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}

adsp_status adsp_normalize_device(adsp_ctx *ctx, const void *in, int64_t n, int64_t rows, int64_t in_stride, double target_peak, void *out,
                                  int64_t out_stride, adsp_precision prec) {
    if (!ctx || !in || !out) return ADSP_ERR_INVALID_ARG;
    if (target_peak < 0) { set_error("normalize target peak must be >= 0"); return ADSP_ERR_INVALID_ARG; }   // generate.go:254-256
    if (n <= 0) { set_error("normalize input must not be empty"); return ADSP_ERR_EMPTY_INPUT; }            // :258-260
    if (rows <= 0 || rows > 65535) { set_error("normalize: 1 .. 65535 rows per call"); return ADSP_ERR_INVALID_ARG; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    ADSP_TRY(ctx->d_small.reserve((size_t)rows * 8 + 64));
    unsigned long long *bits = (unsigned long long *)ctx->d_small.p;
    ADSP_CUDA(cudaMemsetAsync(bits, 0, (size_t)rows * 8, ctx->main));
    dim3 grid(grid_x(n, 256 * 8), (unsigned)rows);
    if (prec == ADSP_F64) {
        absmax_kernel<double><<<grid, 256, 0, ctx->main>>>((const double *)in, n, in_stride, bits);
        normalize_kernel<double><<<grid, 256, 0, ctx->main>>>((const double *)in, (double *)out, n, in_stride, out_stride, bits, target_peak);
    } else {
        absmax_kernel<float><<<grid, 256, 0, ctx->main>>>((const float *)in, n, in_stride, bits);
        normalize_kernel<float><<<grid, 256, 0, ctx->main>>>((const float *)in, (float *)out, n, in_stride, out_stride, bits, target_peak);
    }
    count_launch(ctx, 2);
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}

adsp_status adsp_remove_dc_device(adsp_ctx *ctx, const void *in, int64_t n, int64_t rows, int64_t in_stride, void *out, int64_t out_stride,
                                  adsp_precision prec) {
    if (!ctx || !in || !out) return ADSP_ERR_INVALID_ARG;
    if (n <= 0) { set_error("remove dc input must not be empty"); return ADSP_ERR_EMPTY_INPUT; }             // generate.go:307-309
    if (rows <= 0 || rows > 65535) { set_error("remove dc: 1 .. 65535 rows per call"); return ADSP_ERR_INVALID_ARG; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    ADSP_TRY(ctx->d_small.reserve((size_t)rows * 8 + 64));
    double *sums = (double *)ctx->d_small.p;
    dim3 grid(grid_x(n, 256 * 8), (unsigned)rows);
    if (prec == ADSP_F64) {
        row_sum_kernel<double><<<(unsigned)rows, 1024, 0, ctx->main>>>((const double *)in, n, in_stride, sums);
        remove_dc_kernel<double><<<grid, 256, 0, ctx->main>>>((const double *)in, (double *)out, n, in_stride, out_stride, sums);
    } else {
        row_sum_kernel<float><<<(unsigned)rows, 1024, 0, ctx->main>>>((const float *)in, n, in_stride, sums);
        remove_dc_kernel<float><<<grid, 256, 0, ctx->main>>>((const float *)in, (float *)out, n, in_stride, out_stride, sums);
    }
    count_launch(ctx, 2);
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}

// ---------------------------------------------------------------- host twins (same siggen_core.h arithmetic, fp64)
void adsp_gen_uniform_host(double *out, int64_t n, int64_t seed, int64_t index0) {
    const uint64_t key = adsp_hash_key((uint64_t)seed, 0);
    for (int64_t j = 0; j < n; j++) out[j] = adsp_hash_uniform(key, (uint64_t)(index0 + j));
}
void adsp_gen_white_host(double *out, int64_t n, double amplitude, int64_t seed, int64_t index0) {
    const uint64_t key = adsp_hash_key((uint64_t)seed, 0);
    for (int64_t j = 0; j < n; j++) out[j] = adsp_white_sample(key, (uint64_t)(index0 + j), amplitude);
}
void adsp_gen_pink_host(double *out, int64_t n, double amplitude, int64_t seed, int64_t index0) {
    const uint64_t key = adsp_hash_key((uint64_t)seed, 0);
    double c[5] = {0, 0, 0, 0, 0};
    // state at index0: the last write of every band before it (sequential equivalent of the device carry-in search)
    unsigned found = 0;
    for (int64_t i = index0 - 1; i >= 0 && found != 31u; i--) {
        const int b = adsp_pink_band(adsp_hash_uniform(key, 2 * (uint64_t)i));
        if (b >= 0 && !(found & (1u << b))) { c[b] = adsp_pink_value(key, (uint64_t)i, b); found |= 1u << b; }
    }
    for (int64_t j = 0; j < n; j++) {
        const uint64_t i = (uint64_t)(index0 + j);
        const int b = adsp_pink_band(adsp_hash_uniform(key, 2 * i));
        if (b >= 0) c[b] = adsp_pink_value(key, i, b);
        double sum = 0.0;
        for (int k = 0; k < 5; k++) sum = ADSP_ADD(sum, c[k]);
        out[j] = ADSP_MUL(sum, amplitude);
    }
}
void adsp_gen_decaying_ir_host(double *out, int64_t taps, double decades, int64_t seed) {
    const uint64_t key = adsp_hash_key((uint64_t)seed, 0);
    for (int64_t i = 0; i < taps; i++) out[i] = adsp_decaying_ir_sample(key, (uint64_t)i, (double)taps, decades);
}
void adsp_gen_linear_sweep_host(double *out, int64_t n, int64_t index0, int64_t total_samples, double start_hz, double end_hz, double amplitude,
                                double sample_rate) {
    const double duration = (double)total_samples / sample_rate;
    const double k = (end_hz - start_hz) / duration;
    for (int64_t j = 0; j < n; j++) out[j] = adsp_lin_sweep_sample((uint64_t)(index0 + j), start_hz, k, sample_rate, amplitude);
}
void adsp_gen_log_sweep_host(double *out, int64_t n, int64_t index0, int64_t total_samples, double start_hz, double end_hz, double amplitude,
                             double sample_rate) {
    const double duration = (double)total_samples / sample_rate;
    const double k = log(end_hz / start_hz) / duration;
    for (int64_t j = 0; j < n; j++)
        out[j] = (k == 0) ? adsp_lin_sweep_sample((uint64_t)(index0 + j), start_hz, 0.0, sample_rate, amplitude)
                          : adsp_log_sweep_sample((uint64_t)(index0 + j), start_hz, k, sample_rate, amplitude);
}
int64_t adsp_gen_delay_host(int64_t delay_seed, int64_t row, int64_t delay_mod) {
    return delay_mod > 0 ? (int64_t)(adsp_hash_u64(adsp_hash_key((uint64_t)delay_seed, 1), (uint64_t)row) % (uint64_t)delay_mod) : 0;
}

}  // extern "C"
