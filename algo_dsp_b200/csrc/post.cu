// post.cu -- the consumers right after the dsp/conv path, on device data:
//   measure/ir   SchroederIntegral (measure/ir/ir.go:94-130), FindImpulseStart (:381-404), findPeak (:406-424)
//   measure/sweep LogSweep.Generate (sweep.go:73-94), InverseFilter (:104-155), Deconvolve (:164-239) -- the latter is a full
//                linear convolution with the inverse filter and runs on the library's FFT engine
//   dsp/filter/fir Filter.ProcessBlock (filter.go:64-103), a stateful block FIR on the direct-convolution kernel
//   dsp/resample  Resampler.Process (resample.go:249-292) with the reference's polyphase design (resample_design.go:9-72)
// so a correlation / deconvolution result can be analysed, filtered or rate-converted without leaving HBM.
#include <algorithm>
#include <cmath>

#include "aux_kernels.cuh"
#include "engine.cuh"
#include "siggen_core.h"

namespace adsp {
namespace {

// ================================================================ measure/ir
constexpr int SCH_THREADS = 256, SCH_PER = 8, SCH_TILE = SCH_THREADS * SCH_PER;

// pass 1: energy of every tile of SCH_TILE samples (fixed summation tree)
__global__ void __launch_bounds__(SCH_THREADS) sch_tile_energy(const double *__restrict__ x, long long n, long long stride, long long tiles, double *__restrict__ tsum) {
    __shared__ double sh[SCH_THREADS];
    const double *r = x + (long long)blockIdx.y * stride;
    const long long base = (long long)blockIdx.x * SCH_TILE + (long long)threadIdx.x * SCH_PER;
    double s = 0.0;
#pragma unroll
    for (int i = SCH_PER - 1; i >= 0; i--) { const long long k = base + i; if (k < n) { const double v = r[k]; s = ADSP_ADD(s, ADSP_MUL(v, v)); } }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int w = SCH_THREADS / 2; w; w >>= 1) {
        if ((int)threadIdx.x < w) sh[threadIdx.x] = ADSP_ADD(sh[threadIdx.x], sh[threadIdx.x + w]);
        __syncthreads();
    }
    if (threadIdx.x == 0) tsum[(long long)blockIdx.y * tiles + blockIdx.x] = sh[0];
}
// pass 2: per row, energy BEHIND every tile (exclusive suffix sum over tiles, last tile first) and the total.  One CTA per
// row: thread t owns a contiguous run of tiles (its own suffix sums first, then an exclusive suffix scan over the runs).
__global__ void __launch_bounds__(256) sch_tile_suffix(double *__restrict__ tsum, long long tiles, double *__restrict__ total) {
    __shared__ double sh[256];
    double *t = tsum + (long long)blockIdx.x * tiles;
    const long long per = (tiles + 255) / 256;
    const long long lo = (long long)threadIdx.x * per, hi = lo + per < tiles ? lo + per : tiles;
    double acc = 0.0;
    for (long long b = hi - 1; b >= lo; b--) acc = ADSP_ADD(acc, t[b]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {                      // inclusive suffix scan over the threads' runs
        const double add = ((int)threadIdx.x + d < 256) ? sh[threadIdx.x + d] : 0.0;
        __syncthreads();
        sh[threadIdx.x] = ADSP_ADD(sh[threadIdx.x], add);
        __syncthreads();
    }
    double behind = (threadIdx.x + 1 < 256) ? sh[threadIdx.x + 1] : 0.0;
    if (threadIdx.x == 0) total[blockIdx.x] = sh[0];
    for (long long b = hi - 1; b >= lo; b--) { const double e = t[b]; t[b] = behind; behind = ADSP_ADD(behind, e); }
}
// pass 3: backward cumulative energy inside the tile + energy behind it, normalised, in dB (ir.go:117-127)
__global__ void __launch_bounds__(SCH_THREADS) sch_emit(const double *__restrict__ x, long long n, long long stride, long long tiles, const double *__restrict__ tsuf,
                                                        const double *__restrict__ total, double *__restrict__ out, long long out_stride) {
    __shared__ double sh[SCH_THREADS];
    const double *r = x + (long long)blockIdx.y * stride;
    double *o = out + (long long)blockIdx.y * out_stride;
    const long long base = (long long)blockIdx.x * SCH_TILE + (long long)threadIdx.x * SCH_PER;
    double c[SCH_PER];
    double s = 0.0;
#pragma unroll
    for (int i = SCH_PER - 1; i >= 0; i--) {           // own samples, last first
        const long long k = base + i;
        if (k < n) { const double v = r[k]; s = ADSP_ADD(s, ADSP_MUL(v, v)); }
        c[i] = s;
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    // energy of the threads behind this one (exclusive suffix over the CTA): Hillis-Steele on the reversed order
    for (int d = 1; d < SCH_THREADS; d <<= 1) {
        const double add = ((int)threadIdx.x + d < SCH_THREADS) ? sh[threadIdx.x + d] : 0.0;
        __syncthreads();
        sh[threadIdx.x] = ADSP_ADD(sh[threadIdx.x], add);
        __syncthreads();
    }
    const double behind = ADSP_ADD((threadIdx.x + 1 < SCH_THREADS) ? sh[threadIdx.x + 1] : 0.0, tsuf[(long long)blockIdx.y * tiles + blockIdx.x]);
    const double tot = total[blockIdx.y];
#pragma unroll
    for (int i = 0; i < SCH_PER; i++) {
        const long long k = base + i;
        if (k >= n) continue;
        const double e = ADSP_ADD(c[i], behind);
        double v = e;                                      // totalEnergy <= 0: the raw sums are returned (:112-114)
        if (tot > 0.0) { const double ratio = e / tot; v = ratio <= 0.0 ? -200.0 : 10.0 * log10(ratio); }
        o[k] = v;
    }
}

// abs-max per row (NaN never compares greater: ir.go:387-391, :411-417), as the bit pattern of a non-negative double
__global__ void __launch_bounds__(256) ir_absmax(const double *__restrict__ x, long long n, long long stride, unsigned long long *__restrict__ bits) {
    const double *r = x + (long long)blockIdx.y * stride;
    double m = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double av = fabs(r[i]);
        if (av > m) m = av;
    }
    for (int o = 16; o; o >>= 1) { const double ov = __shfl_xor_sync(0xffffffffu, m, o); if (ov > m) m = ov; }
    if ((threadIdx.x & 31) == 0) atomicMax(&bits[blockIdx.y], (unsigned long long)__double_as_longlong(m));
}
// first index with |x[i]| >= peak * ratio (ir.go:393-398); idx must hold n on entry, rows without a hit are mapped to 0 by ir_first_fix
__global__ void __launch_bounds__(256) ir_first_ge(const double *__restrict__ x, long long n, long long stride, const unsigned long long *__restrict__ bits, double ratio,
                                                   long long *__restrict__ idx) {
    const double *r = x + (long long)blockIdx.y * stride;
    const double thr = ADSP_MUL(__longlong_as_double((long long)bits[blockIdx.y]), ratio);
    long long best = n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n && i < best; i += (long long)gridDim.x * blockDim.x)
        if (fabs(r[i]) >= thr) { best = i; break; }
    for (int o = 16; o; o >>= 1) { const long long ob = __shfl_xor_sync(0xffffffffu, best, o); if (ob < best) best = ob; }
    if ((threadIdx.x & 31) == 0 && best < n) atomicMin((unsigned long long *)&idx[blockIdx.y], (unsigned long long)best);
}
__global__ void ir_first_fix(long long *idx, long long n, long long rows) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows && idx[r] >= n) idx[r] = 0;               // "return 0" when nothing reaches the threshold (:400)
}
__global__ void ir_fill_ll(long long *p, long long v, long long count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) p[i] = v;
}

unsigned gx(long long n, int per) { return (unsigned)std::max<long long>(1, std::min<long long>((n + per - 1) / per, 1 << 16)); }

adsp_status first_ge_rows(adsp_ctx *ctx, const double *x, long long n, long long rows, long long stride, double ratio, long long *idx_dev) {
    ADSP_TRY(ctx->d_small.reserve((size_t)rows * 8 + 64));
    unsigned long long *bits = (unsigned long long *)ctx->d_small.p;
    ADSP_CUDA(cudaMemsetAsync(bits, 0, (size_t)rows * 8, ctx->main));
    ir_fill_ll<<<(unsigned)((rows + 255) / 256), 256, 0, ctx->main>>>(idx_dev, n, rows);
    dim3 grid(gx(n, 256 * 8), (unsigned)rows);
    ir_absmax<<<grid, 256, 0, ctx->main>>>(x, n, stride, bits);
    ir_first_ge<<<grid, 256, 0, ctx->main>>>(x, n, stride, bits, ratio, idx_dev);
    ir_first_fix<<<(unsigned)((rows + 255) / 256), 256, 0, ctx->main>>>(idx_dev, n, rows);
    count_launch(ctx, 4);
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}

// ================================================================ measure/sweep
struct SweepArgs { long long n; double f1, T, lnr, sr, scale; };
// Generate (sweep.go:87-91): phase = 2 pi f1 T / ln(r) * (exp(t/T ln r) - 1), in cycles for the exact reduction
ADSP_HD double logsweep_sample(long long i, const SweepArgs &a) {
    const double t = (double)i / a.sr;
    const double cyc = ADSP_MUL(ADSP_MUL(a.f1, a.T) / a.lnr, ADSP_ADD(adsp_det_exp(ADSP_MUL(t / a.T, a.lnr)), -1.0));
    return adsp_det_sin2pi(cyc);
}
// InverseFilter (sweep.go:126-153): inv[i] = sweep[n-1-i] * f1 / f_inst(t_j) / (T f1 / ln r * sr)
ADSP_HD double loginverse_sample(long long i, const SweepArgs &a) {
    const long long j = a.n - 1 - i;
    const double t = (double)j / a.sr;
    const double finst = ADSP_MUL(a.f1, adsp_det_exp(ADSP_MUL(t / a.T, a.lnr)));
    const double amp = a.f1 / finst;
    return ADSP_MUL(ADSP_MUL(logsweep_sample(j, a), amp), a.scale);
}
template <int INVERSE> __global__ void __launch_bounds__(256) sweep_kernel(double *__restrict__ out, SweepArgs a) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x)
        out[i] = INVERSE ? loginverse_sample(i, a) : logsweep_sample(i, a);
}

adsp_status sweep_args(double f1, double f2, double duration, double sr, SweepArgs *a) {
    // LogSweep.Validate, sweep.go:37-55
    if (!(f1 > 0) || !(f2 > 0)) { set_error("sweep: frequency must be positive"); return ADSP_ERR_INVALID_ARG; }
    if (f1 >= f2) { set_error("sweep: start frequency must be less than end frequency"); return ADSP_ERR_INVALID_ARG; }
    if (!(duration > 0)) { set_error("sweep: duration must be positive"); return ADSP_ERR_INVALID_ARG; }
    if (!(sr > 0)) { set_error("sweep: sample rate must be positive"); return ADSP_ERR_INVALID_ARG; }
    a->n = (long long)llround(duration * sr);              // samples(), :58-60
    a->f1 = f1; a->T = duration; a->lnr = log(f2 / f1); a->sr = sr;
    const double norm = duration * f1 / a->lnr * sr;       // :145-147
    a->scale = norm > 0 ? 1.0 / norm : 1.0;
    if (a->n <= 0) { set_error("sweep: duration shorter than one sample"); return ADSP_ERR_INVALID_ARG; }
    return ADSP_OK;
}

// ================================================================ dsp/resample
// y_g = sum_k taps[ph + k*up] * x[idx - k],  acc = g*down, idx = acc / up, ph = acc % up  (closed form of the phase /
// inputIndex recurrence of resample.go:282-284); samples outside [base, last] are skipped (:272-275), products are added
// in tap order with separate roundings, as the Go loop does (:277)
__global__ void __launch_bounds__(256) resample_kernel(const double *__restrict__ work, long long work_stride, long long base, long long last,
                                                       const double *__restrict__ taps, int ntaps, long long up, long long down, long long g0, long long nout,
                                                       double *__restrict__ out, long long out_stride) {
    const double *w = work + (long long)blockIdx.y * work_stride;
    double *o = out + (long long)blockIdx.y * out_stride;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < nout; j += (long long)gridDim.x * blockDim.x) {
        const long long acc = (g0 + j) * down;
        const long long idx = acc / up;
        const int ph = (int)(acc - idx * up);
        double y = 0.0;
        int k = 0;
        for (int tp = ph; tp < ntaps; tp += (int)up, k++) {
            const long long i = idx - k;
            if (i < base || i > last) continue;
            y = ADSP_ADD(y, ADSP_MUL(taps[tp], w[i - base]));
        }
        o[j] = y;
    }
}

// The same sum from shared memory, R outputs of one polyphase branch per thread.  The kernel above gathers
// taps[ph + k*up] from global memory with a different phase in every lane (32 sectors per warp load) and runs one dependent
// add chain per thread.  Here a CTA takes P*R consecutive outputs of one channel (P = a multiple of `up`, so outputs
// t, t+P, t+2P, ... share their phase and sit exactly c*down input samples apart), stages the prototype (the reference's
// k-major layout: for a fixed k a warp reads phases ph0 + lane*down mod up of one row) and the input window once, and
// every thread carries R independent accumulators: per tap one coefficient read and R sample reads for R outputs.
// The input is [hist (H samples) | blk] so that device callers are read where they are (no work-row copy).
constexpr int RS_R = 8;
__global__ void __launch_bounds__(512) resample_phase_kernel(const double *__restrict__ hist, long long hist_stride, long long H, const double *__restrict__ blk,
                                                             long long blk_stride, long long base, long long last, const double *__restrict__ taps, int ntaps,
                                                             int up, long long down, long long g0, long long nout, double *__restrict__ out,
                                                             long long out_stride, int P, int lmax) {
    extern __shared__ double rs_smem[];
    double *tp_s = rs_smem;                 // [ntaps]
    double *x_s = rs_smem + ntaps;          // window [lo, hi] of absolute input indices
    const double *hs = hist + (long long)blockIdx.y * hist_stride;
    const double *bs = blk + (long long)blockIdx.y * blk_stride;
    double *o = out + (long long)blockIdx.y * out_stride;
    const long long tile = (long long)P * RS_R;
    const long long j0 = (long long)blockIdx.x * tile;
    const long long jn = (nout - j0 < tile) ? (nout - j0) : tile;
    const long long lo = ((g0 + j0) * down) / up - (lmax - 1), hi = ((g0 + j0 + jn - 1) * down) / up;
    for (int i = threadIdx.x; i < ntaps; i += blockDim.x) tp_s[i] = taps[i];
    for (long long i = lo + threadIdx.x; i <= hi; i += blockDim.x) {
        double v = 0.0;                                         // slots outside [base, last] are never used
        if (i >= base && i <= last) { const long long off = i - base; v = off < H ? hs[off] : bs[off - H]; }
        x_s[i - lo] = v;
    }
    __syncthreads();
    const bool interior = lo >= base && hi <= last;
    const int step = (int)((long long)(P / up) * down);         // input distance between a thread's consecutive outputs
    for (int t = threadIdx.x; t < P; t += blockDim.x) {
        if (t >= jn) break;
        const long long acc = (g0 + j0 + t) * down;
        const long long idx = acc / up;
        const int ph = (int)(acc - idx * up);
        const int xi = (int)(idx - lo);
        const int lph = ph < ntaps ? (ntaps - ph + up - 1) / up : 0;          // taps of this branch
        const int nr = (int)((jn - t + P - 1) / P < RS_R ? (jn - t + P - 1) / P : RS_R);   // outputs of this thread inside the block
        double y[RS_R];
#pragma unroll
        for (int r = 0; r < RS_R; r++) y[r] = 0.0;
        if (interior && nr == RS_R) {
            const double *tq = tp_s + ph;
            const double *xq = x_s + xi;
#pragma unroll 4
            for (int k = 0; k < lph; k++) {
                const double c = tq[k * up];
#pragma unroll
                for (int r = 0; r < RS_R; r++) y[r] = ADSP_ADD(y[r], ADSP_MUL(c, xq[r * step - k]));
            }
        } else {
            for (int k = 0; k < lph; k++) {
                const double c = tp_s[ph + k * up];
#pragma unroll
                for (int r = 0; r < RS_R; r++) {
                    const long long i = idx + (long long)r * step - k;
                    if (r < nr && i >= base && i <= last) y[r] = ADSP_ADD(y[r], ADSP_MUL(c, x_s[xi + r * step - k]));
                }
            }
        }
#pragma unroll
        for (int r = 0; r < RS_R; r++)
            if (r < nr) o[j0 + t + (long long)r * P] = y[r];
    }
}

}  // namespace
}  // namespace adsp

using namespace adsp;

// ---------------------------------------------------------------- opaque handles
struct adsp_fir {
    adsp_ctx *ctx = nullptr;
    long long ntaps = 0;
    int channels = 1;
    DevBuf taps;        // coefficients in the order the block path applies them (see adsp_fir_create)
    std::vector<double> h_taps;   // the same on the host: up to 1024 of them travel as a kernel parameter
    DevBuf work;        // [channels][ntaps-1 + block]: history in front of the current block (copy path); block rows (host calls)
    DevBuf full;        // direct-convolution output of the work rows (copy path)
    long long work_cap = 0;
    // in-place path (up to 1024 taps): the last ntaps-1 samples of every channel, double buffered, and the samples in front
    // of every segment of the current block
    bool inplace = false;
    DevBuf hist[2], halo;
    int cur = 0;
};

struct adsp_resampler {
    adsp_ctx *ctx = nullptr;
    int up = 1, down = 1, quality = 1, channels = 1, max_phase_len = 0;
    std::vector<double> taps;
    DevBuf d_taps, work, outbuf;
    long long total_in = 0, out_count = 0, hist_len = 0;    // hist_len samples of every channel sit at the front of `work`
    long long work_stride = 0;
};

namespace {

long long gcd_ll(long long a, long long b) { a = a < 0 ? -a : a; b = b < 0 ? -b : b; while (b) { const long long t = a % b; a = b; b = t; } return a ? a : 1; }

// resample_design.go:131-181
double sincf(double x) { if (fabs(x) < 1e-12) return 1; const double pix = M_PI * x; return sin(pix) / pix; }
double bessel_i0(double x) {
    double sum = 1.0, term = 1.0;
    const double x2 = (x * x) / 4;
    for (int k = 1; k < 64; k++) { term *= x2 / (double)(k * k); sum += term; if (term < 1e-16 * sum) break; }
    return sum;
}
double kaiser(int i, int n, double beta) {
    if (n <= 1 || beta == 0) return 1;
    const double t = 2 * (double)i / (double)(n - 1) - 1;
    const double a = sqrt(std::max(0.0, 1 - t * t));
    return bessel_i0(beta * a) / bessel_i0(beta);
}

long long predict_len(const adsp_resampler *r, long long input_len) {          // resample.go:295-314 in closed form
    if (input_len <= 0) return 0;
    const long long last = r->total_in + input_len - 1;
    // outputs g with (g*down)/up <= last  <=>  g*down <= last*up + up - 1
    const long long gmax = (last * r->up + r->up - 1) / r->down;               // largest such g
    const long long cnt = gmax + 1 - r->out_count;
    return cnt > 0 ? cnt : 0;
}

}  // namespace

extern "C" {

// ================================================================ measure/ir
adsp_status adsp_ir_schroeder_device(adsp_ctx *ctx, const double *ir_dev, int64_t n, int64_t rows, int64_t stride, double *out_dev, int64_t out_stride) {
    if (!ctx) return ADSP_ERR_INVALID_ARG;
    if (n <= 0) { set_error("ir: impulse response is empty"); return ADSP_ERR_EMPTY_IR; }            // ErrEmptyIR, ir.go:95-97
    if (!ir_dev || !out_dev || rows <= 0 || rows > 65535) { set_error("schroeder: 1 .. 65535 rows per call"); return ADSP_ERR_INVALID_ARG; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    const long long tiles = (n + SCH_TILE - 1) / SCH_TILE;
    if (tiles > 0x7fffffffLL) { set_error("schroeder: row too long"); return ADSP_ERR_INVALID_ARG; }
    ADSP_TRY(ctx->d_tmp.reserve((size_t)(rows * tiles + rows) * sizeof(double)));
    double *tsum = (double *)ctx->d_tmp.p, *total = tsum + rows * tiles;
    dim3 grid((unsigned)tiles, (unsigned)rows);
    sch_tile_energy<<<grid, SCH_THREADS, 0, ctx->main>>>(ir_dev, n, stride, tiles, tsum);
    sch_tile_suffix<<<(unsigned)rows, 256, 0, ctx->main>>>(tsum, tiles, total);
    sch_emit<<<grid, SCH_THREADS, 0, ctx->main>>>(ir_dev, n, stride, tiles, tsum, total, out_dev, out_stride);
    count_launch(ctx, 3);
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}

adsp_status adsp_ir_find_impulse_start_device(adsp_ctx *ctx, const double *ir_dev, int64_t n, int64_t rows, int64_t stride, double threshold_ratio,
                                              int64_t *index_dev) {
    if (!ctx) return ADSP_ERR_INVALID_ARG;
    if (n <= 0) { set_error("ir: impulse response is empty"); return ADSP_ERR_EMPTY_IR; }            // ir.go:382-384
    if (!ir_dev || !index_dev || rows <= 0 || rows > 65535) return ADSP_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    return first_ge_rows(ctx, ir_dev, n, rows, stride, threshold_ratio, (long long *)index_dev);
}

// findPeak (ir.go:406-424): index of the absolute maximum, first one wins = the first sample that reaches the peak itself
adsp_status adsp_ir_find_peak_device(adsp_ctx *ctx, const double *ir_dev, int64_t n, int64_t rows, int64_t stride, int64_t *index_dev) {
    return adsp_ir_find_impulse_start_device(ctx, ir_dev, n, rows, stride, 1.0, index_dev);
}

adsp_status adsp_ir_schroeder(adsp_ctx *ctx, const double *ir, int64_t n, double *out) {
    if (!ctx) return ADSP_ERR_INVALID_ARG;
    if (n <= 0) { set_error("ir: impulse response is empty"); return ADSP_ERR_EMPTY_IR; }
    if (!ir || !out) return ADSP_ERR_INVALID_ARG;
    void *din = nullptr, *dout = nullptr;
    adsp_status st = adsp_device_alloc(ctx, (size_t)n * 8, &din);
    if (st == ADSP_OK) st = adsp_device_alloc(ctx, (size_t)n * 8, &dout);
    if (st == ADSP_OK) st = adsp_memcpy_h2d(ctx, din, ir, (size_t)n * 8);
    if (st == ADSP_OK) st = adsp_ir_schroeder_device(ctx, (const double *)din, n, 1, n, (double *)dout, n);
    if (st == ADSP_OK) st = adsp_memcpy_d2h(ctx, out, dout, (size_t)n * 8);
    adsp_device_free(ctx, din);
    adsp_device_free(ctx, dout);
    return st;
}

adsp_status adsp_ir_find_impulse_start(adsp_ctx *ctx, const double *ir, int64_t n, double threshold_ratio, int64_t *index) {
    if (!ctx || !index) return ADSP_ERR_INVALID_ARG;
    if (n <= 0) { set_error("ir: impulse response is empty"); return ADSP_ERR_EMPTY_IR; }
    if (!ir) return ADSP_ERR_INVALID_ARG;
    void *din = nullptr, *didx = nullptr;
    adsp_status st = adsp_device_alloc(ctx, (size_t)n * 8, &din);
    if (st == ADSP_OK) st = adsp_device_alloc(ctx, 8, &didx);
    if (st == ADSP_OK) st = adsp_memcpy_h2d(ctx, din, ir, (size_t)n * 8);
    if (st == ADSP_OK) st = adsp_ir_find_impulse_start_device(ctx, (const double *)din, n, 1, n, threshold_ratio, (int64_t *)didx);
    if (st == ADSP_OK) st = adsp_memcpy_d2h(ctx, index, didx, 8);
    adsp_device_free(ctx, din);
    adsp_device_free(ctx, didx);
    return st;
}

// ================================================================ measure/sweep
int64_t adsp_logsweep_samples(double duration, double sample_rate) { return (int64_t)llround(duration * sample_rate); }

adsp_status adsp_logsweep_generate_device(adsp_ctx *ctx, double *out_dev, double start_hz, double end_hz, double duration, double sample_rate) {
    if (!ctx || !out_dev) return ADSP_ERR_INVALID_ARG;
    SweepArgs a;
    ADSP_TRY(sweep_args(start_hz, end_hz, duration, sample_rate, &a));
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    sweep_kernel<0><<<gx(a.n, 256 * 4), 256, 0, ctx->main>>>(out_dev, a);
    count_launch(ctx);
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}
adsp_status adsp_logsweep_inverse_filter_device(adsp_ctx *ctx, double *out_dev, double start_hz, double end_hz, double duration, double sample_rate) {
    if (!ctx || !out_dev) return ADSP_ERR_INVALID_ARG;
    SweepArgs a;
    ADSP_TRY(sweep_args(start_hz, end_hz, duration, sample_rate, &a));
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    sweep_kernel<1><<<gx(a.n, 256 * 4), 256, 0, ctx->main>>>(out_dev, a);
    count_launch(ctx);
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}
adsp_status adsp_logsweep_generate_host(double *out, double start_hz, double end_hz, double duration, double sample_rate) {
    SweepArgs a;
    ADSP_TRY(sweep_args(start_hz, end_hz, duration, sample_rate, &a));
    for (long long i = 0; i < a.n; i++) out[i] = logsweep_sample(i, a);
    return ADSP_OK;
}
adsp_status adsp_logsweep_inverse_filter_host(double *out, double start_hz, double end_hz, double duration, double sample_rate) {
    SweepArgs a;
    ADSP_TRY(sweep_args(start_hz, end_hz, duration, sample_rate, &a));
    for (long long i = 0; i < a.n; i++) out[i] = loginverse_sample(i, a);
    return ADSP_OK;
}

// LogSweep.Deconvolve (sweep.go:164-239): response (n samples, device) convolved with the inverse filter; out_dev holds
// n + samples - 1 values, the impulse response peaks near index samples - 1.  Responses: `rows` rows, strides in elements.
adsp_status adsp_logsweep_deconvolve_device(adsp_ctx *ctx, const double *response_dev, int64_t n, int64_t rows, int64_t in_stride, double start_hz,
                                            double end_hz, double duration, double sample_rate, double *out_dev, int64_t out_stride) {
    if (!ctx) return ADSP_ERR_INVALID_ARG;
    SweepArgs a;
    ADSP_TRY(sweep_args(start_hz, end_hz, duration, sample_rate, &a));
    if (n <= 0) { set_error("sweep: response signal is empty"); return ADSP_ERR_EMPTY_INPUT; }        // ErrEmptyResponse, :170-172
    if (!response_dev || !out_dev || rows <= 0) return ADSP_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    ADSP_TRY(ctx->d_k.reserve((size_t)a.n * sizeof(double)));
    sweep_kernel<1><<<gx(a.n, 256 * 4), 256, 0, ctx->main>>>((double *)ctx->d_k.p, a);
    count_launch(ctx);
    // full linear convolution (what the zero-padded FFT product of :183-236 computes); inverse filters of at most 64 taps
    // cannot occur for any audible sweep, but the direct kernel covers them
    if (a.n <= 64) return direct_device<double>(ctx, response_dev, n, in_stride, (const double *)ctx->d_k.p, a.n, 0, rows, out_dev, out_stride);
    return fft_convolve_device<double>(ctx, response_dev, n, rows, in_stride, (const double *)ctx->d_k.p, a.n, out_dev, out_stride);
}

adsp_status adsp_logsweep_deconvolve(adsp_ctx *ctx, const double *response, int64_t n, double start_hz, double end_hz, double duration, double sample_rate,
                                     double *out, int64_t out_len) {
    if (!ctx) return ADSP_ERR_INVALID_ARG;
    SweepArgs a;
    ADSP_TRY(sweep_args(start_hz, end_hz, duration, sample_rate, &a));
    if (n <= 0) { set_error("sweep: response signal is empty"); return ADSP_ERR_EMPTY_INPUT; }
    if (!response || !out) return ADSP_ERR_INVALID_ARG;
    if (out_len != n + a.n - 1) { set_error("sweep: output must hold len(response) + samples - 1 values"); return ADSP_ERR_LENGTH_MISMATCH; }
    void *din = nullptr, *dout = nullptr;
    adsp_status st = adsp_device_alloc(ctx, (size_t)n * 8, &din);
    if (st == ADSP_OK) st = adsp_device_alloc(ctx, (size_t)out_len * 8, &dout);
    if (st == ADSP_OK) st = adsp_memcpy_h2d(ctx, din, response, (size_t)n * 8);
    if (st == ADSP_OK) st = adsp_logsweep_deconvolve_device(ctx, (const double *)din, n, 1, n, start_hz, end_hz, duration, sample_rate, (double *)dout, out_len);
    if (st == ADSP_OK) st = adsp_memcpy_d2h(ctx, out, dout, (size_t)out_len * 8);
    adsp_device_free(ctx, din);
    adsp_device_free(ctx, dout);
    return st;
}

// ================================================================ dsp/filter/fir
// New(coeffs) filter.go:18-29.  ProcessBlock (:61-103) has two branches in the reference: below 32 taps it calls
// ProcessSample, y[n] = sum_k h[k] x[n-k]; from 32 taps on it takes vecmath.DotProduct(coeffs, window) over the last n
// samples stored OLDEST FIRST (:93-94), i.e. y[n] = sum_k h[N-1-k] x[n-k].  Both are kept as they are (they agree for
// the symmetric, linear-phase filters the reference's tests and designers produce): the tap order is fixed here once.
adsp_status adsp_fir_create(adsp_ctx *ctx, const double *coeffs, int64_t ntaps, int channels, adsp_fir **out) {
    if (!out) return ADSP_ERR_INVALID_ARG;
    *out = nullptr;
    if (!ctx || channels < 1 || ntaps < 0 || (ntaps > 0 && !coeffs)) return ADSP_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    adsp_fir *f = new adsp_fir();
    f->ctx = ctx; f->ntaps = ntaps; f->channels = channels;
    if (ntaps > 0) {
        std::vector<double> c(coeffs, coeffs + ntaps);
        if (ntaps >= 32) std::reverse(c.begin(), c.end());           // linearizeThreshold, filter.go:70
        f->h_taps = c;
        f->inplace = ntaps <= 16 * DIRECT_MC && channels <= 65535 && env_ll("ADSP_FIR_INPLACE", 1) != 0;
        adsp_status st = f->taps.reserve((size_t)ntaps * 8);
        if (st == ADSP_OK) st = upload(ctx, f->taps.p, c.data(), (size_t)ntaps * 8);
        if (st == ADSP_OK && cudaStreamSynchronize(ctx->main) != cudaSuccess) st = ADSP_ERR_CUDA;
        if (st == ADSP_OK && f->inplace && ntaps > 1) {
            const size_t hb = (size_t)(ntaps - 1) * channels * 8;
            for (int i = 0; i < 2 && st == ADSP_OK; i++) {
                st = f->hist[i].reserve(hb);
                if (st == ADSP_OK && cudaMemsetAsync(f->hist[i].p, 0, hb, ctx->main) != cudaSuccess) st = ADSP_ERR_CUDA;
            }
            if (st == ADSP_OK && cudaStreamSynchronize(ctx->main) != cudaSuccess) st = ADSP_ERR_CUDA;
        }
        if (st != ADSP_OK) { f->taps.release(); f->hist[0].release(); f->hist[1].release(); delete f; return st; }
    }
    *out = f;
    return ADSP_OK;
}

// in place on device rows (see fir_inplace_kernel): halos and the next history first, then one CTA per 8192-sample segment
extern "C++" {
template <int NCH>
static adsp_status fir_inplace_launch(adsp_fir *f, double *d, long long stride, long long n, int nseg, long long seg_len) {
    adsp_ctx *ctx = f->ctx;
    static const bool exact = env_ll("ADSP_DIRECT_EXACT", 0) != 0;
    DirectTapsN<double, NCH> taps;
    memcpy(taps.v, f->h_taps.data(), (size_t)f->ntaps * 8);
    for (long long i = f->ntaps; i < NCH * DIRECT_MC; i++) taps.v[i] = 0.0;
    const int H = (int)f->ntaps - 1;
    const unsigned grid = (unsigned)((long long)nseg * f->channels);
    if (exact) fir_inplace_kernel<double, false, NCH><<<grid, DIRECT_THREADS, 0, ctx->main>>>(d, stride, n, (const double *)f->halo.p, H, taps, (int)f->ntaps, seg_len, nseg);
    else fir_inplace_kernel<double, true, NCH><<<grid, DIRECT_THREADS, 0, ctx->main>>>(d, stride, n, (const double *)f->halo.p, H, taps, (int)f->ntaps, seg_len, nseg);
    count_launch(ctx);
    return ADSP_OK;
}
}  // extern "C++"

static adsp_status fir_run_inplace(adsp_fir *f, double *buf, long long n, long long stride, bool host) {
    adsp_ctx *ctx = f->ctx;
    const int H = (int)f->ntaps - 1;
    const long long seg_len = (long long)FIR_SEG_TILES * DIRECT_TILE;
    const long long nseg_ll = (n + seg_len - 1) / seg_len;
    if (nseg_ll * f->channels > 0x7fffffffLL || nseg_ll > 0x7ffffffeLL) { set_error("fir: block too large"); return ADSP_ERR_INVALID_ARG; }
    const int nseg = (int)nseg_ll;
    double *d = buf;
    long long ds = stride;
    if (host) {
        ds = ((n + 31) / 32) * 32;
        ADSP_TRY(f->work.reserve((size_t)ds * f->channels * 8));
        d = (double *)f->work.p;
        ADSP_TRY(upload2d(ctx, d, (size_t)ds * 8, buf, (size_t)stride * 8, (size_t)n * 8, (size_t)f->channels));
    }
    if (H > 0) {
        ADSP_TRY(f->halo.reserve((size_t)nseg * f->channels * H * 8));
        fir_halo_kernel<double><<<dim3((unsigned)nseg + 1, (unsigned)f->channels), 256, 0, ctx->main>>>(d, ds, n, (const double *)f->hist[f->cur].p, (double *)f->hist[f->cur ^ 1].p,
                                                                                                         (double *)f->halo.p, H, seg_len, nseg);
        count_launch(ctx);
        f->cur ^= 1;
    }
    if (f->ntaps <= 4 * DIRECT_MC) ADSP_TRY(fir_inplace_launch<4>(f, d, ds, n, nseg, seg_len));
    else ADSP_TRY(fir_inplace_launch<16>(f, d, ds, n, nseg, seg_len));
    ADSP_CUDA(cudaGetLastError());
    if (host) {
        ADSP_TRY(download2d(ctx, buf, (size_t)stride * 8, d, (size_t)ds * 8, (size_t)n * 8, (size_t)f->channels));
        ADSP_CUDA(cudaStreamSynchronize(ctx->main));
    }
    return ADSP_OK;
}

static adsp_status fir_reserve(adsp_fir *f, long long n) {
    const long long H = f->ntaps - 1;
    if (n <= f->work_cap) return ADSP_OK;
    adsp_ctx *ctx = f->ctx;
    const long long ws_old = H + f->work_cap, ws_new = ((H + n + 31) / 32) * 32;
    DevBuf nw;
    ADSP_TRY(nw.reserve((size_t)ws_new * f->channels * 8));
    ADSP_CUDA(cudaMemsetAsync(nw.p, 0, (size_t)ws_new * f->channels * 8, ctx->main));
    if (f->work.p && H > 0)   // carry the history over
        ADSP_CUDA(cudaMemcpy2DAsync(nw.p, (size_t)ws_new * 8, f->work.p, (size_t)(((ws_old + 31) / 32) * 32) * 8, (size_t)H * 8, (size_t)f->channels, cudaMemcpyDeviceToDevice, ctx->main));
    ADSP_CUDA(cudaStreamSynchronize(ctx->main));
    f->work.release();
    f->work = nw;
    f->work_cap = ws_new - H;
    ADSP_TRY(f->full.reserve((size_t)(ws_new + f->ntaps) * f->channels * 8));
    return ADSP_OK;
}

static adsp_status fir_run(adsp_fir *f, double *buf, long long n, long long stride, bool host) {
    if (!f) return ADSP_ERR_INVALID_ARG;
    if (n <= 0 || f->ntaps == 0) return ADSP_OK;                         // filter.go:62-65: no coefficients, block untouched
    if (!buf || (f->channels > 1 && stride < n)) return ADSP_ERR_INVALID_ARG;
    adsp_ctx *ctx = f->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    if (f->inplace) return fir_run_inplace(f, buf, n, stride, host);
    ADSP_TRY(fir_reserve(f, n));
    const long long H = f->ntaps - 1, ws = H + f->work_cap, L = H + n, fs = ws + f->ntaps;
    double *work = (double *)f->work.p, *full = (double *)f->full.p;
    // work rows = [history | block]
    if (host) ADSP_TRY(upload2d(ctx, work + H, (size_t)ws * 8, buf, (size_t)stride * 8, (size_t)n * 8, (size_t)f->channels));
    else ADSP_CUDA(cudaMemcpy2DAsync(work + H, (size_t)ws * 8, buf, (size_t)stride * 8, (size_t)n * 8, (size_t)f->channels, cudaMemcpyDeviceToDevice, ctx->main));
    ADSP_TRY(direct_device<double>(ctx, work, L, ws, (const double *)f->taps.p, f->ntaps, 0, f->channels, full, fs, f->h_taps.data()));
    // y[t] = full[H + t]; next history = last H samples of the work row
    if (host) ADSP_TRY(download2d(ctx, buf, (size_t)stride * 8, full + H, (size_t)fs * 8, (size_t)n * 8, (size_t)f->channels));
    else ADSP_CUDA(cudaMemcpy2DAsync(buf, (size_t)stride * 8, full + H, (size_t)fs * 8, (size_t)n * 8, (size_t)f->channels, cudaMemcpyDeviceToDevice, ctx->main));
    if (H > 0) {
        // the history slide goes through `full` when source and destination overlap (block shorter than the history)
        if (n >= H) ADSP_CUDA(cudaMemcpy2DAsync(work, (size_t)ws * 8, work + n, (size_t)ws * 8, (size_t)H * 8, (size_t)f->channels, cudaMemcpyDeviceToDevice, ctx->main));
        else {
            ADSP_CUDA(cudaMemcpy2DAsync(full, (size_t)fs * 8, work + n, (size_t)ws * 8, (size_t)H * 8, (size_t)f->channels, cudaMemcpyDeviceToDevice, ctx->main));
            ADSP_CUDA(cudaMemcpy2DAsync(work, (size_t)ws * 8, full, (size_t)fs * 8, (size_t)H * 8, (size_t)f->channels, cudaMemcpyDeviceToDevice, ctx->main));
        }
    }
    if (host) ADSP_CUDA(cudaStreamSynchronize(ctx->main));
    return ADSP_OK;
}

adsp_status adsp_fir_process_block(adsp_fir *f, double *buf, int64_t n, int64_t stride) { return fir_run(f, buf, n, stride, true); }
adsp_status adsp_fir_process_block_device(adsp_fir *f, double *buf_dev, int64_t n, int64_t stride) { return fir_run(f, buf_dev, n, stride, false); }
int64_t adsp_fir_order(const adsp_fir *f) { return f ? f->ntaps - 1 : 0; }                           // Order(), filter.go
void adsp_fir_reset(adsp_fir *f) {                                                                    // Reset()
    if (!f) return;
    std::lock_guard<std::mutex> lk(f->ctx->mu);
    cudaSetDevice(f->ctx->device);
    if (f->inplace) { if (f->hist[f->cur].p) cudaMemsetAsync(f->hist[f->cur].p, 0, (size_t)(f->ntaps - 1) * f->channels * 8, f->ctx->main); }
    else if (f->work.p) cudaMemsetAsync(f->work.p, 0, f->work.cap, f->ctx->main);
    cudaStreamSynchronize(f->ctx->main);
}
void adsp_fir_destroy(adsp_fir *f) {
    if (!f) return;
    std::lock_guard<std::mutex> lk(f->ctx->mu);
    cudaSetDevice(f->ctx->device);
    cudaStreamSynchronize(f->ctx->main);
    f->taps.release(); f->work.release(); f->full.release(); f->hist[0].release(); f->hist[1].release(); f->halo.release();
    delete f;
}

// ================================================================ dsp/resample
// approximateRatio, resample_design.go:74-113 (continued fractions)
void adsp_resample_approximate_ratio(double v, int max_den, int *num, int *den) {
    int rn = 1, rd = 1;
    if (max_den <= 0) max_den = 4096;
    if (v > 0 && !std::isnan(v) && !std::isinf(v)) {
        const double a0 = floor(v);
        double p0 = 1.0, q0 = 0.0, p1 = a0, q1 = 1.0, x = v;
        for (;;) {
            const double frac = x - floor(x);
            if (frac == 0) break;
            x = 1 / frac;
            const double a = floor(x);
            const double p2 = a * p1 + p0, q2 = a * q1 + q0;
            if (q2 > (double)max_den) break;
            p0 = p1; q0 = q1; p1 = p2; q1 = q2;
        }
        const long long n_ = llround(p1), d_ = llround(q1);
        if (d_ > 0) { const long long g = gcd_ll(n_, d_); rn = (int)(n_ / g); rd = (int)(d_ / g); }
    }
    if (num) *num = rn;
    if (den) *den = rd;
}

// NewRational(up, down, options) resample.go:153-191 + designPolyphaseFIR resample_design.go:9-72.  quality: 0 fast, 1
// balanced, 2 best (QualityProfile :35-44); taps_per_phase / cutoff_scale / kaiser_beta <= 0 take the profile's value.
adsp_status adsp_resampler_create(adsp_ctx *ctx, int up, int down, int quality, int taps_per_phase, double cutoff_scale, double kaiser_beta, int channels,
                                  adsp_resampler **out) {
    if (!out) return ADSP_ERR_INVALID_ARG;
    *out = nullptr;
    if (!ctx || channels < 1) return ADSP_ERR_INVALID_ARG;
    if (up <= 0 || down <= 0) { set_error("resample: invalid ratio"); return ADSP_ERR_INVALID_ARG; }    // ErrInvalidRatio
    const long long g = gcd_ll(up, down);
    up = (int)(up / g); down = (int)(down / g);
    int tpp = quality == 0 ? 16 : quality == 2 ? 64 : 32;
    double cs = quality == 0 ? 0.88 : quality == 2 ? 0.96 : 0.92, kb = quality == 0 ? 5.0 : quality == 2 ? 9.0 : 7.5;
    if (taps_per_phase > 0) tpp = taps_per_phase;
    if (cutoff_scale > 0 && cutoff_scale <= 1) cs = cutoff_scale;
    if (kaiser_beta > 0) kb = kaiser_beta;
    const int ntaps = tpp * up;
    const double fc = (0.5 / (double)std::max(up, down)) * cs;
    if (fc <= 0 || fc >= 0.5) { set_error("resample: invalid cutoff"); return ADSP_ERR_INVALID_ARG; }
    std::vector<double> taps((size_t)ntaps);
    const double center = 0.5 * (double)(ntaps - 1);
    for (int n = 0; n < ntaps; n++) { const double t = (double)n - center; taps[(size_t)n] = 2 * fc * sincf(2 * fc * t) * kaiser(n, ntaps, kb); }
    double sum = 0;
    for (double v : taps) sum += v;
    if (sum == 0) { set_error("resample: designed zero-sum filter"); return ADSP_ERR_INVALID_ARG; }
    const double scale = (double)up / sum;
    for (double &v : taps) v *= scale;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    adsp_resampler *r = new adsp_resampler();
    r->ctx = ctx; r->up = up; r->down = down; r->quality = quality; r->channels = channels; r->taps = taps;
    r->max_phase_len = (ntaps + up - 1) / up;                           // phase 0 is the longest (:54-67)
    adsp_status st = r->d_taps.reserve((size_t)ntaps * 8);
    if (st == ADSP_OK) st = upload(ctx, r->d_taps.p, r->taps.data(), (size_t)ntaps * 8);
    if (st == ADSP_OK && cudaStreamSynchronize(ctx->main) != cudaSuccess) st = ADSP_ERR_CUDA;
    if (st != ADSP_OK) { r->d_taps.release(); delete r; return st; }
    *out = r;
    return ADSP_OK;
}

// NewForRates(inRate, outRate) resample.go:194-213
adsp_status adsp_resampler_create_for_rates(adsp_ctx *ctx, double in_rate, double out_rate, int quality, int max_den, int channels, adsp_resampler **out) {
    if (!(in_rate > 0) || !(out_rate > 0) || std::isnan(in_rate) || std::isnan(out_rate)) { set_error("resample: invalid sample rate"); return ADSP_ERR_INVALID_ARG; }
    int up, down;
    adsp_resample_approximate_ratio(out_rate / in_rate, max_den, &up, &down);
    return adsp_resampler_create(ctx, up, down, quality, 0, 0, 0, channels, out);
}

void adsp_resampler_ratio(const adsp_resampler *r, int *up, int *down) { if (r) { if (up) *up = r->up; if (down) *down = r->down; } }
int adsp_resampler_taps_per_phase(const adsp_resampler *r) { return r ? r->max_phase_len : 0; }
int64_t adsp_resampler_prototype(const adsp_resampler *r, double *out, int64_t cap) {
    if (!r) return 0;
    if (out) for (int64_t i = 0; i < cap && i < (int64_t)r->taps.size(); i++) out[i] = r->taps[(size_t)i];
    return (int64_t)r->taps.size();
}
int64_t adsp_resampler_predict_output_len(const adsp_resampler *r, int64_t input_len) { return r ? predict_len(r, input_len) : 0; }

static adsp_status resample_run(adsp_resampler *r, const double *in, long long n, long long in_stride, double *out, long long out_cap, long long out_stride,
                                int64_t *n_out, bool host) {
    if (!r) return ADSP_ERR_INVALID_ARG;
    if (n_out) *n_out = 0;
    if (n <= 0) return ADSP_OK;                                          // resample.go:250-252
    if (!in || !out) return ADSP_ERR_INVALID_ARG;
    adsp_ctx *ctx = r->ctx;
    const long long nout = predict_len(r, n);
    if (out_cap < nout || (r->channels > 1 && (in_stride < n || out_stride < nout))) { set_error("resample: output buffer too small"); return ADSP_ERR_LENGTH_MISMATCH; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    const long long H = r->hist_len, keep_max = std::max(0, r->max_phase_len - 1);
    // branch-tiled kernel when the prototype and a tile's window fit shared memory, else the gather kernel
    const int ntaps = (int)r->taps.size(), lmax = std::max(1, r->max_phase_len);
    const int cmul = (256 + r->up - 1) / r->up, P = r->up * cmul;              // >= 256 outputs between a thread's outputs
    const int iters = (P + 511) / 512, bd = (((P + iters - 1) / iters) + 31) / 32 * 32;
    const long long tile = (long long)P * RS_R;
    const size_t need = ((size_t)ntaps + (size_t)(tile * r->down / r->up) + (size_t)lmax + 4) * 8;
    const bool use_phase = env_ll("ADSP_RESAMPLE_TILED", 1) != 0 && need <= (size_t)100 * 1024 && (long long)cmul * r->down * RS_R < (1LL << 30) &&
                           (nout + tile - 1) / tile <= 0x7fffffffLL;
    // device rows at least as long as the history are read where they are; otherwise the block is put behind the history
    const bool direct_src = !host && use_phase && n >= keep_max;
    const long long ws = ((keep_max + (direct_src ? 0 : n) + 31) / 32) * 32;
    if (ws > r->work_stride) {                                          // grow, carrying the history over
        DevBuf nw;
        ADSP_TRY(nw.reserve((size_t)ws * r->channels * 8));
        if (r->work.p && H > 0)
            ADSP_CUDA(cudaMemcpy2DAsync(nw.p, (size_t)ws * 8, r->work.p, (size_t)r->work_stride * 8, (size_t)H * 8, (size_t)r->channels, cudaMemcpyDeviceToDevice, ctx->main));
        ADSP_CUDA(cudaStreamSynchronize(ctx->main));
        r->work.release();
        r->work = nw;
        r->work_stride = ws;
    }
    double *work = (double *)r->work.p;
    const long long wst = r->work_stride;
    if (host) ADSP_TRY(upload2d(ctx, work + H, (size_t)wst * 8, in, (size_t)in_stride * 8, (size_t)n * 8, (size_t)r->channels));
    else if (!direct_src) ADSP_CUDA(cudaMemcpy2DAsync(work + H, (size_t)wst * 8, in, (size_t)in_stride * 8, (size_t)n * 8, (size_t)r->channels, cudaMemcpyDeviceToDevice, ctx->main));
    const double *src_blk = direct_src ? in : work + H;
    const long long src_blk_stride = direct_src ? in_stride : wst;
    const long long base = r->total_in - H, last = r->total_in + n - 1;
    double *dout = out;
    long long dstride = out_stride;
    if (host) {
        dstride = ((nout + 31) / 32) * 32;
        ADSP_TRY(r->outbuf.reserve((size_t)std::max<long long>(dstride, 32) * r->channels * 8));
        dout = (double *)r->outbuf.p;
    }
    if (nout > 0) {
        if (use_phase) {
            ADSP_CUDA(cudaFuncSetAttribute(resample_phase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));   // per device
            dim3 grid((unsigned)((nout + tile - 1) / tile), (unsigned)r->channels);
            resample_phase_kernel<<<grid, bd, need, ctx->main>>>(work, wst, H, src_blk, src_blk_stride, base, last, (const double *)r->d_taps.p, ntaps, r->up, r->down,
                                                                 r->out_count, nout, dout, dstride, P, lmax);
        } else {
            dim3 grid(gx(nout, 256), (unsigned)r->channels);
            resample_kernel<<<grid, 256, 0, ctx->main>>>(work, wst, base, last, (const double *)r->d_taps.p, ntaps, r->up, r->down, r->out_count, nout, dout, dstride);
        }
        count_launch(ctx);
        if (host) ADSP_TRY(download2d(ctx, out, (size_t)out_stride * 8, dout, (size_t)dstride * 8, (size_t)nout * 8, (size_t)r->channels));
    }
    // history = the last min(maxPhaseLn - 1, len(work)) samples (:288-290)
    const long long have = H + n, keep = std::min(keep_max, have);
    if (direct_src) {
        if (keep > 0) ADSP_CUDA(cudaMemcpy2DAsync(work, (size_t)wst * 8, in + (n - keep), (size_t)in_stride * 8, (size_t)keep * 8, (size_t)r->channels, cudaMemcpyDeviceToDevice, ctx->main));
    } else if (keep > 0 && have - keep > 0) {
        if (have - keep >= keep) ADSP_CUDA(cudaMemcpy2DAsync(work, (size_t)wst * 8, work + (have - keep), (size_t)wst * 8, (size_t)keep * 8, (size_t)r->channels, cudaMemcpyDeviceToDevice, ctx->main));
        else {   // overlapping slide: through a scratch row set
            ADSP_TRY(ctx->d_tmp.reserve((size_t)keep * r->channels * 8));
            ADSP_CUDA(cudaMemcpy2DAsync(ctx->d_tmp.p, (size_t)keep * 8, work + (have - keep), (size_t)wst * 8, (size_t)keep * 8, (size_t)r->channels, cudaMemcpyDeviceToDevice, ctx->main));
            ADSP_CUDA(cudaMemcpy2DAsync(work, (size_t)wst * 8, ctx->d_tmp.p, (size_t)keep * 8, (size_t)keep * 8, (size_t)r->channels, cudaMemcpyDeviceToDevice, ctx->main));
        }
    }
    r->hist_len = keep;
    r->total_in += n;
    r->out_count += nout;
    if (n_out) *n_out = nout;
    if (host) ADSP_CUDA(cudaStreamSynchronize(ctx->main));
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}

adsp_status adsp_resampler_process(adsp_resampler *r, const double *in, int64_t n, int64_t in_stride, double *out, int64_t out_cap, int64_t out_stride,
                                   int64_t *n_out) {
    return resample_run(r, in, n, in_stride, out, out_cap, out_stride, n_out, true);
}
adsp_status adsp_resampler_process_device(adsp_resampler *r, const double *in_dev, int64_t n, int64_t in_stride, double *out_dev, int64_t out_cap,
                                          int64_t out_stride, int64_t *n_out) {
    return resample_run(r, in_dev, n, in_stride, out_dev, out_cap, out_stride, n_out, false);
}
void adsp_resampler_reset(adsp_resampler *r) { if (r) { r->total_in = 0; r->out_count = 0; r->hist_len = 0; } }      // Reset(), resample.go:241-246
void adsp_resampler_destroy(adsp_resampler *r) {
    if (!r) return;
    std::lock_guard<std::mutex> lk(r->ctx->mu);
    cudaSetDevice(r->ctx->device);
    cudaStreamSynchronize(r->ctx->main);
    r->d_taps.release(); r->work.release(); r->outbuf.release();
    delete r;
}

}  // extern "C"
