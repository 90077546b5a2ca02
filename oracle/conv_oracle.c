/*
 * conv_oracle.c -- CPU oracle for the dsp/conv hot path.   *** TEST INFRASTRUCTURE ***
 *
 * A plain-C restatement of the reference algorithms in /root/reference/dsp/conv
 * (CWBudde/algo-dsp, Go).  It exists so that the CUDA product path can be checked
 * against the reference's results on identical inputs; the Go reference itself cannot
 * run here (no Go toolchain; algo-fft v0.6.10 / algo-vecmath v0.1.0 are un-vendored,
 * go.mod:5-9).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call this code.
 *
 * PARITY PIN: the oracle is pinned against every known-answer vector the reference's
 * own tests hold for this path (tests/golden/reference_kats.json, checked by
 * tests/test_oracle_kats.py): conv_test.go:9-69,343-362, example_test.go:10-127,
 * streaming_overlap_{save,add}_test.go impulse responses, partitioned_test.go stage
 * layouts and the FFT-vs-Direct tolerance ladder (1e-10 / 1e-8).
 *
 * Built with -ffp-contract=off so `a*b + c` is two roundings, as in Go on amd64
 * (conv.go:120,149-152).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum {
    ORC_OK = 0,
    ORC_ERR_EMPTY_INPUT = 1,          /* conv.go:42 ErrEmptyInput */
    ORC_ERR_EMPTY_KERNEL = 2,         /* conv.go:43 ErrEmptyKernel */
    ORC_ERR_LENGTH_MISMATCH = 3,      /* conv.go:44 ErrLengthMismatch */
    ORC_ERR_INVALID_BLOCK_SIZE = 4,   /* conv.go:45 ErrInvalidBlockSize */
    ORC_ERR_INVALID_BLOCK_ORDER = 5,  /* partitioned.go:12 */
    ORC_ERR_EMPTY_IR = 6,             /* partitioned.go:13 */
    ORC_ERR_STAGE_INDEX = 7,          /* partitioned.go:14 */
    ORC_ERR_INVALID_ARG = 8,
    ORC_ERR_DIVISION_BY_ZERO = 9      /* deconvolve.go:14 ErrDivisionByZero */
};

/* nextPowerOf2, dsp/conv/conv.go:250-261. */
int64_t orc_next_pow2(int64_t n) {
    if (n <= 1) return 1;
    int64_t p = 1;
    while (p < n) p *= 2;
    return p;
}

/* isPowerOf2, dsp/conv/conv.go:264-266. */
int orc_is_pow2(int64_t n) { return n > 0 && (n & (n - 1)) == 0; }

/* truncLog2 / bitCountToBits, dsp/conv/partitioned.go:186-203. */
static int orc_trunc_log2(int n) {
    if (n <= 0) return 0;
    int r = 0;
    while (n > 1) { n >>= 1; r++; }
    return r;
}
static int orc_bits(int n) { return (2 << n) - 1; }

/* NewOverlapSave sizing, dsp/conv/overlap_save.go:53-76. */
int orc_ols_sizes(int64_t K, int64_t fftSize, int64_t *fftOut, int64_t *stepOut) {
    if (K <= 0) return ORC_ERR_EMPTY_KERNEL;
    if (fftSize <= 0) {
        fftSize = orc_next_pow2(2 * K);
        if (fftSize < 256) fftSize = 256;
    }
    if (!orc_is_pow2(fftSize)) return ORC_ERR_INVALID_BLOCK_SIZE;
    if (fftSize < 2 * K) fftSize = orc_next_pow2(2 * K);
    *fftOut = fftSize;
    *stepOut = fftSize - K + 1;
    return ORC_OK;
}

/* NewOverlapAdd sizing, dsp/conv/overlap_add.go:44-59. */
int orc_ola_sizes(int64_t K, int64_t blockSize, int64_t *blockOut, int64_t *fftOut) {
    if (K <= 0) return ORC_ERR_EMPTY_KERNEL;
    if (blockSize <= 0) {
        blockSize = orc_next_pow2(K);
        if (blockSize < 256) blockSize = 256;
    }
    *blockOut = blockSize;
    *fftOut = orc_next_pow2(blockSize + K - 1);
    return ORC_OK;
}

/* trimToMode, dsp/conv/conv.go:229-247.  mode: 0 full, 1 same, 2 valid. */
void orc_trim_mode(int64_t lenA, int64_t lenB, int mode, int64_t *start, int64_t *len) {
    const int64_t full = lenA + lenB - 1;
    switch (mode) {
    case 1: *start = (lenB - 1) / 2; *len = lenA; return;
    case 2:
        if (lenA >= lenB) { *start = lenB - 1; *len = lenA - (lenB - 1); }
        else { *start = lenA - 1; *len = lenB - (lenA - 1); }
        return;
    default: *start = 0; *len = full; return;
    }
}

/* LagFromIndex / IndexFromLag, dsp/conv/correlate.go:221-229. */
int64_t orc_lag_from_index(int64_t index, int64_t lenB) { return index - (lenB - 1); }
int64_t orc_index_from_lag(int64_t lag, int64_t lenB) { return lag + (lenB - 1); }

#define REAL double
#define SFX _f64
#include "conv_oracle_impl.h"
#undef REAL
#undef SFX

#define REAL float
#define SFX _f32
#include "conv_oracle_impl.h"
#undef REAL
#undef SFX

/* ------------------------------------------------------------------------------------
 * Timed CPU baseline helpers (bench.py cpu_baseline / --impl reference only).
 * Same algorithmic shape as OverlapSave.Process (overlap_save.go:126-254): a convolver
 * is built once (NewOverlapSave: plan + kernel FFT), then Process runs per channel.
 * dsp/conv itself never spawns goroutines; `threads` > 1 models a Go user running one
 * convolver instance per goroutine over independent channels.
 * ---------------------------------------------------------------------------------- */
#include <pthread.h>
#include <unistd.h>

typedef struct {
    const fftplan_f64 *plan; const cpx_f64 *kfft;
    const double *kernel; int64_t K, fftSize, step;
    const double *in; int64_t n, channels, in_stride;
    double *out; int64_t out_stride;
    int64_t *next; pthread_mutex_t *mu;
} orc_bench_job;

static void *orc_bench_worker(void *arg) {
    orc_bench_job *j = (orc_bench_job *)arg;
    const int64_t K = j->K, n = j->n, step = j->step, fftSize = j->fftSize;
    cpx_f64 *buf = (cpx_f64 *)malloc(sizeof(cpx_f64) * (size_t)fftSize);
    double *history = (double *)calloc((size_t)(K > 1 ? K - 1 : 1), sizeof(double));
    for (;;) {
        pthread_mutex_lock(j->mu);
        int64_t c = (*j->next)++;
        pthread_mutex_unlock(j->mu);
        if (c >= j->channels) break;
        const double *x = j->in + c * j->in_stride;
        double *y = j->out + c * j->out_stride;
        const int64_t outLen = n + K - 1;
        for (int64_t i = 0; i < K - 1; i++) history[i] = 0;
        int64_t inputPos = 0, outputPos = 0;
        while (inputPos < n) {
            int64_t newSamples = step;
            if (inputPos + newSamples > n) newSamples = n - inputPos;
            ols_block_f64(j->plan, j->kfft, buf, history, K, x + inputPos, newSamples);
            for (int64_t i = 0; i < newSamples && outputPos + i < outLen; i++)
                y[outputPos + i] = buf[K - 1 + i].re;
            const int64_t hs = inputPos + newSamples - (K - 1);
            for (int64_t i = 0; i < K - 1; i++) {
                int64_t idx = hs + i;
                history[i] = (idx >= 0 && idx < n) ? x[idx] : 0.0;
            }
            inputPos += newSamples;
            outputPos += newSamples;
        }
        if (outputPos < outLen) {
            ols_block_f64(j->plan, j->kfft, buf, history, K, NULL, 0);
            for (int64_t i = 0; outputPos + i < outLen && K - 1 + i < fftSize; i++)
                y[outputPos + i] = buf[K - 1 + i].re;
        }
    }
    free(buf);
    free(history);
    return NULL;
}

int orc_bench_ols_f64(const double *kernel, int64_t K, int64_t fftSize,
                      const double *in, int64_t n, int64_t channels, int64_t in_stride,
                      double *out, int64_t out_stride, int threads) {
    int64_t step;
    int st = orc_ols_sizes(K, fftSize, &fftSize, &step);
    if (st != ORC_OK) return st;
    if (n <= 0) return ORC_ERR_EMPTY_INPUT;
    fftplan_f64 *plan = fftplan_new_f64((int)fftSize);
    cpx_f64 *kfft = (cpx_f64 *)calloc((size_t)fftSize, sizeof(cpx_f64));
    for (int64_t i = 0; i < K; i++) kfft[i].re = kernel[i];
    fft_exec_f64(plan, kfft, 0);
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    int64_t next = 0;
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    orc_bench_job job = { plan, kfft, kernel, K, fftSize, step, in, n, channels, in_stride,
                          out, out_stride, &next, &mu };
    pthread_t th[256];
    for (int t = 1; t < threads; t++) pthread_create(&th[t], NULL, orc_bench_worker, &job);
    orc_bench_worker(&job);
    for (int t = 1; t < threads; t++) pthread_join(th[t], NULL);
    fftplan_free_f64(plan);
    free(kfft);
    return ORC_OK;
}

/* ------------------------------- deconvolution (dsp/conv/deconvolve.go), float64 only like the reference */

/* variance, deconvolve.go:332-353 (population variance) */
double orc_variance(const double *x, int64_t n) {
    if (n <= 0) return 0;
    double mean = 0;
    for (int64_t i = 0; i < n; i++) mean += x[i];
    mean /= (double)n;
    double sum = 0;
    for (int64_t i = 0; i < n; i++) { double d = x[i] - mean; sum += d * d; }
    return sum / (double)n;
}

int64_t orc_deconv_out_len(int64_t n, int64_t m) { int64_t o = n - m + 1; return o <= 0 ? n : o; }   /* :101-105 */

/* Deconvolve, deconvolve.go:72-330.  method 0 naive (:98-168), 1 regularized (:172-240), 2 Wiener (:244-328).
 * dst holds orc_deconv_out_len(n, m) samples.  *bad_bin = first bin with |H| < 1e-15 (naive only). */
int orc_deconvolve(const double *signal, int64_t n, const double *kernel, int64_t m, int method, double epsilon,
                   double noise_var, double signal_var, double *dst, int64_t *bad_bin) {
    if (n <= 0) return ORC_ERR_EMPTY_INPUT;
    if (m <= 0) return ORC_ERR_EMPTY_KERNEL;
    double reg = 0;
    if (method == 1) reg = epsilon <= 0 ? 1e-6 : epsilon;                                   /* :84-88 */
    else if (method == 2) {                                                                  /* :254-270 */
        if (signal_var <= 0) signal_var = orc_variance(signal, n);
        if (noise_var <= 0) noise_var = signal_var * 0.01;
        reg = noise_var / signal_var;
        if (!(reg > 0)) reg = 1e-6;
    } else if (method != 0) { method = 1; reg = 1e-6; }                                      /* :91-93 */
    const int64_t out_len = orc_deconv_out_len(n, m);
    const int64_t N = orc_next_pow2(n);                                                      /* :107 */
    fftplan_f64 *plan = fftplan_new_f64((int)N);
    cpx_f64 *fs = (cpx_f64 *)calloc((size_t)N, sizeof(cpx_f64));
    cpx_f64 *fk = (cpx_f64 *)calloc((size_t)N, sizeof(cpx_f64));
    for (int64_t i = 0; i < n; i++) fs[i].re = signal[i];
    for (int64_t i = 0; i < m && i < N; i++) fk[i].re = kernel[i];   /* the Go loop would index past fftSize when m > N and panic */
    fft_exec_f64(plan, fs, 0);
    fft_exec_f64(plan, fk, 0);
    int st = ORC_OK;
    for (int64_t i = 0; i < N; i++) {
        const double hr = fk[i].re, hi = fk[i].im;
        if (method == 0) {                                                                   /* :143-151 */
            if (hypot(hr, hi) < 1e-15) { st = ORC_ERR_DIVISION_BY_ZERO; if (bad_bin) *bad_bin = i; break; }
            const double den = hr * hr + hi * hi;
            const double re = (fs[i].re * hr + fs[i].im * hi) / den, im = (fs[i].im * hr - fs[i].re * hi) / den;
            fs[i].re = re; fs[i].im = im;
        } else {                                                                             /* :216-220, :304-308 */
            const double den = hr * hr + hi * hi + reg;
            const double re = (fs[i].re * hr + fs[i].im * hi) / den, im = (fs[i].im * hr - fs[i].re * hi) / den;
            fs[i].re = re; fs[i].im = im;
        }
    }
    if (st == ORC_OK) {
        fft_exec_f64(plan, fs, 1);
        for (int64_t i = 0; i < out_len; i++) dst[i] = fs[i].re;
    }
    fftplan_free_f64(plan);
    free(fs); free(fk);
    return st;
}

/* InverseFilter, deconvolve.go:359-412 */
int orc_inverse_filter(const double *kernel, int64_t m, int64_t length, double epsilon, double *dst) {
    if (m <= 0) return ORC_ERR_EMPTY_KERNEL;
    if (length <= 0) return ORC_OK;
    if (epsilon <= 0) epsilon = 1e-6;
    const int64_t N = orc_next_pow2(length);
    fftplan_f64 *plan = fftplan_new_f64((int)N);
    cpx_f64 *fk = (cpx_f64 *)calloc((size_t)N, sizeof(cpx_f64));
    for (int64_t i = 0; i < m && i < N; i++) fk[i].re = kernel[i];
    fft_exec_f64(plan, fk, 0);
    for (int64_t i = 0; i < N; i++) {
        const double hr = fk[i].re, hi = fk[i].im, den = hr * hr + hi * hi + epsilon;
        fk[i].re = hr / den; fk[i].im = -hi / den;
    }
    fft_exec_f64(plan, fk, 1);
    for (int64_t i = 0; i < length; i++) dst[i] = fk[i].re;
    fftplan_free_f64(plan);
    free(fk);
    return ORC_OK;
}

/* SNR, deconvolve.go:417-434 */
double orc_snr(const double *original, int64_t n, const double *recovered, int64_t n2) {
    if (n != n2 || n == 0) return -INFINITY;
    double sp = 0, np_ = 0;
    for (int64_t i = 0; i < n; i++) { sp += original[i] * original[i]; const double d = original[i] - recovered[i]; np_ += d * d; }
    if (np_ == 0) return INFINITY;
    return 10 * log10(sp / np_);
}

int orc_num_procs(void) { long n = sysconf(_SC_NPROCESSORS_ONLN); return n > 0 ? (int)n : 1; }
