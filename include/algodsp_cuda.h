/*
 * algodsp_cuda.h -- C ABI of libalgodsp_cuda, the B200 (sm_100a) implementation of the
 * dsp/conv hot path of CWBudde/algo-dsp.
 *
 * The reference has no FFI boundary of its own: the boundary it replaces is the exported Go
 * API of package conv (dsp/conv, all .go files).  Each entry point below cites the Go symbol
 * (file:line under /root/reference) whose behaviour it reproduces; a cgo shim with the
 * unchanged Go signatures sits directly on top.
 *
 * Conventions
 *   - plain C, no CUDA or torch types in any signature; device pointers travel as void*.
 *   - every call that takes HOST pointers is synchronous: the memory is only touched during
 *     the call (cgo pointer rule); results are complete when the call returns.
 *   - *_device calls take device pointers, enqueue on the context's streams and return after
 *     enqueueing; call adsp_ctx_sync (or adsp_plan_sync) before reading results.
 *   - every call selects the CUDA device of its context itself (goroutines migrate threads).
 *   - no CPU fallback: if no CUDA device is present adsp_ctx_create fails with ADSP_ERR_CUDA.
 *   - errors: adsp_status; the last error text per thread via adsp_last_error().
 *   - f64 entry points take double*, the _f32 twins take float* (optional fp32 mode).
 */
#ifndef ALGODSP_CUDA_H
#define ALGODSP_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADSP_API __attribute__((visibility("default")))

/* Status codes <-> Go sentinel errors. */
typedef enum adsp_status {
    ADSP_OK = 0,
    ADSP_ERR_EMPTY_INPUT = 1,         /* conv.ErrEmptyInput            dsp/conv/conv.go:42 */
    ADSP_ERR_EMPTY_KERNEL = 2,        /* conv.ErrEmptyKernel           dsp/conv/conv.go:43 */
    ADSP_ERR_LENGTH_MISMATCH = 3,     /* conv.ErrLengthMismatch        dsp/conv/conv.go:44 */
    ADSP_ERR_INVALID_BLOCK_SIZE = 4,  /* conv.ErrInvalidBlockSize      dsp/conv/conv.go:45 */
    ADSP_ERR_INVALID_BLOCK_ORDER = 5, /* conv.ErrInvalidBlockOrder     dsp/conv/partitioned.go:12 */
    ADSP_ERR_EMPTY_IR = 6,            /* conv.ErrEmptyImpulseResponse  dsp/conv/partitioned.go:13 */
    ADSP_ERR_STAGE_INDEX = 7,         /* conv.ErrStageIndexOutOfRange  dsp/conv/partitioned.go:14 */
    ADSP_ERR_INVALID_ARG = 8,         /* nil handle / negative size / bad enum */
    ADSP_ERR_CUDA = 9,                /* CUDA runtime error or no device (text in adsp_last_error) */
    ADSP_ERR_OOM = 10,                /* device or pinned-host allocation failed */
    ADSP_ERR_DIVISION_BY_ZERO = 11    /* conv.ErrDivisionByZero        dsp/conv/deconvolve.go:14 */
} adsp_status;

/* conv.Mode, dsp/conv/conv.go:57-69 */
typedef enum adsp_mode { ADSP_MODE_FULL = 0, ADSP_MODE_SAME = 1, ADSP_MODE_VALID = 2 } adsp_mode;

typedef enum adsp_precision { ADSP_F64 = 0, ADSP_F32 = 1 } adsp_precision;

typedef struct adsp_ctx adsp_ctx;    /* one GPU: streams, twiddle tables, scratch, staging */
typedef struct adsp_plan adsp_plan;  /* a convolver: device-resident IR spectrum + geometry */

/* ---------------------------------------------------------------- library / context */
ADSP_API const char *adsp_version(void);
ADSP_API const char *adsp_status_string(adsp_status st);
/* Copies the calling thread's last error text into buf (NUL terminated); returns its length. */
ADSP_API size_t adsp_last_error(char *buf, size_t buflen);
ADSP_API int adsp_device_count(void);

ADSP_API adsp_status adsp_ctx_create(int device, adsp_ctx **out);
ADSP_API void adsp_ctx_destroy(adsp_ctx *ctx);
ADSP_API adsp_status adsp_ctx_sync(adsp_ctx *ctx);
/* Number of kernels this context has launched since creation (bench "gpu_launches"). */
ADSP_API uint64_t adsp_ctx_launch_count(adsp_ctx *ctx);
/* Per-kernel device timing for the bench roofline: when enabled, every launch is bracketed by a
 * CUDA event pair on its own stream.  kind: 0 cols_fwd, 1 rows, 2 cols_inv, 3 single-kernel FFT
 * conv, 4 direct, 5 other, 6 persistent fused four-step kernel.  adsp_ctx_kernel_time synchronises, returns the accumulated device
 * time and launch count of `kind`, and clears the accumulators when reset != 0. */
ADSP_API void adsp_ctx_kernel_timing(adsp_ctx *ctx, int enable);
ADSP_API adsp_status adsp_ctx_kernel_time(adsp_ctx *ctx, int kind, double *total_ms, uint64_t *launches, int reset);
/* Run-time peaks the bench reports its rooflines against (csrc/diag.cu; not part of the data path): FP64 FMA
 * thread-instructions per second and shared-memory bytes per second (128-bit stores + loads) over the whole GPU at the
 * current clock -- the two per-SM resources that bound the fp64 FFT kernels before HBM does. */
ADSP_API adsp_status adsp_ctx_measure_pipes(adsp_ctx *ctx, double *dfma_per_s, double *smem_bytes_per_s);
/* Raw host-link ceiling: `reps` rounds of one H2D copy (h2d_bytes) and one D2H copy (d2h_bytes) from/to pinned memory,
 * concurrently on two streams (ms_per_round), and each direction alone.  Bounds every host-buffer (end-to-end) figure. */
ADSP_API adsp_status adsp_ctx_copy_ceiling(adsp_ctx *ctx, size_t h2d_bytes, size_t d2h_bytes, int reps, double *ms_per_round,
                                           double *h2d_alone_ms, double *d2h_alone_ms);
/* Raw cudaStream_t of the context's main stream (for event timing by a harness). */
ADSP_API void *adsp_ctx_stream(adsp_ctx *ctx);

/* Host-buffer calls and pageable memory.  Every entry point that takes HOST pointers accepts ordinary pageable memory
 * (a Go []float64, malloc, numpy): the library stages it through its own pinned buffers with a pool of copy threads
 * (ADSP_STAGE_THREADS, default min(8, cores/2)), overlapped with H2D | kernels | D2H (csrc/staging.cu).  Memory from
 * adsp_host_alloc_pinned (or cudaHostRegister'ed by the caller) is DMA'd in place, which saves the two host copies.
 * adsp_ctx_host_profile(ctx, 1) makes the small-call path (one upload, kernels, one download) record a breakdown of the
 * last call: ms[0..5] = {device allocation, upload (stage-in + H2D), kernels, download (D2H + stage-out), 0, total},
 * ms[6..9] = {host copy into the pinned slots, wait for the D2H, host copy out of the pinned slots, pointer-type queries};
 * the phases are serialised by a stream sync each while profiling is on (up to `cap` <= 12 values are written).
 * staged_*_bytes count what went through staging since the context was created. */
ADSP_API void adsp_ctx_host_profile(adsp_ctx *ctx, int enable);
ADSP_API adsp_status adsp_ctx_host_profile_get(adsp_ctx *ctx, double *ms, int cap, uint64_t *staged_in_bytes, uint64_t *staged_out_bytes);
ADSP_API int adsp_host_ptr_is_pinned(const void *p);     /* 1: DMA-able in place, 0: pageable (will be staged) */
ADSP_API int adsp_ctx_stage_threads(adsp_ctx *ctx);      /* size of the context's copy-thread pool (creates it) */

/* Pinned host memory for callers that want zero-copy DMA (Go side: C.malloc replacement). */
ADSP_API adsp_status adsp_host_alloc_pinned(size_t bytes, void **out);
ADSP_API void adsp_host_free_pinned(void *p);
/* Pin caller-owned memory in place (cudaHostRegister): host-pointer calls then DMA from / to it directly instead of staging.
 * For buffers the caller keeps alive and in place (Go: a slice held by runtime.Pinner for the lifetime of the
 * registration); must be unregistered before the memory is freed. */
ADSP_API adsp_status adsp_host_register(void *p, size_t bytes);
ADSP_API adsp_status adsp_host_unregister(void *p);
ADSP_API adsp_status adsp_device_alloc(adsp_ctx *ctx, size_t bytes, void **out);
ADSP_API void adsp_device_free(adsp_ctx *ctx, void *p);
ADSP_API adsp_status adsp_memcpy_h2d(adsp_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
ADSP_API adsp_status adsp_memcpy_d2h(adsp_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);

/* ---------------------------------------------------------------- sizing helpers (pure) */
ADSP_API int64_t adsp_next_pow2(int64_t n);                       /* nextPowerOf2  conv.go:250 */
ADSP_API int adsp_is_pow2(int64_t n);                             /* isPowerOf2    conv.go:264 */
/* NewOverlapSave sizing rules, overlap_save.go:53-76 */
ADSP_API adsp_status adsp_ols_sizes(int64_t kernel_len, int64_t fft_size, int64_t *fft_out, int64_t *step_out);
/* NewOverlapAdd sizing rules, overlap_add.go:44-59 */
ADSP_API adsp_status adsp_ola_sizes(int64_t kernel_len, int64_t block_size, int64_t *block_out, int64_t *fft_out);
/* trimToMode, conv.go:229-247: slice [start, start+len) of the full result */
ADSP_API void adsp_trim_mode(int64_t len_a, int64_t len_b, adsp_mode mode, int64_t *start, int64_t *len);
ADSP_API int64_t adsp_lag_from_index(int64_t index, int64_t len_b);  /* correlate.go:221 */
ADSP_API int64_t adsp_index_from_lag(int64_t lag, int64_t len_b);    /* correlate.go:227 */

/* ---------------------------------------------------------------- one-shot functions (host pointers)
 * out must hold n+m-1 elements.  These mirror the goroutine-safe package-level functions
 * (pooled instances, overlap_save.go:323-333): they are thread-safe per context. */
ADSP_API adsp_status adsp_direct(adsp_ctx *, const double *a, int64_t n, const double *b, int64_t m, double *out);          /* Direct         conv.go:76 */
ADSP_API adsp_status adsp_direct_circular(adsp_ctx *, const double *a, int64_t n, const double *b, int64_t m, double *out); /* DirectCircular conv.go:158 */
ADSP_API adsp_status adsp_convolve(adsp_ctx *, const double *a, int64_t n, const double *b, int64_t m, double *out);        /* Convolve       conv.go:194 */
ADSP_API adsp_status adsp_overlap_add_convolve(adsp_ctx *, const double *sig, int64_t n, const double *k, int64_t m, double *out);  /* overlap_add.go:221 */
ADSP_API adsp_status adsp_overlap_save_convolve(adsp_ctx *, const double *sig, int64_t n, const double *k, int64_t m, double *out); /* overlap_save.go:313 */
ADSP_API adsp_status adsp_correlate(adsp_ctx *, const double *a, int64_t n, const double *b, int64_t m, double *out);        /* Correlate       correlate.go:16 */
ADSP_API adsp_status adsp_correlate_direct(adsp_ctx *, const double *a, int64_t n, const double *b, int64_t m, double *out); /* CorrelateDirect correlate.go:31 */
ADSP_API adsp_status adsp_correlate_fft(adsp_ctx *, const double *a, int64_t n, const double *b, int64_t m, double *out);    /* CorrelateFFT    correlate.go:111 */
ADSP_API adsp_status adsp_correlate_normalized(adsp_ctx *, const double *a, int64_t n, const double *b, int64_t m, double *out); /* correlate.go:86 */
ADSP_API adsp_status adsp_autocorrelate_normalized(adsp_ctx *, const double *a, int64_t n, double *out);                    /* correlate.go:63 */
/* FindPeak, correlate.go:200: signed max, first index wins, (-1, 0) when len == 0. */
ADSP_API adsp_status adsp_find_peak(adsp_ctx *, const double *corr, int64_t len, int64_t *index, double *value);

ADSP_API adsp_status adsp_direct_f32(adsp_ctx *, const float *a, int64_t n, const float *b, int64_t m, float *out);
ADSP_API adsp_status adsp_convolve_f32(adsp_ctx *, const float *a, int64_t n, const float *b, int64_t m, float *out);
ADSP_API adsp_status adsp_correlate_f32(adsp_ctx *, const float *a, int64_t n, const float *b, int64_t m, float *out);

/* ---------------------------------------------------------------- batched one-shots
 * `batch` independent problems laid out with element strides (config 2 and 4 shapes).
 * kernel_stride == 0 shares one kernel across the batch. */
ADSP_API adsp_status adsp_direct_batch(adsp_ctx *, const double *a, int64_t n, int64_t a_stride,
                                       const double *b, int64_t m, int64_t b_stride, int64_t batch,
                                       double *out, int64_t out_stride);
/* Correlate every pair and report the peak (FindPeak rule).  out may be NULL (peaks only). */
ADSP_API adsp_status adsp_correlate_batch(adsp_ctx *, const double *a, int64_t n, int64_t a_stride,
                                          const double *b, int64_t m, int64_t b_stride, int64_t pairs,
                                          double *out, int64_t out_stride, int64_t *peak_index, double *peak_value);
/* Same, all pointers are DEVICE pointers; asynchronous on the context stream. */
ADSP_API adsp_status adsp_direct_batch_device(adsp_ctx *, const void *a, int64_t n, int64_t a_stride,
                                              const void *b, int64_t m, int64_t b_stride, int64_t batch,
                                              void *out, int64_t out_stride, adsp_precision prec);
ADSP_API adsp_status adsp_correlate_batch_device(adsp_ctx *, const void *a, int64_t n, int64_t a_stride,
                                                 const void *b, int64_t m, int64_t b_stride, int64_t pairs,
                                                 void *out, int64_t out_stride, void *peak_index_dev,
                                                 void *peak_value_dev, adsp_precision prec);

/* ---------------------------------------------------------------- reusable convolvers
 * NewOverlapSave(kernel, fftSize) overlap_save.go:53 / NewOverlapAdd(kernel, blockSize) overlap_add.go:44.
 * The kernel spectrum is computed once on the device and cached in the plan. */
ADSP_API adsp_status adsp_overlap_save_create(adsp_ctx *, const void *kernel, int64_t kernel_len, int64_t fft_size,
                                              adsp_precision prec, adsp_plan **out);
ADSP_API adsp_status adsp_overlap_add_create(adsp_ctx *, const void *kernel, int64_t kernel_len, int64_t block_size,
                                             adsp_precision prec, adsp_plan **out);
ADSP_API void adsp_plan_destroy(adsp_plan *plan);
ADSP_API void adsp_plan_reset(adsp_plan *plan);               /* Reset(): overlap_save.go:275 (stateless across Process) */
ADSP_API int64_t adsp_plan_kernel_len(const adsp_plan *plan); /* KernelLen() */
ADSP_API int64_t adsp_plan_fft_size(const adsp_plan *plan);   /* FFTSize(): the reference's formula value */
ADSP_API int64_t adsp_plan_step_size(const adsp_plan *plan);  /* StepSize() (OLS) */
ADSP_API int64_t adsp_plan_block_size(const adsp_plan *plan); /* BlockSize() (OLA) */
/* Internal transform geometry actually used on the GPU (free to differ from the getters). */
ADSP_API void adsp_plan_internal_geometry(const adsp_plan *plan, int64_t *fft_n, int64_t *n1, int64_t *n2,
                                          int64_t *step, int64_t *partitions);
/* Transforms the GPU will run for Process() on n input samples: up to `cap` segments, each written as
 * 4 values {transform length, first output sample, output samples, 1 if a single zero-padded block
 * (nothing discarded) else 0}.  Returns the number of segments.  Diagnostic only: results, lengths and
 * the reference getters (overlap_save.go:115-124) never depend on it. */
ADSP_API int adsp_plan_describe_cover(const adsp_plan *plan, int64_t n, int64_t *out4, int cap);

/* Process(input) / ProcessTo(output, input): out_len must equal n + kernel_len - 1
 * (ErrLengthMismatch otherwise, overlap_save.go:259-262).  Host pointers. */
ADSP_API adsp_status adsp_plan_process(adsp_plan *plan, const void *in, int64_t n, void *out, int64_t out_len);
/* `channels` signals of n samples, strides in elements; out rows hold n + kernel_len - 1. */
ADSP_API adsp_status adsp_plan_process_batch(adsp_plan *plan, const void *in, int64_t n, int64_t channels,
                                             int64_t in_stride, void *out, int64_t out_stride);
/* Device-resident variant (the timed path): pointers are device pointers on the plan's GPU. */
ADSP_API adsp_status adsp_plan_process_device(adsp_plan *plan, const void *in_dev, int64_t n, int64_t channels,
                                              int64_t in_stride, void *out_dev, int64_t out_stride);
ADSP_API adsp_status adsp_plan_sync(adsp_plan *plan);

/* ---------------------------------------------------------------- several GPUs from one host process
 * Shards are independent (no collective): contiguous channel ranges, or time blocks of one long signal with a
 * (kernel_len - 1) halo read from the source -- the rule of overlap_save.go:205-215 at shard granularity, the last
 * shard also emitting the tail (:224-251).  The two helpers are pure arithmetic (also what one-process-per-GPU
 * launchers use: algo_dsp_b200/shard.py); the adsp_plans_* calls drive one plan per GPU from one host thread each. */
ADSP_API void adsp_shard_channel_range(int64_t channels, int rank, int world, int64_t *lo, int64_t *hi);
ADSP_API void adsp_shard_time(int64_t n, int64_t kernel_len, int rank, int world, int64_t *out_lo, int64_t *out_hi,
                              int64_t *in_lo, int64_t *in_hi, int64_t *skip);
/* plans[i]: OverlapSave/OverlapAdd plans of the SAME kernel and precision, one per context/device; host pointers */
ADSP_API adsp_status adsp_plans_process_batch(adsp_plan *const *plans, int nplans, const void *in, int64_t n, int64_t channels,
                                              int64_t in_stride, void *out, int64_t out_stride);
ADSP_API adsp_status adsp_plans_process_long(adsp_plan *const *plans, int nplans, const void *in, int64_t n, void *out,
                                             int64_t out_len);

/* ---------------------------------------------------------------- deconvolution, float64 like the reference
 * Deconvolve(signal, kernel, opts) deconvolve.go:72: circular spectral division at N = nextPow2(len(signal));
 * method 0 = DeconvNaive (ErrDivisionByZero when a bin has |H| < 1e-15; the bin index is in adsp_last_error),
 * 1 = DeconvRegularized (epsilon <= 0 -> 1e-6), 2 = DeconvWiener (variances <= 0 are estimated as the reference does),
 * anything else = regularized with 1e-6.  out_len must equal adsp_deconv_out_len(n, m) (n - m + 1, or n if that is <= 0). */
ADSP_API int64_t adsp_deconv_out_len(int64_t n, int64_t m);
ADSP_API adsp_status adsp_deconvolve(adsp_ctx *, const double *signal, int64_t n, const double *kernel, int64_t m, int method,
                                     double epsilon, double noise_variance, double signal_variance, double *out, int64_t out_len);
/* InverseFilter(kernel, length, epsilon) deconvolve.go:359 */
ADSP_API adsp_status adsp_inverse_filter(adsp_ctx *, const double *kernel, int64_t m, int64_t length, double epsilon, double *out);
/* SNR(original, recovered) deconvolve.go:417 (host arithmetic; -Inf on length mismatch or empty input) */
ADSP_API double adsp_snr(const double *original, int64_t n, const double *recovered, int64_t n2);
/* `batch` independent problems, device pointers, strides in elements (k_stride 0 = one kernel for all); reg < 0 = naive. */
ADSP_API adsp_status adsp_deconvolve_batch_device(adsp_ctx *, const double *signal_dev, int64_t n, int64_t s_stride,
                                                  const double *kernel_dev, int64_t m, int64_t k_stride, int64_t batch,
                                                  double reg, double *out_dev, int64_t out_stride);

/* ---------------------------------------------------------------- dsp/signal generators on the device
 * The step BEFORE the path: inputs are generated in HBM (csrc/siggen.cu).  Formulas of dsp/signal/generate.go --
 * WhiteNoise :188, PinkNoise :210 (Voss-McCartney, tables :220-221), LinearSweep :134, LogSweep :157, Normalize :253,
 * RemoveDC :306 -- on uniforms from a STATELESS hash of (seed, stream, index) instead of Go's math/rand (whose seeding
 * table is not in the tree): sample i depends only on (seed, i), so any shard of a stream can be produced anywhere
 * (index0 = first stream index of this call).  Row r of a batch uses seed0 + r * seed_step.  out: DEVICE pointer, `rows`
 * rows of n samples, stride in elements; asynchronous on the context stream.  The *_host twins run the same arithmetic
 * (csrc/siggen_core.h) on the CPU and are bit-identical to the fp64 device output -- for feeding a CPU reference the
 * very same signal.  Errors mirror the reference's argument checks (ADSP_ERR_INVALID_ARG + text). */
ADSP_API adsp_status adsp_gen_uniform_device(adsp_ctx *, void *out, int64_t n, int64_t rows, int64_t stride, int64_t seed0,
                                             int64_t seed_step, int64_t index0, adsp_precision prec);
ADSP_API adsp_status adsp_gen_white_device(adsp_ctx *, void *out, int64_t n, int64_t rows, int64_t stride, double amplitude,
                                           int64_t seed0, int64_t seed_step, int64_t index0, adsp_precision prec);
ADSP_API adsp_status adsp_gen_pink_device(adsp_ctx *, void *out, int64_t n, int64_t rows, int64_t stride, double amplitude,
                                          int64_t seed0, int64_t seed_step, int64_t index0, adsp_precision prec);
/* synthetic decaying IR of SURVEY 8d: h[i] = (u_i*2-1) * 10^(-decades*i/taps)  (decades = 3: -60 dB at the last tap) */
ADSP_API adsp_status adsp_gen_decaying_ir_device(adsp_ctx *, void *out, int64_t taps, int64_t rows, int64_t stride, double decades,
                                                 int64_t seed0, int64_t seed_step, adsp_precision prec);
/* samples [index0, index0+n) of a sweep that is total_samples long */
ADSP_API adsp_status adsp_gen_linear_sweep_device(adsp_ctx *, void *out, int64_t n, int64_t index0, int64_t total_samples,
                                                  double start_hz, double end_hz, double amplitude, double sample_rate, adsp_precision prec);
ADSP_API adsp_status adsp_gen_log_sweep_device(adsp_ctx *, void *out, int64_t n, int64_t index0, int64_t total_samples,
                                               double start_hz, double end_hz, double amplitude, double sample_rate, adsp_precision prec);
/* config 4 responses: out[r][i] = (i >= d_r ? src[i-d_r] : 0) + white(seed0 + r*seed_step)[i] * noise_amplitude,
 * d_r = adsp_gen_delay_host(delay_seed, r, delay_mod); delays_dev (nullable) receives d_r as int64. rows <= 65535. */
ADSP_API adsp_status adsp_gen_delay_mix_device(adsp_ctx *, void *out, int64_t n, int64_t rows, int64_t stride, const void *src_dev,
                                               double noise_amplitude, int64_t seed0, int64_t seed_step, int64_t delay_seed,
                                               int64_t delay_mod, int64_t *delays_dev, adsp_precision prec);
/* Normalize(data, targetPeak) generate.go:253 / RemoveDC(data) :306, row-wise on device data (rows <= 65535). */
ADSP_API adsp_status adsp_normalize_device(adsp_ctx *, const void *in_dev, int64_t n, int64_t rows, int64_t in_stride, double target_peak,
                                           void *out_dev, int64_t out_stride, adsp_precision prec);
ADSP_API adsp_status adsp_remove_dc_device(adsp_ctx *, const void *in_dev, int64_t n, int64_t rows, int64_t in_stride, void *out_dev,
                                           int64_t out_stride, adsp_precision prec);
ADSP_API void adsp_gen_uniform_host(double *out, int64_t n, int64_t seed, int64_t index0);
ADSP_API void adsp_gen_white_host(double *out, int64_t n, double amplitude, int64_t seed, int64_t index0);
ADSP_API void adsp_gen_pink_host(double *out, int64_t n, double amplitude, int64_t seed, int64_t index0);
ADSP_API void adsp_gen_decaying_ir_host(double *out, int64_t taps, double decades, int64_t seed);
ADSP_API void adsp_gen_linear_sweep_host(double *out, int64_t n, int64_t index0, int64_t total_samples, double start_hz, double end_hz,
                                         double amplitude, double sample_rate);
ADSP_API void adsp_gen_log_sweep_host(double *out, int64_t n, int64_t index0, int64_t total_samples, double start_hz, double end_hz,
                                      double amplitude, double sample_rate);
ADSP_API int64_t adsp_gen_delay_host(int64_t delay_seed, int64_t row, int64_t delay_mod);

/* ---------------------------------------------------------------- the step AFTER the path: csrc/post.cu
 * Consumers of correlation / deconvolution results, on device data so that nothing returns to the host in between.
 * float64 like the reference.  Rows: `rows` signals of n samples, strides in elements (rows <= 65535 per call).
 *
 * measure/ir: SchroederIntegral (measure/ir/ir.go:94-130): backward cumulative energy, normalised, in dB (-200 dB floor;
 * an all-zero IR returns the raw zero sums); FindImpulseStart (:381-404): first sample with |x| >= ratio * peak (the
 * reference uses ratio 0.1), 0 when none; findPeak (:406-424): index of the absolute maximum, first one wins.
 * n == 0 -> ADSP_ERR_EMPTY_IR (ir.ErrEmptyIR). */
ADSP_API adsp_status adsp_ir_schroeder_device(adsp_ctx *, const double *ir_dev, int64_t n, int64_t rows, int64_t stride, double *out_dev, int64_t out_stride);
ADSP_API adsp_status adsp_ir_find_impulse_start_device(adsp_ctx *, const double *ir_dev, int64_t n, int64_t rows, int64_t stride, double threshold_ratio,
                                                       int64_t *index_dev);
ADSP_API adsp_status adsp_ir_find_peak_device(adsp_ctx *, const double *ir_dev, int64_t n, int64_t rows, int64_t stride, int64_t *index_dev);
ADSP_API adsp_status adsp_ir_schroeder(adsp_ctx *, const double *ir, int64_t n, double *out);                                       /* host pointers */
ADSP_API adsp_status adsp_ir_find_impulse_start(adsp_ctx *, const double *ir, int64_t n, double threshold_ratio, int64_t *index); /* host pointers */
/* measure/sweep LogSweep (measure/sweep/sweep.go): Generate :73-94, InverseFilter :104-155 (time-reversed sweep with a
 * 6 dB/octave envelope, normalised by T*f1/ln(f2/f1)*sampleRate), Deconvolve :164-239 = full linear convolution of the
 * response with the inverse filter (n + samples - 1 values; the IR peaks near index samples - 1).  samples =
 * round(duration * sampleRate).  Validation as LogSweep.Validate (:37-55): ADSP_ERR_INVALID_ARG + the reference's text;
 * an empty response -> ADSP_ERR_EMPTY_INPUT (sweep.ErrEmptyResponse).  The *_host twins are bit-identical to the device. */
ADSP_API int64_t adsp_logsweep_samples(double duration, double sample_rate);
ADSP_API adsp_status adsp_logsweep_generate_device(adsp_ctx *, double *out_dev, double start_hz, double end_hz, double duration, double sample_rate);
ADSP_API adsp_status adsp_logsweep_inverse_filter_device(adsp_ctx *, double *out_dev, double start_hz, double end_hz, double duration, double sample_rate);
ADSP_API adsp_status adsp_logsweep_generate_host(double *out, double start_hz, double end_hz, double duration, double sample_rate);
ADSP_API adsp_status adsp_logsweep_inverse_filter_host(double *out, double start_hz, double end_hz, double duration, double sample_rate);
ADSP_API adsp_status adsp_logsweep_deconvolve_device(adsp_ctx *, const double *response_dev, int64_t n, int64_t rows, int64_t in_stride, double start_hz,
                                                     double end_hz, double duration, double sample_rate, double *out_dev, int64_t out_stride);
ADSP_API adsp_status adsp_logsweep_deconvolve(adsp_ctx *, const double *response, int64_t n, double start_hz, double end_hz, double duration,
                                              double sample_rate, double *out, int64_t out_len);
/* dsp/filter/fir Filter (dsp/filter/fir/filter.go): New(coeffs) :18, ProcessBlock(buf) :61-103 in place with the delay
 * line carried across calls, Reset, Order.  `channels` independent filters with the same coefficients, one per row.
 * Tap order exactly as the reference applies it: below 32 taps y[n] = sum_k h[k] x[n-k] (ProcessSample, :36-59); from 32
 * taps on the block path dots the coefficients with the window stored oldest first (:93-94), y[n] = sum_k h[N-1-k] x[n-k]
 * -- identical for the symmetric filters the reference designs and tests. */
typedef struct adsp_fir adsp_fir;
ADSP_API adsp_status adsp_fir_create(adsp_ctx *, const double *coeffs, int64_t ntaps, int channels, adsp_fir **out);
ADSP_API adsp_status adsp_fir_process_block(adsp_fir *, double *buf, int64_t n, int64_t stride);             /* host pointers */
ADSP_API adsp_status adsp_fir_process_block_device(adsp_fir *, double *buf_dev, int64_t n, int64_t stride);
ADSP_API int64_t adsp_fir_order(const adsp_fir *);
ADSP_API void adsp_fir_reset(adsp_fir *);
ADSP_API void adsp_fir_destroy(adsp_fir *);
/* dsp/resample Resampler (dsp/resample/resample.go, resample_design.go): NewRational(up, down, options) :153 with the
 * Kaiser-windowed-sinc polyphase design of designPolyphaseFIR; quality 0 fast / 1 balanced / 2 best (QualityProfile :35),
 * taps_per_phase / cutoff_scale / kaiser_beta <= 0 keep the profile's value; NewForRates :194 (continued-fraction ratio,
 * max_den <= 0 -> 4096).  Process :249-292 keeps phase, input index and history across calls; n_out receives the number
 * of samples written per row (= adsp_resampler_predict_output_len before the call); out_cap smaller than that ->
 * ADSP_ERR_LENGTH_MISMATCH.  up or down <= 0 -> ADSP_ERR_INVALID_ARG ("resample: invalid ratio"). */
typedef struct adsp_resampler adsp_resampler;
ADSP_API void adsp_resample_approximate_ratio(double v, int max_den, int *num, int *den);
ADSP_API adsp_status adsp_resampler_create(adsp_ctx *, int up, int down, int quality, int taps_per_phase, double cutoff_scale, double kaiser_beta,
                                           int channels, adsp_resampler **out);
ADSP_API adsp_status adsp_resampler_create_for_rates(adsp_ctx *, double in_rate, double out_rate, int quality, int max_den, int channels,
                                                     adsp_resampler **out);
ADSP_API void adsp_resampler_ratio(const adsp_resampler *, int *up, int *down);
ADSP_API int adsp_resampler_taps_per_phase(const adsp_resampler *);
ADSP_API int64_t adsp_resampler_prototype(const adsp_resampler *, double *out, int64_t cap);
ADSP_API int64_t adsp_resampler_predict_output_len(const adsp_resampler *, int64_t input_len);
ADSP_API adsp_status adsp_resampler_process(adsp_resampler *, const double *in, int64_t n, int64_t in_stride, double *out, int64_t out_cap,
                                            int64_t out_stride, int64_t *n_out);                             /* host pointers */
ADSP_API adsp_status adsp_resampler_process_device(adsp_resampler *, const double *in_dev, int64_t n, int64_t in_stride, double *out_dev,
                                                   int64_t out_cap, int64_t out_stride, int64_t *n_out);
ADSP_API void adsp_resampler_reset(adsp_resampler *);
ADSP_API void adsp_resampler_destroy(adsp_resampler *);

/* ---------------------------------------------------------------- partitioned (long IR, streaming)
 * NewPartitionedConvolution(kernel, minBlockOrder, maxBlockOrder) partitioned.go:212,335.
 * ProcessBlock(input, output): equal lengths, output delayed by Latency() = 2^minBlockOrder. */
ADSP_API adsp_status adsp_partitioned_create(adsp_ctx *, const void *kernel, int64_t kernel_len, int min_block_order,
                                             int max_block_order, adsp_precision prec, adsp_plan **out);
ADSP_API adsp_status adsp_partitioned_process_block(adsp_plan *plan, const void *in, int64_t n, void *out, int64_t n_out);
ADSP_API int adsp_partitioned_latency(const adsp_plan *plan);      /* partitioned.go:410 */
ADSP_API int adsp_partitioned_stage_count(const adsp_plan *plan);  /* partitioned.go:420 */
ADSP_API adsp_status adsp_partitioned_stage_info(const adsp_plan *plan, int index, int *part_size, int *block_count); /* :426 */

/* Many channels per launch: `channels` independent streams through one IR, each row behaving exactly
 * like its own PartitionedConvolution (partitioned.go:348-396): row c of `out` = full linear convolution of row c
 * of `in`, delayed by Latency(), arbitrary n per call.  Runs on a device-resident frequency-domain delay line with
 * a fused spectral multiply-accumulate (csrc/fdl.cu); needs min_block_order >= 3.  Strides in elements. */
ADSP_API adsp_status adsp_partitioned_create_batch(adsp_ctx *, const void *kernel, int64_t kernel_len, int min_block_order,
                                                   int max_block_order, int channels, adsp_precision prec, adsp_plan **out);
ADSP_API adsp_status adsp_partitioned_process_block_batch(adsp_plan *plan, const void *in, int64_t n, int64_t in_stride,
                                                          void *out, int64_t out_stride);
ADSP_API adsp_status adsp_partitioned_process_block_batch_device(adsp_plan *plan, const void *in_dev, int64_t n, int64_t in_stride,
                                                                 void *out_dev, int64_t out_stride);
/* ConvolutionReverb.SetWetDry / ProcessInPlace (dsp/effects/reverb/convolution.go:51-83), the mix fused into the
 * output kernel: block[c][i] = dry * block[c][i] + wet * reverb(block[c])[i]. */
ADSP_API adsp_status adsp_partitioned_set_wet_dry(adsp_plan *plan, double wet, double dry);
ADSP_API adsp_status adsp_partitioned_process_in_place_batch(adsp_plan *plan, void *block, int64_t n, int64_t stride);
ADSP_API adsp_status adsp_partitioned_process_in_place_batch_device(adsp_plan *plan, void *block_dev, int64_t n, int64_t stride);
ADSP_API int adsp_partitioned_channels(const adsp_plan *plan);
/* Internal stage layout of the delay-line engine (diagnostic; StageCount/StageInfo report the reference's layout). */
ADSP_API int adsp_partitioned_plan_layout(int64_t kernel_len, int min_block_order, int max_block_order, int *part_size, int *count,
                                          int64_t *ir_offset, int cap);   /* same layout, no plan and no GPU needed */
ADSP_API int adsp_partitioned_internal_stage_count(const adsp_plan *plan);
ADSP_API adsp_status adsp_partitioned_internal_stage_info(const adsp_plan *plan, int index, int *part_size, int *count,
                                                          int64_t *ir_offset);

/* ---------------------------------------------------------------- fixed-block streaming convolvers
 * NewStreamingOverlapAdd(kernel, blockSize) streaming_overlap_add.go:41 / NewStreamingOverlapSave
 * streaming_overlap_save.go:44 (+ the float32 twins): ProcessBlock(input[blockSize]) -> output[blockSize],
 * state carried across calls; wrong input or output length -> ErrLengthMismatch; blockSize <= 0 ->
 * ADSP_ERR_INVALID_ARG.  adsp_plan_block_size / adsp_plan_fft_size / adsp_plan_kernel_len /
 * adsp_plan_reset serve as BlockSize() / FFTSize() / KernelLen() / Reset(). */
ADSP_API adsp_status adsp_streaming_create(adsp_ctx *, const void *kernel, int64_t kernel_len, int64_t block_size,
                                           int overlap_save, adsp_precision prec, adsp_plan **out);
ADSP_API adsp_status adsp_streaming_process_block(adsp_plan *plan, const void *in, int64_t n, void *out, int64_t n_out);

#ifdef __cplusplus
}
#endif
#endif /* ALGODSP_CUDA_H */
