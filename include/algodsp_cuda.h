/*
 * algodsp_cuda.h -- C ABI of libalgodsp_cuda, the B200 (sm_100a) implementation of the
 * dsp/conv hot path of CWBudde/algo-dsp.
 *
 * The reference has no FFI boundary of its own: the boundary it replaces is the exported Go
 * API of package conv (dsp/conv/*.go).  Each entry point below cites the Go symbol
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
/* Raw cudaStream_t of the context's main stream (for event timing by a harness). */
ADSP_API void *adsp_ctx_stream(adsp_ctx *ctx);

/* Host-buffer calls and pageable memory.  Every entry point that takes HOST pointers accepts ordinary pageable memory
 * (a Go []float64, malloc, numpy): the library stages it through its own pinned buffers with a pool of copy threads
 * (ADSP_STAGE_THREADS, default min(8, cores/2)), overlapped with H2D | kernels | D2H (csrc/staging.cu).  Memory from
 * adsp_host_alloc_pinned (or cudaHostRegister'ed by the caller) is DMA'd in place, which saves the two host copies.
 * adsp_ctx_host_profile(ctx, 1) makes the small-call path (one upload, kernels, one download) record a breakdown of the
 * last call: ms6 = {device allocation, upload (stage-in + H2D), kernels, download (D2H + stage-out), 0, total}; the
 * phases are serialised by a stream sync each while profiling is on.  staged_*_bytes count what went through staging. */
ADSP_API void adsp_ctx_host_profile(adsp_ctx *ctx, int enable);
ADSP_API adsp_status adsp_ctx_host_profile_get(adsp_ctx *ctx, double *ms6, uint64_t *staged_in_bytes, uint64_t *staged_out_bytes);
ADSP_API int adsp_host_ptr_is_pinned(const void *p);     /* 1: DMA-able in place, 0: pageable (will be staged) */
ADSP_API int adsp_ctx_stage_threads(adsp_ctx *ctx);      /* size of the context's copy-thread pool (creates it) */

/* Pinned host memory for callers that want zero-copy DMA (Go side: C.malloc replacement). */
ADSP_API adsp_status adsp_host_alloc_pinned(size_t bytes, void **out);
ADSP_API void adsp_host_free_pinned(void *p);
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
