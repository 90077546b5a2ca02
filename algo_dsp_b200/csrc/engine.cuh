// engine.cuh -- host-side engine shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/algodsp_cuda.h"
#include "conv_kernels.cuh"
#include "conv_kernels_mr.cuh"
#include "conv_kernels_mrp.cuh"

namespace adsp {

void set_error(const std::string &msg);
adsp_status cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define ADSP_CUDA(call)                                                        \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) return ::adsp::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)
#define ADSP_TRY(call)                       \
    do {                                     \
        adsp_status s__ = (call);            \
        if (s__ != ADSP_OK) return s__;      \
    } while (0)

constexpr int kWorkerStreams = 8;   // created per context; ADSP_STREAMS (default 3) of them are used per call

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    adsp_status reserve(size_t bytes);
    void release();
};

struct PinnedBuf {
    void *p = nullptr;
    size_t cap = 0;
    adsp_status reserve(size_t bytes);
    void release();
};

// ------------------------------------------------------------------ pinned host staging (staging.cu)
// A Go []float64 (overlap_save.go:132-133) -- like any malloc'ed buffer -- is PAGEABLE: the DMA engines cannot read it, and
// the CUDA runtime's own pageable path is a synchronous single-threaded bounce copy.  The library therefore stages
// pageable caller memory through its own pinned buffers with a small pool of copy threads, overlapped with the
// H2D | kernels | D2H pipeline.  Pinned or registered caller memory (adsp_host_alloc_pinned) is DMA'd in place.
constexpr int kPipeSlots = 3;       // device / pinned slots per direction of the host-buffer pipeline

class StagePool {
public:
    struct Job { std::atomic<long long> remaining{0}; std::mutex m; std::condition_variable cv; };
    using Ticket = std::shared_ptr<Job>;
    explicit StagePool(int nthreads);
    ~StagePool();
    int threads() const { return (int)workers_.size(); }
    // copies `rows` rows of `width` bytes (row pitches in bytes), split into slices run by the pool; returns at once
    // flush_src: evict every source line from the CPU caches once read (stage-out: the source is the next DMA target)
    Ticket copy2d_async(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows, bool flush_src = false);
    void wait(const Ticket &t);          // the waiting thread helps: it runs queued slices until its job is done
    void copy2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows) { wait(copy2d_async(dst, dpitch, src, spitch, width, rows)); }
private:
    struct Slice { char *dst; const char *src; size_t dpitch, spitch, width, rows; Ticket job; bool flush_src = false; };
    void run();
    static void exec(Slice &s);
    std::vector<std::thread> workers_;
    std::deque<Slice> q_;
    std::mutex m_;
    std::condition_variable cv_;
    bool stop_ = false;
};

// true when the DMA engines can access `p` directly (cudaMallocHost / cudaHostRegister / managed memory)
bool host_ptr_is_pinned(const void *p);

// Internal transform geometry chosen for a kernel length (free to differ from the reference's
// FFTSize()/StepSize(), which are reported from the reference formulas).
struct FftChoice {
    long long N = 0;       // transform length
    int N1 = 1, N2 = 0;    // N = N1*N2 (N1 == 1: single-kernel path)
    int lgN = 0;
    int P = 1;             // odd factor of N (mixed-radix columns); 1: power-of-two transform
    int M = 1;             // length of the in-register column DFT: P or 2P; N1 = 16*M
    long long D = 0, S = 0;  // discard / step per block
    int parts = 1;           // IR partitions (only when K-1 exceeds half the largest transform)
    long long part_len = 0;  // taps per partition
};

FftChoice choose_fft(long long K);
FftChoice make_choice(long long Kp, long long N);
long long env_ll(const char *name, long long dflt);

}  // namespace adsp

struct adsp_ctx {
    int device = 0;
    int sm_count = 0;
    size_t l2_bytes = 0;
    cudaStream_t main = nullptr;
    cudaStream_t worker[adsp::kWorkerStreams] = {};
    cudaEvent_t ev_fork = nullptr;
    cudaEvent_t ev_join[adsp::kWorkerStreams] = {};
    std::mutex mu;  // serialises calls that use the context's staging / scratch buffers
    std::map<std::pair<int, int>, void *> tw_tables;                     // (L, prec) -> device table
    std::map<std::pair<int, int>, std::pair<void *, void *>> tw4_tables;  // (lgN, prec) -> (hi, lo)
    adsp::DevBuf scratch;              // four-step intermediates (L2 resident by construction)
    adsp::DevBuf d_in, d_out, d_k, d_tmp, d_small, d_counters;
    adsp::DevBuf spec_cache;           // spectra of transient convolvers (one-shot calls, shared-b correlation): no cudaMalloc / cudaFree per call
    adsp::PinnedBuf h_in[adsp::kPipeSlots], h_out[adsp::kPipeSlots];   // pinned staging of pageable caller memory
    std::unique_ptr<adsp::StagePool> pool;                               // copy threads (created on the first pageable call)
    // host-path profile (adsp_ctx_host_profile): ms of {stage-in, H2D, kernels, D2H, stage-out, total} of the last small call
    bool host_profile = false;
    double host_prof_ms[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // [6] stage-in copy, [7] wait for D2H, [8] stage-out copy, [9] pointer queries
    uint64_t staged_bytes_in = 0, staged_bytes_out = 0;                  // bytes that went through pinned staging (diagnostic)
    std::atomic<uint64_t> launches{0};
    size_t scratch_budget = 0;  // bytes of scratch allowed in flight (fits L2)
    // host-staged pipeline (H2D | compute | D2H overlapped over channel chunks)
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t ev_in[adsp::kPipeSlots] = {}, ev_comp[adsp::kPipeSlots] = {}, ev_out[adsp::kPipeSlots] = {};
    adsp::DevBuf pipe_in[adsp::kPipeSlots], pipe_out[adsp::kPipeSlots];
    // optional per-kernel timing (bench roofline): event pairs around every launch of a kind
    bool timing = false;
    struct TimedLaunch { int kind; cudaEvent_t e0, e1; };
    std::vector<TimedLaunch> timed;
    std::vector<cudaEvent_t> event_pool;
    double kind_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint64_t kind_launches[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

namespace adsp {

// Device-resident FFT convolver for one (partition of an) impulse response.
template <typename T> struct FftConv {
    adsp_ctx *ctx = nullptr;
    long long K = 0;  // taps of this partition
    FftChoice ch;
    cpx<T> *H = nullptr;  // cached spectrum (four-step order, scaled 1/N)
    const cpx<T> *tw_rows = nullptr, *tw_cols = nullptr, *tw_hi = nullptr, *tw_lo = nullptr;
    bool no_discard = false;  // next run(): single zero-padded block, D = 0, S = N (set by the caller per segment)
    bool owns_H = true;       // false: H lives in a buffer of the caller (transient convolvers reuse the context's spectrum cache)

    // H_ext != nullptr: build the spectrum into that buffer (ch.N complex elements) instead of allocating one
    adsp_status init(adsp_ctx *c, const T *d_kernel, long long K_, const FftChoice &choice, cpx<T> *H_ext = nullptr);
    void destroy();
    // y[ch][out_shift + o] (=|+=) conv(x[ch], h)[o + in_shift'], see ConvGeom
    adsp_status run(const T *d_x, long long n, long long channels, long long in_stride, T *d_y,
                    long long out_stride, long long out_len, long long in_shift, long long out_shift,
                    bool accumulate);
};

template <typename T> adsp_status get_tw_table(adsp_ctx *ctx, int L, const cpx<T> **out);
template <typename T> adsp_status get_tw4_tables(adsp_ctx *ctx, long long N, const cpx<T> **hi, const cpx<T> **lo);
template <typename T> adsp_status get_twp_table(adsp_ctx *ctx, int M, const cpx<T> **out);
bool fft_size_supported(long long N);

// full linear convolution of `channels` signals with one kernel (device pointers), any K:
// splits long kernels into partitions.  Builds the spectra on the fly (one-shot use).
template <typename T>
adsp_status fft_convolve_device(adsp_ctx *ctx, const T *d_x, long long n, long long channels, long long in_stride,
                                const T *d_k, long long K, T *d_y, long long out_stride);

template <typename T>
adsp_status fft_correlate_pairs_device(adsp_ctx *ctx, const T *a, long long n, long long a_stride, const T *b, long long m,
                                       long long b_stride, long long pairs, T *out, long long out_stride, T *peak_v, long long *peak_i, bool *done);

template <typename T>
adsp_status fft_deconvolve_device(adsp_ctx *ctx, const T *sig, long long n, long long s_stride, const T *ker, long long m,
                                  long long k_stride, long long batch, T *out, long long out_stride, long long out_len, T reg,
                                  long long *d_bad);

template <typename T>
adsp_status direct_device(adsp_ctx *ctx, const T *d_a, long long n, long long a_stride, const T *d_b, long long m,
                          long long b_stride, long long batch, T *d_out, long long out_stride, const T *h_b = nullptr);

// Streaming partitioned convolution on a frequency-domain delay line, many channels per launch (fdl.cu)
struct FdlEngine;
struct FdlStage { int part_size; int count; long long ir_offset; };
bool fdl_supported(int min_order);
std::vector<FdlStage> fdl_layout(long long K, int min_order, int max_order);
adsp_status fdl_create(adsp_ctx *ctx, const void *d_kernel, long long K, int min_order, int max_order, int channels,
                       adsp_precision prec, FdlEngine **out);
void fdl_destroy(FdlEngine *e);
adsp_status fdl_reset(FdlEngine *e);
void fdl_set_wet_dry(FdlEngine *e, double wet, double dry);
int fdl_channels(const FdlEngine *e);
int fdl_stage_count(const FdlEngine *e);
void fdl_stage_info(const FdlEngine *e, int i, int *part, int *count, long long *off);
// in/out: `channels` rows of n samples (strides in elements), host or device pointers; mix: out = dry*in + wet*y
adsp_status fdl_process(FdlEngine *e, const void *in, long long n, long long in_stride, void *out, long long out_stride, bool host_ptrs,
                        bool mix);

// host <-> device transfers that stage pageable memory through the context's pinned buffers (staging.cu).
// upload: returns once `src` has been consumed (the device copy may still be in flight on ctx->main);
// download: enqueues after everything already on ctx->main, returns when `dst` holds the data.
StagePool *stage_pool(adsp_ctx *ctx);
adsp_status upload(adsp_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
adsp_status download(adsp_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);
adsp_status upload2d(adsp_ctx *ctx, void *dst_dev, size_t dpitch, const void *src_host, size_t spitch, size_t width, size_t rows);
adsp_status download2d(adsp_ctx *ctx, void *dst_host, size_t dpitch, const void *src_dev, size_t spitch, size_t width, size_t rows);

inline void count_launch(adsp_ctx *ctx, int n = 1) { ctx->launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

enum KernelKind { KK_COLS_FWD = 0, KK_ROWS = 1, KK_COLS_INV = 2, KK_FULL = 3, KK_DIRECT = 4, KK_OTHER = 5, KK_FUSED = 6 };

// RAII bracket: when ctx->timing is on, records an event pair around one launch
struct LaunchTimer {
    adsp_ctx *ctx; cudaStream_t st; int kind; cudaEvent_t e0 = nullptr;
    static cudaEvent_t get(adsp_ctx *c) {
        if (!c->event_pool.empty()) { cudaEvent_t e = c->event_pool.back(); c->event_pool.pop_back(); return e; }
        cudaEvent_t e = nullptr; cudaEventCreate(&e); return e;
    }
    LaunchTimer(adsp_ctx *c, cudaStream_t s, int k) : ctx(c), st(s), kind(k) {
        if (ctx->timing) { e0 = get(ctx); cudaEventRecord(e0, st); }
    }
    ~LaunchTimer() {
        if (e0) { cudaEvent_t e1 = get(ctx); cudaEventRecord(e1, st); ctx->timed.push_back({kind, e0, e1}); }
    }
};

}  // namespace adsp
