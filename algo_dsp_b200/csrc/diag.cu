// diag.cu -- run-time measurements the bench reports its rooflines against (not part of the data path):
//   * per-SM pipe peaks of THIS GPU at THIS clock: FP64 FMA issue rate and shared-memory (LSU) bandwidth, the two
//     resources that bound the fp64 FFT kernels before HBM does;
//   * the raw pinned-memory copy ceiling of the host link: concurrent H2D and D2H cudaMemcpyAsync of given sizes,
//     which bounds any end-to-end (host buffers in, host buffers out) figure.
#include "engine.cuh"

namespace adsp {
namespace {

__global__ void __launch_bounds__(256, 2) pipe_dfma(double *out, int iters) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
    const double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fma(a[i], b, c);              // 64 independent-enough DFMA per iteration
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    out[(size_t)blockIdx.x * 256 + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256, 2) pipe_smem(double *out, int iters) {
    extern __shared__ double2 sm[];
    const int t = threadIdx.x;
    double2 v[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { v[i].x = t; v[i].y = i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {                                         // 4 x (4 STS.128 + 4 LDS.128) = 512 B per thread
#pragma unroll
            for (int i = 0; i < 4; i++) sm[i * 256 + t] = v[i];
#pragma unroll
            for (int i = 0; i < 4; i++) { const double2 q = sm[i * 256 + ((t + 32) & 255)]; v[i].x += q.x; v[i].y = q.y; }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) s += v[i].x + v[i].y;
    out[(size_t)blockIdx.x * 256 + t] = s;
}

}  // namespace
}  // namespace adsp

using namespace adsp;

extern "C" {

// dfma_per_s: FP64 FMA thread-instructions per second over the whole GPU; smem_bytes_per_s: shared-memory bytes moved
// (stores + loads, 128-bit accesses) per second over the whole GPU.  Each kernel runs ~1-2 ms after a warm-up launch.
adsp_status adsp_ctx_measure_pipes(adsp_ctx *ctx, double *dfma_per_s, double *smem_bytes_per_s) {
    if (!ctx) return ADSP_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    const int grid = 2 * ctx->sm_count, iters = 4000;
    ADSP_TRY(ctx->d_tmp.reserve((size_t)grid * 256 * sizeof(double)));
    cudaEvent_t e0, e1;
    ADSP_CUDA(cudaEventCreate(&e0));
    ADSP_CUDA(cudaEventCreate(&e1));
    float ms = 0;
    pipe_dfma<<<grid, 256, 0, ctx->main>>>((double *)ctx->d_tmp.p, 200);
    ADSP_CUDA(cudaEventRecord(e0, ctx->main));
    pipe_dfma<<<grid, 256, 0, ctx->main>>>((double *)ctx->d_tmp.p, iters);
    ADSP_CUDA(cudaEventRecord(e1, ctx->main));
    ADSP_CUDA(cudaStreamSynchronize(ctx->main));
    ADSP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (dfma_per_s) *dfma_per_s = (double)grid * 256.0 * 64.0 * iters / (ms * 1e-3);
    pipe_smem<<<grid, 256, 16384, ctx->main>>>((double *)ctx->d_tmp.p, 200);
    ADSP_CUDA(cudaEventRecord(e0, ctx->main));
    pipe_smem<<<grid, 256, 16384, ctx->main>>>((double *)ctx->d_tmp.p, iters);
    ADSP_CUDA(cudaEventRecord(e1, ctx->main));
    ADSP_CUDA(cudaStreamSynchronize(ctx->main));
    ADSP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (smem_bytes_per_s) *smem_bytes_per_s = (double)grid * 256.0 * 512.0 * iters / (ms * 1e-3);
    count_launch(ctx, 4);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}

// `reps` rounds of one H2D copy of h2d_bytes and one D2H copy of d2h_bytes running at the same time on two streams,
// from/to freshly allocated pinned buffers.  ms_per_round: wall time per round (both directions overlapped);
// h2d_alone_ms / d2h_alone_ms: each direction on its own.
adsp_status adsp_ctx_copy_ceiling(adsp_ctx *ctx, size_t h2d_bytes, size_t d2h_bytes, int reps, double *ms_per_round, double *h2d_alone_ms,
                                  double *d2h_alone_ms) {
    if (!ctx || reps < 1 || (h2d_bytes == 0 && d2h_bytes == 0)) return ADSP_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    void *hi = nullptr, *ho = nullptr, *di = nullptr, *dout = nullptr;
    adsp_status st = ADSP_OK;
    auto fail = [&](cudaError_t e) { if (e != cudaSuccess && st == ADSP_OK) st = cuda_fail(e, "copy ceiling", __FILE__, __LINE__); return e != cudaSuccess; };
    using clk = std::chrono::steady_clock;
    do {
        if (h2d_bytes && (fail(cudaMallocHost(&hi, h2d_bytes)) || fail(cudaMalloc(&di, h2d_bytes)))) break;
        if (d2h_bytes && (fail(cudaMallocHost(&ho, d2h_bytes)) || fail(cudaMalloc(&dout, d2h_bytes)))) break;
        if (hi) memset(hi, 1, h2d_bytes);
        if (dout && fail(cudaMemsetAsync(dout, 1, d2h_bytes, ctx->main))) break;
        if (fail(cudaStreamSynchronize(ctx->main))) break;
        auto round = [&](bool in, bool out, int n) -> double {
            const auto t0 = clk::now();
            for (int r = 0; r < n; r++) {
                if (in && hi) cudaMemcpyAsync(di, hi, h2d_bytes, cudaMemcpyHostToDevice, ctx->copy_in);
                if (out && ho) cudaMemcpyAsync(ho, dout, d2h_bytes, cudaMemcpyDeviceToHost, ctx->copy_out);
            }
            cudaStreamSynchronize(ctx->copy_in);
            cudaStreamSynchronize(ctx->copy_out);
            return std::chrono::duration<double, std::milli>(clk::now() - t0).count() / n;
        };
        round(true, true, 1);                                                 // warm-up
        if (ms_per_round) *ms_per_round = round(true, true, reps);
        if (h2d_alone_ms) *h2d_alone_ms = round(true, false, reps);
        if (d2h_alone_ms) *d2h_alone_ms = round(false, true, reps);
        fail(cudaGetLastError());
    } while (0);
    if (hi) cudaFreeHost(hi);
    if (ho) cudaFreeHost(ho);
    if (di) cudaFree(di);
    if (dout) cudaFree(dout);
    return st;
}

}  // extern "C"
