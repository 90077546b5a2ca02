// Microbenchmark: can the FP64 pipe and the shared-memory (LSU) pipe of a B200 SM run at full rate
// at the same time?  Decides whether DP|LSU phase overlap is worth engineering for.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void dp_work(double (&a)[8], double b, double c) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = fma(a[i], b, c);
}

// mode 0: all warps DFMA; 1: all warps LDS/STS.128; 2: even warps DFMA, odd warps smem; 3: every warp interleaves both
template <int MODE>
__global__ void __launch_bounds__(256, 2) pipes(double *out, int iters, long long *cycles) {
    extern __shared__ double2 sm[];
    const int t = threadIdx.x, w = t >> 5;
    double a[8];
    for (int i = 0; i < 8; i++) a[i] = 1.0 + t * 1e-9 + i;
    double2 v[4];
    for (int i = 0; i < 4; i++) { v[i].x = t; v[i].y = i; }
    const double b = 1.0000001, c = 1e-9;
    __syncthreads();
    const long long t0 = clock64();
    const bool do_dp = (MODE == 0) || (MODE == 3) || (MODE == 2 && (w & 1) == 0);
    const bool do_sm = (MODE == 1) || (MODE == 3) || (MODE == 2 && (w & 1) == 1);
    for (int it = 0; it < iters; it++) {
        if (do_dp) {
#pragma unroll
            for (int r = 0; r < 8; r++) dp_work(a, b, c);   // 64 DFMA per iteration
        }
        if (do_sm) {
#pragma unroll
            for (int r = 0; r < 4; r++) {                    // 4 x (STS.128 x4 + LDS.128 x4) per iteration = 512 B per thread
#pragma unroll
                for (int i = 0; i < 4; i++) sm[(i * 256 + t)] = v[i];
#pragma unroll
                for (int i = 0; i < 4; i++) { double2 q = sm[(i * 256 + ((t + 32) & 255))]; v[i].x += q.x; v[i].y = q.y; }
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 8; i++) s += a[i];
    for (int i = 0; i < 4; i++) s += v[i].x + v[i].y;
    out[blockIdx.x * 256 + t] = s;
    if (t == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE> int run(const char *name, int sms, double *out, long long *cyc) {
    const int iters = 2000, grid = 2 * sms;
    cudaFuncSetAttribute(pipes<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    pipes<MODE><<<grid, 256, 16384>>>(out, iters, cyc);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    pipes<MODE><<<grid, 256, 16384>>>(out, iters, cyc);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[1024]; CK(cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < grid; i++) avg += h[i]; avg /= grid;
    // per SM: 2 CTAs x 256 threads
    const double dp_threads = (MODE == 2) ? 256.0 : 512.0, sm_threads = (MODE == 2) ? 256.0 : 512.0;
    const double dfma = (MODE == 1) ? 0 : dp_threads * 64.0 * iters;          // per SM
    const double bytes = (MODE == 0) ? 0 : sm_threads * 512.0 * iters;        // per SM (st + ld)
    printf("%-34s %8.3f ms  cycles/CTA %10.0f  DFMA/clk/SM %6.1f (peak 64)  smem B/clk/SM %6.1f (peak 128)\n", name, ms, avg, dfma / avg, bytes / avg);
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    double *out; long long *cyc;
    CK(cudaMalloc(&out, sizeof(double) * 2 * p.multiProcessorCount * 256)); CK(cudaMalloc(&cyc, sizeof(long long) * 1024));
    run<0>("all warps DFMA", p.multiProcessorCount, out, cyc);
    run<1>("all warps LDS/STS.128", p.multiProcessorCount, out, cyc);
    run<2>("even warps DFMA, odd warps smem", p.multiProcessorCount, out, cyc);
    run<3>("every warp interleaves both", p.multiProcessorCount, out, cyc);
    return 0;
}
