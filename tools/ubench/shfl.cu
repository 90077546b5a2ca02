// Microbenchmark: would warp-shuffle butterflies beat the shared-memory Stockham exchange for fp64 complex data?
// Moves the same 16-byte elements (a) through shared memory (STS.128 + LDS.128) and (b) through shfl_xor
// (4 x SHFL.32 per element), and reports bytes per clock per SM.  The north star sketches "warp-shuffle butterflies";
// this is why fft_core.cuh exchanges through shared memory instead.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int MODE>   // 0: shared memory, 1: shuffle
__global__ void __launch_bounds__(256, 2) xchg(double *out, int iters) {
    extern __shared__ double2 sm[];
    const int t = threadIdx.x;
    double2 v[8];
    for (int i = 0; i < 8; i++) { v[i].x = t + i; v[i].y = i * 0.5; }
    __syncthreads();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) sm[i * 256 + t] = v[i];
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; i++) { const double2 q = sm[i * 256 + (t ^ (1 + (it & 15)))]; v[i].x += q.x * 1e-9; v[i].y = q.y; }
            __syncwarp();
        } else {
            const int mask = 1 + (it & 15);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const double qx = __shfl_xor_sync(0xffffffffu, v[i].x, mask), qy = __shfl_xor_sync(0xffffffffu, v[i].y, mask);
                v[i].x += qx * 1e-9; v[i].y = qy;
            }
        }
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += v[i].x + v[i].y;
    out[blockIdx.x * 256 + t] = s;
}

template <int MODE> int run(const char *name, int sms, double *out, double ghz) {
    const int iters = 4000, grid = 2 * sms;
    cudaFuncSetAttribute(xchg<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    xchg<MODE><<<grid, 256, 32768>>>(out, iters);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    xchg<MODE><<<grid, 256, 32768>>>(out, iters);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double elems = (double)grid * 256 * 8 * iters;            // 16-byte elements exchanged
    const double bytes_per_clk_sm = elems * 16 / (ms * 1e-3) / (ghz * 1e9) / sms;
    printf("%-14s %8.3f ms  %6.1f G elements/s  %6.1f B of payload per clock per SM (shared memory counts write+read once)\n", name, ms,
           elems / ms / 1e6, bytes_per_clk_sm);
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    double *out; CK(cudaMalloc(&out, sizeof(double) * 2 * p.multiProcessorCount * 256));
    const double ghz = p.clockRate * 1e-6;
    if (run<0>("shared memory", p.multiProcessorCount, out, ghz)) return 1;
    if (run<1>("shfl_xor", p.multiProcessorCount, out, ghz)) return 1;
    return 0;
}
