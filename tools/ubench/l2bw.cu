// L2 / HBM bandwidth microbenchmark (read-modify-write of a buffer with 128-bit accesses).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void rmw(double2 *p, size_t n, int reps) {
    for (int r = 0; r < reps; r++)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            double2 v = __ldcg(p + i); v.x += 1.0; __stcg(p + i, v);
        }
}
__global__ void rd(const double2 *p, size_t n, int reps, double *out) {
    double acc = 0;
    for (int r = 0; r < reps; r++)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            double2 v = __ldcg(p + i); acc += v.x + v.y;
        }
    if (acc == 12345.678) out[0] = acc;
}
int main() {
    double2 *p; double *o; cudaMalloc(&p, 1ull << 30); cudaMalloc(&o, 8); cudaMemset(p, 0, 1ull << 30);
    for (size_t mb : {16, 32, 48, 64, 96, 128, 512}) {
        size_t n = (mb << 20) / 16; int reps = mb <= 128 ? 40 : 8;
        for (int mode = 0; mode < 2; mode++) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            if (mode == 0) rmw<<<148 * 8, 256>>>(p, n, 2); else rd<<<148 * 8, 256>>>(p, n, 2, o);
            cudaEventRecord(e0);
            if (mode == 0) rmw<<<148 * 8, 256>>>(p, n, reps); else rd<<<148 * 8, 256>>>(p, n, reps, o);
            cudaEventRecord(e1); cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double bytes = (double)n * 16 * reps * (mode == 0 ? 2 : 1);
            printf("%4zu MB %s: %8.1f GB/s\n", mb, mode == 0 ? "read+write" : "read only ", bytes / ms / 1e6);
        }
    }
    return 0;
}
