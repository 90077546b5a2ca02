// How much does it cost just to launch the bench's CTAs?  ~147 000 CTAs of 128 threads with 36 KB of
// dynamic shared memory per step, in ~195 kernels over 3 streams (empty bodies).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(128, 4) empty(int *p) { extern __shared__ int s[]; if (p && threadIdx.x == 12345) p[0] = s[0]; }
__global__ void __launch_bounds__(128, 4) barriers(int *p, int nb) {
    extern __shared__ int s[];
    for (int i = 0; i < nb; i++) { s[threadIdx.x] = i; __syncthreads(); }
    if (p && threadIdx.x == 12345) p[0] = s[0];
}
int main() {
    cudaStream_t st[3]; for (auto &s : st) cudaStreamCreate(&s);
    cudaFuncSetAttribute(empty, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
    cudaFuncSetAttribute(barriers, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
    for (int mode = 0; mode < 3; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaDeviceSynchronize();
            cudaEventRecord(e0, st[0]);
            cudaEvent_t f; cudaEventCreate(&f); cudaEventRecord(f, st[0]);
            cudaStreamWaitEvent(st[1], f); cudaStreamWaitEvent(st[2], f);
            for (int k = 0; k < 195; k++) {
                if (mode == 0) empty<<<768, 128, 36864, st[k % 3]>>>(nullptr);
                else barriers<<<768, 128, 36864, st[k % 3]>>>(nullptr, mode == 1 ? 10 : 20);
            }
            cudaEvent_t j1, j2; cudaEventCreate(&j1); cudaEventCreate(&j2);
            cudaEventRecord(j1, st[1]); cudaEventRecord(j2, st[2]);
            cudaStreamWaitEvent(st[0], j1); cudaStreamWaitEvent(st[0], j2);
            cudaEventRecord(e1, st[0]);
            cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) printf("mode %d (%s): 195 kernels x 768 CTAs: %.3f ms  (%.1f ns per CTA chip-wide)\n", mode,
                            mode == 0 ? "empty" : (mode == 1 ? "10 barriers" : "20 barriers"), ms, ms * 1e6 / (195.0 * 768));
        }
    }
    return 0;
}
