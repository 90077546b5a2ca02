// Microbenchmark: TMA tile-load throughput per SM against box geometry and element type (B200).
// One persistent CTA per SM re-loads boxes of an L2-resident tensor into shared memory, `depth` boxes in flight, no compute.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/tma_bw tools/ubench/tma_bw.cu && tools/ubench/tma_bw
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) tma_loop(const __grid_constant__ CUtensorMap tm, int box_bytes, int depth, int iters, int ncol_tiles, int nrow_tiles,
                                                   int inner_el, int box_rows, long long *cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar[8];
    if (threadIdx.x == 0) {
        for (int d = 0; d < depth; d++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[d])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    auto issue = [&](int k) {
        const int d = k % depth;
        const int tile = (blockIdx.x * 7 + k) % (ncol_tiles * nrow_tiles);
        const int c0 = (tile % ncol_tiles) * inner_el, r0 = (tile / ncol_tiles) * box_rows;   // element coordinates
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[d])), "r"(box_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                         smem_u32(smem + (size_t)d * box_bytes)),
                     "l"(&tm), "r"(c0), "r"(r0), "r"(smem_u32(&bar[d]))
                     : "memory");
    };
    const long long t0 = clock64();
    for (int k = 0; k < depth && k < iters; k++) issue(k);
    for (int k = 0; k < iters; k++) {
        const int d = k % depth;
        const unsigned parity = (k / depth) & 1;
        unsigned ok = 0;
        while (!ok) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar[d])), "r"(parity) : "memory");
        }
        if (k + depth < iters) issue(k + depth);
    }
    cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    void *fnp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fnp;
    const size_t row_bytes = 32768, rows = 1024;       // 32 MiB tensor: L2 resident
    void *buf; CK(cudaMalloc(&buf, row_bytes * rows)); CK(cudaMemset(buf, 1, row_bytes * rows));
    long long *cyc; CK(cudaMalloc(&cyc, sizeof(long long) * 1024));
    CK(cudaFuncSetAttribute(tma_loop, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    struct Ty { const char *name; CUtensorMapDataType t; int es; } types[] = {{"u8", CU_TENSOR_MAP_DATA_TYPE_UINT8, 1}, {"f32", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4}, {"f64", CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8}};
    const int inner_bytes[] = {64, 128, 256, 1024};
    const int box_rows[] = {16, 144, 256};
    const int depths[] = {1, 2, 4};
    printf("%-5s %6s %5s %5s | %9s %10s\n", "type", "innerB", "rows", "depth", "B/clk/SM", "GB/s chip");
    for (auto &ty : types)
        for (int ib : inner_bytes)
            for (int br : box_rows)
                for (int depth : depths) {
                    const int inner_el = ib / ty.es;
                    if (inner_el > 256) continue;
                    const int box_bytes = ib * br;
                    if ((size_t)box_bytes * depth > 190 * 1024) continue;
                    CUtensorMap tm;
                    cuuint64_t dims[2] = {row_bytes / ty.es, rows}, strides[1] = {row_bytes};
                    cuuint32_t box[2] = {(cuuint32_t)inner_el, (cuuint32_t)br}, es[2] = {1, 1};
                    CUresult r = enc(&tm, ty.t, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
                    const int iters = 400;
                    const int ncol = (int)(row_bytes / ib), nrow = (int)(rows / br);
                    struct L { static void run(const CUtensorMap &tm, int bb, int d, int it, int nc, int nr, int ie, int brr, long long *cy, int grid, int smem) { tma_loop<<<grid, 128, smem>>>(tm, bb, d, it, nc, nr, ie, brr, cy); } };
                    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                    L::run(tm, box_bytes, depth, 20, ncol, nrow, inner_el, br, cyc, p.multiProcessorCount, box_bytes * depth);
                    CK(cudaDeviceSynchronize());
                    cudaEventRecord(e0);
                    L::run(tm, box_bytes, depth, iters, ncol, nrow, inner_el, br, cyc, p.multiProcessorCount, box_bytes * depth);
                    cudaEventRecord(e1);
                    CK(cudaDeviceSynchronize());
                    float ms; cudaEventElapsedTime(&ms, e0, e1);
                    long long h[256]; CK(cudaMemcpy(h, cyc, sizeof(long long) * p.multiProcessorCount, cudaMemcpyDeviceToHost));
                    double avg = 0; for (int i = 0; i < p.multiProcessorCount; i++) avg += h[i]; avg /= p.multiProcessorCount;
                    printf("%-5s %6d %5d %5d | %9.2f %10.1f\n", ty.name, ib, br, depth, (double)box_bytes * iters / avg, (double)box_bytes * iters * p.multiProcessorCount / (ms * 1e6));
                }
    return 0;
}
