// tma.cuh -- Tensor Memory Accelerator plumbing: tensor-map construction on the host (driver entry point looked up at
// run time, so the library does not link against libcuda) and the PTX wrappers the kernels use (mbarrier completion,
// cp.async.bulk.tensor tile loads).  SASS of the loads: UTMALDG.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace adsp {

// ---------------------------------------------------------------- host: tensor maps
// rank-`rank` tiled map over fp64 / fp32 elements; dims / strides innermost first (strides[i] in BYTES, for dims 1..rank-1),
// OOB elements of a box read as zero.  Returns false when the driver refuses (alignment, sizes) -- callers fall back.
bool tma_encode(CUtensorMap *map, bool f64, int rank, const void *base, const uint64_t *dims, const uint64_t *strides_bytes, const uint32_t *box);

// ---------------------------------------------------------------- device: mbarrier + tile loads
#if defined(__CUDACC__)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned tx_bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(tx_bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// makes the mbarrier initialisation visible to the async proxy (the TMA unit)
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// orders this thread's earlier generic-proxy accesses to shared memory before later async-proxy (TMA) ones
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *m) { asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory"); }

// box of a 2-D / 3-D tiled tensor map -> shared memory (dense, innermost dimension fastest); completion = box bytes on `bar`
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
#endif

}  // namespace adsp
