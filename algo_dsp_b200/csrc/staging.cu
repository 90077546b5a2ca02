// staging.cu -- pinned host staging for pageable caller memory (north star: "pinned host staging").
//
// The Go API hands the library plain []float64 slices (overlap_save.go:132-133 allocates the result with make); such
// memory is pageable, so cudaMemcpyAsync on it degrades to the runtime's synchronous bounce copy and nothing overlaps.
// Here pageable buffers are copied to/from the context's own pinned slots by a small pool of host threads (a single
// memcpy thread moves ~10 GB/s, PCIe Gen5 x16 needs ~50 GB/s per direction) while the DMA engines and the SMs work on
// the neighbouring chunks.  Memory that is already pinned/registered is DMA'd in place.
#include <algorithm>
#include <chrono>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "engine.cuh"
#include "tma.cuh"

namespace adsp {

// ------------------------------------------------------------------ TMA tensor maps (tma.cuh)
// cuTensorMapEncodeTiled lives in the driver library; it is looked up once through the runtime so that libalgodsp_cuda
// needs no link-time dependency on libcuda.
using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                   const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return (EncodeTiledFn)p;
    }();
    return fn;
}

bool tma_encode(CUtensorMap *map, bool f64, int rank, const void *base, const uint64_t *dims, const uint64_t *strides_bytes, const uint32_t *box) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || rank < 1 || rank > 5) return false;
    cuuint64_t d[5], s[5];
    cuuint32_t b[5], es[5];
    for (int i = 0; i < rank; i++) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; if (dims[i] == 0 || box[i] == 0 || box[i] > 256) return false; }
    for (int i = 0; i + 1 < rank; i++) { s[i] = strides_bytes[i]; if (s[i] % 16 != 0) return false; }
    if ((uintptr_t)base % 16 != 0) return false;
    // L2 promotion: the granule the TMA unit requests from L2.  Without promotion it asks sector by sector (32 B), which
    // caps a box of narrow rows at ~8 B/clk/SM on B200 (measured, profiles/r02_d_mrp_tensor_ab.log); 128-byte requests
    // match the cache line.  ADSP_TMA_L2PROMO = 0 none, 1 64 B, 2 128 B (default), 3 256 B.
    static const long long promo = env_ll("ADSP_TMA_L2PROMO", 2);
    const CUtensorMapL2promotion pr = promo <= 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                      : promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    const CUresult r = fn(map, f64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void *>(base), d, s, b, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// ------------------------------------------------------------------ copy-thread pool
StagePool::StagePool(int nthreads) {
    if (nthreads < 1) nthreads = 1;
    for (int i = 0; i < nthreads; i++) workers_.emplace_back([this] { run(); });
}

StagePool::~StagePool() {
    {
        std::lock_guard<std::mutex> lk(m_);
        stop_ = true;
    }
    cv_.notify_all();
    for (auto &t : workers_) t.join();
}

// Bulk copy for staging.  Both directions stream: a pinned slot is written by the CPU and then read only by the DMA
// engine, the caller's output is far larger than the caches -- so the stores bypass the cache (no read-for-ownership, a
// third less DRAM traffic than memcpy's cached stores below its non-temporal threshold) when the CPU has AVX2.
#if defined(__x86_64__)
__attribute__((target("avx2"))) static void stream_copy_avx2(char *dst, const char *src, size_t n) {
    const size_t head = (32 - ((uintptr_t)dst & 31)) & 31;
    if (head) { const size_t h = head < n ? head : n; memcpy(dst, src, h); dst += h; src += h; n -= h; }
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256((const __m256i *)(src + i)), b = _mm256_loadu_si256((const __m256i *)(src + i + 32));
        const __m256i c = _mm256_loadu_si256((const __m256i *)(src + i + 64)), d = _mm256_loadu_si256((const __m256i *)(src + i + 96));
        _mm256_stream_si256((__m256i *)(dst + i), a); _mm256_stream_si256((__m256i *)(dst + i + 32), b);
        _mm256_stream_si256((__m256i *)(dst + i + 64), c); _mm256_stream_si256((__m256i *)(dst + i + 96), d);
    }
    for (; i + 32 <= n; i += 32) _mm256_stream_si256((__m256i *)(dst + i), _mm256_loadu_si256((const __m256i *)(src + i)));
    _mm_sfence();
    if (i < n) memcpy(dst + i, src + i, n - i);
}
// Stage-OUT variant (pinned slot -> caller memory): the same streaming copy, and every source line is flushed from the CPU
// caches once it has been read.  The slot is the target of the NEXT device-to-host DMA; lines the copy threads left in
// their caches make that DMA 6x slower (measured on the B200 host, 4.6 MB: 0.61 ms against 0.09 ms into memory no CPU has
// touched -- every inbound write has to invalidate the cached copies first).
__attribute__((target("avx2,clflushopt"))) static void stream_copy_flush_avx2(char *dst, const char *src, size_t n) {
    const size_t head = (64 - ((uintptr_t)src & 63)) & 63;          // align the SOURCE to cache lines here (it is the one flushed)
    if (head) { const size_t h = head < n ? head : n; memcpy(dst, src, h); _mm_clflushopt((void *)src); dst += h; src += h; n -= h; }
    size_t i = 0;
    const bool dst_al = (((uintptr_t)dst) & 31) == 0;
    for (; i + 64 <= n; i += 64) {
        const __m256i a = _mm256_load_si256((const __m256i *)(src + i)), b = _mm256_load_si256((const __m256i *)(src + i + 32));
        if (dst_al) { _mm256_stream_si256((__m256i *)(dst + i), a); _mm256_stream_si256((__m256i *)(dst + i + 32), b); }
        else { _mm256_storeu_si256((__m256i *)(dst + i), a); _mm256_storeu_si256((__m256i *)(dst + i + 32), b); }
        _mm_clflushopt((void *)(src + i));
    }
    if (i < n) { memcpy(dst + i, src + i, n - i); _mm_clflushopt((void *)(src + i)); }
    _mm_sfence();
}
static bool cpu_has_clflushopt() {
    unsigned a = 7, b = 0, c = 0, d = 0;   // leaf 7, sub-leaf 0
    __asm__ __volatile__("cpuid" : "+a"(a), "=b"(b), "+c"(c), "=d"(d));
    return (b >> 23) & 1;                                           // CPUID.(EAX=7,ECX=0):EBX bit 23
}
static const bool g_use_stream_copy = __builtin_cpu_supports("avx2") && env_ll("ADSP_STAGE_NT", 1) != 0;
static const bool g_use_flush = __builtin_cpu_supports("avx2") && cpu_has_clflushopt() && env_ll("ADSP_STAGE_FLUSH", 1) != 0;
#else
static const bool g_use_stream_copy = false, g_use_flush = false;
static void stream_copy_avx2(char *, const char *, size_t) {}
static void stream_copy_flush_avx2(char *, const char *, size_t) {}
#endif

static inline void bulk_copy(char *dst, const char *src, size_t n, bool flush_src) {
    if (flush_src && g_use_flush && n >= 64) stream_copy_flush_avx2(dst, src, n);
    else if (g_use_stream_copy && n >= 4096) stream_copy_avx2(dst, src, n);
    else memcpy(dst, src, n);
}

void StagePool::exec(Slice &s) {
    if (s.dpitch == s.width && s.spitch == s.width) bulk_copy(s.dst, s.src, s.width * s.rows, s.flush_src);
    else for (size_t r = 0; r < s.rows; r++) bulk_copy(s.dst + r * s.dpitch, s.src + r * s.spitch, s.width, s.flush_src);
    if (s.job->remaining.fetch_sub(1, std::memory_order_acq_rel) == 1) {
        std::lock_guard<std::mutex> lk(s.job->m);
        s.job->cv.notify_all();
    }
}

void StagePool::run() {
    for (;;) {
        Slice s;
        {
            std::unique_lock<std::mutex> lk(m_);
            cv_.wait(lk, [this] { return stop_ || !q_.empty(); });
            if (q_.empty()) return;   // stop requested and nothing left
            s = std::move(q_.front());
            q_.pop_front();
        }
        exec(s);
    }
}

StagePool::Ticket StagePool::copy2d_async(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows, bool flush_src) {
    Ticket job = std::make_shared<Job>();
    if (rows == 0 || width == 0) return job;
    // slices of about 256 KiB, at most 8 per worker per job (tiny copies must not pay for the queue, large ones balance)
    const size_t total = width * rows;
    size_t nsl = std::min<size_t>((total + (256u << 10) - 1) >> 18, (size_t)workers_.size() * 8);
    if (nsl < 1) nsl = 1;
    std::vector<Slice> sl;
    if (rows >= nsl) {                       // split by rows
        const size_t per = (rows + nsl - 1) / nsl;
        for (size_t r0 = 0; r0 < rows; r0 += per)
            sl.push_back({(char *)dst + r0 * dpitch, (const char *)src + r0 * spitch, dpitch, spitch, width, std::min(per, rows - r0), job, flush_src});
    } else {                                 // few long rows: split every row by columns (64-byte aligned cuts)
        const size_t per_row = (nsl + rows - 1) / rows;
        size_t seg = (width + per_row - 1) / per_row;
        seg = (seg + 63) & ~(size_t)63;
        for (size_t r = 0; r < rows; r++)
            for (size_t c0 = 0; c0 < width; c0 += seg)
                sl.push_back({(char *)dst + r * dpitch + c0, (const char *)src + r * spitch + c0, 0, 0, std::min(seg, width - c0), 1, job, flush_src});
        for (auto &s : sl) { s.dpitch = s.width; s.spitch = s.width; }
    }
    job->remaining.store((long long)sl.size(), std::memory_order_release);
    {
        std::lock_guard<std::mutex> lk(m_);
        for (auto &s : sl) q_.push_back(std::move(s));
    }
    if (sl.size() >= workers_.size()) cv_.notify_all();
    else for (size_t i = 0; i < sl.size(); i++) cv_.notify_one();
    return job;
}

// The waiting thread works too: it takes slices (of any job) off the queue until its own job is complete, so a small
// copy does not depend on how fast the sleeping workers wake up.
void StagePool::wait(const Ticket &t) {
    if (!t) return;
    while (t->remaining.load(std::memory_order_acquire) > 0) {
        Slice s;
        bool have = false;
        {
            std::lock_guard<std::mutex> lk(m_);
            if (!q_.empty()) { s = std::move(q_.front()); q_.pop_front(); have = true; }
        }
        if (have) { exec(s); continue; }
        std::unique_lock<std::mutex> lk(t->m);
        t->cv.wait(lk, [&] { return t->remaining.load(std::memory_order_acquire) <= 0; });
    }
}

bool host_ptr_is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type != cudaMemoryTypeUnregistered;
}

StagePool *stage_pool(adsp_ctx *ctx) {
    if (!ctx->pool) {
        const unsigned hw = std::thread::hardware_concurrency();
        long long n = env_ll("ADSP_STAGE_THREADS", 0);
        if (n <= 0) n = std::max(1u, std::min(12u, hw ? (hw * 3) / 4 : 4u));   // the calling thread copies too while it waits
        ctx->pool.reset(new StagePool((int)n));
    }
    return ctx->pool.get();
}

// ------------------------------------------------------------------ staged transfers on ctx->main
// Pageable transfers below this size go through the runtime's own bounce path (one call, no pool wake-up).
static constexpr size_t kStageMin = 256u << 10;

static size_t stage_chunk_bytes() {
    long long mb = env_ll("ADSP_STAGE_CHUNK_MB", 8);
    if (mb < 1) mb = 1;
    return (size_t)mb << 20;
}

adsp_status upload2d(adsp_ctx *ctx, void *dst_dev, size_t dpitch, const void *src_host, size_t spitch, size_t width, size_t rows) {
    if (rows == 0 || width == 0) return ADSP_OK;
    const bool dense = (rows == 1) || (dpitch == width && spitch == width);
    if (host_ptr_is_pinned(src_host) || width * rows < kStageMin) {
        if (dense) ADSP_CUDA(cudaMemcpyAsync(dst_dev, src_host, width * rows, cudaMemcpyHostToDevice, ctx->main));
        else ADSP_CUDA(cudaMemcpy2DAsync(dst_dev, dpitch, src_host, spitch, width, rows, cudaMemcpyHostToDevice, ctx->main));
        return ADSP_OK;
    }
    StagePool *pool = stage_pool(ctx);
    const size_t chunk = stage_chunk_bytes();
    for (int s = 0; s < kPipeSlots; s++) ADSP_TRY(ctx->h_in[s].reserve(chunk));
    ctx->staged_bytes_in += width * rows;
    int i = 0;
    auto send = [&](char *d, const char *h, size_t dp, size_t sp, size_t w, size_t r) -> adsp_status {
        const int s = i++ % kPipeSlots;
        ADSP_CUDA(cudaEventSynchronize(ctx->ev_in[s]));                  // the DMA that last read this slot is done
        const auto c0 = std::chrono::steady_clock::now();
        pool->copy2d(ctx->h_in[s].p, w, h, sp, w, r);                    // user memory -> pinned slot (dense rows)
        if (ctx->host_profile) ctx->host_prof_ms[6] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - c0).count();
        if (r == 1) ADSP_CUDA(cudaMemcpyAsync(d, ctx->h_in[s].p, w, cudaMemcpyHostToDevice, ctx->main));
        else ADSP_CUDA(cudaMemcpy2DAsync(d, dp, ctx->h_in[s].p, w, w, r, cudaMemcpyHostToDevice, ctx->main));
        ADSP_CUDA(cudaEventRecord(ctx->ev_in[s], ctx->main));
        return ADSP_OK;
    };
    if (width >= chunk || rows == 1) {
        for (size_t r = 0; r < rows; r++)
            for (size_t c0 = 0; c0 < width; c0 += chunk)
                ADSP_TRY(send((char *)dst_dev + r * dpitch + c0, (const char *)src_host + r * spitch + c0, 0, 0, std::min(chunk, width - c0), 1));
    } else {
        const size_t per = std::max<size_t>(1, chunk / width);
        for (size_t r0 = 0; r0 < rows; r0 += per)
            ADSP_TRY(send((char *)dst_dev + r0 * dpitch, (const char *)src_host + r0 * spitch, dpitch, spitch, width, std::min(per, rows - r0)));
    }
    return ADSP_OK;
}

adsp_status download2d(adsp_ctx *ctx, void *dst_host, size_t dpitch, const void *src_dev, size_t spitch, size_t width, size_t rows) {
    if (rows == 0 || width == 0) return ADSP_OK;
    const bool dense = (rows == 1) || (dpitch == width && spitch == width);
    if (host_ptr_is_pinned(dst_host) || width * rows < kStageMin) {
        if (dense) ADSP_CUDA(cudaMemcpyAsync(dst_host, src_dev, width * rows, cudaMemcpyDeviceToHost, ctx->main));
        else ADSP_CUDA(cudaMemcpy2DAsync(dst_host, dpitch, src_dev, spitch, width, rows, cudaMemcpyDeviceToHost, ctx->main));
        ADSP_CUDA(cudaStreamSynchronize(ctx->main));
        return ADSP_OK;
    }
    StagePool *pool = stage_pool(ctx);
    const size_t chunk = stage_chunk_bytes();
    for (int s = 0; s < kPipeSlots; s++) ADSP_TRY(ctx->h_out[s].reserve(chunk));
    ctx->staged_bytes_out += width * rows;
    struct Piece { char *h; const char *d; size_t hp, dp, w, r; };
    std::vector<Piece> pieces;
    if (width >= chunk || rows == 1) {
        for (size_t r = 0; r < rows; r++)
            for (size_t c0 = 0; c0 < width; c0 += chunk)
                pieces.push_back({(char *)dst_host + r * dpitch + c0, (const char *)src_dev + r * spitch + c0, 0, 0, std::min(chunk, width - c0), 1});
    } else {
        const size_t per = std::max<size_t>(1, chunk / width);
        for (size_t r0 = 0; r0 < rows; r0 += per)
            pieces.push_back({(char *)dst_host + r0 * dpitch, (const char *)src_dev + r0 * spitch, dpitch, spitch, width, std::min(per, rows - r0)});
    }
    StagePool::Ticket tk[kPipeSlots];
    struct Drain {   // no copy into caller memory may outlive this call, error paths included
        StagePool *pool; StagePool::Ticket *t;
        ~Drain() { for (int i = 0; i < kPipeSlots; i++) pool->wait(t[i]); }
    } drain{pool, tk};
    const size_t np = pieces.size();
    for (size_t i = 0; i <= np; i++) {
        if (i < np) {
            const int s = (int)(i % kPipeSlots);
            pool->wait(tk[s]);                                      // the host copy that last read this slot is done
            const Piece &pc = pieces[i];
            if (pc.r == 1) ADSP_CUDA(cudaMemcpyAsync(ctx->h_out[s].p, pc.d, pc.w, cudaMemcpyDeviceToHost, ctx->main));
            else ADSP_CUDA(cudaMemcpy2DAsync(ctx->h_out[s].p, pc.w, pc.d, pc.dp, pc.w, pc.r, cudaMemcpyDeviceToHost, ctx->main));
            ADSP_CUDA(cudaEventRecord(ctx->ev_out[s], ctx->main));
        }
        if (i >= 1) {
            const size_t j = i - 1;
            const int s = (int)(j % kPipeSlots);
            const auto c0 = std::chrono::steady_clock::now();
            ADSP_CUDA(cudaEventSynchronize(ctx->ev_out[s]));
            if (ctx->host_profile) ctx->host_prof_ms[7] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - c0).count();
            const Piece &pc = pieces[j];
            tk[s] = pool->copy2d_async(pc.h, pc.r == 1 ? pc.w : pc.hp, ctx->h_out[s].p, pc.w, pc.w, pc.r, true);
        }
    }
    const auto c1 = std::chrono::steady_clock::now();
    for (int s = 0; s < kPipeSlots; s++) pool->wait(tk[s]);
    if (ctx->host_profile) ctx->host_prof_ms[8] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - c1).count();
    return ADSP_OK;
}

adsp_status upload(adsp_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes) {
    return upload2d(ctx, dst_dev, bytes, src_host, bytes, bytes, 1);
}
adsp_status download(adsp_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes) {
    return download2d(ctx, dst_host, bytes, src_dev, bytes, bytes, 1);
}

}  // namespace adsp
