// api.cu -- the C ABI of libalgodsp_cuda (include/algodsp_cuda.h): context, plans, one-shot
// functions.  Host-side bookkeeping only; all arithmetic runs in the sm_100a kernels of
// conv_kernels.cuh / aux_kernels.cuh.  There is no CPU fallback anywhere in this file.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <thread>

#include "aux_kernels.cuh"
#include "engine.cuh"

namespace adsp {

static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }

adsp_status cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    g_last_error = buf;
    cudaGetLastError();  // clear non-sticky error state
    return (e == cudaErrorMemoryAllocation) ? ADSP_ERR_OOM : ADSP_ERR_CUDA;
}

// Bumped whenever a DevBuf is freed or re-allocated: captured CUDA graphs bake device pointers of the shared scratch
// buffers in, so a graph captured under an older generation must not be replayed (adsp_plan_process_device).
std::atomic<uint64_t> g_alloc_generation{1};

adsp_status DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return ADSP_OK;
    g_alloc_generation.fetch_add(1, std::memory_order_relaxed);
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    const size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); if (e != cudaSuccess) { p = nullptr; return cuda_fail(e, "cudaMalloc", __FILE__, __LINE__); } cap = bytes; return ADSP_OK; }
    cap = want;
    return ADSP_OK;
}
void DevBuf::release() { if (p) { g_alloc_generation.fetch_add(1, std::memory_order_relaxed); cudaFree(p); } p = nullptr; cap = 0; }

adsp_status PinnedBuf::reserve(size_t bytes) {
    if (bytes <= cap) return ADSP_OK;
    if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
    cudaError_t e = cudaMallocHost(&p, bytes);
    if (e != cudaSuccess) { p = nullptr; return cuda_fail(e, "cudaMallocHost", __FILE__, __LINE__); }
    cap = bytes;
    return ADSP_OK;
}
void PinnedBuf::release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }

static long long next_pow2_ll(long long n) {
    if (n <= 1) return 1;
    long long p = 1;
    while (p < n) p *= 2;
    return p;
}

long long env_ll(const char *name, long long dflt) {
    const char *v = getenv(name);
    if (!v || !*v) return dflt;
    return atoll(v);
}

// Internal transform size for K taps.  The API only fixes results, so the GPU is free to use a
// larger transform than the reference's nextPow2(2K): a longer block wastes less of each
// transform on the K-1 discarded samples.
// geometry for a given transform length N and (partition) kernel length Kp
static int odd_part(long long N) { while (N > 1 && (N & 1) == 0) N >>= 1; return (int)N; }

// transform lengths the engine has kernels for: powers of two 256 .. 2^22, and P * 2^k with P in {3,5,7,9},
// k = 12 .. 17, as (16*M) x N2 four-step transforms (conv_kernels_mr.cuh): M = P with N2 = 2^(k-4) up to k = 15,
// M = 2P with N2 = 2^(k-5) for k = 16, 17 (keeps the rows at 2048 / 4096 points)
bool fft_size_supported(long long N) {
    if (N < 256) return false;
    const int P = odd_part(N);
    if (P == 1) return N <= (1LL << 22);
    if (P != 3 && P != 5 && P != 7 && P != 9) return false;
    const long long pow2 = N / P;
    return pow2 >= (1LL << 12) && pow2 <= (1LL << 17);
}

FftChoice make_choice(long long Kp, long long N) {
    FftChoice c;
    c.part_len = Kp;
    c.N = N;
    int lg = 0;
    while ((1LL << lg) < N) lg++;
    c.lgN = lg;
    c.P = odd_part(N);
    if (c.P > 1) {
        const long long pow2 = N / c.P;   // 2^17 needs M = 2P (rows stop at 4096 points); 2^16 prefers it (2048-point rows)
        c.M = (pow2 >= (1LL << 17) || (pow2 == (1LL << 16) && env_ll("ADSP_MR_NO_2P", 0) == 0)) ? 2 * c.P : c.P;
        c.N1 = 16 * c.M; c.N2 = (int)(N / c.N1); c.lgN = 0;
    }
    else if (N <= 4096) { c.N1 = 1; c.N2 = (int)N; }
    else {
        // rows of 2048 points run as 128-thread CTAs (4 per SM); 2^20-point transforms keep 4096-point
        // rows so that the column transforms stay at N1 = 256
        // 2^16 = 256 x 256 needs only four radix-16 passes (measured +6 % over 32 x 2048)
        long long n2 = env_ll("ADSP_FFT_N2", N >= (1LL << 20) ? 4096 : (N == (1LL << 16) ? 256 : 2048));
        if (n2 > N / 16) n2 = N / 16;
        if (n2 < 256) n2 = 256;
        while (N / n2 > 1024) n2 *= 2;
        c.N2 = (int)n2;
        c.N1 = (int)(N / n2);
    }
    // discard count rounded up to 32 samples so block starts stay 256-byte aligned
    long long D = ((Kp - 1 + 31) / 32) * 32;
    if (D >= N) D = Kp - 1;
    c.D = D;
    c.S = N - D;
    return c;
}

FftChoice choose_fft(long long K) {
    // Preferred largest transform 2^20 (N1 = 256 columns, the fastest kernels).  A kernel that does not fit half
    // of it used to be split into partitions of 2^19 taps; measured on B200 one 2^22-point transform
    // (N1 = 1024) is 2.6x faster for 2^20 taps (48.7 vs 18.9 Gsamples/s), so long kernels take N = 2^22 and
    // are only partitioned beyond 2^21 taps.  ADSP_MAX_FFT forces the old single-size rule (tests, tuning).
    const long long forced_max = env_ll("ADSP_MAX_FFT", 0);
    const long long NMAX = forced_max > 0 ? forced_max : (1LL << 20);
    const long long NBIG = 1LL << 22;
    long long Kp = K;
    int parts = 1;
    long long N;
    if (K - 1 <= NMAX / 2 || forced_max > 0) {
        if (K - 1 > NMAX / 2) {
            parts = (int)((K + NMAX / 2 - 1) / (NMAX / 2));
            Kp = (K + parts - 1) / parts;
        }
        N = next_pow2_ll(8 * Kp);
        if (N < 256) N = 256;
        if (N > NMAX) N = NMAX;
        // up to ~1600 taps the single-kernel 4096-point transform beats the smallest four-step ones although less of it
        // is output (measured, 256 ch x 2^20: K=1000 150 vs 121, K=1500 130 vs 122, K=2000 107 vs 119 Gsamples/s)
        if (N > 4096 && Kp <= 1600 && NMAX >= 4096) N = 4096;
    } else {
        parts = (int)((K + NBIG / 2 - 1) / (NBIG / 2));
        Kp = (K + parts - 1) / parts;
        N = NBIG;
    }
    const long long forced = env_ll("ADSP_FFT_N", 0);
    if (forced > 0) N = forced;
    while (N < 2 * Kp && N < (1LL << 22)) N *= 2;
    FftChoice c = make_choice(Kp, N);
    c.parts = parts;
    return c;
}

struct Segment { long long N; long long off; long long len; bool no_discard; };

// relative cost of one transform of length N (points x passes; the mixed-radix column kernels idle some lanes)
static double fft_cost(long long N) {
    int lg = 0;
    while ((1LL << lg) < N) lg++;
    return (double)N * (lg + 4) * (odd_part(N) > 1 ? 1.1 : 1.0);
}

// Cover `out_len` output samples of a K-tap convolution.  Two shapes compete on cost:
//  (A) overlap-save: full blocks of the plan's largest transform, then the cheapest one or two smaller
//      transforms for the remainder (a 96k-tap IR on a 480k-sample signal: 2^19 + 2^18 instead of 2^20);
//  (B) when the whole result fits one transform: a single zero-padded block, nothing discarded
//      (the same signal: 9 * 2^16 = 589 824 points).
static std::vector<Segment> plan_segments(long long out_len, long long K, const FftChoice &big) {
    std::vector<Segment> segs;
    const long long full = out_len / big.S;
    const long long rem = out_len - full * big.S;
    const bool mixed = env_ll("ADSP_FFT_N", 0) <= 0 && env_ll("ADSP_NO_MIXED", 0) <= 0;
    if (!mixed) {
        if (full > 0) segs.push_back({big.N, 0, full * big.S, false});
        if (rem > 0) segs.push_back({big.N, full * big.S, rem, false});
        return segs;
    }
    const bool odd_ok = env_ll("ADSP_NO_ODD", 0) <= 0;
    std::vector<long long> cand;   // candidate lengths <= big.N, ascending
    for (long long b = 256; b <= big.N; b <<= 1) {
        cand.push_back(b);
        if (odd_ok)
            for (long long N : {b / 8 * 9, b / 4 * 5, b / 2 * 3, b / 4 * 7})
                if (N < big.N && fft_size_supported(N)) cand.push_back(N);
    }
    std::sort(cand.begin(), cand.end());
    // (A) remainder cover
    double cost_a = (double)full * fft_cost(big.N);
    long long ba = 0, bb = 0;
    if (rem > 0) {
        const long long nmin = std::max<long long>(2 * K, 256);
        double best = fft_cost(big.N);
        ba = big.N;
        for (size_t ia = cand.size(); ia-- > 0;) {
            const long long Na = cand[ia];
            if (Na < nmin) break;
            const long long Sa = make_choice(K, Na).S;
            if (Sa >= rem) { if (fft_cost(Na) < best) { best = fft_cost(Na); ba = Na; bb = 0; } continue; }
            for (size_t ib = ia + 1; ib-- > 0;) {
                const long long Nb = cand[ib];
                if (Nb < nmin) break;
                if (Sa + make_choice(K, Nb).S < rem) break;
                if (fft_cost(Na) + fft_cost(Nb) < best) { best = fft_cost(Na) + fft_cost(Nb); ba = Na; bb = Nb; }
            }
        }
        cost_a += best;
    }
    // (B) single zero-padded block
    for (long long N : cand)
        if (N >= out_len) {
            if (fft_cost(N) <= cost_a) { segs.push_back({N, 0, out_len, true}); return segs; }
            break;
        }
    if (full > 0) segs.push_back({big.N, 0, full * big.S, false});
    if (rem > 0) {
        const long long off = full * big.S;
        const long long la = std::min(rem, make_choice(K, ba).S);
        segs.push_back({ba, off, la, false});
        if (bb > 0 && rem > la) segs.push_back({bb, off + la, rem - la, false});
    }
    return segs;
}

// taps as a kernel parameter, NCH chunks of 64 (aux_kernels.cuh); h_b: the caller's host copy of the taps, if it has one
template <typename T, int NCH>
static adsp_status direct_ctapsn(adsp_ctx *ctx, const T *d_a, long long n, long long a_stride, const T *d_b, const T *h_b, long long m, T *d_out,
                                 long long out_stride, long long tiles, unsigned grid, bool exact) {
    DirectTapsN<T, NCH> taps;
    if (h_b) memcpy(taps.v, h_b, (size_t)m * sizeof(T));
    else {
        ADSP_CUDA(cudaMemcpyAsync(taps.v, d_b, (size_t)m * sizeof(T), cudaMemcpyDeviceToHost, ctx->main));
        ADSP_CUDA(cudaStreamSynchronize(ctx->main));
    }
    for (long long i = m; i < NCH * DIRECT_MC; i++) taps.v[i] = (T)0;
    LaunchTimer lt(ctx, ctx->main, KK_DIRECT);
    if (exact) direct_conv_ctapsn_kernel<T, false, NCH><<<grid, DIRECT_THREADS, 0, ctx->main>>>(d_a, n, a_stride, taps, (int)m, d_out, out_stride, tiles);
    else direct_conv_ctapsn_kernel<T, true, NCH><<<grid, DIRECT_THREADS, 0, ctx->main>>>(d_a, n, a_stride, taps, (int)m, d_out, out_stride, tiles);
    count_launch(ctx);
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}

template <typename T>
adsp_status direct_device(adsp_ctx *ctx, const T *d_a, long long n, long long a_stride, const T *d_b, long long m,
                          long long b_stride, long long batch, T *d_out, long long out_stride, const T *h_b) {
    const long long out_len = n + m - 1;
    static const bool exact = env_ll("ADSP_DIRECT_EXACT", 0) != 0;
    const long long tiles = (out_len + DIRECT_TILE - 1) / DIRECT_TILE;
    const long long grid = tiles * batch;
    if (grid <= 0) return ADSP_OK;
    if (grid > 0x7fffffffLL) { set_error("direct: grid too large"); return ADSP_ERR_INVALID_ARG; }
    // at most 64 taps shared by every channel (the auto-select case): taps as a kernel parameter (constant-bank operands)
    static const bool ctaps = env_ll("ADSP_DIRECT_CTAPS", 1) != 0;
    if (ctaps && m <= DIRECT_MC && (b_stride == 0 || batch == 1)) {
        DirectTaps<T> taps;
        for (int i = 0; i < DIRECT_MC; i++) taps.v[i] = (T)0;
        if (h_b) memcpy(taps.v, h_b, (size_t)m * sizeof(T));
        else {
            ADSP_CUDA(cudaMemcpyAsync(taps.v, d_b, (size_t)m * sizeof(T), cudaMemcpyDeviceToHost, ctx->main));
            ADSP_CUDA(cudaStreamSynchronize(ctx->main));
        }
        LaunchTimer lt(ctx, ctx->main, KK_DIRECT);
        const unsigned g = (unsigned)grid;
        if (m == DIRECT_MC) {
            if (exact) direct_conv_ctaps_kernel<T, false, false><<<g, DIRECT_THREADS, 0, ctx->main>>>(d_a, n, a_stride, taps, (int)m, d_out, out_stride, tiles);
            else direct_conv_ctaps_kernel<T, true, false><<<g, DIRECT_THREADS, 0, ctx->main>>>(d_a, n, a_stride, taps, (int)m, d_out, out_stride, tiles);
        } else {
            if (exact) direct_conv_ctaps_kernel<T, false, true><<<g, DIRECT_THREADS, 0, ctx->main>>>(d_a, n, a_stride, taps, (int)m, d_out, out_stride, tiles);
            else direct_conv_ctaps_kernel<T, true, true><<<g, DIRECT_THREADS, 0, ctx->main>>>(d_a, n, a_stride, taps, (int)m, d_out, out_stride, tiles);
        }
        count_launch(ctx);
        ADSP_CUDA(cudaGetLastError());
        return ADSP_OK;
    }
    if (ctaps && m <= 16 * DIRECT_MC && (b_stride == 0 || batch == 1)) {
        if (m <= 4 * DIRECT_MC) return direct_ctapsn<T, 4>(ctx, d_a, n, a_stride, d_b, h_b, m, d_out, out_stride, tiles, (unsigned)grid, exact);
        return direct_ctapsn<T, 16>(ctx, d_a, n, a_stride, d_b, h_b, m, d_out, out_stride, tiles, (unsigned)grid, exact);
    }
    LaunchTimer lt(ctx, ctx->main, KK_DIRECT);
    if (exact)
        direct_conv_kernel<T, false><<<(unsigned)grid, DIRECT_THREADS, 0, ctx->main>>>(d_a, n, a_stride, d_b, m, b_stride, d_out, out_stride, tiles);
    else
        direct_conv_kernel<T, true><<<(unsigned)grid, DIRECT_THREADS, 0, ctx->main>>>(d_a, n, a_stride, d_b, m, b_stride, d_out, out_stride, tiles);
    count_launch(ctx);
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}
template adsp_status direct_device<double>(adsp_ctx *, const double *, long long, long long, const double *, long long, long long, long long, double *, long long, const double *);
template adsp_status direct_device<float>(adsp_ctx *, const float *, long long, long long, const float *, long long, long long, long long, float *, long long, const float *);

// FindPeak on device vectors; results land in d_v[batch], d_i[batch]
template <typename T>
static adsp_status peak_device(adsp_ctx *ctx, const T *d_x, long long len, long long stride, long long batch, T *d_v,
                               long long *d_i) {
    if (batch <= 0) return ADSP_OK;
    int nparts = (int)std::min<long long>(64, (len + 256 * 8 - 1) / (256 * 8));
    if (nparts < 1) nparts = 1;
    const size_t need = (size_t)batch * nparts * (sizeof(T) + sizeof(long long));
    ADSP_TRY(ctx->d_small.reserve(need + 64));
    long long *pi = (long long *)ctx->d_small.p;
    T *pv = (T *)((char *)ctx->d_small.p + (size_t)batch * nparts * sizeof(long long));
    for (long long b0 = 0; b0 < batch; b0 += 65535) {   // grid.y is limited to 65535 rows per launch
        const long long nb = std::min<long long>(65535, batch - b0);
        dim3 g1((unsigned)nparts, (unsigned)nb);
        peak_partial_kernel<T><<<g1, 256, 0, ctx->main>>>(d_x + b0 * stride, len, stride, pv + b0 * nparts, pi + b0 * nparts);
        count_launch(ctx);
    }
    peak_final_kernel<T><<<(unsigned)batch, 32, 0, ctx->main>>>(d_x, stride, pv, pi, nparts, batch, d_v, d_i);
    count_launch(ctx);
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}

}  // namespace adsp

using namespace adsp;

// ================================================================ plans
enum PlanKind { PLAN_OLS = 0, PLAN_OLA = 1, PLAN_PART = 2, PLAN_STREAM = 3 };

struct PartStage { int part_size; int count; int start_pos; };

struct adsp_plan {
    adsp_ctx *ctx = nullptr;
    PlanKind kind = PLAN_OLS;
    adsp_precision prec = ADSP_F64;
    long long K = 0;
    long long ref_fft = 0, ref_step = 0, ref_block = 0;  // what the Go getters report
    FftChoice ch;
    std::vector<FftConv<double>> fc64;   // one per IR partition at the main transform length
    std::vector<FftConv<float>> fc32;
    DevBuf d_kernel;                     // the IR stays on the device so other transform lengths can be built lazily
    std::map<long long, FftConv<double>> extra64;   // engines for remainder blocks, keyed by N
    std::map<long long, FftConv<float>> extra32;
    // partitioned (streaming) state
    int latency = 0, min_order = 0, max_order = 0;
    std::vector<PartStage> stages;
    DevBuf hist[2];      // last K-1+latency input samples, ping-pong
    int hist_cur = 0;
    long long hist_len = 0;
    DevBuf blk_in, blk_out;
    FdlEngine *fdl = nullptr;   // frequency-domain delay-line engine (partitioned plans with minBlockOrder >= 3)
    int channels = 1;
    // CUDA-graph cache of whole device-resident calls (same pointers/sizes as a previous call)
    struct GraphEntry {
        const void *in; void *out; long long n, channels, in_stride, out_stride;
        cudaGraphExec_t exec; uint64_t launches; int seen;
        uint64_t alloc_gen;   // g_alloc_generation at capture: any later (re)allocation of a shared buffer invalidates the graph
    };
    std::vector<GraphEntry> graphs;
};

template <typename T> static std::vector<FftConv<T>> &plan_fc(adsp_plan *p);
template <> std::vector<FftConv<double>> &plan_fc<double>(adsp_plan *p) { return p->fc64; }
template <> std::vector<FftConv<float>> &plan_fc<float>(adsp_plan *p) { return p->fc32; }
template <typename T> static std::map<long long, FftConv<T>> &plan_extra(adsp_plan *p);
template <> std::map<long long, FftConv<double>> &plan_extra<double>(adsp_plan *p) { return p->extra64; }
template <> std::map<long long, FftConv<float>> &plan_extra<float>(adsp_plan *p) { return p->extra32; }

template <typename T> static adsp_status plan_build(adsp_plan *p, const T *host_kernel) {
    adsp_ctx *ctx = p->ctx;
    ADSP_TRY(p->d_kernel.reserve((size_t)p->K * sizeof(T)));
    ADSP_TRY(upload(ctx, p->d_kernel.p, host_kernel, (size_t)p->K * sizeof(T)));
    p->ch = choose_fft(p->K);
    auto &v = plan_fc<T>(p);
    v.resize((size_t)p->ch.parts);
    for (int i = 0; i < p->ch.parts; i++) {
        const long long k0 = (long long)i * p->ch.part_len;
        const long long kp = std::min<long long>(p->ch.part_len, p->K - k0);
        ADSP_TRY(v[(size_t)i].init(ctx, (const T *)p->d_kernel.p + k0, kp, p->ch));
    }
    return ADSP_OK;
}

template <typename T>
static adsp_status plan_run_device(adsp_plan *p, const T *d_in, long long n, long long channels, long long in_stride,
                                   T *d_out, long long out_stride) {
    auto &v = plan_fc<T>(p);
    const long long out_len = n + p->K - 1;
    if (v.size() == 1) {
        for (const Segment &sg : plan_segments(out_len, p->K, p->ch)) {
            FftConv<T> *fc = &v[0];
            if (sg.N != p->ch.N) {
                auto &ex = plan_extra<T>(p);
                auto it = ex.find(sg.N);
                if (it == ex.end()) {
                    FftConv<T> f;
                    ADSP_TRY(f.init(p->ctx, (const T *)p->d_kernel.p, p->K, make_choice(p->K, sg.N)));
                    it = ex.emplace(sg.N, f).first;
                }
                fc = &it->second;
            }
            fc->no_discard = sg.no_discard;
            ADSP_TRY(fc->run(d_in, n, channels, in_stride, d_out, out_stride, sg.len, sg.off, sg.off, false));
        }
        return ADSP_OK;
    }
    ADSP_CUDA(cudaMemset2DAsync(d_out, (size_t)out_stride * sizeof(T), 0, (size_t)out_len * sizeof(T), (size_t)channels, p->ctx->main));
    for (size_t i = 0; i < v.size(); i++) {
        const long long k0 = (long long)i * p->ch.part_len;
        ADSP_TRY(v[i].run(d_in, n, channels, in_stride, d_out, out_stride, n + v[i].K - 1, 0, k0, true));
    }
    return ADSP_OK;
}

// host-pointer batch.  Small calls: upload, compute, download (pageable memory staged through the pinned slots by the
// copy-thread pool, staging.cu).  Large multi-channel calls: channel chunks flow through a five-stage pipeline
//   stage-in (pool: caller memory -> pinned slot) | H2D (copy_in) | kernels (main + workers) | D2H (copy_out) |
//   stage-out (pool: pinned slot -> caller memory)
// with kPipeSlots slots per direction, so both PCIe directions, the SMs and the host copy threads all overlap.  The two
// staging stages disappear for pinned/registered caller memory (adsp_host_alloc_pinned), which is DMA'd in place.
template <typename T>
static adsp_status plan_run_host(adsp_plan *p, const T *in, long long n, long long channels, long long in_stride,
                                 T *out, long long out_stride) {
    adsp_ctx *ctx = p->ctx;
    const long long out_len = n + p->K - 1;
    // device layout: dense rows padded to 32 elements so every channel starts 256-byte aligned
    const long long dis = ((n + 31) / 32) * 32, dos = ((out_len + 31) / 32) * 32;
    const size_t row_bytes = (size_t)(dis + dos) * sizeof(T);
    const auto tq = std::chrono::steady_clock::now();
    const bool stage_in = !host_ptr_is_pinned(in), stage_out = !host_ptr_is_pinned(out);
    if (ctx->host_profile) {
        for (int i = 6; i < 12; i++) ctx->host_prof_ms[i] = 0;
        ctx->host_prof_ms[9] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tq).count();
    }
    const bool staged = stage_in || stage_out;
    // chunk size: 96 MB measured best for DMA straight from pinned caller memory (tools/e2e_sweep.py); staged chunks are
    // smaller so that the five stages fill sooner and the pinned slots stay modest (kPipeSlots x 2 x chunk)
    const long long chunk_target = (staged ? env_ll("ADSP_STAGE_PIPE_CHUNK_MB", 32) : env_ll("ADSP_PIPE_CHUNK_MB", 96)) << 20;
    long long cc = (long long)(chunk_target / (long long)row_bytes);
    if (cc < 1) cc = 1;
    if (channels < 4 || (size_t)channels * row_bytes < (size_t)(32u << 20) || cc >= channels) {
        using clk = std::chrono::steady_clock;
        const bool prof = ctx->host_profile;
        auto t0 = clk::now();
        auto lap = [&](int slot) {   // profile mode serialises the phases (one extra stream sync each)
            if (!prof) return;
            cudaStreamSynchronize(ctx->main);
            const auto t1 = clk::now();
            ctx->host_prof_ms[slot] = std::chrono::duration<double, std::milli>(t1 - t0).count();
            t0 = t1;
        };
        const auto tstart = t0;
        ADSP_TRY(ctx->d_in.reserve((size_t)dis * channels * sizeof(T)));
        ADSP_TRY(ctx->d_out.reserve((size_t)dos * channels * sizeof(T)));
        lap(0);   // allocation (zero after the first call of a shape)
        ADSP_TRY(upload2d(ctx, ctx->d_in.p, (size_t)dis * sizeof(T), in, (size_t)in_stride * sizeof(T), (size_t)n * sizeof(T), (size_t)channels));
        lap(1);   // stage-in + H2D
        ADSP_TRY(plan_run_device<T>(p, (const T *)ctx->d_in.p, n, channels, dis, (T *)ctx->d_out.p, dos));
        lap(2);   // kernels
        ADSP_TRY(download2d(ctx, out, (size_t)out_stride * sizeof(T), ctx->d_out.p, (size_t)dos * sizeof(T), (size_t)out_len * sizeof(T),
                            (size_t)channels));
        ADSP_CUDA(cudaStreamSynchronize(ctx->main));
        lap(3);   // D2H + stage-out
        if (prof) ctx->host_prof_ms[5] = std::chrono::duration<double, std::milli>(clk::now() - tstart).count();
        return ADSP_OK;
    }
    constexpr int NS = kPipeSlots;
    StagePool *pool = staged ? stage_pool(ctx) : nullptr;
    // Flushing the pinned slot's lines after the stage-out read (see staging.cu) takes a mono call from 1.0 to 0.35 ms, but
    // costs the bulk pipeline 15 % (39.7 -> 45.9 ms per 2.2 GB step: with 32 MB chunks the inbound DMA mostly lands in lines
    // that have left the caches anyway, and the flushes compete with the copies): off here, on in the small-call path.
    static const bool flush_pipe = env_ll("ADSP_STAGE_FLUSH_PIPE", 0) != 0;
    for (int s = 0; s < NS; s++) {
        ADSP_TRY(ctx->pipe_in[s].reserve((size_t)dis * cc * sizeof(T)));
        ADSP_TRY(ctx->pipe_out[s].reserve((size_t)dos * cc * sizeof(T)));
        if (stage_in) ADSP_TRY(ctx->h_in[s].reserve((size_t)n * cc * sizeof(T)));
        if (stage_out) ADSP_TRY(ctx->h_out[s].reserve((size_t)out_len * cc * sizeof(T)));
    }
    if (stage_in) ctx->staged_bytes_in += (uint64_t)channels * n * sizeof(T);
    if (stage_out) ctx->staged_bytes_out += (uint64_t)channels * out_len * sizeof(T);
    const long long nchunks = (channels + cc - 1) / cc;
    auto chunk_c0 = [&](long long i) { return i * cc; };
    auto chunk_nc = [&](long long i) { return std::min(cc, channels - i * cc); };
    StagePool::Ticket tk_in[NS], tk_out[NS];
    // The copy threads read and write CALLER memory: whatever way this function is left (an error in the middle of the
    // pipeline included), no copy may still be running when it returns (cgo pointer rule: the memory is only ours during
    // the call), and no DMA either.
    struct Drain {
        adsp_ctx *ctx; StagePool *pool; StagePool::Ticket *a, *b; int n;
        ~Drain() {
            if (pool) for (int i = 0; i < n; i++) { pool->wait(a[i]); pool->wait(b[i]); }
            cudaStreamSynchronize(ctx->copy_in); cudaStreamSynchronize(ctx->copy_out); cudaStreamSynchronize(ctx->main);
        }
    } drain{ctx, pool, tk_in, tk_out, NS};
    // iteration i: stage-in of chunk i | enqueue H2D + kernels + D2H of chunk i-1 | stage-out of chunk i-2
    for (long long i = 0; i < nchunks + 2; i++) {
        if (i < nchunks && stage_in) {
            const int s = (int)(i % NS);
            ADSP_CUDA(cudaEventSynchronize(ctx->ev_in[s]));              // the H2D that last read this pinned slot is done
            tk_in[s] = pool->copy2d_async(ctx->h_in[s].p, (size_t)n * sizeof(T), in + chunk_c0(i) * in_stride, (size_t)in_stride * sizeof(T),
                                          (size_t)n * sizeof(T), (size_t)chunk_nc(i));
        }
        if (i >= 1 && i - 1 < nchunks) {
            const long long j = i - 1, c0 = chunk_c0(j), nc = chunk_nc(j);
            const int s = (int)(j % NS);
            // H2D (device slot free once the kernels that last read it are done)
            if (j >= NS) ADSP_CUDA(cudaStreamWaitEvent(ctx->copy_in, ctx->ev_comp[s], 0));
            if (stage_in) {
                pool->wait(tk_in[s]);
                ADSP_CUDA(cudaMemcpy2DAsync(ctx->pipe_in[s].p, (size_t)dis * sizeof(T), ctx->h_in[s].p, (size_t)n * sizeof(T), (size_t)n * sizeof(T),
                                            (size_t)nc, cudaMemcpyHostToDevice, ctx->copy_in));
            } else {
                ADSP_CUDA(cudaMemcpy2DAsync(ctx->pipe_in[s].p, (size_t)dis * sizeof(T), in + c0 * in_stride, (size_t)in_stride * sizeof(T),
                                            (size_t)n * sizeof(T), (size_t)nc, cudaMemcpyHostToDevice, ctx->copy_in));
            }
            ADSP_CUDA(cudaEventRecord(ctx->ev_in[s], ctx->copy_in));
            // kernels (output slot free once its previous D2H is done)
            ADSP_CUDA(cudaStreamWaitEvent(ctx->main, ctx->ev_in[s], 0));
            if (j >= NS) ADSP_CUDA(cudaStreamWaitEvent(ctx->main, ctx->ev_out[s], 0));
            ADSP_TRY(plan_run_device<T>(p, (const T *)ctx->pipe_in[s].p, n, nc, dis, (T *)ctx->pipe_out[s].p, dos));
            ADSP_CUDA(cudaEventRecord(ctx->ev_comp[s], ctx->main));
            // D2H (pinned slot free once the pool has copied its previous contents out)
            ADSP_CUDA(cudaStreamWaitEvent(ctx->copy_out, ctx->ev_comp[s], 0));
            if (stage_out) {
                pool->wait(tk_out[s]);
                ADSP_CUDA(cudaMemcpy2DAsync(ctx->h_out[s].p, (size_t)out_len * sizeof(T), ctx->pipe_out[s].p, (size_t)dos * sizeof(T),
                                            (size_t)out_len * sizeof(T), (size_t)nc, cudaMemcpyDeviceToHost, ctx->copy_out));
            } else {
                ADSP_CUDA(cudaMemcpy2DAsync(out + c0 * out_stride, (size_t)out_stride * sizeof(T), ctx->pipe_out[s].p, (size_t)dos * sizeof(T),
                                            (size_t)out_len * sizeof(T), (size_t)nc, cudaMemcpyDeviceToHost, ctx->copy_out));
            }
            ADSP_CUDA(cudaEventRecord(ctx->ev_out[s], ctx->copy_out));
        }
        if (i >= 2 && stage_out) {
            const long long j = i - 2;
            const int s = (int)(j % NS);
            ADSP_CUDA(cudaEventSynchronize(ctx->ev_out[s]));
            tk_out[s] = pool->copy2d_async(out + chunk_c0(j) * out_stride, (size_t)out_stride * sizeof(T), ctx->h_out[s].p, (size_t)out_len * sizeof(T),
                                           (size_t)out_len * sizeof(T), (size_t)chunk_nc(j), flush_pipe);
        }
    }
    for (int s = 0; s < NS; s++) pool->wait(tk_out[s]);
    ADSP_CUDA(cudaStreamSynchronize(ctx->copy_out));
    ADSP_CUDA(cudaStreamSynchronize(ctx->main));
    return ADSP_OK;
}

// ================================================================ C ABI
extern "C" {

const char *adsp_version(void) { return "algodsp_cuda 0.1 (sm_100a)"; }

const char *adsp_status_string(adsp_status st) {
    switch (st) {
    case ADSP_OK: return "ok";
    case ADSP_ERR_EMPTY_INPUT: return "conv: empty input";
    case ADSP_ERR_EMPTY_KERNEL: return "conv: empty kernel";
    case ADSP_ERR_LENGTH_MISMATCH: return "conv: buffer length mismatch";
    case ADSP_ERR_INVALID_BLOCK_SIZE: return "conv: invalid block size";
    case ADSP_ERR_INVALID_BLOCK_ORDER: return "conv: invalid block order";
    case ADSP_ERR_EMPTY_IR: return "conv: empty impulse response";
    case ADSP_ERR_STAGE_INDEX: return "conv: stage index out of range";
    case ADSP_ERR_INVALID_ARG: return "algodsp: invalid argument";
    case ADSP_ERR_CUDA: return "algodsp: CUDA error";
    case ADSP_ERR_OOM: return "algodsp: out of memory";
    case ADSP_ERR_DIVISION_BY_ZERO: return "conv: division by zero in deconvolution";
    }
    return "algodsp: unknown status";
}

size_t adsp_last_error(char *buf, size_t buflen) {
    const std::string &s = g_last_error;
    if (buf && buflen) {
        const size_t n = std::min(buflen - 1, s.size());
        memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return s.size();
}

int adsp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

adsp_status adsp_ctx_create(int device, adsp_ctx **out) {
    if (!out) return ADSP_ERR_INVALID_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no CUDA device available: libalgodsp_cuda has no CPU fallback");
        cudaGetLastError();
        return ADSP_ERR_CUDA;
    }
    if (device < 0 || device >= n) { set_error("device index out of range"); return ADSP_ERR_INVALID_ARG; }
    ADSP_CUDA(cudaSetDevice(device));
    adsp_ctx *c = new adsp_ctx();
    c->device = device;
    cudaDeviceProp prop;
    ADSP_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    c->l2_bytes = (size_t)prop.l2CacheSize;
    long long budget_mb = env_ll("ADSP_SCRATCH_MB", 0);
    // four-step intermediates in flight across all worker streams.  Measured optimum on B200 (126 MB L2, two
    // partitions): ~40 MB; beyond ~56 MB the streaming signals start evicting scratch lines (tools/hint_sweep.sh)
    if (budget_mb <= 0) budget_mb = (long long)(c->l2_bytes >> 20) * 32 / 100;
    if (budget_mb < 8) budget_mb = 8;
    c->scratch_budget = (size_t)budget_mb << 20;
    ADSP_CUDA(cudaStreamCreateWithFlags(&c->main, cudaStreamNonBlocking));
    ADSP_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    for (int i = 0; i < kWorkerStreams; i++) {
        ADSP_CUDA(cudaStreamCreateWithFlags(&c->worker[i], cudaStreamNonBlocking));
        ADSP_CUDA(cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming));
    }
    ADSP_CUDA(cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking));
    ADSP_CUDA(cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking));
    for (int i = 0; i < kPipeSlots; i++) {
        ADSP_CUDA(cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
        ADSP_CUDA(cudaEventCreateWithFlags(&c->ev_comp[i], cudaEventDisableTiming));
        ADSP_CUDA(cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming));
    }
    *out = c;
    return ADSP_OK;
}

void adsp_ctx_destroy(adsp_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (auto &kv : c->tw_tables) cudaFree(kv.second);
    for (auto &kv : c->tw4_tables) { cudaFree(kv.second.first); cudaFree(kv.second.second); }
    c->spec_cache.release();
    c->scratch.release(); c->d_in.release(); c->d_out.release(); c->d_k.release(); c->d_tmp.release(); c->d_small.release(); c->d_counters.release();
    c->pool.reset();
    for (int i = 0; i < kPipeSlots; i++) {
        c->h_in[i].release(); c->h_out[i].release(); c->pipe_in[i].release(); c->pipe_out[i].release();
        cudaEventDestroy(c->ev_in[i]); cudaEventDestroy(c->ev_comp[i]); cudaEventDestroy(c->ev_out[i]);
    }
    for (auto &t : c->timed) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); }
    for (auto e : c->event_pool) cudaEventDestroy(e);
    cudaStreamDestroy(c->copy_in); cudaStreamDestroy(c->copy_out);
    for (int i = 0; i < kWorkerStreams; i++) { cudaStreamDestroy(c->worker[i]); cudaEventDestroy(c->ev_join[i]); }
    cudaEventDestroy(c->ev_fork);
    cudaStreamDestroy(c->main);
    delete c;
}

adsp_status adsp_ctx_sync(adsp_ctx *c) {
    if (!c) return ADSP_ERR_INVALID_ARG;
    ADSP_CUDA(cudaSetDevice(c->device));
    ADSP_CUDA(cudaStreamSynchronize(c->main));
    return ADSP_OK;
}

uint64_t adsp_ctx_launch_count(adsp_ctx *c) { return c ? c->launches.load() : 0; }

void adsp_ctx_kernel_timing(adsp_ctx *c, int enable) {
    if (!c) return;
    c->timing = enable != 0;
}

adsp_status adsp_ctx_kernel_time(adsp_ctx *c, int kind, double *total_ms, uint64_t *launches, int reset) {
    if (!c || kind < 0 || kind >= 8) return ADSP_ERR_INVALID_ARG;
    ADSP_CUDA(cudaSetDevice(c->device));
    ADSP_CUDA(cudaDeviceSynchronize());
    for (auto &t : c->timed) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, t.e0, t.e1) == cudaSuccess) { c->kind_ms[t.kind] += ms; c->kind_launches[t.kind]++; }
        c->event_pool.push_back(t.e0); c->event_pool.push_back(t.e1);
    }
    c->timed.clear();
    if (total_ms) *total_ms = c->kind_ms[kind];
    if (launches) *launches = c->kind_launches[kind];
    if (reset) for (int k = 0; k < 8; k++) { c->kind_ms[k] = 0; c->kind_launches[k] = 0; }
    return ADSP_OK;
}
void *adsp_ctx_stream(adsp_ctx *c) { return c ? (void *)c->main : nullptr; }

void adsp_ctx_host_profile(adsp_ctx *c, int enable) {
    if (!c) return;
    std::lock_guard<std::mutex> lk(c->mu);
    c->host_profile = enable != 0;
    for (double &v : c->host_prof_ms) v = 0;
}
adsp_status adsp_ctx_host_profile_get(adsp_ctx *c, double *ms, int cap, uint64_t *staged_in_bytes, uint64_t *staged_out_bytes) {
    if (!c) return ADSP_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    if (ms) for (int i = 0; i < cap && i < 12; i++) ms[i] = c->host_prof_ms[i];
    if (staged_in_bytes) *staged_in_bytes = c->staged_bytes_in;
    if (staged_out_bytes) *staged_out_bytes = c->staged_bytes_out;
    return ADSP_OK;
}
int adsp_host_ptr_is_pinned(const void *p) { return host_ptr_is_pinned(p) ? 1 : 0; }
int adsp_ctx_stage_threads(adsp_ctx *c) {
    if (!c) return 0;
    std::lock_guard<std::mutex> lk(c->mu);
    return stage_pool(c)->threads();
}

adsp_status adsp_host_alloc_pinned(size_t bytes, void **out) {
    if (!out) return ADSP_ERR_INVALID_ARG;
    ADSP_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
    return ADSP_OK;
}
void adsp_host_free_pinned(void *p) { if (p) cudaFreeHost(p); }

// In-place pinning of caller memory the caller keeps alive (a long-lived Go slice held by a runtime.Pinner, a C buffer):
// afterwards every host-pointer call DMAs from / to it directly, like memory from adsp_host_alloc_pinned.  The range MUST be
// unregistered before it is freed or moved.
adsp_status adsp_host_register(void *p, size_t bytes) {
    if (!p || bytes == 0) return ADSP_ERR_INVALID_ARG;
    ADSP_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return ADSP_OK;
}
adsp_status adsp_host_unregister(void *p) {
    if (!p) return ADSP_ERR_INVALID_ARG;
    ADSP_CUDA(cudaHostUnregister(p));
    return ADSP_OK;
}

adsp_status adsp_device_alloc(adsp_ctx *c, size_t bytes, void **out) {
    if (!c || !out) return ADSP_ERR_INVALID_ARG;
    ADSP_CUDA(cudaSetDevice(c->device));
    ADSP_CUDA(cudaMalloc(out, bytes ? bytes : 1));
    return ADSP_OK;
}
void adsp_device_free(adsp_ctx *c, void *p) { if (c && p) { cudaSetDevice(c->device); cudaFree(p); } }

adsp_status adsp_memcpy_h2d(adsp_ctx *c, void *dst, const void *src, size_t bytes) {
    if (!c) return ADSP_ERR_INVALID_ARG;
    ADSP_CUDA(cudaSetDevice(c->device));
    ADSP_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->main));
    ADSP_CUDA(cudaStreamSynchronize(c->main));
    return ADSP_OK;
}
adsp_status adsp_memcpy_d2h(adsp_ctx *c, void *dst, const void *src, size_t bytes) {
    if (!c) return ADSP_ERR_INVALID_ARG;
    ADSP_CUDA(cudaSetDevice(c->device));
    ADSP_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->main));
    ADSP_CUDA(cudaStreamSynchronize(c->main));
    return ADSP_OK;
}

// ---------------------------------------------------------------- sizing helpers
int64_t adsp_next_pow2(int64_t n) { return next_pow2_ll(n); }
int adsp_is_pow2(int64_t n) { return n > 0 && (n & (n - 1)) == 0; }

adsp_status adsp_ols_sizes(int64_t K, int64_t fft, int64_t *fft_out, int64_t *step_out) {
    if (K <= 0) return ADSP_ERR_EMPTY_KERNEL;
    if (fft <= 0) fft = std::max<int64_t>(next_pow2_ll(2 * K), 256);
    if (!adsp_is_pow2(fft)) {
        set_error("conv: invalid block size: fftSize must be power of 2, got " + std::to_string(fft));
        return ADSP_ERR_INVALID_BLOCK_SIZE;
    }
    if (fft < 2 * K) fft = next_pow2_ll(2 * K);
    if (fft_out) *fft_out = fft;
    if (step_out) *step_out = fft - K + 1;
    return ADSP_OK;
}

adsp_status adsp_ola_sizes(int64_t K, int64_t block, int64_t *block_out, int64_t *fft_out) {
    if (K <= 0) return ADSP_ERR_EMPTY_KERNEL;
    if (block <= 0) block = std::max<int64_t>(next_pow2_ll(K), 256);
    if (block_out) *block_out = block;
    if (fft_out) *fft_out = next_pow2_ll(block + K - 1);
    return ADSP_OK;
}

void adsp_trim_mode(int64_t la, int64_t lb, adsp_mode mode, int64_t *start, int64_t *len) {
    int64_t s = 0, l = la + lb - 1;
    if (mode == ADSP_MODE_SAME) { s = (lb - 1) / 2; l = la; }
    else if (mode == ADSP_MODE_VALID) {
        if (la >= lb) { s = lb - 1; l = la - s; }
        else { s = la - 1; l = lb - s; }
    }
    if (start) *start = s;
    if (len) *len = l;
}

int64_t adsp_lag_from_index(int64_t index, int64_t len_b) { return index - (len_b - 1); }
int64_t adsp_index_from_lag(int64_t lag, int64_t len_b) { return lag + (len_b - 1); }

}  // extern "C"

extern "C" { static void plan_free(adsp_plan *p); }   // frees a plan; the caller holds the context lock

// ---------------------------------------------------------------- one-shot implementations
namespace {

enum OneShot { OS_DIRECT, OS_CONVOLVE, OS_FFT, OS_CORRELATE, OS_CORRELATE_DIRECT, OS_CORRELATE_FFT };

// a (n), b (m) host -> out (n+m-1) host.  `post`: 0 none, 1 divide by ||a||*||b||, 2 divide by out[n-1]
template <typename T>
adsp_status oneshot(adsp_ctx *ctx, OneShot op, const T *a, int64_t n, const T *b, int64_t m, T *out, int post) {
    if (!ctx) return ADSP_ERR_INVALID_ARG;
    const bool corr = (op == OS_CORRELATE || op == OS_CORRELATE_DIRECT || op == OS_CORRELATE_FFT);
    if (corr) { if (n <= 0 || m <= 0) return ADSP_ERR_EMPTY_INPUT; }   // correlate.go:17-19
    else {
        if (n <= 0) return ADSP_ERR_EMPTY_INPUT;                       // conv.go:77-83,195-201
        if (m <= 0) return ADSP_ERR_EMPTY_KERNEL;
    }
    if (!a || !b || !out) return ADSP_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    const int64_t out_len = n + m - 1;
    ADSP_TRY(ctx->d_in.reserve((size_t)n * sizeof(T)));
    ADSP_TRY(ctx->d_tmp.reserve((size_t)m * sizeof(T) * 2));
    ADSP_TRY(ctx->d_out.reserve((size_t)out_len * sizeof(T)));
    T *da = (T *)ctx->d_in.p, *db = (T *)ctx->d_tmp.p, *dbr = db + m, *dout = (T *)ctx->d_out.p;
    ADSP_TRY(upload(ctx, da, a, (size_t)n * sizeof(T)));
    ADSP_TRY(upload(ctx, db, b, (size_t)m * sizeof(T)));
    const T *dk = db;
    if (corr) {  // reverse b: correlate.go:22-25
        reverse_kernel<T><<<(unsigned)((m + 255) / 256), 256, 0, ctx->main>>>(db, m, m, dbr, m, 1);
        count_launch(ctx);
        dk = dbr;
    }
    bool use_direct = (op == OS_DIRECT || op == OS_CORRELATE_DIRECT);
    const T *sig = da; int64_t sn = n; const T *ker = dk; int64_t km = m;
    if (op == OS_CONVOLVE || op == OS_CORRELATE) {
        if (m > n) { sig = dk; sn = m; ker = da; km = n; }   // conv.go:204-206 swap so the signal is longer
        if (km <= 64) use_direct = true;                     // conv.go:209-211 (code wins over doc: <= 64)
    }
    if (use_direct) ADSP_TRY(direct_device<T>(ctx, sig, sn, 0, ker, km, 0, 1, dout, 0));
    else {
        bool done = false;
        // long correlations: one packed transform of (a, reverse b) instead of a generic long-kernel convolution
        if (corr && env_ll("ADSP_CORR_GENERIC", 0) == 0)
            ADSP_TRY(fft_correlate_pairs_device<T>(ctx, da, n, n, db, m, m, 1, dout, out_len, (T *)nullptr, (long long *)nullptr, &done));
        if (!done) ADSP_TRY(fft_convolve_device<T>(ctx, sig, sn, 1, 0, ker, km, dout, 0));
    }
    if (post) {
        ADSP_TRY(ctx->d_small.reserve(64));
        T *dden = (T *)ctx->d_small.p;
        if (post == 1) { norm_product_kernel<T><<<1, 1024, 0, ctx->main>>>(da, n, db, m, dden); count_launch(ctx); }
        else ADSP_CUDA(cudaMemcpyAsync(dden, dout + (n - 1), sizeof(T), cudaMemcpyDeviceToDevice, ctx->main));
        scale_by_inverse_kernel<T><<<(unsigned)((out_len + 255) / 256), 256, 0, ctx->main>>>(dout, out_len, dden);
        count_launch(ctx);
    }
    ADSP_TRY(download(ctx, out, dout, (size_t)out_len * sizeof(T)));
    ADSP_CUDA(cudaStreamSynchronize(ctx->main));
    return ADSP_OK;
}

template <typename T>
adsp_status plan_create(adsp_ctx *ctx, PlanKind kind, const T *kernel, int64_t K, int64_t size_arg, adsp_plan **out) {
    if (!out) return ADSP_ERR_INVALID_ARG;
    *out = nullptr;
    if (!ctx) return ADSP_ERR_INVALID_ARG;
    if (K <= 0 || !kernel) return ADSP_ERR_EMPTY_KERNEL;   // overlap_save.go:54-56, overlap_add.go:45-47
    int64_t f = 0, s = 0, bsz = 0;
    if (kind == PLAN_OLS) ADSP_TRY(adsp_ols_sizes(K, size_arg, &f, &s));
    else ADSP_TRY(adsp_ola_sizes(K, size_arg, &bsz, &f));
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    adsp_plan *p = new adsp_plan();
    p->ctx = ctx; p->kind = kind; p->prec = sizeof(T) == 8 ? ADSP_F64 : ADSP_F32; p->K = K;
    p->ref_fft = f; p->ref_step = s; p->ref_block = bsz;
    adsp_status st = plan_build<T>(p, kernel);
    if (st != ADSP_OK) { plan_free(p); return st; }
    *out = p;
    return ADSP_OK;
}

}  // namespace

extern "C" {

adsp_status adsp_direct(adsp_ctx *c, const double *a, int64_t n, const double *b, int64_t m, double *out) { return oneshot<double>(c, OS_DIRECT, a, n, b, m, out, 0); }
adsp_status adsp_convolve(adsp_ctx *c, const double *a, int64_t n, const double *b, int64_t m, double *out) { return oneshot<double>(c, OS_CONVOLVE, a, n, b, m, out, 0); }
adsp_status adsp_overlap_add_convolve(adsp_ctx *c, const double *s, int64_t n, const double *k, int64_t m, double *out) {
    if (m <= 0) return ADSP_ERR_EMPTY_KERNEL;   // overlap_add.go:222-224 (kernel checked first)
    return oneshot<double>(c, OS_FFT, s, n, k, m, out, 0);
}
adsp_status adsp_overlap_save_convolve(adsp_ctx *c, const double *s, int64_t n, const double *k, int64_t m, double *out) {
    if (m <= 0) return ADSP_ERR_EMPTY_KERNEL;   // overlap_save.go:314-316
    return oneshot<double>(c, OS_FFT, s, n, k, m, out, 0);
}
adsp_status adsp_correlate(adsp_ctx *c, const double *a, int64_t n, const double *b, int64_t m, double *out) { return oneshot<double>(c, OS_CORRELATE, a, n, b, m, out, 0); }
adsp_status adsp_correlate_direct(adsp_ctx *c, const double *a, int64_t n, const double *b, int64_t m, double *out) { return oneshot<double>(c, OS_CORRELATE_DIRECT, a, n, b, m, out, 0); }
// ---------------------------------------------------------------- deconvolution (deconvolve.go)
int64_t adsp_deconv_out_len(int64_t n, int64_t m) { const int64_t o = n - m + 1; return o <= 0 ? n : o; }   // deconvolve.go:101-105

static adsp_status deconv_run(adsp_ctx *ctx, const double *d_sig, int64_t n, int64_t ss, const double *d_ker, int64_t m, int64_t ks,
                              int64_t batch, double reg, double *d_out, int64_t os, int64_t out_len) {
    ADSP_TRY(ctx->d_small.reserve(64));
    long long *d_bad = (long long *)ctx->d_small.p;
    const long long none = 0x7fffffffffffffffLL;
    ADSP_CUDA(cudaMemcpyAsync(d_bad, &none, sizeof none, cudaMemcpyHostToDevice, ctx->main));
    ADSP_TRY(fft_deconvolve_device<double>(ctx, d_sig, n, ss, d_ker, m, ks, batch, d_out, os, out_len, reg, d_bad));
    if (reg < 0) {
        long long bad = none;
        ADSP_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, ctx->main));
        ADSP_CUDA(cudaStreamSynchronize(ctx->main));
        if (bad != none) {                                                  // deconvolve.go:146-148
            set_error("conv: division by zero in deconvolution: at frequency bin " + std::to_string(bad));
            return ADSP_ERR_DIVISION_BY_ZERO;
        }
    }
    return ADSP_OK;
}

adsp_status adsp_deconvolve(adsp_ctx *ctx, const double *signal, int64_t n, const double *kernel, int64_t m, int method, double epsilon,
                            double noise_variance, double signal_variance, double *out, int64_t out_len) {
    if (!ctx) return ADSP_ERR_INVALID_ARG;
    if (n <= 0) return ADSP_ERR_EMPTY_INPUT;                                 // deconvolve.go:73-79
    if (m <= 0) return ADSP_ERR_EMPTY_KERNEL;
    if (!signal || !kernel || !out) return ADSP_ERR_INVALID_ARG;
    if (out_len != adsp_deconv_out_len(n, m)) { set_error("conv: buffer length mismatch"); return ADSP_ERR_LENGTH_MISMATCH; }
    double reg;
    if (method == 0) reg = -1.0;
    else if (method == 2) {                                                  // deconvolve.go:254-270
        double sv = signal_variance, nv = noise_variance;
        if (sv <= 0) {                                                       // variance(), deconvolve.go:332-353
            double mean = 0;
            for (int64_t i = 0; i < n; i++) mean += signal[i];
            mean /= (double)n;
            double sum = 0;
            for (int64_t i = 0; i < n; i++) { const double d = signal[i] - mean; sum += d * d; }
            sv = sum / (double)n;
        }
        if (nv <= 0) nv = sv * 0.01;
        reg = nv / sv;
        if (!(reg > 0)) reg = 1e-6;
    } else if (method == 1) reg = epsilon <= 0 ? 1e-6 : epsilon;             // deconvolve.go:84-88
    else reg = 1e-6;                                                         // deconvolve.go:91-93
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    ADSP_TRY(ctx->d_in.reserve((size_t)n * sizeof(double)));
    ADSP_TRY(ctx->d_tmp.reserve((size_t)m * sizeof(double)));
    ADSP_TRY(ctx->d_out.reserve((size_t)out_len * sizeof(double)));
    ADSP_TRY(upload(ctx, ctx->d_in.p, signal, (size_t)n * sizeof(double)));
    ADSP_TRY(upload(ctx, ctx->d_tmp.p, kernel, (size_t)m * sizeof(double)));
    ADSP_TRY(deconv_run(ctx, (const double *)ctx->d_in.p, n, n, (const double *)ctx->d_tmp.p, m, m, 1, reg, (double *)ctx->d_out.p, out_len, out_len));
    ADSP_TRY(download(ctx, out, ctx->d_out.p, (size_t)out_len * sizeof(double)));
    ADSP_CUDA(cudaStreamSynchronize(ctx->main));
    return ADSP_OK;
}

adsp_status adsp_inverse_filter(adsp_ctx *ctx, const double *kernel, int64_t m, int64_t length, double epsilon, double *out) {
    if (!ctx) return ADSP_ERR_INVALID_ARG;
    if (m <= 0) return ADSP_ERR_EMPTY_KERNEL;                                // deconvolve.go:360-362
    if (length <= 0) return ADSP_OK;
    if (!kernel || !out) return ADSP_ERR_INVALID_ARG;
    if (epsilon <= 0) epsilon = 1e-6;                                        // :364-366
    int64_t N = 1;
    while (N < length) N *= 2;
    const int64_t mm = std::min(m, N);                                       // :377-379 kernel truncated to the transform
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    // the signal is a unit impulse: S[k] = 1, so R = conj(H) / (|H|^2 + eps)  (:391-395)
    ADSP_TRY(ctx->d_in.reserve((size_t)length * sizeof(double)));
    ADSP_TRY(ctx->d_tmp.reserve((size_t)mm * sizeof(double)));
    ADSP_TRY(ctx->d_out.reserve((size_t)length * sizeof(double)));
    ADSP_CUDA(cudaMemsetAsync(ctx->d_in.p, 0, (size_t)length * sizeof(double), ctx->main));
    const double one = 1.0;
    ADSP_CUDA(cudaMemcpyAsync(ctx->d_in.p, &one, sizeof one, cudaMemcpyHostToDevice, ctx->main));
    ADSP_CUDA(cudaMemcpyAsync(ctx->d_tmp.p, kernel, (size_t)mm * sizeof(double), cudaMemcpyHostToDevice, ctx->main));
    ADSP_TRY(deconv_run(ctx, (const double *)ctx->d_in.p, length, length, (const double *)ctx->d_tmp.p, mm, mm, 1, epsilon, (double *)ctx->d_out.p, length, length));
    ADSP_CUDA(cudaMemcpyAsync(out, ctx->d_out.p, (size_t)length * sizeof(double), cudaMemcpyDeviceToHost, ctx->main));
    ADSP_CUDA(cudaStreamSynchronize(ctx->main));
    return ADSP_OK;
}

double adsp_snr(const double *original, int64_t n, const double *recovered, int64_t n2) {     // deconvolve.go:417-434
    if (n != n2 || n <= 0 || !original || !recovered) return -INFINITY;
    double sp = 0, np = 0;
    for (int64_t i = 0; i < n; i++) { sp += original[i] * original[i]; const double d = original[i] - recovered[i]; np += d * d; }
    if (np == 0) return INFINITY;
    return 10 * log10(sp / np);
}

adsp_status adsp_deconvolve_batch_device(adsp_ctx *ctx, const double *sig, int64_t n, int64_t ss, const double *ker, int64_t m, int64_t ks,
                                         int64_t batch, double reg, double *out, int64_t os) {
    if (!ctx) return ADSP_ERR_INVALID_ARG;
    if (n <= 0) return ADSP_ERR_EMPTY_INPUT;
    if (m <= 0) return ADSP_ERR_EMPTY_KERNEL;
    if (!sig || !ker || !out || batch <= 0) return ADSP_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    return deconv_run(ctx, sig, n, ss, ker, m, ks, batch, reg, out, os, adsp_deconv_out_len(n, m));
}

adsp_status adsp_correlate_fft(adsp_ctx *c, const double *a, int64_t n, const double *b, int64_t m, double *out) { return oneshot<double>(c, OS_CORRELATE_FFT, a, n, b, m, out, 0); }
adsp_status adsp_correlate_normalized(adsp_ctx *c, const double *a, int64_t n, const double *b, int64_t m, double *out) { return oneshot<double>(c, OS_CORRELATE, a, n, b, m, out, 1); }
adsp_status adsp_autocorrelate_normalized(adsp_ctx *c, const double *a, int64_t n, double *out) { return oneshot<double>(c, OS_CORRELATE, a, n, a, n, out, 2); }
adsp_status adsp_direct_f32(adsp_ctx *c, const float *a, int64_t n, const float *b, int64_t m, float *out) { return oneshot<float>(c, OS_DIRECT, a, n, b, m, out, 0); }
adsp_status adsp_convolve_f32(adsp_ctx *c, const float *a, int64_t n, const float *b, int64_t m, float *out) { return oneshot<float>(c, OS_CONVOLVE, a, n, b, m, out, 0); }
adsp_status adsp_correlate_f32(adsp_ctx *c, const float *a, int64_t n, const float *b, int64_t m, float *out) { return oneshot<float>(c, OS_CORRELATE, a, n, b, m, out, 0); }

adsp_status adsp_direct_circular(adsp_ctx *ctx, const double *a, int64_t n, const double *b, int64_t m, double *out) {
    if (!ctx) return ADSP_ERR_INVALID_ARG;
    if (n <= 0 || m <= 0) return ADSP_ERR_EMPTY_INPUT;        // conv.go:159-161
    if (n != m) return ADSP_ERR_LENGTH_MISMATCH;              // conv.go:163-165
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    ADSP_TRY(ctx->d_in.reserve((size_t)n * 2 * sizeof(double)));
    ADSP_TRY(ctx->d_out.reserve((size_t)n * sizeof(double)));
    double *da = (double *)ctx->d_in.p, *db = da + n, *dout = (double *)ctx->d_out.p;
    ADSP_CUDA(cudaMemcpyAsync(da, a, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->main));
    ADSP_CUDA(cudaMemcpyAsync(db, b, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->main));
    direct_circular_kernel<double><<<(unsigned)((n + 127) / 128), 128, 0, ctx->main>>>(da, db, dout, n);
    count_launch(ctx);
    ADSP_CUDA(cudaMemcpyAsync(out, dout, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->main));
    ADSP_CUDA(cudaStreamSynchronize(ctx->main));
    return ADSP_OK;
}

adsp_status adsp_find_peak(adsp_ctx *ctx, const double *corr, int64_t len, int64_t *index, double *value) {
    if (!ctx || !index || !value) return ADSP_ERR_INVALID_ARG;
    if (len <= 0) { *index = -1; *value = 0; return ADSP_OK; }   // correlate.go:201-203
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    ADSP_TRY(ctx->d_in.reserve((size_t)len * sizeof(double)));
    ADSP_TRY(ctx->d_tmp.reserve(64));
    ADSP_TRY(upload(ctx, ctx->d_in.p, corr, (size_t)len * sizeof(double)));
    double *dv = (double *)ctx->d_tmp.p;
    long long *di = (long long *)(dv + 1);
    ADSP_TRY(peak_device<double>(ctx, (const double *)ctx->d_in.p, len, len, 1, dv, di));
    long long hi = -1; double hv = 0;
    ADSP_CUDA(cudaMemcpyAsync(&hv, dv, sizeof(double), cudaMemcpyDeviceToHost, ctx->main));
    ADSP_CUDA(cudaMemcpyAsync(&hi, di, sizeof(long long), cudaMemcpyDeviceToHost, ctx->main));
    ADSP_CUDA(cudaStreamSynchronize(ctx->main));
    *index = hi; *value = hv;
    return ADSP_OK;
}

// ---------------------------------------------------------------- batched one-shots
adsp_status adsp_direct_batch_device(adsp_ctx *ctx, const void *a, int64_t n, int64_t a_stride, const void *b, int64_t m,
                                     int64_t b_stride, int64_t batch, void *out, int64_t out_stride, adsp_precision prec) {
    if (!ctx) return ADSP_ERR_INVALID_ARG;
    if (n <= 0) return ADSP_ERR_EMPTY_INPUT;
    if (m <= 0) return ADSP_ERR_EMPTY_KERNEL;
    if (batch <= 0) return ADSP_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    if (prec == ADSP_F64) return direct_device<double>(ctx, (const double *)a, n, a_stride, (const double *)b, m, b_stride, batch, (double *)out, out_stride);
    return direct_device<float>(ctx, (const float *)a, n, a_stride, (const float *)b, m, b_stride, batch, (float *)out, out_stride);
}

adsp_status adsp_direct_batch(adsp_ctx *ctx, const double *a, int64_t n, int64_t a_stride, const double *b, int64_t m,
                              int64_t b_stride, int64_t batch, double *out, int64_t out_stride) {
    if (!ctx) return ADSP_ERR_INVALID_ARG;
    if (n <= 0) return ADSP_ERR_EMPTY_INPUT;
    if (m <= 0) return ADSP_ERR_EMPTY_KERNEL;
    if (batch <= 0) return ADSP_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    const int64_t out_len = n + m - 1;
    const int64_t nb = b_stride == 0 ? 1 : batch;
    ADSP_TRY(ctx->d_in.reserve((size_t)n * batch * sizeof(double)));
    ADSP_TRY(ctx->d_tmp.reserve((size_t)m * nb * sizeof(double)));
    ADSP_TRY(ctx->d_out.reserve((size_t)out_len * batch * sizeof(double)));
    ADSP_TRY(upload2d(ctx, ctx->d_in.p, (size_t)n * 8, a, (size_t)a_stride * 8, (size_t)n * 8, (size_t)batch));
    ADSP_TRY(upload2d(ctx, ctx->d_tmp.p, (size_t)m * 8, b, (size_t)(b_stride ? b_stride : m) * 8, (size_t)m * 8, (size_t)nb));
    ADSP_TRY(direct_device<double>(ctx, (const double *)ctx->d_in.p, n, n, (const double *)ctx->d_tmp.p, m, b_stride ? m : 0, batch,
                                   (double *)ctx->d_out.p, out_len));
    ADSP_TRY(download2d(ctx, out, (size_t)out_stride * 8, ctx->d_out.p, (size_t)out_len * 8, (size_t)out_len * 8, (size_t)batch));
    ADSP_CUDA(cudaStreamSynchronize(ctx->main));
    return ADSP_OK;
}

}  // extern "C"

// correlate batch: every pair has its own kernel (reversed b), so spectra are per pair.
// v1: loop over pairs with the one-shot FFT engine (sharing the transform tables); the peak search
// runs batched at the end when the full outputs are kept, per pair otherwise.
namespace {
template <typename T>
adsp_status correlate_batch_dev(adsp_ctx *ctx, const T *a, int64_t n, int64_t a_stride, const T *b, int64_t m,
                                int64_t b_stride, int64_t pairs, T *out, int64_t out_stride, long long *peak_i, T *peak_v) {
    const int64_t out_len = n + m - 1;
    const bool want_peaks = peak_i && peak_v;
    // ONE b for every pair (b_stride == 0: the measurement case, every response against the same excitation sweep):
    // Correlate(a_p, b) = Convolve(a_p, reverse(b)) (correlate.go:16-28) is then a batched convolution with one cached
    // kernel spectrum -- the overlap-save engine itself, two pairs per complex transform, no mirror-bin extraction at all
    // (half the forward transforms of the per-pair path below).
    if (b_stride == 0 && pairs >= 2 && std::min(n, m) > 64 && env_ll("ADSP_CORR_GENERIC", 0) == 0 && env_ll("ADSP_CORR_SHARED_B", 1) != 0) {
        const FftChoice big = choose_fft(m);
        if (big.parts == 1) {
            ADSP_TRY(ctx->d_k.reserve((size_t)m * sizeof(T)));
            T *dk = (T *)ctx->d_k.p;
            reverse_kernel<T><<<(unsigned)((m + 255) / 256), 256, 0, ctx->main>>>(b, m, m, dk, m, 1);
            count_launch(ctx);
            struct Seg { Segment sg; FftConv<T> fc; };
            std::vector<Seg> segs;
            adsp_status st = ADSP_OK;
            const std::vector<Segment> cover = plan_segments(out_len, m, big);
            size_t spec_elems = 0;
            for (const Segment &sg : cover) spec_elems += (size_t)sg.N;
            ADSP_TRY(ctx->spec_cache.reserve(spec_elems * sizeof(cpx<T>)));   // the spectra live in the context's cache: no cudaMalloc / cudaFree per call
            cpx<T> *hc = (cpx<T> *)ctx->spec_cache.p;
            for (const Segment &sg : cover) {
                segs.push_back({sg, FftConv<T>()});
                st = segs.back().fc.init(ctx, dk, m, make_choice(m, sg.N), hc);
                hc += sg.N;
                if (st != ADSP_OK) break;
            }
            const int64_t tstride = ((out_len + 31) / 32) * 32;
            const int64_t chunk = out ? pairs : std::min<int64_t>(pairs, 64);
            if (st == ADSP_OK && !out) st = ctx->d_tmp.reserve((size_t)chunk * (size_t)tstride * sizeof(T));
            for (int64_t c0 = 0; c0 < pairs && st == ADSP_OK; c0 += chunk) {
                const int64_t np = std::min(chunk, pairs - c0);
                T *o = out ? out + c0 * out_stride : (T *)ctx->d_tmp.p;
                const int64_t os = out ? out_stride : tstride;
                for (Seg &sgm : segs) {
                    sgm.fc.no_discard = sgm.sg.no_discard;
                    st = sgm.fc.run(a + c0 * a_stride, n, np, a_stride, o, os, sgm.sg.len, sgm.sg.off, sgm.sg.off, false);
                    if (st != ADSP_OK) break;
                }
                if (st == ADSP_OK && want_peaks) st = peak_device<T>(ctx, o, out_len, os, np, peak_v + c0, peak_i + c0);
            }
            for (Seg &sgm : segs) sgm.fc.destroy();          // (the spectra stay in the context's cache: nothing to wait for)
            return st;
        }
    }
    // long operands: one packed transform per pair + one shared inverse per two pairs, the peak search in the epilogue of the
    // inverse column pass; with out == NULL the correlation itself is never written
    if (std::min(n, m) > 64 && env_ll("ADSP_CORR_GENERIC", 0) == 0) {
        bool done = false;
        ADSP_TRY(fft_correlate_pairs_device<T>(ctx, a, n, a_stride, b, m, b_stride, pairs, out, out_stride, want_peaks ? peak_v : (T *)nullptr,
                                               want_peaks ? peak_i : (long long *)nullptr, &done));
        if (done) return ADSP_OK;
    }
    // generic path: Convolve(a, reverse(b)) per pair (direct for short operands, partitioned FFT beyond 2^22 points)
    ADSP_TRY(ctx->d_tmp.reserve((size_t)(m + (out ? 0 : out_len)) * sizeof(T)));
    T *dbr = (T *)ctx->d_tmp.p;
    T *tmp_out = out ? nullptr : dbr + m;
    for (int64_t p = 0; p < pairs; p++) {
        const T *ap = a + p * a_stride;
        const T *bp = b + p * b_stride;
        T *op = out ? out + p * out_stride : tmp_out;
        reverse_kernel<T><<<(unsigned)((m + 255) / 256), 256, 0, ctx->main>>>(bp, m, m, dbr, m, 1);
        count_launch(ctx);
        const T *sig = ap; int64_t sn = n; const T *ker = dbr; int64_t km = m;
        if (m > n) { sig = dbr; sn = m; ker = ap; km = n; }
        if (km <= 64) ADSP_TRY(direct_device<T>(ctx, sig, sn, 0, ker, km, 0, 1, op, 0));
        else ADSP_TRY(fft_convolve_device<T>(ctx, sig, sn, 1, 0, ker, km, op, 0));
        if (want_peaks && !out) ADSP_TRY(peak_device<T>(ctx, op, out_len, out_len, 1, peak_v + p, peak_i + p));
    }
    if (want_peaks && out) ADSP_TRY(peak_device<T>(ctx, out, out_len, out_stride, pairs, peak_v, peak_i));
    return ADSP_OK;
}
}  // namespace

extern "C" {

adsp_status adsp_correlate_batch_device(adsp_ctx *ctx, const void *a, int64_t n, int64_t a_stride, const void *b, int64_t m,
                                        int64_t b_stride, int64_t pairs, void *out, int64_t out_stride,
                                        void *peak_index_dev, void *peak_value_dev, adsp_precision prec) {
    if (!ctx) return ADSP_ERR_INVALID_ARG;
    if (n <= 0 || m <= 0) return ADSP_ERR_EMPTY_INPUT;
    if (pairs <= 0) return ADSP_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    if (prec == ADSP_F64)
        return correlate_batch_dev<double>(ctx, (const double *)a, n, a_stride, (const double *)b, m, b_stride, pairs,
                                           (double *)out, out_stride, (long long *)peak_index_dev, (double *)peak_value_dev);
    return correlate_batch_dev<float>(ctx, (const float *)a, n, a_stride, (const float *)b, m, b_stride, pairs, (float *)out,
                                      out_stride, (long long *)peak_index_dev, (float *)peak_value_dev);
}

adsp_status adsp_correlate_batch(adsp_ctx *ctx, const double *a, int64_t n, int64_t a_stride, const double *b, int64_t m,
                                 int64_t b_stride, int64_t pairs, double *out, int64_t out_stride, int64_t *peak_index,
                                 double *peak_value) {
    if (!ctx) return ADSP_ERR_INVALID_ARG;
    if (n <= 0 || m <= 0) return ADSP_ERR_EMPTY_INPUT;
    if (pairs <= 0) return ADSP_OK;
    const int64_t out_len = n + m - 1;
    void *da = nullptr, *db = nullptr, *dout = nullptr, *dpi = nullptr, *dpv = nullptr;
    adsp_status st = ADSP_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);   // the staged transfers use the context's pinned slots
    ADSP_CUDA(cudaSetDevice(ctx->device));
    do {
        if ((st = adsp_device_alloc(ctx, (size_t)n * pairs * 8, &da)) != ADSP_OK) break;
        if ((st = adsp_device_alloc(ctx, (size_t)m * pairs * 8, &db)) != ADSP_OK) break;
        if (out && (st = adsp_device_alloc(ctx, (size_t)out_len * pairs * 8, &dout)) != ADSP_OK) break;
        if ((st = adsp_device_alloc(ctx, (size_t)pairs * 8, &dpi)) != ADSP_OK) break;
        if ((st = adsp_device_alloc(ctx, (size_t)pairs * 8, &dpv)) != ADSP_OK) break;
        if ((st = upload2d(ctx, da, (size_t)n * 8, a, (size_t)a_stride * 8, (size_t)n * 8, (size_t)pairs)) != ADSP_OK) break;
        if ((st = upload2d(ctx, db, (size_t)m * 8, b, (size_t)b_stride * 8, (size_t)m * 8, (size_t)pairs)) != ADSP_OK) break;
        st = correlate_batch_dev<double>(ctx, (const double *)da, n, n, (const double *)db, m, m, pairs, (double *)dout, out_len,
                                         (peak_index && peak_value) ? (long long *)dpi : nullptr, (peak_index && peak_value) ? (double *)dpv : nullptr);
        if (st != ADSP_OK) break;
        if (out && (st = download2d(ctx, out, (size_t)out_stride * 8, dout, (size_t)out_len * 8, (size_t)out_len * 8, (size_t)pairs)) != ADSP_OK) break;
        if (peak_index && peak_value) {
            cudaMemcpyAsync(peak_index, dpi, (size_t)pairs * 8, cudaMemcpyDeviceToHost, ctx->main);
            cudaMemcpyAsync(peak_value, dpv, (size_t)pairs * 8, cudaMemcpyDeviceToHost, ctx->main);
        }
        cudaError_t e = cudaStreamSynchronize(ctx->main);
        if (e != cudaSuccess) st = cuda_fail(e, "sync", __FILE__, __LINE__);
    } while (0);
    adsp_device_free(ctx, da); adsp_device_free(ctx, db); adsp_device_free(ctx, dout); adsp_device_free(ctx, dpi); adsp_device_free(ctx, dpv);
    return st;
}

// ---------------------------------------------------------------- plans
adsp_status adsp_overlap_save_create(adsp_ctx *ctx, const void *kernel, int64_t K, int64_t fft_size, adsp_precision prec, adsp_plan **out) {
    if (prec == ADSP_F64) return plan_create<double>(ctx, PLAN_OLS, (const double *)kernel, K, fft_size, out);
    return plan_create<float>(ctx, PLAN_OLS, (const float *)kernel, K, fft_size, out);
}
adsp_status adsp_overlap_add_create(adsp_ctx *ctx, const void *kernel, int64_t K, int64_t block_size, adsp_precision prec, adsp_plan **out) {
    if (prec == ADSP_F64) return plan_create<double>(ctx, PLAN_OLA, (const double *)kernel, K, block_size, out);
    return plan_create<float>(ctx, PLAN_OLA, (const float *)kernel, K, block_size, out);
}

void adsp_plan_destroy(adsp_plan *p) {
    if (!p) return;
    // finalizers run on arbitrary threads (runtime.SetFinalizer in the Go shim): serialise with calls in flight on the context
    std::lock_guard<std::mutex> lk(p->ctx->mu);
    plan_free(p);
}

// caller holds the context lock (error paths of the constructors, adsp_plan_destroy)
static void plan_free(adsp_plan *p) {
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->main);
    for (auto &f : p->fc64) f.destroy();
    for (auto &f : p->fc32) f.destroy();
    for (auto &g : p->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    for (auto &kv : p->extra64) kv.second.destroy();
    for (auto &kv : p->extra32) kv.second.destroy();
    if (p->fdl) fdl_destroy(p->fdl);
    p->d_kernel.release();
    p->hist[0].release(); p->hist[1].release(); p->blk_in.release(); p->blk_out.release();
    delete p;
}

int64_t adsp_plan_kernel_len(const adsp_plan *p) { return p ? p->K : 0; }
int64_t adsp_plan_fft_size(const adsp_plan *p) { return p ? p->ref_fft : 0; }
int64_t adsp_plan_step_size(const adsp_plan *p) { return p ? p->ref_step : 0; }
int64_t adsp_plan_block_size(const adsp_plan *p) { return p ? p->ref_block : 0; }
void adsp_plan_internal_geometry(const adsp_plan *p, int64_t *fft_n, int64_t *n1, int64_t *n2, int64_t *step, int64_t *parts) {
    if (!p) return;
    if (fft_n) *fft_n = p->ch.N;
    if (n1) *n1 = p->ch.N1;
    if (n2) *n2 = p->ch.N2;
    if (step) *step = p->ch.S;
    if (parts) *parts = p->ch.parts;
}

int adsp_plan_describe_cover(const adsp_plan *p, int64_t n, int64_t *out4, int cap) {
    if (!p || n <= 0) return 0;
    const long long out_len = n + p->K - 1;
    std::vector<Segment> segs;
    if (p->ch.parts == 1) segs = plan_segments(out_len, p->K, p->ch);
    else segs.push_back({p->ch.N, 0, out_len, false});   // partitioned IR: every partition runs the plan's transform
    int k = 0;
    for (const Segment &sg : segs) {
        if (out4 && k < cap) { out4[4 * k] = sg.N; out4[4 * k + 1] = sg.off; out4[4 * k + 2] = sg.len; out4[4 * k + 3] = sg.no_discard ? 1 : 0; }
        k++;
    }
    return k;
}

static adsp_status plan_run_device_any(adsp_plan *p, const void *in, int64_t n, int64_t channels, int64_t in_stride,
                                       void *out, int64_t out_stride) {
    if (p->prec == ADSP_F64) return plan_run_device<double>(p, (const double *)in, n, channels, in_stride, (double *)out, out_stride);
    return plan_run_device<float>(p, (const float *)in, n, channels, in_stride, (float *)out, out_stride);
}

adsp_status adsp_plan_process_device(adsp_plan *p, const void *in, int64_t n, int64_t channels, int64_t in_stride,
                                     void *out, int64_t out_stride) {
    if (!p || p->kind == PLAN_PART || p->kind == PLAN_STREAM) return ADSP_ERR_INVALID_ARG;
    if (n <= 0) return ADSP_ERR_EMPTY_INPUT;
    if (channels <= 0) return ADSP_OK;
    if (!in || !out) return ADSP_ERR_INVALID_ARG;
    adsp_ctx *ctx = p->ctx;
    // enqueue under the context lock: the launch sequence mutates ctx->scratch, the table maps, the fork/join events
    // and p->graphs, all shared with the host-pointer entry points (the lock is NOT held across any device wait)
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    // A repeated call (same buffers and sizes) replays a captured CUDA graph of the whole launch
    // sequence (kernels on the worker streams, fork/join events): ~200 launches become one submission.
    // First call runs eagerly (allocations, spectra, attributes), second call captures, later calls replay.
    static const bool use_graphs = env_ll("ADSP_GRAPHS", 1) != 0;
    if (!use_graphs || ctx->timing) return plan_run_device_any(p, in, n, channels, in_stride, out, out_stride);
    adsp_plan::GraphEntry *ent = nullptr;
    for (auto &g : p->graphs)
        if (g.in == in && g.out == out && g.n == n && g.channels == channels && g.in_stride == in_stride && g.out_stride == out_stride) { ent = &g; break; }
    if (ent && ent->exec && ent->alloc_gen != g_alloc_generation.load(std::memory_order_relaxed)) {
        // a shared device buffer (scratch, staging) was re-allocated since the capture: the graph's baked-in pointers may
        // be stale.  Drop it; this call runs eagerly (and re-reserves what it needs), the next one captures again.
        cudaGraphExecDestroy(ent->exec);
        p->graphs.erase(p->graphs.begin() + (ent - p->graphs.data()));
        ent = nullptr;
    }
    if (ent && ent->exec) {
        ADSP_CUDA(cudaGraphLaunch(ent->exec, ctx->main));
        count_launch(ctx, (int)ent->launches);
        return ADSP_OK;
    }
    if (!ent) {
        if (p->graphs.size() >= 8) {   // small cache: drop the oldest
            if (p->graphs.front().exec) cudaGraphExecDestroy(p->graphs.front().exec);
            p->graphs.erase(p->graphs.begin());
        }
        p->graphs.push_back({in, out, n, channels, in_stride, out_stride, nullptr, 0, 1, 0});
        return plan_run_device_any(p, in, n, channels, in_stride, out, out_stride);
    }
    // second sighting: capture
    const uint64_t l0 = ctx->launches.load();
    const uint64_t gen0 = g_alloc_generation.load(std::memory_order_relaxed);
    cudaGraph_t graph = nullptr;
    ADSP_CUDA(cudaStreamBeginCapture(ctx->main, cudaStreamCaptureModeThreadLocal));
    adsp_status st = plan_run_device_any(p, in, n, channels, in_stride, out, out_stride);
    cudaError_t ce = cudaStreamEndCapture(ctx->main, &graph);
    // an allocation during the capture (a buffer grew) means pointers recorded before it are stale: do not keep the graph
    const bool moved = g_alloc_generation.load(std::memory_order_relaxed) != gen0;
    if (st != ADSP_OK || ce != cudaSuccess || !graph || moved) {
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        p->graphs.erase(p->graphs.begin() + (ent - p->graphs.data()));
        return plan_run_device_any(p, in, n, channels, in_stride, out, out_stride);
    }
    cudaGraphExec_t exec = nullptr;
    ce = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) { cudaGetLastError(); return plan_run_device_any(p, in, n, channels, in_stride, out, out_stride); }
    ent->exec = exec;
    ent->alloc_gen = gen0;
    ent->launches = ctx->launches.load() - l0;
    ctx->launches.store(l0);
    ADSP_CUDA(cudaGraphLaunch(ent->exec, ctx->main));
    count_launch(ctx, (int)ent->launches);
    return ADSP_OK;
}

adsp_status adsp_plan_process_batch(adsp_plan *p, const void *in, int64_t n, int64_t channels, int64_t in_stride,
                                    void *out, int64_t out_stride) {
    if (!p || p->kind == PLAN_PART || p->kind == PLAN_STREAM) return ADSP_ERR_INVALID_ARG;
    if (n <= 0) return ADSP_ERR_EMPTY_INPUT;        // overlap_save.go:127-129
    if (channels <= 0) return ADSP_OK;
    if (!in || !out) return ADSP_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(p->ctx->mu);
    ADSP_CUDA(cudaSetDevice(p->ctx->device));
    if (p->prec == ADSP_F64) return plan_run_host<double>(p, (const double *)in, n, channels, in_stride, (double *)out, out_stride);
    return plan_run_host<float>(p, (const float *)in, n, channels, in_stride, (float *)out, out_stride);
}

adsp_status adsp_plan_process(adsp_plan *p, const void *in, int64_t n, void *out, int64_t out_len) {
    if (!p) return ADSP_ERR_INVALID_ARG;
    if (out_len != n + p->K - 1 && n > 0) {           // overlap_save.go:259-262
        set_error("conv: buffer length mismatch: expected " + std::to_string(n + p->K - 1) + ", got " + std::to_string(out_len));
        return ADSP_ERR_LENGTH_MISMATCH;
    }
    return adsp_plan_process_batch(p, in, n, 1, n, out, out_len);
}

adsp_status adsp_plan_sync(adsp_plan *p) { return p ? adsp_ctx_sync(p->ctx) : ADSP_ERR_INVALID_ARG; }

// ---------------------------------------------------------------- several GPUs from one host process
// Shards are independent: contiguous channel ranges, or time blocks of one long signal with a (K-1) halo read from
// the source (the rule of overlap_save.go:205-215 at shard granularity; the last shard also emits the K-1 tail,
// :224-251).  No collective; one host thread per plan (= per GPU) drives that plan's own streams.
void adsp_shard_channel_range(int64_t channels, int rank, int world, int64_t *lo, int64_t *hi) {
    if (world < 1) world = 1;
    const int64_t base = channels / world, rem = channels % world;
    const int64_t l = rank * base + std::min<int64_t>(rank, rem);
    if (lo) *lo = l;
    if (hi) *hi = l + base + (rank < rem ? 1 : 0);
}

void adsp_shard_time(int64_t n, int64_t kernel_len, int rank, int world, int64_t *out_lo, int64_t *out_hi, int64_t *in_lo,
                     int64_t *in_hi, int64_t *skip) {
    if (world < 1) world = 1;
    const int64_t out_len = n + kernel_len - 1;
    int64_t S = (n + world - 1) / world;
    S = (S + 31) / 32 * 32;                                                    // 256-byte aligned shard starts
    const int64_t last_rank = std::min<int64_t>(world - 1, S > 0 ? (n + S - 1) / S - 1 : 0);   // owns the K-1 tail
    int64_t olo = out_len, ohi = out_len, ilo = n, ihi = n, sk = 0;              // ranks past the signal own nothing
    if (rank <= last_rank) {
        const int64_t lo = std::min<int64_t>((int64_t)rank * S, n), hi = std::min<int64_t>((int64_t)(rank + 1) * S, n);
        olo = lo;
        ohi = (rank == last_rank) ? out_len : hi;
        ilo = std::max<int64_t>(0, lo - (kernel_len - 1));
        ihi = hi;
        sk = lo - ilo;
    }
    if (out_lo) *out_lo = olo;
    if (out_hi) *out_hi = ohi;
    if (in_lo) *in_lo = ilo;
    if (in_hi) *in_hi = ihi;
    if (skip) *skip = sk;
}

adsp_status adsp_plans_process_batch(adsp_plan *const *plans, int nplans, const void *in, int64_t n, int64_t channels,
                                     int64_t in_stride, void *out, int64_t out_stride) {
    if (!plans || nplans < 1) return ADSP_ERR_INVALID_ARG;
    for (int i = 0; i < nplans; i++)
        if (!plans[i] || plans[i]->prec != plans[0]->prec || plans[i]->K != plans[0]->K) return ADSP_ERR_INVALID_ARG;
    if (n <= 0) return ADSP_ERR_EMPTY_INPUT;
    const size_t es = plans[0]->prec == ADSP_F64 ? 8 : 4;
    std::vector<adsp_status> st((size_t)nplans, ADSP_OK);
    std::vector<std::thread> th;
    for (int r = 0; r < nplans; r++) {
        int64_t lo, hi;
        adsp_shard_channel_range(channels, r, nplans, &lo, &hi);
        if (hi <= lo) continue;
        th.emplace_back([=, &st] {
            st[(size_t)r] = adsp_plan_process_batch(plans[r], (const char *)in + (size_t)lo * (size_t)in_stride * es, n, hi - lo, in_stride,
                                                    (char *)out + (size_t)lo * (size_t)out_stride * es, out_stride);
        });
    }
    for (auto &t : th) t.join();
    for (adsp_status s : st) if (s != ADSP_OK) return s;
    return ADSP_OK;
}

adsp_status adsp_plans_process_long(adsp_plan *const *plans, int nplans, const void *in, int64_t n, void *out, int64_t out_len) {
    if (!plans || nplans < 1) return ADSP_ERR_INVALID_ARG;
    for (int i = 0; i < nplans; i++)
        if (!plans[i] || plans[i]->prec != plans[0]->prec || plans[i]->K != plans[0]->K) return ADSP_ERR_INVALID_ARG;
    if (n <= 0) return ADSP_ERR_EMPTY_INPUT;
    const int64_t K = plans[0]->K;
    if (out_len != n + K - 1) { set_error("conv: buffer length mismatch"); return ADSP_ERR_LENGTH_MISMATCH; }
    const size_t es = plans[0]->prec == ADSP_F64 ? 8 : 4;
    std::vector<adsp_status> st((size_t)nplans, ADSP_OK);
    std::vector<std::thread> th;
    for (int r = 0; r < nplans; r++) {
        int64_t olo, ohi, ilo, ihi, skip;
        adsp_shard_time(n, K, r, nplans, &olo, &ohi, &ilo, &ihi, &skip);
        if (ohi <= olo) continue;
        th.emplace_back([=, &st] {
            const int64_t seg = ihi - ilo;
            std::vector<char> tmp((size_t)(seg + K - 1) * es);          // full convolution of the shard's segment (halo included)
            adsp_status s = adsp_plan_process(plans[r], (const char *)in + (size_t)ilo * es, seg, tmp.data(), seg + K - 1);
            if (s == ADSP_OK) memcpy((char *)out + (size_t)olo * es, tmp.data() + (size_t)skip * es, (size_t)(ohi - olo) * es);
            st[(size_t)r] = s;
        });
    }
    for (auto &t : th) t.join();
    for (adsp_status s : st) if (s != ADSP_OK) return s;
    return ADSP_OK;
}

}  // extern "C"

// ================================================================ partitioned (streaming) convolution
namespace {

int trunc_log2(long long n) { int r = 0; while (n > 1) { n >>= 1; r++; } return r; }
long long bits_upto(int n) { return (2LL << n) - 1; }

// Stage layout of the reference's non-uniform partitioning (reported by StageCount/StageInfo).
// Follows the arithmetic of partitionIR, dsp/conv/partitioned.go:269-332.
std::vector<PartStage> partition_layout(long long kernel_len_padded, int min_order, int max_order) {
    const long long min_block = 1LL << min_order;
    int max_ir = trunc_log2(kernel_len_padded + min_block) - 1;
    long long res = kernel_len_padded - (bits_upto(max_ir) - bits_upto(min_order - 1));
    if (res > 0 && ((res >> max_ir) & 1) == 0 && max_ir > min_order) max_ir--;
    if (max_ir > max_order) max_ir = max_order;
    res = kernel_len_padded - (bits_upto(max_ir) - bits_upto(min_order - 1));
    std::vector<PartStage> st;
    long long start = 0;
    for (int order = min_order; order < max_ir; order++) {
        const int count = 1 + (int)((res >> order) & 1);
        st.push_back({1 << order, count, (int)start});
        start += (long long)count << order;
        res -= (long long)(count - 1) << order;
    }
    int count = 1;
    if (max_ir > 0) count = (int)std::max<long long>(1, 1 + res / (1LL << max_ir));
    st.push_back({1 << max_ir, count, (int)start});
    return st;
}

// ProcessBlock semantics (partitioned.go:348-396): output sample t of the stream equals the full
// linear convolution at t - latency (zero before).  The device keeps the last K-1+latency input
// samples; a call with n samples evaluates n "valid" outputs from [history | block].
template <typename T>
adsp_status part_process(adsp_plan *p, const T *in, int64_t n, T *out) {
    adsp_ctx *ctx = p->ctx;
    const long long HL = p->hist_len;  // K - 1 + latency
    ADSP_TRY(p->blk_in.reserve((size_t)(HL + n) * sizeof(T)));
    ADSP_TRY(p->blk_out.reserve((size_t)n * sizeof(T)));
    T *buf = (T *)p->blk_in.p;
    T *hist = (T *)p->hist[p->hist_cur].p;
    T *hist_next = (T *)p->hist[p->hist_cur ^ 1].p;
    ADSP_CUDA(cudaMemcpyAsync(buf, hist, (size_t)HL * sizeof(T), cudaMemcpyDeviceToDevice, ctx->main));
    ADSP_CUDA(cudaMemcpyAsync(buf + HL, in, (size_t)n * sizeof(T), cudaMemcpyHostToDevice, ctx->main));
    // outputs y[t0 + o], o < n, t0 = pos - latency; buf[0] is x[t0 - (K-1)]
    auto &v = plan_fc<T>(p);
    T *dout = (T *)p->blk_out.p;
    if (v.size() == 1) {
        ADSP_TRY(v[0].run(buf, HL + n, 1, 0, dout, 0, n, p->K - 1, 0, false));
    } else {
        ADSP_CUDA(cudaMemsetAsync(dout, 0, (size_t)n * sizeof(T), ctx->main));
        for (size_t i = 0; i < v.size(); i++) {
            const long long k0 = (long long)i * p->ch.part_len;
            // partition i sees the input delayed by k0: shift the window start back by k0
            ADSP_TRY(v[i].run(buf, HL + n, 1, 0, dout, 0, n, p->K - 1 - k0, 0, true));
        }
    }
    ADSP_CUDA(cudaMemcpyAsync(out, dout, (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, ctx->main));
    ADSP_CUDA(cudaMemcpyAsync(hist_next, buf + n, (size_t)HL * sizeof(T), cudaMemcpyDeviceToDevice, ctx->main));
    p->hist_cur ^= 1;
    ADSP_CUDA(cudaStreamSynchronize(ctx->main));
    return ADSP_OK;
}

template <typename T>
adsp_status part_create(adsp_ctx *ctx, const T *kernel, int64_t K, int min_order, int max_order, int channels, adsp_plan **out) {
    if (!out) return ADSP_ERR_INVALID_ARG;
    *out = nullptr;
    if (!ctx) return ADSP_ERR_INVALID_ARG;
    if (K <= 0 || !kernel) return ADSP_ERR_EMPTY_IR;                          // partitioned.go:215-217
    if (min_order < 1) { set_error("conv: invalid block order: minBlockOrder must be >= 1"); return ADSP_ERR_INVALID_BLOCK_ORDER; }
    if (max_order < min_order) { set_error("conv: invalid block order: maxBlockOrder must be >= minBlockOrder"); return ADSP_ERR_INVALID_BLOCK_ORDER; }
    if (min_order > 24) { set_error("conv: invalid block order: too large"); return ADSP_ERR_INVALID_BLOCK_ORDER; }
    if (channels < 1) { set_error("partitioned: channels must be >= 1"); return ADSP_ERR_INVALID_ARG; }
    if (channels > 131070) { set_error("partitioned: at most 131070 channels per plan (65535 channel pairs per launch); split the batch over several plans"); return ADSP_ERR_INVALID_ARG; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    adsp_plan *p = new adsp_plan();
    p->ctx = ctx; p->kind = PLAN_PART; p->prec = sizeof(T) == 8 ? ADSP_F64 : ADSP_F32; p->K = K; p->channels = channels;
    p->latency = 1 << min_order; p->min_order = min_order; p->max_order = max_order;
    const long long padded = ((K + p->latency - 1) / p->latency) * p->latency;
    p->stages = partition_layout(padded, min_order, max_order);
    p->hist_len = K - 1 + p->latency;
    adsp_status st = ADSP_OK;
    const bool use_fdl = fdl_supported(min_order) && env_ll("ADSP_NO_FDL", 0) == 0;
    if (use_fdl) {
        // streaming engine on a frequency-domain delay line (fdl.cu); the IR only has to reach the device
        st = p->d_kernel.reserve((size_t)K * sizeof(T));
        if (st == ADSP_OK && cudaMemcpyAsync(p->d_kernel.p, kernel, (size_t)K * sizeof(T), cudaMemcpyHostToDevice, ctx->main) != cudaSuccess) st = ADSP_ERR_CUDA;
        if (st == ADSP_OK && cudaStreamSynchronize(ctx->main) != cudaSuccess) st = ADSP_ERR_CUDA;
        p->ch = choose_fft(K);
        if (st == ADSP_OK) st = fdl_create(ctx, p->d_kernel.p, K, min_order, max_order, channels, p->prec, &p->fdl);
    } else {
        if (channels != 1) { set_error("partitioned: multi-channel plans need minBlockOrder >= 3"); st = ADSP_ERR_INVALID_ARG; }
        if (st == ADSP_OK) st = plan_build<T>(p, kernel);
        for (int i = 0; i < 2 && st == ADSP_OK; i++) {
            st = p->hist[i].reserve((size_t)p->hist_len * sizeof(T));
            if (st == ADSP_OK && cudaMemsetAsync(p->hist[i].p, 0, (size_t)p->hist_len * sizeof(T), ctx->main) != cudaSuccess) st = ADSP_ERR_CUDA;
        }
    }
    if (st != ADSP_OK) { plan_free(p); return st; }
    *out = p;
    return ADSP_OK;
}

}  // namespace

extern "C" {

adsp_status adsp_partitioned_create(adsp_ctx *ctx, const void *kernel, int64_t K, int min_order, int max_order,
                                    adsp_precision prec, adsp_plan **out) {
    if (prec == ADSP_F64) return part_create<double>(ctx, (const double *)kernel, K, min_order, max_order, 1, out);
    return part_create<float>(ctx, (const float *)kernel, K, min_order, max_order, 1, out);
}

adsp_status adsp_partitioned_create_batch(adsp_ctx *ctx, const void *kernel, int64_t K, int min_order, int max_order, int channels,
                                          adsp_precision prec, adsp_plan **out) {
    if (prec == ADSP_F64) return part_create<double>(ctx, (const double *)kernel, K, min_order, max_order, channels, out);
    return part_create<float>(ctx, (const float *)kernel, K, min_order, max_order, channels, out);
}

static adsp_status part_batch(adsp_plan *p, const void *in, int64_t n, int64_t in_stride, void *out, int64_t out_stride, bool host, bool mix) {
    if (!p || p->kind != PLAN_PART || !p->fdl) return ADSP_ERR_INVALID_ARG;
    if (n <= 0) return ADSP_OK;
    if (!in || !out) return ADSP_ERR_INVALID_ARG;
    if (p->channels > 1 && (in_stride < n || out_stride < n)) { set_error("partitioned: stride shorter than the block"); return ADSP_ERR_INVALID_ARG; }
    std::lock_guard<std::mutex> lk(p->ctx->mu);
    ADSP_CUDA(cudaSetDevice(p->ctx->device));
    return fdl_process(p->fdl, in, n, in_stride, out, out_stride, host, mix);
}

adsp_status adsp_partitioned_process_block(adsp_plan *p, const void *in, int64_t n, void *out, int64_t n_out) {
    if (!p || p->kind != PLAN_PART) return ADSP_ERR_INVALID_ARG;
    if (n != n_out) {                                                        // partitioned.go:349-352
        set_error("conv: buffer length mismatch: input length " + std::to_string(n) + " != output length " + std::to_string(n_out));
        return ADSP_ERR_LENGTH_MISMATCH;
    }
    if (n <= 0) return ADSP_OK;
    if (!in || !out) return ADSP_ERR_INVALID_ARG;
    if (p->fdl) {
        if (p->channels != 1) { set_error("partitioned: use the batch entry points on a multi-channel plan"); return ADSP_ERR_INVALID_ARG; }
        return part_batch(p, in, n, n, out, n, true, false);
    }
    std::lock_guard<std::mutex> lk(p->ctx->mu);
    ADSP_CUDA(cudaSetDevice(p->ctx->device));
    if (p->prec == ADSP_F64) return part_process<double>(p, (const double *)in, n, (double *)out);
    return part_process<float>(p, (const float *)in, n, (float *)out);
}

adsp_status adsp_partitioned_process_block_batch(adsp_plan *p, const void *in, int64_t n, int64_t in_stride, void *out, int64_t out_stride) {
    return part_batch(p, in, n, in_stride, out, out_stride, true, false);
}
adsp_status adsp_partitioned_process_block_batch_device(adsp_plan *p, const void *in, int64_t n, int64_t in_stride, void *out,
                                                        int64_t out_stride) {
    return part_batch(p, in, n, in_stride, out, out_stride, false, false);
}
adsp_status adsp_partitioned_set_wet_dry(adsp_plan *p, double wet, double dry) {
    if (!p || p->kind != PLAN_PART || !p->fdl) return ADSP_ERR_INVALID_ARG;
    fdl_set_wet_dry(p->fdl, wet, dry);
    return ADSP_OK;
}
adsp_status adsp_partitioned_process_in_place_batch(adsp_plan *p, void *block, int64_t n, int64_t stride) {
    return part_batch(p, block, n, stride, block, stride, true, true);
}
adsp_status adsp_partitioned_process_in_place_batch_device(adsp_plan *p, void *block, int64_t n, int64_t stride) {
    return part_batch(p, block, n, stride, block, stride, false, true);
}
// stage layout the delay-line engine would use for (K, minBlockOrder, maxBlockOrder); no GPU needed (host arithmetic)
int adsp_partitioned_plan_layout(int64_t kernel_len, int min_order, int max_order, int *part_size, int *count, int64_t *ir_offset, int cap) {
    if (kernel_len <= 0 || !fdl_supported(min_order) || max_order < min_order) return 0;
    const std::vector<FdlStage> st = fdl_layout(kernel_len, min_order, max_order);
    for (int i = 0; i < (int)st.size() && i < cap; i++) {
        if (part_size) part_size[i] = st[(size_t)i].part_size;
        if (count) count[i] = st[(size_t)i].count;
        if (ir_offset) ir_offset[i] = st[(size_t)i].ir_offset;
    }
    return (int)st.size();
}
int adsp_partitioned_channels(const adsp_plan *p) { return (p && p->kind == PLAN_PART) ? p->channels : 0; }
int adsp_partitioned_internal_stage_count(const adsp_plan *p) { return (p && p->fdl) ? fdl_stage_count(p->fdl) : 0; }
adsp_status adsp_partitioned_internal_stage_info(const adsp_plan *p, int index, int *part_size, int *count, int64_t *ir_offset) {
    if (!p || !p->fdl || index < 0 || index >= fdl_stage_count(p->fdl)) return ADSP_ERR_STAGE_INDEX;
    long long off = 0;
    fdl_stage_info(p->fdl, index, part_size, count, &off);
    if (ir_offset) *ir_offset = off;
    return ADSP_OK;
}

void adsp_plan_reset(adsp_plan *p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(p->ctx->mu);
    if (p->fdl) { cudaSetDevice(p->ctx->device); fdl_reset(p->fdl); return; }   // partitioned.go:399-407
    if (p->kind == PLAN_PART || p->kind == PLAN_STREAM) {                    // partitioned.go:399-407, streaming_overlap_save.go:167-169
        cudaSetDevice(p->ctx->device);
        const size_t es = p->prec == ADSP_F64 ? 8 : 4;
        for (int i = 0; i < 2; i++) cudaMemsetAsync(p->hist[i].p, 0, (size_t)p->hist_len * es, p->ctx->main);
        cudaStreamSynchronize(p->ctx->main);
    }
    // OverlapSave/OverlapAdd keep no state across Process calls (overlap_save.go:136-138, overlap_add.go:185-187)
}

// ---------------------------------------------------------------- fixed-block streaming convolvers
// NewStreamingOverlapAdd / NewStreamingOverlapSave (streaming_overlap_add.go:41, streaming_overlap_save.go:44):
// ProcessBlock(input[blockSize]) -> output[blockSize], state carried across calls, no latency.  Both
// algorithms produce the same stream (streaming_test.go:122-176), so one device engine serves both:
// the partitioned engine with latency 0 (history = last K-1 input samples on the device).
static adsp_status stream_create(adsp_ctx *ctx, const void *kernel, int64_t K, int64_t block_size, adsp_precision prec, adsp_plan **out) {
    if (!out) return ADSP_ERR_INVALID_ARG;
    *out = nullptr;
    if (!ctx) return ADSP_ERR_INVALID_ARG;
    if (K <= 0 || !kernel) return ADSP_ERR_EMPTY_KERNEL;
    if (block_size <= 0) { set_error("conv: blockSize must be positive, got " + std::to_string(block_size)); return ADSP_ERR_INVALID_ARG; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    ADSP_CUDA(cudaSetDevice(ctx->device));
    adsp_plan *p = new adsp_plan();
    p->ctx = ctx; p->kind = PLAN_STREAM; p->prec = prec; p->K = K;
    p->latency = 0;
    p->ref_block = block_size;
    p->ref_fft = adsp_next_pow2(block_size + K - 1);          // streaming_overlap_save.go:57-58
    p->hist_len = K - 1;
    const size_t es = prec == ADSP_F64 ? 8 : 4;
    adsp_status st = prec == ADSP_F64 ? plan_build<double>(p, (const double *)kernel) : plan_build<float>(p, (const float *)kernel);
    for (int i = 0; i < 2 && st == ADSP_OK; i++) {
        st = p->hist[i].reserve((size_t)std::max<long long>(p->hist_len, 1) * es);
        if (st == ADSP_OK && cudaMemsetAsync(p->hist[i].p, 0, (size_t)std::max<long long>(p->hist_len, 1) * es, ctx->main) != cudaSuccess) st = ADSP_ERR_CUDA;
    }
    if (st != ADSP_OK) { plan_free(p); return st; }
    *out = p;
    return ADSP_OK;
}

adsp_status adsp_streaming_create(adsp_ctx *ctx, const void *kernel, int64_t K, int64_t block_size, int overlap_save,
                                  adsp_precision prec, adsp_plan **out) {
    (void)overlap_save;   // same results either way; kept so the binding can mirror both constructors
    return stream_create(ctx, kernel, K, block_size, prec, out);
}

adsp_status adsp_streaming_process_block(adsp_plan *p, const void *in, int64_t n, void *out, int64_t n_out) {
    if (!p || p->kind != PLAN_STREAM) return ADSP_ERR_INVALID_ARG;
    if (n != p->ref_block) {                                   // streaming_overlap_save.go:139-141
        set_error("conv: buffer length mismatch: expected " + std::to_string(p->ref_block) + " samples, got " + std::to_string(n));
        return ADSP_ERR_LENGTH_MISMATCH;
    }
    if (n_out != p->ref_block) {                               // streaming_overlap_save.go:156-158
        set_error("conv: buffer length mismatch: expected " + std::to_string(p->ref_block) + " output samples, got " + std::to_string(n_out));
        return ADSP_ERR_LENGTH_MISMATCH;
    }
    if (!in || !out) return ADSP_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(p->ctx->mu);
    ADSP_CUDA(cudaSetDevice(p->ctx->device));
    if (p->prec == ADSP_F64) return part_process<double>(p, (const double *)in, n, (double *)out);
    return part_process<float>(p, (const float *)in, n, (float *)out);
}

int adsp_partitioned_latency(const adsp_plan *p) { return p ? p->latency : 0; }
int adsp_partitioned_stage_count(const adsp_plan *p) { return p ? (int)p->stages.size() : 0; }
adsp_status adsp_partitioned_stage_info(const adsp_plan *p, int index, int *part_size, int *block_count) {
    if (!p) return ADSP_ERR_INVALID_ARG;
    if (index < 0 || index >= (int)p->stages.size()) {
        set_error("conv: stage index out of range: index " + std::to_string(index) + ", have " + std::to_string(p->stages.size()) + " stages");
        return ADSP_ERR_STAGE_INDEX;
    }
    if (part_size) *part_size = p->stages[(size_t)index].part_size;
    if (block_count) *block_count = p->stages[(size_t)index].count;
    return ADSP_OK;
}

}  // extern "C"
