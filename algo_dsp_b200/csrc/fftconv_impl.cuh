// fftconv_impl.cuh -- FftConv<T>: table construction, IR-spectrum caching and kernel dispatch.
// Included by fftconv_f64.cu / fftconv_f32.cu with ADSP_REAL defined (one TU per precision so
// the template-heavy kernels compile in parallel).
#pragma once
#include <algorithm>
#include <cmath>

#include "engine.cuh"

namespace adsp {

// ------------------------------------------------------------------ twiddle tables (host, long double)
static inline void unit_root(long long num, long long den, long double *re, long double *im) {
    // exp(-2*pi*i*num/den), exact octant reduction so that table entries are correctly rounded
    num %= den;
    const long double PI = 3.14159265358979323846264338327950288L;
    // reduce to [0, den/8] using symmetries when den is a multiple of 8
    long double c, s;
    if (den % 8 == 0) {
        const long long e = den / 8;
        const long long oct = num / e;
        const long long rem = num % e;
        long double a;
        // angle within octant, measured from the nearest axis
        if (oct % 2 == 0) a = 2.0L * PI * (long double)rem / (long double)den;
        else a = 2.0L * PI * (long double)(e - rem) / (long double)den;
        const long double ca = cosl(a), sa = sinl(a);
        switch (oct) {
        case 0: c = ca;  s = sa;  break;
        case 1: c = sa;  s = ca;  break;
        case 2: c = -sa; s = ca;  break;
        case 3: c = -ca; s = sa;  break;
        case 4: c = -ca; s = -sa; break;
        case 5: c = -sa; s = -ca; break;
        case 6: c = sa;  s = -ca; break;
        default: c = ca; s = -sa; break;
        }
    } else {
        const long double a = 2.0L * PI * (long double)num / (long double)den;
        c = cosl(a); s = sinl(a);
    }
    *re = c;
    *im = -s;
}

template <typename T> adsp_status get_tw_table(adsp_ctx *ctx, int L, const cpx<T> **out) {
    const int prec = sizeof(T) == 8 ? 0 : 1;
    auto key = std::make_pair(L, prec);
    auto it = ctx->tw_tables.find(key);
    if (it != ctx->tw_tables.end()) { *out = (const cpx<T> *)it->second; return ADSP_OK; }
    int lg = 0;
    while ((1 << lg) < L) lg++;
    const int P = (lg - 1) / 4;
    const int R0 = L >> (4 * P);
    const int entries = tw_total_entries(R0, P);
    std::vector<cpx<T>> h((size_t)(entries > 0 ? entries : 1));
    int off = 0, ns = R0;
    for (int t = 1; t <= P; t++) {
        const int rmax = (ns <= 16) ? 15 : 1;  // compact layout, see tw_pass_entries()
        for (int r = 1; r <= rmax; r++)
            for (int k = 0; k < ns; k++) {
                long double re, im;
                unit_root((long long)r * k, 16LL * ns, &re, &im);
                h[(size_t)off + (size_t)(r - 1) * ns + k].x = (T)re;
                h[(size_t)off + (size_t)(r - 1) * ns + k].y = (T)im;
            }
        off += tw_pass_entries(ns);
        ns *= 16;
    }
    void *d = nullptr;
    ADSP_CUDA(cudaMalloc(&d, h.size() * sizeof(cpx<T>)));
    ADSP_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(cpx<T>), cudaMemcpyHostToDevice));
    ctx->tw_tables[key] = d;
    *out = (const cpx<T> *)d;
    return ADSP_OK;
}

// mixed-radix column twiddles W_(16M)^(j*km) at [km*16 + j]
template <typename T> adsp_status get_twp_table(adsp_ctx *ctx, int P, const cpx<T> **out) {
    const int prec = sizeof(T) == 8 ? 0 : 1;
    auto key = std::make_pair(-P, prec);
    auto it = ctx->tw_tables.find(key);
    if (it != ctx->tw_tables.end()) { *out = (const cpx<T> *)it->second; return ADSP_OK; }
    std::vector<cpx<T>> h((size_t)16 * P);
    for (int kp = 0; kp < P; kp++)
        for (int j = 0; j < 16; j++) {
            long double re, im;
            unit_root((long long)j * kp, 16LL * P, &re, &im);
            h[(size_t)kp * 16 + j].x = (T)re;
            h[(size_t)kp * 16 + j].y = (T)im;
        }
    void *d = nullptr;
    ADSP_CUDA(cudaMalloc(&d, h.size() * sizeof(cpx<T>)));
    ADSP_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(cpx<T>), cudaMemcpyHostToDevice));
    ctx->tw_tables[key] = d;
    *out = (const cpx<T> *)d;
    return ADSP_OK;
}

template <typename T> adsp_status get_tw4_tables(adsp_ctx *ctx, long long N, const cpx<T> **hi, const cpx<T> **lo) {
    const int prec = sizeof(T) == 8 ? 0 : 1;
    auto key = std::make_pair((int)(N >> 8), prec);
    auto it = ctx->tw4_tables.find(key);
    if (it != ctx->tw4_tables.end()) {
        *hi = (const cpx<T> *)it->second.first;
        *lo = (const cpx<T> *)it->second.second;
        return ADSP_OK;
    }
    const long long nhi = (N >> 10) > 0 ? (N >> 10) : 1;
    std::vector<cpx<T>> hhi((size_t)nhi), hlo(1024);
    for (long long a = 0; a < nhi; a++) {
        long double re, im;
        unit_root(a << 10, N, &re, &im);
        hhi[(size_t)a].x = (T)re; hhi[(size_t)a].y = (T)im;
    }
    for (long long b = 0; b < 1024; b++) {
        long double re, im;
        unit_root(b, N, &re, &im);
        hlo[(size_t)b].x = (T)re; hlo[(size_t)b].y = (T)im;
    }
    void *dhi = nullptr, *dlo = nullptr;
    ADSP_CUDA(cudaMalloc(&dhi, hhi.size() * sizeof(cpx<T>)));
    ADSP_CUDA(cudaMalloc(&dlo, hlo.size() * sizeof(cpx<T>)));
    ADSP_CUDA(cudaMemcpy(dhi, hhi.data(), hhi.size() * sizeof(cpx<T>), cudaMemcpyHostToDevice));
    ADSP_CUDA(cudaMemcpy(dlo, hlo.data(), hlo.size() * sizeof(cpx<T>), cudaMemcpyHostToDevice));
    ctx->tw4_tables[key] = std::make_pair(dhi, dlo);
    *hi = (const cpx<T> *)dhi;
    *lo = (const cpx<T> *)dlo;
    return ADSP_OK;
}

// ------------------------------------------------------------------ kernel dispatch
template <typename K> static adsp_status set_smem(K kern, size_t bytes) {
    if (bytes > 48 * 1024) ADSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return ADSP_OK;
}
// Grid for the flattened tile loops: one CTA per tile by default (the hardware block scheduler balances
// them); ADSP_TILES_PER_CTA = k > 1 makes CTAs walk k tiles (but never fewer CTAs than fill the machine).
static int persistent_grid(adsp_ctx *ctx, int ntiles, int ctas_per_sm) {
    const long long k = env_ll("ADSP_TILES_PER_CTA", 1);
    if (k <= 1) return ntiles;
    const long long resident = (long long)ctx->sm_count * ctas_per_sm;
    long long grid = (ntiles + k - 1) / k;
    if (grid < resident) grid = std::min<long long>(resident, ntiles);
    return (int)grid;
}

// the opt-in shared-memory attribute is per (function, device): remember it per device
struct AttrOnce {
    bool done[64] = {};
    bool need(int dev) { if (dev < 0 || dev >= 64) return true; if (done[dev]) return false; done[dev] = true; return true; }
};

struct OccPerDev {
    int v[64] = {};
    int *at(int dev) { return &v[(dev >= 0 && dev < 64) ? dev : 0]; }
};

template <typename T, int L, bool SPEC>
static adsp_status launch_full_t(adsp_ctx *ctx, cudaStream_t st, const ConvGeom &g, const T *x, T *y, const cpx<T> *H,
                                 cpx<T> *spec, T scale, const cpx<T> *tw, long long npairs) {
    constexpr int THREADS = rows_cta_threads(L);
    constexpr int ROWS = THREADS / FftShape<L>::TPF;
    const size_t smem = ((size_t)ROWS * L + FftShape<L>::TW_ENTRIES) * sizeof(cpx<T>);
    static AttrOnce once;
    if (once.need(ctx->device)) ADSP_TRY(set_smem(fftconv_full<T, L, SPEC>, smem));
    const long long grid = (npairs + ROWS - 1) / ROWS;
    LaunchTimer lt(ctx, st, KK_FULL);
    fftconv_full<T, L, SPEC><<<(unsigned)grid, THREADS, smem, st>>>(g, x, y, H, spec, scale, tw, npairs);
    count_launch(ctx);
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}

template <typename T, bool SPEC>
static adsp_status launch_full(adsp_ctx *ctx, cudaStream_t st, int L, const ConvGeom &g, const T *x, T *y,
                               const cpx<T> *H, cpx<T> *spec, T scale, const cpx<T> *tw, long long npairs) {
    switch (L) {
    case 256:  return launch_full_t<T, 256, SPEC>(ctx, st, g, x, y, H, spec, scale, tw, npairs);
    case 512:  return launch_full_t<T, 512, SPEC>(ctx, st, g, x, y, H, spec, scale, tw, npairs);
    case 1024: return launch_full_t<T, 1024, SPEC>(ctx, st, g, x, y, H, spec, scale, tw, npairs);
    case 2048: return launch_full_t<T, 2048, SPEC>(ctx, st, g, x, y, H, spec, scale, tw, npairs);
    case 4096: return launch_full_t<T, 4096, SPEC>(ctx, st, g, x, y, H, spec, scale, tw, npairs);
    default: set_error("unsupported single-kernel FFT length"); return ADSP_ERR_INVALID_ARG;
    }
}

template <typename T, int L, int SPEC>
static adsp_status launch_rows_t(adsp_ctx *ctx, cudaStream_t st, cpx<T> *scratch, const cpx<T> *H, cpx<T> *spec,
                                 T scale, int N1, const cpx<T> *tw, int pairs) {
    constexpr int THREADS = rows_cta_threads(L);
    constexpr int ROWS = THREADS / FftShape<L>::TPF;
    const size_t smem = ((size_t)ROWS * L + FftShape<L>::TW_ENTRIES) * sizeof(cpx<T>);
    static AttrOnce once;
    if (once.need(ctx->device)) ADSP_TRY(set_smem(fftconv_rows<T, L, SPEC>, smem));
    const int tiles_per_pair = (N1 / ROWS > 0 ? N1 / ROWS : 1);
    dim3 grid((unsigned)tiles_per_pair, (unsigned)pairs);
    int ntiles = 0;
    if (SPEC == 0) {   // convolution mode: flattened 1-D grid, optionally persistent (ADSP_TILES_PER_CTA > 1)
        ntiles = tiles_per_pair * pairs;
        grid = dim3((unsigned)persistent_grid(ctx, ntiles, rows_min_ctas(L)), 1);
    }
    LaunchTimer lt(ctx, st, KK_ROWS);
    fftconv_rows<T, L, SPEC><<<grid, THREADS, smem, st>>>(scratch, H, spec, scale, N1, tw, ntiles);
    count_launch(ctx);
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}

template <typename T, int SPEC>
static adsp_status launch_rows(adsp_ctx *ctx, cudaStream_t st, int L, cpx<T> *scratch, const cpx<T> *H, cpx<T> *spec,
                               T scale, int N1, const cpx<T> *tw, int pairs) {
    switch (L) {
    case 256:  return launch_rows_t<T, 256, SPEC>(ctx, st, scratch, H, spec, scale, N1, tw, pairs);
    case 512:  return launch_rows_t<T, 512, SPEC>(ctx, st, scratch, H, spec, scale, N1, tw, pairs);
    case 1024: return launch_rows_t<T, 1024, SPEC>(ctx, st, scratch, H, spec, scale, N1, tw, pairs);
    case 2048: return launch_rows_t<T, 2048, SPEC>(ctx, st, scratch, H, spec, scale, N1, tw, pairs);
    case 4096: return launch_rows_t<T, 4096, SPEC>(ctx, st, scratch, H, spec, scale, N1, tw, pairs);
    default: set_error("unsupported row FFT length"); return ADSP_ERR_INVALID_ARG;
    }
}

template <typename T, int N1>
static adsp_status launch_cols_t(adsp_ctx *ctx, cudaStream_t st, bool inverse, const ConvGeom &g, const T *x, T *y,
                                 cpx<T> *scratch, int N2, int lgN, const cpx<T> *tw, const cpx<T> *hi,
                                 const cpx<T> *lo, long long pair0, int pairs) {
    using CS = ColShape<N1>;
    const size_t smem = (FftShape<N1>::P > 0) ? ((size_t)CS::SMEM_ELEMS + FftShape<N1>::TW_ENTRIES) * sizeof(cpx<T>) : 16;
    static AttrOnce once;
    if (once.need(ctx->device)) {
        ADSP_TRY(set_smem(fftconv_cols_fwd<T, N1>, smem));
        ADSP_TRY(set_smem(fftconv_cols_inv<T, N1>, smem));
    }
    const int ntiles = (N2 / CS::TC) * pairs;
    const unsigned grid = (unsigned)persistent_grid(ctx, ntiles, CS::MIN_CTAS);
    LaunchTimer lt(ctx, st, inverse ? KK_COLS_INV : KK_COLS_FWD);
    if (!inverse)
        fftconv_cols_fwd<T, N1><<<grid, CS::THREADS, smem, st>>>(g, x, scratch, N2, lgN, tw, hi, lo, pair0, ntiles);
    else
        fftconv_cols_inv<T, N1><<<grid, CS::THREADS, smem, st>>>(g, scratch, x, y, N2, lgN, tw, hi, lo, pair0, ntiles);
    count_launch(ctx);
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}

template <typename T>
static adsp_status launch_cols(adsp_ctx *ctx, cudaStream_t st, int N1, bool inverse, const ConvGeom &g, const T *x, T *y,
                               cpx<T> *scratch, int N2, int lgN, const cpx<T> *tw, const cpx<T> *hi, const cpx<T> *lo,
                               long long pair0, int pairs) {
    switch (N1) {
    case 16:   return launch_cols_t<T, 16>(ctx, st, inverse, g, x, y, scratch, N2, lgN, tw, hi, lo, pair0, pairs);
    case 32:   return launch_cols_t<T, 32>(ctx, st, inverse, g, x, y, scratch, N2, lgN, tw, hi, lo, pair0, pairs);
    case 64:   return launch_cols_t<T, 64>(ctx, st, inverse, g, x, y, scratch, N2, lgN, tw, hi, lo, pair0, pairs);
    case 128:  return launch_cols_t<T, 128>(ctx, st, inverse, g, x, y, scratch, N2, lgN, tw, hi, lo, pair0, pairs);
    case 256:  return launch_cols_t<T, 256>(ctx, st, inverse, g, x, y, scratch, N2, lgN, tw, hi, lo, pair0, pairs);
    case 512:  return launch_cols_t<T, 512>(ctx, st, inverse, g, x, y, scratch, N2, lgN, tw, hi, lo, pair0, pairs);
    case 1024: return launch_cols_t<T, 1024>(ctx, st, inverse, g, x, y, scratch, N2, lgN, tw, hi, lo, pair0, pairs);
    default: set_error("unsupported column FFT length"); return ADSP_ERR_INVALID_ARG;
    }
}

// ------------------------------------------------------------------ mixed-radix columns (N1 = 16*P)
// Tensor maps of one FftConv::run call for the persistent TMA-fed column kernels (conv_kernels_mrp.cuh); built on the
// host per call (the input pointer is the caller's) and per scratch slot.  fwd_ok / inv_ok say which side may use them.
struct MrpLaunch {
    bool fwd_ok = false, inv_ok = false;
    int r_part = -1;                 // row of the N1 x N2 view the signal ends in (-1: it ends on a row boundary)
    int box_rows = 0, nbox = 0;      // forward: nbox boxes of box_rows signal rows per real block
    CUtensorMap main, part;          // input: (columns, full rows, channels) and the narrower (n mod N2, 1, channels)
    CUtensorMap scr;                 // scratch slot: (2*N2 reals, N1 rows x pairs)
};

template <typename T>
static void mrp_build_input_maps(MrpLaunch *m, const FftChoice &ch, const ConvGeom &g, const T *x, long long channels) {
    m->fwd_ok = false;
    // ADSP_MRP: 0 (default) plain kernels, 1 both persistent TMA-fed kernels, 2 forward only, 3 inverse only.  Off by default:
    // alone on the GPU the TMA-fed forward kernel is 11 % faster than the plain one (0.404 vs 0.453 ms per step), but in the
    // real schedule -- one block pair per launch on four streams, so that the intermediates stay in L2 -- a launch has
    // fewer tiles (256) than the machine has CTA slots (444), nothing is left to prefetch, and the step is 6-11 % slower in
    // every streams x group x tiles-per-CTA setting tried (profiles/r02_i_mrp_schedule_sweep.log, r02_h_mrp_v3_ab.log).
    static const bool use_mrp = env_ll("ADSP_MRP", 0) == 1 || env_ll("ADSP_MRP", 0) == 2;
    if (!use_mrp || ch.P <= 1 || g.D != 0 || g.nblk != 1 || g.in_shift != 0 || g.n > ch.N) return;   // single zero-padded block plans only
    const long long rows_full = g.n / ch.N2, rem = g.n % ch.N2;
    if (rows_full < 1 || channels < 1) return;
    const int TC = ADSP_MR_TC;
    const bool f64 = sizeof(T) == 8;
    // only the rows that hold signal are staged: one box of up to 256 rows, two beyond that
    // (box heights in multiples of 128 bytes of shared memory: a TMA tile destination must be 128-byte aligned)
    const int nbox = rows_full > 256 ? 2 : 1;
    const int ralign = 128 / (TC * (int)sizeof(T));
    const int BR = (int)(((rows_full + nbox - 1) / nbox + ralign - 1) / ralign * ralign);
    m->box_rows = BR; m->nbox = nbox;
    const uint64_t dims[3] = {(uint64_t)ch.N2, (uint64_t)rows_full, (uint64_t)channels};
    const uint64_t strides[2] = {(uint64_t)ch.N2 * sizeof(T), (uint64_t)g.in_stride * sizeof(T)};
    const uint32_t box[3] = {(uint32_t)TC, (uint32_t)BR, 1};
    if (!tma_encode(&m->main, f64, 3, x, dims, strides, box)) return;
    m->r_part = -1;
    if (rem > 0) {
        const uint64_t pd[3] = {(uint64_t)rem, 1, (uint64_t)channels};
        const uint32_t pb[3] = {(uint32_t)TC, 1, 1};
        if (!tma_encode(&m->part, f64, 3, x + rows_full * ch.N2, pd, strides, pb)) return;
        m->r_part = (int)rows_full;
    } else m->part = m->main;
    m->fwd_ok = true;
}

template <typename T>
static void mrp_build_scratch_map(MrpLaunch *m, const FftChoice &ch, const cpx<T> *slot, long long pairs) {
    m->inv_ok = false;
    static const bool use_mrp = env_ll("ADSP_MRP", 0) == 1 || env_ll("ADSP_MRP", 0) == 3;
    if (!use_mrp || ch.P <= 1 || pairs < 1) return;
    const int TC = ADSP_MR_TC, BR = ch.N1 <= 256 ? ch.N1 : ch.N1 / 2;
    const uint64_t dims[2] = {(uint64_t)ch.N2 * 2, (uint64_t)ch.N1 * (uint64_t)pairs};
    const uint64_t strides[1] = {(uint64_t)ch.N2 * sizeof(cpx<T>)};
    const uint32_t box[2] = {(uint32_t)TC * 2, (uint32_t)BR};
    m->inv_ok = tma_encode(&m->scr, sizeof(T) == 8, 2, slot, dims, strides, box);
}

template <typename T, int P>
static adsp_status launch_cols_mr_t(adsp_ctx *ctx, cudaStream_t st, bool inverse, const ConvGeom &g, const T *x, T *y,
                                    cpx<T> *scratch, int N2, long long N, const cpx<T> *tw, const cpx<T> *hi, const cpx<T> *lo,
                                    long long pair0, int pairs, const MrpLaunch *mrp) {
    using CS = ColShapeMR<P>;
    const int ntiles = (N2 / CS::TC) * pairs;
    // persistent TMA-fed kernels (conv_kernels_mrp.cuh) where the call has tensor maps for them (opt-in: ADSP_MRP=1)
    if (mrp && (inverse ? mrp->inv_ok : mrp->fwd_ok)) {
        using PS = ColShapeMRP<T, P>;
        static AttrOnce once_p;
        if (once_p.need(ctx->device)) {
            ADSP_TRY(set_smem(fftconv_cols_fwd_mrp<T, P>, PS::smem_fwd(PS::N1)));
            ADSP_TRY(set_smem(fftconv_cols_inv_mrp<T, P>, PS::SMEM_INV));
        }
        const size_t smem_p = inverse ? PS::SMEM_INV : PS::smem_fwd(mrp->box_rows * mrp->nbox);
        int per_sm = 0;   // resident CTAs for this shared-memory size (the forward stage depends on the signal length)
        if (!inverse) ADSP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fftconv_cols_fwd_mrp<T, P>, CS::THREADS, smem_p));
        else ADSP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fftconv_cols_inv_mrp<T, P>, CS::THREADS, smem_p));
        if (per_sm < 1) { set_error("persistent column kernel does not fit on an SM"); return ADSP_ERR_CUDA; }
        const long long cap = env_ll("ADSP_MRP_GRID_CTAS", 0);   // tuning: resident CTAs per SM to use
        if (cap > 0 && cap < per_sm) per_sm = (int)cap;
        int grid = (int)std::min<long long>((long long)per_sm * ctx->sm_count, ntiles);
        // launches that do not fill the machine (one or two pairs per group in the multi-stream schedule): k tiles per CTA, so
        // that every CTA still overlaps the copies of its next tile with the current one
        const long long k = env_ll("ADSP_MRP_TILES_PER_CTA", 2);
        if (k > 0 && (long long)grid * k > ntiles) grid = (int)std::max<long long>(1, (ntiles + k - 1) / k);
        LaunchTimer lt(ctx, st, inverse ? KK_COLS_INV : KK_COLS_FWD);
        if (!inverse)
            fftconv_cols_fwd_mrp<T, P><<<grid, CS::THREADS, smem_p, st>>>(mrp->main, mrp->part, mrp->r_part, mrp->box_rows, mrp->nbox, scratch, N2, (unsigned)N, tw,
                                                                         hi, lo, pair0, ntiles);
        else
            fftconv_cols_inv_mrp<T, P><<<grid, CS::THREADS, smem_p, st>>>(g, mrp->scr, x, y, N2, (unsigned)N, tw, hi, lo, pair0, ntiles);
        count_launch(ctx);
        ADSP_CUDA(cudaGetLastError());
        return ADSP_OK;
    }
    const size_t smem = ((size_t)CS::SMEM_ELEMS + CS::TW_ENTRIES) * sizeof(cpx<T>);
    static AttrOnce once;
    if (once.need(ctx->device)) {
        ADSP_TRY(set_smem(fftconv_cols_fwd_mr<T, P>, smem));
        ADSP_TRY(set_smem(fftconv_cols_inv_mr<T, P>, smem));
    }
    LaunchTimer lt(ctx, st, inverse ? KK_COLS_INV : KK_COLS_FWD);
    if (!inverse)
        fftconv_cols_fwd_mr<T, P><<<ntiles, CS::THREADS, smem, st>>>(g, x, scratch, N2, (unsigned)N, tw, hi, lo, pair0, ntiles);
    else
        fftconv_cols_inv_mr<T, P><<<ntiles, CS::THREADS, smem, st>>>(g, scratch, x, y, N2, (unsigned)N, tw, hi, lo, pair0, ntiles);
    count_launch(ctx);
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}

template <typename T>
static adsp_status launch_cols_mr(adsp_ctx *ctx, cudaStream_t st, int P, bool inverse, const ConvGeom &g, const T *x, T *y,
                                  cpx<T> *scratch, int N2, long long N, const cpx<T> *tw, const cpx<T> *hi, const cpx<T> *lo,
                                  long long pair0, int pairs, const MrpLaunch *mrp) {
    switch (P) {   // P here is M, the in-register DFT length (odd P, or 2P)
    case 6: return launch_cols_mr_t<T, 6>(ctx, st, inverse, g, x, y, scratch, N2, N, tw, hi, lo, pair0, pairs, mrp);
    case 10: return launch_cols_mr_t<T, 10>(ctx, st, inverse, g, x, y, scratch, N2, N, tw, hi, lo, pair0, pairs, mrp);
    case 14: return launch_cols_mr_t<T, 14>(ctx, st, inverse, g, x, y, scratch, N2, N, tw, hi, lo, pair0, pairs, mrp);
    case 18: return launch_cols_mr_t<T, 18>(ctx, st, inverse, g, x, y, scratch, N2, N, tw, hi, lo, pair0, pairs, mrp);
    case 3: return launch_cols_mr_t<T, 3>(ctx, st, inverse, g, x, y, scratch, N2, N, tw, hi, lo, pair0, pairs, mrp);
    case 5: return launch_cols_mr_t<T, 5>(ctx, st, inverse, g, x, y, scratch, N2, N, tw, hi, lo, pair0, pairs, mrp);
    case 7: return launch_cols_mr_t<T, 7>(ctx, st, inverse, g, x, y, scratch, N2, N, tw, hi, lo, pair0, pairs, mrp);
    case 9: return launch_cols_mr_t<T, 9>(ctx, st, inverse, g, x, y, scratch, N2, N, tw, hi, lo, pair0, pairs, mrp);
    default: set_error("unsupported odd column factor"); return ADSP_ERR_INVALID_ARG;
    }
}

// column pass of either kind, chosen by the transform geometry
template <typename T>
static adsp_status launch_cols_any(adsp_ctx *ctx, cudaStream_t st, const FftChoice &ch, bool inverse, const ConvGeom &g, const T *x,
                                   T *y, cpx<T> *scratch, const cpx<T> *tw, const cpx<T> *hi, const cpx<T> *lo, long long pair0,
                                   int pairs, const MrpLaunch *mrp = nullptr) {
    if (ch.P > 1) return launch_cols_mr<T>(ctx, st, ch.M, inverse, g, x, y, scratch, ch.N2, ch.N, tw, hi, lo, pair0, pairs, mrp);
    return launch_cols<T>(ctx, st, ch.N1, inverse, g, x, y, scratch, ch.N2, ch.lgN, tw, hi, lo, pair0, pairs);
}

// ------------------------------------------------------------------ FftConv
template <typename T>
adsp_status FftConv<T>::init(adsp_ctx *c, const T *d_kernel, long long K_, const FftChoice &choice, cpx<T> *H_ext) {
    ctx = c;
    K = K_;
    ch = choice;
    const size_t hbytes = (size_t)ch.N * sizeof(cpx<T>);
    if (H_ext) { H = H_ext; owns_H = false; }
    else ADSP_CUDA(cudaMalloc((void **)&H, hbytes));
    ADSP_TRY(get_tw_table<T>(ctx, ch.N2, &tw_rows));
    // geometry that presents the kernel as one zero-padded block (im part absent -> 0)
    ConvGeom g{};
    g.n = K; g.out_len = ch.N; g.in_stride = 0; g.out_stride = 0; g.S = ch.N; g.D = 0;
    g.total_blocks = 1; g.in_shift = 0; g.out_shift = 0; g.nblk = 1; g.accumulate = 0;
    const T scale = (T)(1.0L / (long double)ch.N);
    if (ch.N1 == 1) {
        ADSP_TRY((launch_full<T, true>(ctx, ctx->main, ch.N2, g, d_kernel, (T *)nullptr, (const cpx<T> *)nullptr, H,
                                       scale, tw_rows, 1)));
    } else {
        if (ch.P > 1) ADSP_TRY(get_twp_table<T>(ctx, ch.M, &tw_cols));
        else ADSP_TRY(get_tw_table<T>(ctx, ch.N1, &tw_cols));
        ADSP_TRY(get_tw4_tables<T>(ctx, ch.N, &tw_hi, &tw_lo));
        ADSP_TRY(ctx->scratch.reserve(hbytes));
        cpx<T> *scr = (cpx<T> *)ctx->scratch.p;
        ADSP_TRY(launch_cols_any<T>(ctx, ctx->main, ch, false, g, d_kernel, (T *)nullptr, scr, tw_cols, tw_hi, tw_lo, 0, 1));
        ADSP_TRY((launch_rows<T, true>(ctx, ctx->main, ch.N2, scr, (const cpx<T> *)nullptr, H, scale, ch.N1, tw_rows, 1)));
    }
    ADSP_CUDA(cudaStreamSynchronize(ctx->main));
    return ADSP_OK;
}

template <typename T> void FftConv<T>::destroy() {
    if (H && owns_H) cudaFree(H);
    H = nullptr;
}

template <typename T>
adsp_status FftConv<T>::run(const T *d_x, long long n, long long channels, long long in_stride, T *d_y,
                            long long out_stride, long long out_len, long long in_shift, long long out_shift,
                            bool accumulate) {
#ifdef ADSP_PHASE_DEBUG
    { const int f = (int)env_ll("ADSP_PHASE_SKIP", 0); cudaMemcpyToSymbol(g_phase_skip, &f, sizeof f);
      const int a = (int)env_ll("ADSP_SCRATCH_ALIAS", 0); cudaMemcpyToSymbol(g_scratch_alias, &a, sizeof a); }
#endif
    ConvGeom g{};
    g.n = n; g.out_len = out_len; g.in_stride = in_stride; g.out_stride = out_stride;
    g.S = ch.S; g.D = ch.D;
    if (no_discard) { g.S = ch.N; g.D = 0; no_discard = false; }   // single zero-padded block
    g.nblk = (int)((out_len + g.S - 1) / g.S);
    g.total_blocks = channels * (long long)g.nblk;
    g.in_shift = in_shift; g.out_shift = out_shift; g.accumulate = accumulate ? 1 : 0;
    const long long npairs = (g.total_blocks + 1) / 2;
    if (npairs <= 0) return ADSP_OK;

    if (ch.N1 == 1)
        return launch_full<T, false>(ctx, ctx->main, ch.N2, g, d_x, d_y, H, (cpx<T> *)nullptr, (T)0, tw_rows, npairs);

    // four-step, three kernels per group of pairs sized so that the intermediates of all in-flight groups stay in L2
    const size_t per_pair = (size_t)ch.N * sizeof(cpx<T>);
    size_t budget = ctx->scratch_budget;
    const long long mb = env_ll("ADSP_SCRATCH_MB", 0);  // tuning override
    if (mb > 0) budget = (size_t)mb << 20;

    // four groups in flight; three once a pair's intermediate reaches 16 MB (N = 2^20 in fp64), where a fourth evicts
    // spectrum and scratch lines from L2 (measured round 2, 16 ch x 14.4 M x 288k taps: 71.9 vs 69.0 Gsamples/s)
    int nstreams = (int)env_ll("ADSP_STREAMS", per_pair >= ((size_t)16 << 20) ? 3 : 4);
    if (nstreams < 1) nstreams = 1;
    if (nstreams > kWorkerStreams) nstreams = kWorkerStreams;
    if (per_pair * 2 >= budget && nstreams > 2) nstreams = 2;   // pairs that alone fill the budget (N >= 2^21): two in flight
    int nslots_max = nstreams;
    long long G = (long long)(budget / nslots_max / per_pair);
    const long long forced_g = env_ll("ADSP_GROUP_PAIRS", 0);   // tuning override
    if (forced_g > 0) G = forced_g;
    if (G < 1) G = 1;
    if (G > npairs) G = npairs;
    if (G > 32768) G = 32768;
    const int nslots = (npairs > G) ? nslots_max : 1;
    ADSP_TRY(ctx->scratch.reserve((size_t)nslots * (size_t)G * per_pair));
    cpx<T> *scr = (cpx<T> *)ctx->scratch.p;

    // tensor maps for the persistent TMA-fed column kernels (mixed-radix transforms): one pair over the caller's input,
    // one per scratch slot
    MrpLaunch mrp[kWorkerStreams];
    if (ch.P > 1) {
        mrp_build_input_maps<T>(&mrp[0], ch, g, d_x, channels);
        for (int s = 0; s < nslots; s++) {
            if (s > 0) { mrp[s] = mrp[0]; }
            mrp_build_scratch_map<T>(&mrp[s], ch, scr + (size_t)s * (size_t)G * (size_t)ch.N, G);
        }
    }
    if (nslots > 1) {
        ADSP_CUDA(cudaEventRecord(ctx->ev_fork, ctx->main));
        for (int s = 0; s < nslots; s++) ADSP_CUDA(cudaStreamWaitEvent(ctx->worker[s], ctx->ev_fork, 0));
    }
    long long grp = 0;
    for (long long pair0 = 0; pair0 < npairs; pair0 += G, grp++) {
        const int gp = (int)((npairs - pair0 < G) ? (npairs - pair0) : G);
        const int slot = (int)(grp % nslots);
        cudaStream_t st = (nslots > 1) ? ctx->worker[slot] : ctx->main;
        cpx<T> *sl = scr + (size_t)slot * (size_t)G * (size_t)ch.N;
        ADSP_TRY(launch_cols_any<T>(ctx, st, ch, false, g, d_x, d_y, sl, tw_cols, tw_hi, tw_lo, pair0, gp, ch.P > 1 ? &mrp[slot] : nullptr));
        ADSP_TRY((launch_rows<T, false>(ctx, st, ch.N2, sl, H, (cpx<T> *)nullptr, (T)0, ch.N1, tw_rows, gp)));
        ADSP_TRY(launch_cols_any<T>(ctx, st, ch, true, g, d_x, d_y, sl, tw_cols, tw_hi, tw_lo, pair0, gp, ch.P > 1 ? &mrp[slot] : nullptr));
    }
    if (nslots > 1) {
        for (int s = 0; s < nslots; s++) {
            ADSP_CUDA(cudaEventRecord(ctx->ev_join[s], ctx->worker[s]));
            ADSP_CUDA(cudaStreamWaitEvent(ctx->main, ctx->ev_join[s], 0));
        }
    }
    return ADSP_OK;
}

template <typename T>
adsp_status fft_convolve_device(adsp_ctx *ctx, const T *d_x, long long n, long long channels, long long in_stride,
                                const T *d_k, long long K, T *d_y, long long out_stride) {
    const FftChoice ch = choose_fft(K);
    const long long out_len = n + K - 1;
    // the spectrum of a one-shot call lives in the context's cache buffer: no cudaMalloc / cudaFree (a device-wide
    // synchronisation, with millisecond spikes) per call
    ADSP_TRY(ctx->spec_cache.reserve((size_t)ch.N * sizeof(cpx<T>)));
    cpx<T> *Hc = (cpx<T> *)ctx->spec_cache.p;
    if (ch.parts == 1) {
        FftConv<T> fc;
        adsp_status st = fc.init(ctx, d_k, K, ch, Hc);
        if (st == ADSP_OK) st = fc.run(d_x, n, channels, in_stride, d_y, out_stride, out_len, 0, 0, false);
        if (st == ADSP_OK) { cudaError_t e = cudaStreamSynchronize(ctx->main); if (e != cudaSuccess) st = cuda_fail(e, "sync", __FILE__, __LINE__); }
        fc.destroy();
        return st;
    }
    // long kernel: sum over IR partitions, each delayed by p*part_len
    ADSP_CUDA(cudaMemset2DAsync(d_y, (size_t)out_stride * sizeof(T), 0, (size_t)out_len * sizeof(T), (size_t)channels, ctx->main));
    for (int p = 0; p < ch.parts; p++) {
        const long long k0 = (long long)p * ch.part_len;
        const long long kp = (K - k0 < ch.part_len) ? (K - k0) : ch.part_len;
        if (kp <= 0) break;
        FftConv<T> fc;
        adsp_status st = fc.init(ctx, d_k + k0, kp, ch, Hc);
        if (st == ADSP_OK) st = fc.run(d_x, n, channels, in_stride, d_y, out_stride, n + kp - 1, 0, k0, true);
        if (st == ADSP_OK) { cudaError_t e = cudaStreamSynchronize(ctx->main); if (e != cudaSuccess) st = cuda_fail(e, "sync", __FILE__, __LINE__); }
        fc.destroy();
        if (st != ADSP_OK) return st;
    }
    return ADSP_OK;
}

// ------------------------------------------------------------------ pairwise FFT correlation
template <typename T, int N1>
static adsp_status launch_corr_cols_t(adsp_ctx *ctx, cudaStream_t st, const T *a, long long n, long long a_stride, const T *b,
                                      long long m, long long b_stride, cpx<T> *scratch, int N2, int lgN, const cpx<T> *tw,
                                      const cpx<T> *hi, const cpx<T> *lo, long long pair0, int pairs, int reverse_b = 1) {
    using CS = ColShape<N1>;
    const size_t smem = (FftShape<N1>::P > 0) ? ((size_t)CS::SMEM_ELEMS + FftShape<N1>::TW_ENTRIES) * sizeof(cpx<T>) : 16;
    static AttrOnce once;
    if (once.need(ctx->device)) ADSP_TRY(set_smem(corr_cols_fwd<T, N1>, smem));
    dim3 grid((unsigned)(N2 / CS::TC), (unsigned)pairs);
    LaunchTimer lt(ctx, st, KK_COLS_FWD);
    corr_cols_fwd<T, N1><<<grid, CS::THREADS, smem, st>>>(a, n, a_stride, b, m, b_stride, scratch, N2, lgN, tw, hi, lo, pair0, reverse_b);
    count_launch(ctx);
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}

template <typename T, int L>
static adsp_status launch_corr_rows_fused_t(adsp_ctx *ctx, cudaStream_t st, cpx<T> *Z, int npairs, int N1, T scale, const cpx<T> *tw) {
    constexpr int THREADS = 2 * (L / 16);
    const size_t smem = ((size_t)2 * L + FftShape<L>::TW_ENTRIES) * sizeof(cpx<T>);
    static AttrOnce once;
    if (once.need(ctx->device)) ADSP_TRY(set_smem(corr_rows_fused<T, L>, smem));
    dim3 grid((unsigned)(N1 / 2), (unsigned)((npairs + 1) / 2));
    LaunchTimer lt(ctx, st, KK_ROWS);
    corr_rows_fused<T, L><<<grid, THREADS, smem, st>>>(Z, npairs, N1, scale, tw);
    count_launch(ctx);
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}

template <typename T, int N1, bool STORE>
static adsp_status launch_corr_cols_inv_t(adsp_ctx *ctx, cudaStream_t st, const ConvGeom &g, const cpx<T> *scratch, T *y, int N2, int lgN, const cpx<T> *tw,
                                          const cpx<T> *hi, const cpx<T> *lo, long long pair0, int nq, T *part_v, long long *part_i, T *first_v) {
    using CS = ColShape<N1>;
    const size_t smem = (FftShape<N1>::P > 0) ? ((size_t)CS::SMEM_ELEMS + FftShape<N1>::TW_ENTRIES) * sizeof(cpx<T>) : 16;
    static AttrOnce once;
    if (once.need(ctx->device)) ADSP_TRY(set_smem(corr_cols_inv_peak<T, N1, STORE>, smem));
    dim3 grid((unsigned)(N2 / CS::TC), (unsigned)nq);
    LaunchTimer lt(ctx, st, KK_COLS_INV);
    corr_cols_inv_peak<T, N1, STORE><<<grid, CS::THREADS, smem, st>>>(g, scratch, 2, y, N2, lgN, tw, hi, lo, pair0, part_v, part_i, first_v);
    count_launch(ctx);
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}

// out[p][k] = sum_i a[p][i] * b[p][i - (k - (m-1))], k < n+m-1, for `pairs` pairs (device pointers); out may be null
// (peak search only).  peak_v / peak_i (nullable, device): FindPeak of every pair (correlate.go:200-216).
// Three launches per group of pairs: packed forward columns, fused rows (forward, spectral product of two pairs, inverse),
// inverse columns with the peak search in their epilogue; one final reduction per call.
// Returns *done=false when the shape is outside the transform sizes of this path.
template <typename T>
adsp_status fft_correlate_pairs_device(adsp_ctx *ctx, const T *a, long long n, long long a_stride, const T *b, long long m,
                                       long long b_stride, long long pairs, T *out, long long out_stride, T *peak_v, long long *peak_i, bool *done) {
    *done = false;
    const long long out_len = n + m - 1;
    long long N = 8192;
    while (N < out_len) N *= 2;
    if (N > (1LL << 22) || pairs <= 0) return ADSP_OK;
    FftChoice ch = make_choice(1, N);     // geometry only (N1, N2, lgN)
    const cpx<T> *tw_rows, *tw_cols, *tw_hi, *tw_lo;
    ADSP_TRY(get_tw_table<T>(ctx, ch.N2, &tw_rows));
    ADSP_TRY(get_tw_table<T>(ctx, ch.N1, &tw_cols));
    ADSP_TRY(get_tw4_tables<T>(ctx, ch.N, &tw_hi, &tw_lo));
    // groups of an even number of pairs whose spectra (one slot each; Q of two pairs reuses the first one's slot) fit the L2
    // scratch budget, never fewer than two
    const size_t slot = (size_t)N * sizeof(cpx<T>);
    long long budget = (long long)ctx->scratch_budget;
    const long long mb = env_ll("ADSP_CORR_SCRATCH_MB", 0);
    if (mb > 0) budget = mb << 20;
    long long G = budget / (long long)slot;
    G = std::max<long long>(2, G - (G & 1));
    G = std::min<long long>(G, std::min<long long>(pairs + (pairs & 1), 8192));
    ADSP_TRY(ctx->scratch.reserve((size_t)G * slot));
    cpx<T> *Z = (cpx<T> *)ctx->scratch.p;
    // peak candidates: one (value, index) per output row and column tile, plus each row's first sample (the NaN rule of
    // FindPeak needs corr[0])
    const int tiles = ch.N2 / (ch.N1 <= 256 ? (ADSP_COLS_CTA_THREADS / (ch.N1 / 16)) : (ch.N1 == 512 ? ADSP_COLS_TC_512 : ADSP_COLS_TC_1024));
    const long long rows_pad = pairs + (pairs & 1);
    ADSP_TRY(ctx->d_small.reserve((size_t)rows_pad * tiles * (sizeof(T) + sizeof(long long)) + (size_t)rows_pad * sizeof(T) + 256));
    long long *part_i = (long long *)ctx->d_small.p;
    T *part_v = (T *)(part_i + rows_pad * tiles);
    T *first_v = part_v + rows_pad * tiles;
    ConvGeom g{};                          // output side: block 2q -> pair 2q (re), block 2q+1 -> pair 2q+1 (im)
    g.n = N; g.out_len = out_len; g.in_stride = 0; g.out_stride = out_stride; g.S = N; g.D = 0;
    g.total_blocks = pairs; g.in_shift = 0; g.out_shift = 0; g.nblk = 1; g.accumulate = 0;
    cudaStream_t st = ctx->main;
    const T scale = (T)(1.0L / (long double)N);
    for (long long p0 = 0; p0 < pairs; p0 += G) {
        const int np = (int)std::min<long long>(G, pairs - p0), nq = (np + 1) / 2;
#define ADSP_CORR_COLS(n1) case n1: ADSP_TRY((launch_corr_cols_t<T, n1>(ctx, st, a, n, a_stride, b, m, b_stride, Z, ch.N2, ch.lgN, tw_cols, tw_hi, tw_lo, p0, np))); break;
        switch (ch.N1) {
            ADSP_CORR_COLS(16) ADSP_CORR_COLS(32) ADSP_CORR_COLS(64) ADSP_CORR_COLS(128) ADSP_CORR_COLS(256) ADSP_CORR_COLS(512) ADSP_CORR_COLS(1024)
        default: set_error("correlate: unsupported transform shape"); return ADSP_ERR_INVALID_ARG;
        }
#undef ADSP_CORR_COLS
        switch (ch.N2) {
#define ADSP_CORR_ROWS(l) case l: ADSP_TRY((launch_corr_rows_fused_t<T, l>(ctx, st, Z, np, ch.N1, scale, tw_rows))); break;
            ADSP_CORR_ROWS(256) ADSP_CORR_ROWS(512) ADSP_CORR_ROWS(1024) ADSP_CORR_ROWS(2048) ADSP_CORR_ROWS(4096)
#undef ADSP_CORR_ROWS
        default: set_error("correlate: unsupported row length"); return ADSP_ERR_INVALID_ARG;
        }
#define ADSP_CORR_INV(n1) \
    case n1: \
        if (out) ADSP_TRY((launch_corr_cols_inv_t<T, n1, true>(ctx, st, g, Z, out, ch.N2, ch.lgN, tw_cols, tw_hi, tw_lo, p0 / 2, nq, part_v, part_i, first_v))); \
        else ADSP_TRY((launch_corr_cols_inv_t<T, n1, false>(ctx, st, g, Z, out, ch.N2, ch.lgN, tw_cols, tw_hi, tw_lo, p0 / 2, nq, part_v, part_i, first_v))); \
        break;
        switch (ch.N1) {
            ADSP_CORR_INV(16) ADSP_CORR_INV(32) ADSP_CORR_INV(64) ADSP_CORR_INV(128) ADSP_CORR_INV(256) ADSP_CORR_INV(512) ADSP_CORR_INV(1024)
        default: set_error("correlate: unsupported transform shape"); return ADSP_ERR_INVALID_ARG;
        }
#undef ADSP_CORR_INV
    }
    if (peak_v && peak_i) {
        peak_final_kernel<T><<<(unsigned)pairs, 32, 0, st>>>(first_v, 1, part_v, part_i, tiles, pairs, peak_v, peak_i);
        count_launch(ctx);
    }
    ADSP_CUDA(cudaGetLastError());
    *done = true;
    return ADSP_OK;
}
// ------------------------------------------------------------------ deconvolution (deconvolve.go:72-412)
// out[p][i], i < out_len = IFFT( S * conj(H) / (|H|^2 + reg) )[i] with N = nextPow2(n) (circular, exactly the reference's
// transform length); reg < 0: naive division, *d_bad receives the smallest bin with |H| < 1e-15.  Device pointers.
template <typename T, int L>
static adsp_status launch_deconv_small(adsp_ctx *ctx, const T *sig, long long n, long long ss, const T *ker, long long m, long long ks,
                                       long long batch, T *out, long long os, long long out_len, T reg, long long *d_bad) {
    const cpx<T> *tw = nullptr;
    ADSP_TRY(get_tw_table<T>(ctx, L, &tw));
    const size_t smem = ((size_t)L + FftShape<L>::TW_ENTRIES) * sizeof(cpx<T>);
    static AttrOnce once;
    if (once.need(ctx->device)) ADSP_TRY(set_smem(deconv_small<T, L>, smem));
    LaunchTimer lt(ctx, ctx->main, KK_OTHER);
    deconv_small<T, L><<<(unsigned)batch, FftShape<L>::TPF, smem, ctx->main>>>(sig, n, ss, ker, m, ks, out, os, out_len, reg, tw, d_bad);
    count_launch(ctx);
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}

template <typename T>
adsp_status fft_deconvolve_device(adsp_ctx *ctx, const T *sig, long long n, long long s_stride, const T *ker, long long m,
                                  long long k_stride, long long batch, T *out, long long out_stride, long long out_len, T reg,
                                  long long *d_bad) {
    long long N = 1;
    while (N < n) N *= 2;
    if (batch <= 0) { set_error("deconvolve: empty batch"); return ADSP_ERR_INVALID_ARG; }
    if (N > (1LL << 22)) {
        set_error("deconvolve: signals longer than 2^22 = 4194304 samples are not supported (the reference's transform length nextPow2(n) = " +
                  std::to_string(N) + " exceeds the largest single transform of this library)");
        return ADSP_ERR_INVALID_ARG;
    }
    if (m > N) { set_error("deconvolve: kernel (" + std::to_string(m) + " taps) longer than the transform length nextPow2(n) = " + std::to_string(N)); return ADSP_ERR_INVALID_ARG; }
    if (N < 16) {
        LaunchTimer lt(ctx, ctx->main, KK_OTHER);
        deconv_tiny<T><<<(unsigned)batch, 1, 0, ctx->main>>>(sig, n, s_stride, ker, m, k_stride, out, out_stride, out_len, (int)N, reg, d_bad);
        count_launch(ctx);
        ADSP_CUDA(cudaGetLastError());
        return ADSP_OK;
    }
    switch (N) {
#define ADSP_DS(l) case l: return launch_deconv_small<T, l>(ctx, sig, n, s_stride, ker, m, k_stride, batch, out, out_stride, out_len, reg, d_bad);
        ADSP_DS(16) ADSP_DS(32) ADSP_DS(64) ADSP_DS(128) ADSP_DS(256) ADSP_DS(512) ADSP_DS(1024) ADSP_DS(2048) ADSP_DS(4096)
#undef ADSP_DS
    default: break;
    }
    FftChoice ch = make_choice(1, N);     // geometry only (N1, N2, lgN)
    const cpx<T> *tw_rows, *tw_cols, *tw_hi, *tw_lo;
    ADSP_TRY(get_tw_table<T>(ctx, ch.N2, &tw_rows));
    ADSP_TRY(get_tw_table<T>(ctx, ch.N1, &tw_cols));
    ADSP_TRY(get_tw4_tables<T>(ctx, ch.N, &tw_hi, &tw_lo));
    // groups of problems whose spectra (one slot each) and shared inverse transforms (one slot per two problems) fit the
    // L2 scratch budget; two problems share an inverse transform (Q = R_A + i*R_B, results in re / im)
    const size_t slot = (size_t)N * sizeof(cpx<T>);
    long long G = (long long)(ctx->scratch_budget / slot) * 2 / 3;
    G = std::max<long long>(2, G - (G & 1));
    G = std::min<long long>(G, std::min<long long>(batch + (batch & 1), 4096));
    const long long GQ = (G + 1) / 2;
    const bool in_place = (G == 2);        // one inverse per group: Q can live in slot 0 (a thread reads its inputs before it writes)
    ADSP_TRY(ctx->scratch.reserve((size_t)(in_place ? 2 : G + GQ) * slot));
    cpx<T> *Z = (cpx<T> *)ctx->scratch.p;
    cpx<T> *Q = in_place ? Z : Z + (size_t)G * N;
    ConvGeom g{};
    g.n = N; g.out_len = out_len; g.in_stride = 0; g.out_stride = out_stride; g.S = N; g.D = 0;
    g.total_blocks = batch; g.in_shift = 0; g.out_shift = 0; g.nblk = 1; g.accumulate = 0;
    cudaStream_t st = ctx->main;
    const T scale = (T)(1.0L / (long double)N);
    for (long long p0 = 0; p0 < batch; p0 += G) {
        const int np = (int)std::min<long long>(G, batch - p0), nq = (np + 1) / 2;
#define ADSP_DC_COLS(n1) case n1: ADSP_TRY((launch_corr_cols_t<T, n1>(ctx, st, sig, n, s_stride, ker, m, k_stride, Z, ch.N2, ch.lgN, tw_cols, tw_hi, tw_lo, p0, np, 0))); break;
        switch (ch.N1) {
            ADSP_DC_COLS(16) ADSP_DC_COLS(32) ADSP_DC_COLS(64) ADSP_DC_COLS(128) ADSP_DC_COLS(256) ADSP_DC_COLS(512) ADSP_DC_COLS(1024)
        default: set_error("deconvolve: unsupported transform shape"); return ADSP_ERR_INVALID_ARG;
        }
#undef ADSP_DC_COLS
        ADSP_TRY((launch_rows<T, 1>(ctx, st, ch.N2, Z, (const cpx<T> *)nullptr, Z, (T)1, ch.N1, tw_rows, np)));
        {
            LaunchTimer lt(ctx, st, KK_OTHER);
            dim3 grid((unsigned)((N + 255) / 256), (unsigned)nq);
            deconv_pointwise<T><<<grid, 256, 0, st>>>(Z, Q, np, ch.N1, ch.N2, scale, reg, d_bad);
            count_launch(ctx);
        }
        ADSP_TRY((launch_rows<T, 2>(ctx, st, ch.N2, Q, (const cpx<T> *)nullptr, Q, (T)1, ch.N1, tw_rows, nq)));
        // block 2*pair -> problem (re), 2*pair + 1 -> next problem (im): p0 is even, so pair0 = p0 / 2 lines the rows up
        ADSP_TRY(launch_cols<T>(ctx, st, ch.N1, true, g, (const T *)nullptr, out, Q, ch.N2, ch.lgN, tw_cols, tw_hi, tw_lo, p0 / 2, nq));
    }
    ADSP_CUDA(cudaGetLastError());
    return ADSP_OK;
}
template adsp_status fft_deconvolve_device<ADSP_REAL>(adsp_ctx *, const ADSP_REAL *, long long, long long, const ADSP_REAL *, long long, long long,
                                                      long long, ADSP_REAL *, long long, long long, ADSP_REAL, long long *);

template adsp_status fft_correlate_pairs_device<ADSP_REAL>(adsp_ctx *, const ADSP_REAL *, long long, long long, const ADSP_REAL *, long long,
                                                           long long, long long, ADSP_REAL *, long long, ADSP_REAL *, long long *, bool *);

template struct FftConv<ADSP_REAL>;
template adsp_status get_tw_table<ADSP_REAL>(adsp_ctx *, int, const cpx<ADSP_REAL> **);
template adsp_status get_tw4_tables<ADSP_REAL>(adsp_ctx *, long long, const cpx<ADSP_REAL> **, const cpx<ADSP_REAL> **);
template adsp_status get_twp_table<ADSP_REAL>(adsp_ctx *, int, const cpx<ADSP_REAL> **);
template adsp_status fft_convolve_device<ADSP_REAL>(adsp_ctx *, const ADSP_REAL *, long long, long long, long long,
                                                    const ADSP_REAL *, long long, ADSP_REAL *, long long);

}  // namespace adsp
