// conv_kernels_mrp.cuh -- persistent, TMA-fed versions of the mixed-radix column kernels (conv_kernels_mr.cuh).
//
// Same arithmetic, same tile shape (TC columns x N1 = 16*M rows), same scratch layout.  What changes is how a tile's
// inputs reach the SM.  In the one-tile-per-CTA kernels every tile pays, in sequence: the launch of its CTA, the copy of
// the twiddle table into shared memory, one exposed round trip for the four-step twiddle seeds and one for the data
// (36 scalar 8-byte loads per thread in the forward kernel) -- ncu attributes 21 % of the forward kernel's stall samples
// to those round trips and another 22 % to warps waiting at barriers for the warps that sit in them
// (profiles/r01_h_steady_ncu_summary.csv).  Here
//   * CTAs are persistent: grid = resident CTAs, a CTA walks tiles t, t + gridDim.x, ...
//   * the inputs of tile t+1 are fetched by the TMA unit (cp.async.bulk.tensor: one box of TC columns x up to 256 rows per
//     request, completion counted in bytes on an mbarrier) into a shared-memory stage while tile t is being transformed;
//     the zero padding behind the signal is the tensor map's out-of-bounds fill, the one row the signal ends in comes
//     from a second, narrower map.  (A first version issued one 64-byte cp.async.bulk per tile row: 576 requests per
//     tile, measured 11-20 cycles per request and SM whatever the number of resident CTAs -- the TMA unit wants few, large
//     requests: profiles/r02_c_mrp_rowcopies_ab.log.)
//   * the four-step twiddle seeds of tile t+1 are loaded into registers during tile t,
//   * the column twiddles W_N1^(j*km) are copied to shared memory once per persistent CTA.  To keep three CTAs of the
//     288-row shape resident next to them, the inverse kernel uses ONE buffer as TMA stage and as exchange buffer (a third
//     barrier per tile), and the forward kernel stages only the rows that hold signal (a single-block plan always ends in
//     at least K-1 zeros, which are never fetched).  (With separate stage and exchange buffers and the column twiddles read through L1 the
//     kernels were 5-10 % SLOWER than the plain ones: 25 % of all stall samples sat on those loads, the L1 carve-out
//     left by 3 x 74 KB of shared memory being too small to hold the tables: profiles/r02_g_mrp_v2_ncu_summary.csv.)
// The forward kernel serves the single zero-padded block plan (D = 0, one block per channel: the headline shape) with a
// 16-byte aligned channel stride; the launcher keeps the plain kernels for everything else.  The inverse kernel reads
// the library's own scratch and has no such restriction.
#pragma once
#include "conv_kernels_mr.cuh"
#include "tma.cuh"

namespace adsp {

#ifndef ADSP_MRP_CTAS
#define ADSP_MRP_CTAS 3     // wide columns (M > 16, 144-thread CTAs): stage + exchange = 72 KB per CTA
#endif
// resident CTAs the register allocator must allow: four 128-thread CTAs (128 registers: five would spill) where the
// shared memory (2 * N1 * TC complex) leaves room for them
template <int M> constexpr int mrp_min_ctas() { return M > 16 ? ADSP_MRP_CTAS : 4; }
// fp32 halves registers and shared memory per point: six 128-thread CTAs, four of the wide (144-thread) ones
template <typename T, int M> constexpr int mrp_ctas() { return sizeof(T) == 8 ? mrp_min_ctas<M>() : (M > 16 ? 4 : 6); }

// Four-step twiddle seeds of a tile: W_N^(n2*j) and W_N^(n2*M), each the product of a (hi, lo) table pair.  The persistent
// kernels PREFETCH the four table lines of the next tile into L1 half a tile ahead (no registers held) and read them at
// the top of the tile, where the round trip is an L1 hit that overlaps the wait for the TMA copies.
template <typename T>
__device__ __forceinline__ void mrp_seed_prefetch(const cpx<T> *__restrict__ tw_hi, const cpx<T> *__restrict__ tw_lo, unsigned n2, unsigned j, unsigned M,
                                                  unsigned N) {
    const unsigned mb = (n2 * j) % N, mr = (n2 * M) % N;
    prefetch_l1(&tw_hi[mb >> 10]); prefetch_l1(&tw_lo[mb & 1023u]);
    prefetch_l1(&tw_hi[mr >> 10]); prefetch_l1(&tw_lo[mr & 1023u]);
}

template <typename T, int M> struct ColShapeMRP {
    using CS = ColShapeMR<M>;
    static constexpr int N1 = CS::N1, TC = CS::TC, THREADS = CS::THREADS;
    static constexpr int BR = N1 <= 256 ? N1 : N1 / 2;                       // rows per TMA box (box dimensions are limited to 256)
    static constexpr int NBOX = N1 / BR;
    static constexpr size_t TILE_BYTES = (size_t)N1 * TC * sizeof(cpx<T>);   // one tile of complex points
    static constexpr size_t TW_BYTES = (size_t)CS::TW_ENTRIES * sizeof(cpx<T>);
    static constexpr size_t PART_BYTES = 256;                                // forward: the row cut by the end of the signal, per block (128 B apart)
    // inverse: ONE tile buffer serves as TMA stage and as exchange buffer (a third barrier per tile separates "everyone has
    // read the stage" from the exchange writes)
    static constexpr size_t SMEM_INV = TILE_BYTES + TW_BYTES + 16;
    // forward: exchange buffer + a stage of `rows` signal rows per real block (rows behind the end of the signal are zero
    // and never staged: a single-block plan always has at least K-1 of them)
    static constexpr size_t smem_fwd(int rows) { return TILE_BYTES + TW_BYTES + PART_BYTES + 128 + 2 * (size_t)rows * TC * sizeof(T); }   // 128: the mbarrier
    static constexpr unsigned PART_TX = 2u * TC * sizeof(T);
    static constexpr unsigned INV_TX = (unsigned)(N1 * TC * sizeof(cpx<T>));
    static_assert(N1 % BR == 0 && BR <= 256, "box rows");
};

// Forward tile: both real blocks of the pair (channels 2*pair, 2*pair+1 of the single-block plan) as `nbox` boxes of
// TC columns x box_rows rows each from the 3-D map (columns, full rows, channels); rows past the signal inside the last
// box and the channel past an odd batch are out of bounds of the map and arrive as zeros.  The row the signal ends in
// comes from a second map whose row is only (n mod N2) wide, into a 128-byte side buffer per block.
template <typename T, int M>
__device__ __forceinline__ void mrp_fwd_issue(const CUtensorMap *tm_main, const CUtensorMap *tm_part, int r_part, int box_rows, int nbox, int ch0, int col0,
                                              T *stage, T *part, uint64_t *bar) {
    using P = ColShapeMRP<T, M>;
    const int rows = box_rows * nbox;
    mbar_arrive_expect_tx(bar, 2u * (unsigned)rows * P::TC * (unsigned)sizeof(T) + (r_part >= 0 ? P::PART_TX : 0u));
#pragma unroll
    for (int blk = 0; blk < 2; blk++) {
        for (int rb = 0; rb < nbox; rb++)
            tma_load_3d(stage + (size_t)(blk * rows + rb * box_rows) * P::TC, tm_main, col0, rb * box_rows, ch0 + blk, bar);
        if (r_part >= 0) tma_load_3d(part + blk * (128 / (int)sizeof(T)), tm_part, col0, 0, ch0 + blk, bar);
    }
}

template <typename T, int M>
__global__ void __launch_bounds__(ColShapeMR<M>::THREADS, mrp_ctas<T, M>())
fftconv_cols_fwd_mrp(const __grid_constant__ CUtensorMap tm_main, const __grid_constant__ CUtensorMap tm_part, int r_part, int box_rows, int nbox,
                     cpx<T> *__restrict__ scratch, int N2, unsigned N, const cpx<T> *__restrict__ tw,
                     const cpx<T> *__restrict__ tw_hi, const cpx<T> *__restrict__ tw_lo, long long pair0, int ntiles) {
    using C = cpx<T>;
    using P = ColShapeMRP<T, M>;
    constexpr int TC = P::TC, N1 = P::N1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    C *buf = reinterpret_cast<C *>(smem_raw);                               // exchange buffer
    C *stw = reinterpret_cast<C *>(smem_raw + P::TILE_BYTES);
    T *part = reinterpret_cast<T *>(smem_raw + P::TILE_BYTES + P::TW_BYTES);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + P::TILE_BYTES + P::TW_BYTES + P::PART_BYTES);
    T *stage = reinterpret_cast<T *>(smem_raw + P::TILE_BYTES + P::TW_BYTES + P::PART_BYTES + 128);   // [2][rows][TC] reals
    const int rows = box_rows * nbox;                                       // staged rows per block; rows behind them are zero
    const int c = threadIdx.x % TC;
    const int j = threadIdx.x / TC;
    const int tiles_per_pair = N2 / TC;
    const size_t pair_elems = (size_t)N1 * N2;
    const uint64_t keep = l2_policy_keep();
    for (int i = threadIdx.x; i < P::CS::TW_ENTRIES; i += P::THREADS) stw[i] = tw[i];   // once per persistent CTA
    if (threadIdx.x == 0) {
        prefetch_tmap(&tm_main);
        prefetch_tmap(&tm_part);
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    int t = blockIdx.x;
    if (t < ntiles && threadIdx.x == 0)
        mrp_fwd_issue<T, M>(&tm_main, &tm_part, r_part, box_rows, nbox, (int)(2 * (pair0 + t / tiles_per_pair)), (t % tiles_per_pair) * TC, stage, part, bar);
    unsigned parity = 0;
    for (; t < ntiles; t += gridDim.x) {
        const int tn = t + gridDim.x;
        const int pl = t / tiles_per_pair, tile = t - pl * tiles_per_pair;
        const int n2 = tile * TC + c;
        mbar_wait(bar, parity);
        parity ^= 1;
        if (j < 16) {
            C e[M];
#pragma unroll
            for (int i = 0; i < M; i++) {
                const int row = i * 16 + j;
                const bool cut = row == r_part;            // the row the signal ends in: its valid part sits in the side buffer
                const bool in = row < rows;
                e[i].x = cut ? part[c] : (in ? stage[(size_t)row * TC + c] : (T)0);
                e[i].y = cut ? part[128 / (int)sizeof(T) + c] : (in ? stage[(size_t)(rows + row) * TC + c] : (T)0);
            }
            small_dft<M, false>(e);
#pragma unroll
            for (int km = 1; km < M; km++) e[km] = cmul_tw<false>(e[km], stw[km * 16 + j]);
            // (the barrier after the previous tile's exchange reads guarantees its readers are done with buf)
#pragma unroll
            for (int km = 0; km < M; km++) buf[(km * 16 + j) * TC + c] = e[km];
        }
        __syncthreads();                                   // exchange written; stage consumed by everyone
        if (tn < ntiles) {
            if (threadIdx.x == 0) {
                fence_proxy_async_smem();
                mrp_fwd_issue<T, M>(&tm_main, &tm_part, r_part, box_rows, nbox, (int)(2 * (pair0 + tn / tiles_per_pair)), (tn % tiles_per_pair) * TC, stage, part,
                                    bar);
            }
            if (j < M) mrp_seed_prefetch<T>(tw_hi, tw_lo, (unsigned)((tn % tiles_per_pair) * TC + c), (unsigned)j, (unsigned)M, N);
        }
        C f[16];
        if (j < M) {
#pragma unroll
            for (int jj = 0; jj < 16; jj++) f[jj] = buf[(j * 16 + jj) * TC + c];
        }
        __syncthreads();                                   // exchange consumed: the next tile may overwrite it
        if (j < M) {
            // four-step twiddle seeds: L1 hits (prefetched during the previous tile), in flight during the butterflies
            const C w_base = twiddle_any<T>(tw_hi, tw_lo, ((unsigned)n2 * (unsigned)j) % N);
            const C w_rho = twiddle_any<T>(tw_hi, tw_lo, ((unsigned)n2 * (unsigned)M) % N);
            Dft<16, 1, false, C>::run(&f[0]);
            apply_geometric16<false>(f, w_base, w_rho);
            C *dst = scratch + (size_t)ADSP_ALIAS(pl) * pair_elems + n2;
#pragma unroll
            for (int r = 0; r < 16; r++) st_scratch(&dst[(size_t)(j + M * r) * N2], f[r], keep);
        }
    }
}

// Inverse tile: the scratch rows of the tile ([N1][TC] complex) as NBOX boxes from the 2-D map over the slot
// (2*N2 reals per row, N1 rows per pair, pairs stacked)
template <typename T, int M>
__device__ __forceinline__ void mrp_inv_issue(const CUtensorMap *tm_scr, int pl, int col0, cpx<T> *stage, uint64_t *bar) {
    using P = ColShapeMRP<T, M>;
    const uint64_t policy = l2_policy_drop();   // last use of these scratch lines
    mbar_arrive_expect_tx(bar, P::INV_TX);
#pragma unroll
    for (int rb = 0; rb < P::NBOX; rb++)
        tma_load_2d_hint(stage + (size_t)rb * P::BR * P::TC, tm_scr, 2 * col0, pl * P::N1 + rb * P::BR, bar, policy);
}

template <typename T, int M>
__global__ void __launch_bounds__(ColShapeMR<M>::THREADS, mrp_ctas<T, M>())
fftconv_cols_inv_mrp(ConvGeom g, const __grid_constant__ CUtensorMap tm_scr, const T *__restrict__ x, T *__restrict__ y, int N2, unsigned N,
                     const cpx<T> *__restrict__ tw, const cpx<T> *__restrict__ tw_hi, const cpx<T> *__restrict__ tw_lo, long long pair0, int ntiles) {
    using C = cpx<T>;
    using P = ColShapeMRP<T, M>;
    constexpr int TC = P::TC;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    C *buf = reinterpret_cast<C *>(smem_raw);                               // exchange buffer and TMA stage: same bytes
    C *stage = reinterpret_cast<C *>(smem_raw);
    C *stw = reinterpret_cast<C *>(smem_raw + P::TILE_BYTES);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + P::TILE_BYTES + P::TW_BYTES);
    const int c = threadIdx.x % TC;
    const int j = threadIdx.x / TC;
    const int tiles_per_pair = N2 / TC;
    for (int i = threadIdx.x; i < P::CS::TW_ENTRIES; i += P::THREADS) stw[i] = tw[i];   // once per persistent CTA
    if (threadIdx.x == 0) {
        prefetch_tmap(&tm_scr);
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    int t = blockIdx.x;
    if (t < ntiles && threadIdx.x == 0) mrp_inv_issue<T, M>(&tm_scr, t / tiles_per_pair, (t % tiles_per_pair) * TC, stage, bar);
    unsigned parity = 0;
    for (; t < ntiles; t += gridDim.x) {
        const int tn = t + gridDim.x;
        const int pl = t / tiles_per_pair, tile = t - pl * tiles_per_pair;
        const int n2 = tile * TC + c;
        mbar_wait(bar, parity);
        parity ^= 1;
        C f[16], w_base, w_rho;
        if (j < M) {
#pragma unroll
            for (int r = 0; r < 16; r++) f[r] = stage[(size_t)(j + M * r) * TC + c];
            // four-step twiddle seeds: L1 hits (prefetched during the previous tile), in flight across the barrier
            w_base = twiddle_any<T>(tw_hi, tw_lo, ((unsigned)n2 * (unsigned)j) % N);
            w_rho = twiddle_any<T>(tw_hi, tw_lo, ((unsigned)n2 * (unsigned)M) % N);
        }
        __syncthreads();                                   // stage consumed by everyone: it becomes the exchange buffer
        if (j < M) {
            apply_geometric16<true>(f, w_base, w_rho);
            Dft<16, 1, true, C>::run(&f[0]);
#pragma unroll
            for (int jj = 0; jj < 16; jj++) buf[(j * 16 + jj) * TC + c] = f[jj];
        }
        __syncthreads();                                   // exchange written
        C e[M];
        if (j < 16) {
#pragma unroll
            for (int km = 0; km < M; km++) e[km] = buf[(km * 16 + j) * TC + c];
        }
        __syncthreads();                                   // exchange consumed: the buffer is free for the next tile's copies
        if (tn < ntiles) {
            if (threadIdx.x == 0) {
                fence_proxy_async_smem();
                mrp_inv_issue<T, M>(&tm_scr, tn / tiles_per_pair, (tn % tiles_per_pair) * TC, stage, bar);
            }
            if (j < M) mrp_seed_prefetch<T>(tw_hi, tw_lo, (unsigned)((tn % tiles_per_pair) * TC + c), (unsigned)j, (unsigned)M, N);
        }
        if (j >= 16) continue;                             // helper threads of wide columns (M > 16)
#pragma unroll
        for (int km = 1; km < M; km++) e[km] = cmul_tw<true>(e[km], stw[km * 16 + j]);
        small_dft<M, true>(e);
        const TileOut<T> a = tile_out<T>(block_io<T>(g, x, y, 2 * (pair0 + pl)), (long long)N);
        const TileOut<T> b = tile_out<T>(block_io<T>(g, x, y, 2 * (pair0 + pl) + 1), (long long)N);
        const int o0 = j * N2 + n2 - (int)g.D;
        const bool acc = g.accumulate != 0;
#pragma unroll
        for (int i = 0; i < M; i++) {
            const int o = o0 + i * 16 * N2;
            if ((unsigned)o < a.cnt) { if (acc) a.p[o] += e[i].x; else __stcs(a.p + o, e[i].x); }
            if ((unsigned)o < b.cnt) { if (acc) b.p[o] += e[i].y; else __stcs(b.p + o, e[i].y); }
        }
    }
}

}  // namespace adsp
