// conv_kernels_il.cuh -- "interleaved" row kernel: every thread carries TWO row tiles (A and B), one slot apart,
// so that inside one instruction stream the shared-memory traffic of one tile overlaps the FP64 butterflies of
// the other, and each Stockham exchange costs one CTA barrier instead of two.
//
// Same arithmetic as rows_tile (conv_kernels.cuh): per row an N2-point FFT, multiplication by the cached IR
// spectrum, N2-point inverse FFT, in place on the L2-resident scratch (the middle of the reference's per-block
// loop, dsp/conv/overlap_save.go:166-177).  A tile's life is a chain of 4P+2 stages that alternate
//      R-type (only loads: global load, or the shared-memory read of an exchange)   and
//      C-type (the butterflies of a pass, then the shared-memory write of the next exchange / the global store).
// In slot s tile A runs stage s and tile B stage s-1, so every slot pairs an R-type stage of one tile with a
// C-type stage of the other; one __syncthreads closes the slot.  A tile's exchange write (C-type, slot k) and
// the read that follows it (R-type, slot k+1) are separated by that barrier, and so are the read and the next
// write into the same buffer (slot k+1 -> k+2): one buffer per tile, no hazard.
// Why: ncu on the one-tile kernels shows LSU 69 % / FP64 49 % with 4 barrier-phased CTAs per SM overlapping
// only by chance; here the overlap is in the instruction stream.
#pragma once
#include "conv_kernels.cuh"

namespace adsp {

// ---- pieces of cta_fft as separate steps (same index conventions as fft_core.cuh) ----------------------
template <typename T, int L> struct PassInfo {
    using Sh = FftShape<L>;
    static constexpr int ns(int t) { int n = Sh::R0; for (int i = 1; i < t; i++) n *= 16; return n; }      // sub-transform length before pass t
    static constexpr int off(int t) { int o = 0, n = Sh::R0; for (int i = 1; i < t; i++) { o += tw_pass_entries(n); n *= 16; } return o; }
};

template <typename T, int L, bool INV> __device__ __forceinline__ void il_pass0(cpx<T> (&e)[16]) {
    using C = cpx<T>;
    constexpr int R0 = FftShape<L>::R0, S0 = 16 / R0;
#pragma unroll
    for (int u = 0; u < S0; u++) Dft<R0, S0, INV, C>::run(&e[u]);
}

// twiddles + radix-16 butterflies of pass t (1..P)
template <typename T, int L, bool INV, int t> __device__ __forceinline__ void il_pass(cpx<T> (&e)[16], const cpx<T> *stw, int j) {
    using C = cpx<T>;
    constexpr int ns = PassInfo<T, L>::ns(t), off = PassInfo<T, L>::off(t);
    const int k = j & (ns - 1);
    if (ns <= 16 && !ADSP_TW_TREE_ALL) {
        const C *twp = stw + off + k;
#pragma unroll
        for (int r = 1; r < 16; r++) e[r] = cmul_tw<INV>(e[r], twp[(r - 1) * ns]);
    } else {
        const C w1 = stw[off + k];
        const C w2 = cmul(w1, w1), w4 = cmul(w2, w2), w8 = cmul(w4, w4);
        e[1] = cmul_tw<INV>(e[1], w1);
        e[2] = cmul_tw<INV>(e[2], w2);
        e[3] = cmul_tw<INV>(e[3], cmul(w2, w1));
        e[4] = cmul_tw<INV>(e[4], w4);
        e[5] = cmul_tw<INV>(e[5], cmul(w4, w1));
        const C w6 = cmul(w4, w2);
        e[6] = cmul_tw<INV>(e[6], w6);
        e[7] = cmul_tw<INV>(e[7], cmul(w6, w1));
        e[8] = cmul_tw<INV>(e[8], w8);
        e[9] = cmul_tw<INV>(e[9], cmul(w8, w1));
        const C w10 = cmul(w8, w2);
        e[10] = cmul_tw<INV>(e[10], w10);
        e[11] = cmul_tw<INV>(e[11], cmul(w10, w1));
        const C w12 = cmul(w8, w4);
        e[12] = cmul_tw<INV>(e[12], w12);
        e[13] = cmul_tw<INV>(e[13], cmul(w12, w1));
        const C w14 = cmul(w12, w2);
        e[14] = cmul_tw<INV>(e[14], w14);
        e[15] = cmul_tw<INV>(e[15], cmul(w14, w1));
    }
    Dft<16, 1, INV, C>::run(&e[0]);
}

// exchange after pass t (0..P-1): write side
template <typename T, int L, int t, typename Addr>
__device__ __forceinline__ void il_write(const cpx<T> (&e)[16], cpx<T> *buf, const Addr &addr, int j) {
    using Sh = FftShape<L>;
    constexpr int R0 = Sh::R0, TPF = Sh::TPF, S0 = 16 / R0;
    if (t == 0) {
#pragma unroll
        for (int u = 0; u < S0; u++) {
            const int b = j + u * TPF;
#pragma unroll
            for (int r = 0; r < R0; r++) buf[addr.at(R0 * b + r, 0)] = e[u + r * S0];
        }
    } else {
        constexpr int ns = PassInfo<T, L>::ns(t);
        const int k = j & (ns - 1);
        const int j0 = (j - k) * 16 + k;
#pragma unroll
        for (int r = 0; r < 16; r++) buf[addr.at(j0 + r * ns, t)] = e[r];
    }
}
// exchange before pass t (1..P): read side
template <typename T, int L, int t, typename Addr>
__device__ __forceinline__ void il_read(cpx<T> (&e)[16], const cpx<T> *buf, const Addr &addr, int j) {
    constexpr int TPF = FftShape<L>::TPF;
#pragma unroll
    for (int q = 0; q < 16; q++) e[q] = buf[addr.at(j + q * TPF, t - 1)];
}

// ---- one tile's stage machine -----------------------------------------------------------------------------
template <typename T, int L> struct RowTileIL {
    using C = cpx<T>;
    using Sh = FftShape<L>;
    static constexpr int P = Sh::P, TPF = Sh::TPF;
    static constexpr int NSTAGES = 4 * P + 2;
    static_assert(P >= 1, "interleaved rows need at least one exchange");

    C e[16];
    C *buf;                 // this tile's exchange buffer (L elements)
    C *gp;                  // scratch row + j (global), valid when active
    const C *hp;            // spectrum row + j
    bool active;

    // stage S of the chain (compile-time).  Forward half: even stages read (global / exchange), odd stages run a
    // pass and write the next exchange.  Stage 2P+1 runs the last forward pass, the spectral multiply and inverse
    // pass 0 but must NOT write: other threads are still reading their spectrum values out of the buffer in this
    // slot.  So in the inverse half the roles shift: even stages only write an exchange, odd stages read it
    // (next slot, after the barrier) and run the pass.
    template <int S> __device__ __forceinline__ void stage(const RowAddr<T, Sh::R0> &addr, const C *stw, int j, uint64_t keep) {
        if constexpr (S == 0) {                                   // global load
            if (active) {
#pragma unroll
                for (int q = 0; q < 16; q++) e[q] = ld_scratch(&gp[q * TPF], keep);
            } else {
#pragma unroll
                for (int q = 0; q < 16; q++) { e[q].x = (T)0; e[q].y = (T)0; }
            }
        } else if constexpr (S <= 2 * P && S % 2 == 0) {          // forward exchange read before pass S/2
            il_read<T, L, S / 2>(e, buf, addr, j);
            if constexpr (S == 2 * P) {                           // last forward read: spectrum into the emptied slots
                if (active) {
#pragma unroll
                    for (int q = 0; q < 16; q++) cp_async_elem_keep(&buf[addr.at(j + q * TPF, P - 1)], &hp[q * TPF], keep);
                }
                cp_async_commit();
            }
        } else if constexpr (S < 2 * P + 1) {                     // forward pass (S-1)/2, then exchange write
            constexpr int t = (S - 1) / 2;
            if constexpr (t == 0) il_pass0<T, L, false>(e); else il_pass<T, L, false, t>(e, stw, j);
            il_write<T, L, t>(e, buf, addr, j);
        } else if constexpr (S == 2 * P + 1) {                    // last forward pass, * H, inverse pass 0
            il_pass<T, L, false, P>(e, stw, j);
            cp_async_wait_all();
            if (active) {
#pragma unroll
                for (int q = 0; q < 16; q++) e[q] = cmul(e[q], buf[addr.at(j + q * TPF, P - 1)]);
            }
            il_pass0<T, L, true>(e);
        } else if constexpr (S % 2 == 0) {                        // inverse exchange write after pass (S-2P-2)/2
            il_write<T, L, (S - 2 * P - 2) / 2>(e, buf, addr, j);
        } else {                                                  // inverse exchange read + pass (S-2P-1)/2 (+ store)
            constexpr int t = (S - 2 * P - 1) / 2;
            il_read<T, L, t>(e, buf, addr, j);
            il_pass<T, L, true, t>(e, stw, j);
            if constexpr (t == P) {
                if (active) {
#pragma unroll
                    for (int q = 0; q < 16; q++) st_scratch(&gp[q * TPF], e[q], keep);
                }
            }
        }
    }
};

// compile-time slot loop: in slot S tile A runs stage S, tile B stage S-1 (B's stage NSTAGES-1 of the PREVIOUS
// iteration runs in slot 0); FIRST / LAST trim the pipeline's head and tail.
template <typename T, int L, int S, bool FIRST> struct SlotLoop {
    using TL = RowTileIL<T, L>;
    static __device__ __forceinline__ void run(TL &A, TL &B, const RowAddr<T, FftShape<L>::R0> &addrA,
                                               const RowAddr<T, FftShape<L>::R0> &addrB, const cpx<T> *stw, int j, uint64_t keep) {
        if constexpr (S < TL::NSTAGES) {
            // the load-only / store-only stage of the slot goes first so that its memory operations are in flight
            // while the other tile's butterflies issue
            if constexpr (S % 2 == 0) {
                A.template stage<S>(addrA, stw, j, keep);
                if constexpr (S > 0) B.template stage<S - 1>(addrB, stw, j, keep);
                else if constexpr (!FIRST) B.template stage<TL::NSTAGES - 1>(addrB, stw, j, keep);
            } else {
                B.template stage<S - 1>(addrB, stw, j, keep);
                A.template stage<S>(addrA, stw, j, keep);
            }
            __syncthreads();
            SlotLoop<T, L, S + 1, FIRST>::run(A, B, addrA, addrB, stw, j, keep);
        }
    }
};

// Rows, two tiles per CTA.  A "tile" is ROWS rows (ROWS = threads / (L/16)); the CTA owns tile pairs
// (2u, 2u+1), u = blockIdx.x, blockIdx.x + gridDim.x, ...
template <typename T, int L>
__global__ void __launch_bounds__(rows_cta_threads(L), (rows_cta_threads(L) == 128) ? 2 : 1)
fftconv_rows_il(cpx<T> *__restrict__ scratch, const cpx<T> *__restrict__ H, int N1, const cpx<T> *__restrict__ tw, int ntiles) {
    using C = cpx<T>;
    using Sh = FftShape<L>;
    using TL = RowTileIL<T, L>;
    constexpr int TPF = Sh::TPF, THREADS = rows_cta_threads(L), ROWS = THREADS / TPF;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *bufA = reinterpret_cast<C *>(smem_raw);
    C *bufB = bufA + ROWS * L;
    C *stw = bufB + ROWS * L;
    load_tw_smem<T, L>(stw, tw, threadIdx.x, THREADS);
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    const int row = threadIdx.x / TPF;
    const int j = threadIdx.x % TPF;
    RowAddr<T, Sh::R0> addr{row * L};      // same offsets inside bufA and bufB
    const uint64_t keep = l2_policy_keep();
    const int tiles_per_pair = N1 / ROWS;
    const size_t pair_elems = (size_t)N1 * L;
    TL A, B;
    A.buf = bufA; B.buf = bufB;
    auto bind = [&](TL &X, int t) {
        X.active = t < ntiles;
        const int tt = X.active ? t : 0;
        const int pl = tt / tiles_per_pair, tile = tt - pl * tiles_per_pair;
        const size_t hoff = ((size_t)tile * ROWS + row) * L + j;
        X.gp = scratch + (size_t)ADSP_ALIAS(pl) * pair_elems + hoff;
        X.hp = H + hoff;
    };
    const int npairs_t = (ntiles + 1) / 2;
    int u = blockIdx.x;
    if (u >= npairs_t) return;
    bind(A, 2 * u);
    bind(B, 2 * u + 1);
    SlotLoop<T, L, 0, true>::run(A, B, addr, addr, stw, j, keep);
    for (u += gridDim.x; u < npairs_t; u += gridDim.x) {
        // slot 0 of this iteration still finishes B's previous tile: rebind A now, B after slot 0
        bind(A, 2 * u);
        A.template stage<0>(addr, stw, j, keep);
        B.template stage<TL::NSTAGES - 1>(addr, stw, j, keep);
        __syncthreads();
        bind(B, 2 * u + 1);
        SlotLoop<T, L, 1, false>::run(A, B, addr, addr, stw, j, keep);
    }
    B.template stage<TL::NSTAGES - 1>(addr, stw, j, keep);   // tail: B's last stage
}

}  // namespace adsp
