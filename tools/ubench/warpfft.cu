// Prototype for the next step of the row kernel: ONE WARP per 1024-point row, 32 complex fp64 points per
// thread, two radix-32 passes with a single warp-private exchange (no CTA barrier anywhere):
//     forward FFT -> multiply by a cached spectrum -> inverse FFT, in place on an L2-resident scratch,
// i.e. the work of fftconv_rows<double, 1024>.  Reports ns per row and ps per complex point, and checks the
// result (identity spectrum: output must equal input) so that the timing is of a correct transform.
//   1024 = 32 x 32, n = 32*n1 + n2, k = k1 + 32*k2:  lane n2 runs the 32-point DFT over n1, multiplies by W_1024^(n2*k1),
//   the 32 x 32 transpose goes through 16 KB of shared memory private to the warp (XOR-swizzled, conflict free),
//   lane k1 runs the 32-point DFT over n2 and ends up owning X[k1 + 32*k2] -- the ownership it started with.
#include <cmath>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

typedef double2 C;
__device__ __forceinline__ C cadd(C a, C b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ C csub(C a, C b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ C cmul(C a, C b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
template <bool INV> __device__ __forceinline__ C ctw(C a, C w) {   // a*w (forward) or a*conj(w)
    return INV ? make_double2(a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y) : make_double2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
}
template <bool INV> __device__ __forceinline__ C mul_mi(C a) { return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x); }
template <bool INV> __device__ __forceinline__ void dft4(C &a0, C &a1, C &a2, C &a3) {
    C t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_mi<INV>(csub(a1, a3));
    a0 = cadd(t0, t2); a1 = cadd(t1, t3); a2 = csub(t0, t2); a3 = csub(t1, t3);
}
template <bool INV> __device__ __forceinline__ C kc(double re, double im) { return make_double2(re, INV ? -im : im); }

// 16-point DFT on v[0], v[S], ..., natural order in and out
template <int S, bool INV> __device__ __forceinline__ void dft16(C *v) {
    constexpr double h = 0.70710678118654752440, c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;
    C y[4][4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        y[q][0] = v[q * S]; y[q][1] = v[(q + 4) * S]; y[q][2] = v[(q + 8) * S]; y[q][3] = v[(q + 12) * S];
        dft4<INV>(y[q][0], y[q][1], y[q][2], y[q][3]);
    }
    y[1][1] = cmul(y[1][1], kc<INV>(c1, -s1)); y[1][2] = cmul(y[1][2], kc<INV>(h, -h)); y[1][3] = cmul(y[1][3], kc<INV>(s1, -c1));
    y[2][1] = cmul(y[2][1], kc<INV>(h, -h));   y[2][2] = mul_mi<INV>(y[2][2]);          y[2][3] = cmul(y[2][3], kc<INV>(-h, -h));
    y[3][1] = cmul(y[3][1], kc<INV>(s1, -c1)); y[3][2] = cmul(y[3][2], kc<INV>(-h, -h)); y[3][3] = cmul(y[3][3], kc<INV>(-c1, s1));
#pragma unroll
    for (int k = 0; k < 4; k++) {
        dft4<INV>(y[0][k], y[1][k], y[2][k], y[3][k]);
        v[k * S] = y[0][k]; v[(k + 4) * S] = y[1][k]; v[(k + 8) * S] = y[2][k]; v[(k + 12) * S] = y[3][k];
    }
}

__constant__ double2 c_w32[16];   // W_32^k, k = 0..15

// 32-point DFT in registers, natural order in and out: X[k] = E[k] + W32^k O[k], X[k+16] = E[k] - W32^k O[k]
template <bool INV> __device__ __forceinline__ void dft32(C (&e)[32]) {
    dft16<2, INV>(&e[0]);
    dft16<2, INV>(&e[1]);
    C x[32];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const C t = (k == 0) ? e[1] : ctw<INV>(e[2 * k + 1], c_w32[k]);
        x[k] = cadd(e[2 * k], t);
        x[k + 16] = csub(e[2 * k], t);
    }
#pragma unroll
    for (int k = 0; k < 32; k++) e[k] = x[k];
}

// e[k1] *= w1^k1 (forward) / conj (inverse), k1 = 0..31: w^k = w^(k & 7) * w^(k & 24), 10 powers live instead of 31
template <bool INV> __device__ __forceinline__ void twiddle32(C (&e)[32], C w1) {
    C lo[8], hi[4];
    lo[1] = w1; lo[2] = cmul(w1, w1); lo[3] = cmul(lo[2], w1); lo[4] = cmul(lo[2], lo[2]);
    lo[5] = cmul(lo[4], w1); lo[6] = cmul(lo[4], lo[2]); lo[7] = cmul(lo[4], lo[3]);
    hi[1] = cmul(lo[4], lo[4]); hi[2] = cmul(hi[1], hi[1]); hi[3] = cmul(hi[2], hi[1]);
#pragma unroll
    for (int k = 1; k < 8; k++) e[k] = ctw<INV>(e[k], lo[k]);
#pragma unroll
    for (int g = 1; g < 4; g++) {
        e[8 * g] = ctw<INV>(e[8 * g], hi[g]);
#pragma unroll
        for (int k = 1; k < 8; k++) e[8 * g + k] = ctw<INV>(e[8 * g + k], cmul(hi[g], lo[k]));
    }
}

// lane l holds A[k1] in e[k1]; afterwards lane k1 holds A_{n2}[k1] in e[n2]
__device__ __forceinline__ void transpose32(C (&e)[32], C *wbuf, int lane) {
#pragma unroll
    for (int k = 0; k < 32; k++) wbuf[k * 32 + (lane ^ k)] = e[k];
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < 32; n2++) e[n2] = wbuf[lane * 32 + (n2 ^ lane)];
    __syncwarp();
}

template <bool INV> __device__ __forceinline__ void fft1024(C (&e)[32], C *wbuf, int lane, C w1) {
    dft32<INV>(e);
    twiddle32<INV>(e, w1);
    transpose32(e, wbuf, lane);
    dft32<INV>(e);
}

__global__ void __launch_bounds__(128, 2) rows_warp(C *__restrict__ scratch, const C *__restrict__ H, const C *__restrict__ tw1024, int nrows, int hrows) {
    extern __shared__ C smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    C *wbuf = smem + warp * 1024;
    const C w1 = tw1024[lane];                       // W_1024^lane
    const int warps_total = gridDim.x * (blockDim.x >> 5);
    for (int row = blockIdx.x * (blockDim.x >> 5) + warp; row < nrows; row += warps_total) {
        C *p = scratch + (size_t)row * 1024 + lane;
        const C *hp = H + (size_t)(row % hrows) * 1024 + lane;
        C e[32];
#pragma unroll
        for (int q = 0; q < 32; q++) e[q] = __ldcg(&p[q * 32]);
        fft1024<false>(e, wbuf, lane, w1);
#pragma unroll
        for (int q = 0; q < 32; q++) e[q] = cmul(e[q], __ldg(&hp[q * 32]));
        fft1024<true>(e, wbuf, lane, w1);
#pragma unroll
        for (int q = 0; q < 32; q++) __stcg(&p[q * 32], e[q]);
    }
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const int nrows = 32768, hrows = 576;            // 32 MB of scratch (L2 resident), a 9 MB spectrum
    std::vector<C> hx((size_t)nrows * 1024), hH((size_t)hrows * 1024), htw(1024);
    for (size_t i = 0; i < hx.size(); i++) hx[i] = make_double2(sin(0.001 * (double)(i % 7919)) + 0.3, cos(0.0007 * (double)(i % 6131)));
    for (size_t i = 0; i < hH.size(); i++) hH[i] = make_double2(1.0 / 1024.0, 0.0);      // identity: out == in
    for (int k = 0; k < 1024; k++) htw[k] = make_double2(cos(2 * M_PI * k / 1024.0), -sin(2 * M_PI * k / 1024.0));
    C w32[16];
    for (int k = 0; k < 16; k++) w32[k] = make_double2(cos(2 * M_PI * k / 32.0), -sin(2 * M_PI * k / 32.0));
    CK(cudaMemcpyToSymbol(c_w32, w32, sizeof w32));
    C *dx, *dH, *dtw;
    CK(cudaMalloc(&dx, hx.size() * sizeof(C))); CK(cudaMalloc(&dH, hH.size() * sizeof(C))); CK(cudaMalloc(&dtw, htw.size() * sizeof(C)));
    CK(cudaMemcpy(dx, hx.data(), hx.size() * sizeof(C), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dH, hH.data(), hH.size() * sizeof(C), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dtw, htw.data(), htw.size() * sizeof(C), cudaMemcpyHostToDevice));
    const int smem = 4 * 1024 * sizeof(C);           // 16 KB per warp
    CK(cudaFuncSetAttribute(rows_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rows_warp, 128, smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, rows_warp));
    printf("rows_warp: %d registers, %zu B local (spill), %d CTAs of 4 warps per SM\n", fa.numRegs, fa.localSizeBytes, per_sm);
    const int grid = sms * per_sm;
    rows_warp<<<grid, 128, smem>>>(dx, dH, dtw, nrows, hrows);
    CK(cudaDeviceSynchronize());
    std::vector<C> hy(hx.size());
    CK(cudaMemcpy(hy.data(), dx, hy.size() * sizeof(C), cudaMemcpyDeviceToHost));
    double num = 0, den = 0;
    for (size_t i = 0; i < hx.size(); i++) { num += (hy[i].x - hx[i].x) * (hy[i].x - hx[i].x) + (hy[i].y - hx[i].y) * (hy[i].y - hx[i].y); den += hx[i].x * hx[i].x + hx[i].y * hx[i].y; }
    printf("identity-spectrum round trip: rel. L2 error %.3e (must be ~1e-16)\n", sqrt(num / den));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20;
    cudaEventRecord(e0);
    for (int i = 0; i < iters; i++) rows_warp<<<grid, 128, smem>>>(dx, dH, dtw, nrows, hrows);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    ms /= iters;
    printf("%d rows x 1024 points: %.3f ms per pass = %.2f ps per complex point (fftconv_rows<2048> steady state: 8.7 ps)\n", nrows, ms,
           ms * 1e9 / ((double)nrows * 1024));
    return 0;
}
