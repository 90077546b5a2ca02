// siggen_core.h -- arithmetic shared, instruction for instruction, by the device generators (siggen.cu kernels) and
// their host twins (adsp_gen_*_host): a stateless 64-bit hash PRNG and deterministic exp / sin(2*pi*x), so that host and
// device produce BIT-IDENTICAL synthetic inputs ("PRNG = stateless 64-bit hash of (seed, stream, index) ->
// uniform [0,1), same code host+device").  The formulas on top follow the reference's dsp/signal/generate.go:
// white :188-205, pink :210-250, linear sweep :134-154, log sweep :157-185.
//
// Every floating-point operation is spelled as an explicitly rounded multiply, add or fused multiply-add, so neither
// nvcc's nor gcc's contraction heuristics can make the two sides differ (the host TU is also built with
// -ffp-contract=off).
#pragma once
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define ADSP_HD __host__ __device__ inline
#define ADSP_MUL(a, b) __dmul_rn((a), (b))
#define ADSP_ADD(a, b) __dadd_rn((a), (b))
#define ADSP_FMA(a, b, c) __fma_rn((a), (b), (c))
#define ADSP_FLOOR(a) floor(a)
#define ADSP_RINT(a) rint(a)
#else
#include <math.h>
#include <string.h>
#if defined(__CUDACC__)
#define ADSP_HD __host__ __device__ inline
#else
#define ADSP_HD static inline
#endif
#define ADSP_MUL(a, b) ((a) * (b))
#define ADSP_ADD(a, b) ((a) + (b))
#define ADSP_FMA(a, b, c) fma((a), (b), (c))
#define ADSP_FLOOR(a) floor(a)
#define ADSP_RINT(a) rint(a)
#endif

// ---------------------------------------------------------------- hash PRNG
// splitmix64 finaliser (Steele, Lea, Flood 2014): a bijection of 64-bit words with full avalanche
ADSP_HD uint64_t adsp_mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27; z *= 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z;
}
// key of a (seed, stream) pair; computed once per call
ADSP_HD uint64_t adsp_hash_key(uint64_t seed, uint64_t stream) {
    return adsp_mix64(adsp_mix64(seed + 0x9E3779B97F4A7C15ULL) ^ (stream * 0xD1B54A32D192ED03ULL + 0x2545F4914F6CDD1DULL));
}
ADSP_HD uint64_t adsp_hash_u64(uint64_t key, uint64_t index) { return adsp_mix64(key + (index + 1) * 0x9E3779B97F4A7C15ULL); }
// uniform in [0, 1): the top 53 bits (what Go's rng.Float64() also yields: a multiple of 2^-53 below 1)
ADSP_HD double adsp_hash_uniform(uint64_t key, uint64_t index) { return ADSP_MUL((double)(adsp_hash_u64(key, index) >> 11), 1.1102230246251565e-16); }

// ---------------------------------------------------------------- deterministic elementary functions
// exp(x), |x| <= 700: k = round(x / ln 2), r = x - k ln 2 (two-term Cody-Waite), degree-13 Taylor polynomial on
// |r| <= 0.347 (truncation 5e-18), scaled by 2^k through the exponent field.  Error <= 2 ulp; what matters here is
// that host and device agree bit for bit.
ADSP_HD double adsp_det_exp(double x) {
    const double inv_ln2 = 1.4426950408889634, ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double kf = ADSP_RINT(ADSP_MUL(x, inv_ln2));
    double r = ADSP_FMA(-kf, ln2_hi, x);
    r = ADSP_FMA(-kf, ln2_lo, r);
    double p = 1.0 / 6227020800.0;                       // 1/13!
    p = ADSP_FMA(p, r, 1.0 / 479001600.0);
    p = ADSP_FMA(p, r, 1.0 / 39916800.0);
    p = ADSP_FMA(p, r, 1.0 / 3628800.0);
    p = ADSP_FMA(p, r, 1.0 / 362880.0);
    p = ADSP_FMA(p, r, 1.0 / 40320.0);
    p = ADSP_FMA(p, r, 1.0 / 5040.0);
    p = ADSP_FMA(p, r, 1.0 / 720.0);
    p = ADSP_FMA(p, r, 1.0 / 120.0);
    p = ADSP_FMA(p, r, 1.0 / 24.0);
    p = ADSP_FMA(p, r, 1.0 / 6.0);
    p = ADSP_FMA(p, r, 0.5);
    p = ADSP_FMA(p, r, 1.0);
    p = ADSP_FMA(p, r, 1.0);
    // 2^k in two factors so that k in [-1022*2, 1023*2] never leaves the normal range inside a factor
    const long long k = (long long)kf;
    const long long k1 = k / 2, k2 = k - k1;
    uint64_t b1 = (uint64_t)(k1 + 1023) << 52, b2 = (uint64_t)(k2 + 1023) << 52;
    double s1, s2;
#if defined(__CUDA_ARCH__)
    s1 = __longlong_as_double((long long)b1); s2 = __longlong_as_double((long long)b2);
#else
    memcpy(&s1, &b1, 8); memcpy(&s2, &b2, 8);
#endif
    return ADSP_MUL(ADSP_MUL(p, s1), s2);
}

// sin(2*pi*c) for c given in CYCLES: c - floor(c) is exact, so the argument reduction loses nothing however many
// cycles the sweep has run through (the reference forms phase = 2*pi*cycles first, generate.go:148,180, and leaves
// the reduction to math.Sin; same value up to the rounding of that product).
ADSP_HD double adsp_det_sin2pi(double c) {
    const double f = ADSP_ADD(c, -ADSP_FLOOR(c));        // [0, 1)
    const double q = ADSP_RINT(ADSP_MUL(f, 4.0));        // 0 .. 4
    const double r = ADSP_FMA(q, -0.25, f);              // [-1/8, 1/8], exact
    const double a = ADSP_MUL(r, 6.283185307179586476925);   // [-pi/4, pi/4]
    const double a2 = ADSP_MUL(a, a);
    // sin a = a * (1 - a2/3! + ... - a2^7/15!)     cos a = 1 - a2/2! + ... + a2^8/16!
    double s = -1.0 / 1307674368000.0;
    s = ADSP_FMA(s, a2, 1.0 / 6227020800.0);
    s = ADSP_FMA(s, a2, -1.0 / 39916800.0);
    s = ADSP_FMA(s, a2, 1.0 / 362880.0);
    s = ADSP_FMA(s, a2, -1.0 / 5040.0);
    s = ADSP_FMA(s, a2, 1.0 / 120.0);
    s = ADSP_FMA(s, a2, -1.0 / 6.0);
    s = ADSP_FMA(ADSP_MUL(s, a2), a, a);
    double co = 1.0 / 20922789888000.0;
    co = ADSP_FMA(co, a2, -1.0 / 87178291200.0);
    co = ADSP_FMA(co, a2, 1.0 / 479001600.0);
    co = ADSP_FMA(co, a2, -1.0 / 3628800.0);
    co = ADSP_FMA(co, a2, 1.0 / 40320.0);
    co = ADSP_FMA(co, a2, -1.0 / 720.0);
    co = ADSP_FMA(co, a2, 1.0 / 24.0);
    co = ADSP_FMA(co, a2, -0.5);
    co = ADSP_FMA(co, a2, 1.0);
    const int qi = ((int)q) & 3;
    return qi == 0 ? s : (qi == 1 ? co : (qi == 2 ? -s : -co));
}

// ---------------------------------------------------------------- per-sample formulas (generate.go)
// WhiteNoise :199-202   out[i] = (rng.Float64()*2 - 1) * amplitude
ADSP_HD double adsp_white_sample(uint64_t key, uint64_t i, double amplitude) {
    return ADSP_MUL(ADSP_ADD(ADSP_MUL(adsp_hash_uniform(key, i), 2.0), -1.0), amplitude);
}
// PinkNoise :220-232: per sample two uniforms (ur1 picks the band, ur2 the value); band b of sample i, or -1
ADSP_HD int adsp_pink_band(double ur1) {
    return ur1 <= 0.00198 ? 0 : ur1 <= 0.01478 ? 1 : ur1 <= 0.06378 ? 2 : ur1 <= 0.23378 ? 3 : ur1 <= 0.91578 ? 4 : -1;
}
ADSP_HD double adsp_pink_weight(int b) { return b == 0 ? 0.23980 : b == 1 ? 0.18727 : b == 2 ? 0.16380 : b == 3 ? 0.194685 : 0.214463; }
// contribution written at sample i if its band is b: val * pA[b], val = ur2*2 - 1
ADSP_HD double adsp_pink_value(uint64_t key, uint64_t i, int b) {
    const double ur2 = adsp_hash_uniform(key, 2 * i + 1);
    return ADSP_MUL(ADSP_ADD(ADSP_MUL(ur2, 2.0), -1.0), adsp_pink_weight(b));
}
// LinearSweep :144-151   t = i/sr; phase = 2 pi (f0 t + 0.5 k t^2), k = (f1 - f0)/duration
ADSP_HD double adsp_lin_sweep_sample(uint64_t i, double f0, double k, double sr, double amplitude) {
    const double t = (double)i / sr;
    const double cyc = ADSP_ADD(ADSP_MUL(f0, t), ADSP_MUL(ADSP_MUL(ADSP_MUL(0.5, k), t), t));
    return ADSP_MUL(amplitude, adsp_det_sin2pi(cyc));
}
// LogSweep :170-181      phase = 2 pi f0 (exp(k t) - 1)/k, k = ln(f1/f0)/duration
ADSP_HD double adsp_log_sweep_sample(uint64_t i, double f0, double k, double sr, double amplitude) {
    const double t = (double)i / sr;
    const double cyc = ADSP_MUL(f0, ADSP_ADD(adsp_det_exp(ADSP_MUL(k, t)), -1.0) / k);
    return ADSP_MUL(amplitude, adsp_det_sin2pi(cyc));
}
// synthetic decaying IR: h[i] = (u_i*2 - 1) * 10^(-decades * i / K)
ADSP_HD double adsp_decaying_ir_sample(uint64_t key, uint64_t i, double K, double decades) {
    const double env = adsp_det_exp(ADSP_MUL(ADSP_MUL(-decades, 2.302585092994045684), (double)i / K));
    return ADSP_MUL(ADSP_ADD(ADSP_MUL(adsp_hash_uniform(key, i), 2.0), -1.0), env);
}
