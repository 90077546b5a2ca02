/*
 * conv_oracle_impl.h -- precision-generic body of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  Included twice by conv_oracle.c, once with
 * REAL=double (suffix _f64) and once with REAL=float (suffix _f32).  It restates,
 * in plain C, the algorithms of the reference package dsp/conv (CWBudde/algo-dsp).
 * Every function cites the reference file:line it follows.  Nothing in the product
 * path (algo_dsp_b200/, include/) may include, link or call this file.
 *
 * The FFT arithmetic of the reference lives in the third-party module
 * github.com/cwbudde/algo-fft v0.6.10 (go.mod:5-9), whose source is NOT under
 * /root/reference.  Its published contract, which the reference's own tests pin
 * (FFT paths compared against the time-domain Direct, conv_test.go:100-219,463-485),
 * is: complex DFT, forward unnormalised, inverse scaled by 1/N.  The restatement
 * below is an iterative radix-2 decimation-in-time transform with that contract.
 */

#ifndef REAL
#error "define REAL and SFX before including"
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

typedef struct { REAL re, im; } FN(cpx);

/* ------------------------------------------------------------------ FFT plan */

typedef struct {
    int n;
    int log2n;
    FN(cpx) *tw;   /* tw[k] = exp(-2*pi*i*k/n), k < n/2 */
    int *rev;      /* bit-reversal permutation */
} FN(fftplan);

static FN(fftplan) *FN(fftplan_new)(int n) {
    FN(fftplan) *p = (FN(fftplan) *)calloc(1, sizeof(*p));
    p->n = n;
    int l = 0;
    while ((1 << l) < n) l++;
    p->log2n = l;
    p->tw = (FN(cpx) *)malloc(sizeof(FN(cpx)) * (size_t)(n / 2 + 1));
    p->rev = (int *)malloc(sizeof(int) * (size_t)n);
    for (int k = 0; k < n / 2; k++) {
        long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)n;
        p->tw[k].re = (REAL)cosl(ang);
        p->tw[k].im = (REAL)sinl(ang);
    }
    for (int i = 0; i < n; i++) {
        int r = 0;
        for (int b = 0; b < l; b++)
            if (i & (1 << b)) r |= 1 << (l - 1 - b);
        p->rev[i] = r;
    }
    return p;
}

static void FN(fftplan_free)(FN(fftplan) *p) {
    if (!p) return;
    free(p->tw);
    free(p->rev);
    free(p);
}

/* In-place transform.  inverse!=0 -> conjugate twiddles and scale by 1/n
 * (algo-fft Plan.Forward / Plan.Inverse contract, see header comment). */
static void FN(fft_exec)(const FN(fftplan) *p, FN(cpx) *x, int inverse) {
    const int n = p->n;
    for (int i = 0; i < n; i++) {
        int r = p->rev[i];
        if (r > i) { FN(cpx) t = x[i]; x[i] = x[r]; x[r] = t; }
    }
    for (int len = 2; len <= n; len <<= 1) {
        const int half = len >> 1;
        const int step = n / len;
        for (int i = 0; i < n; i += len) {
            for (int k = 0; k < half; k++) {
                FN(cpx) w = p->tw[k * step];
                if (inverse) w.im = -w.im;
                FN(cpx) a = x[i + k], b = x[i + k + half];
                REAL tr = b.re * w.re - b.im * w.im;
                REAL ti = b.re * w.im + b.im * w.re;
                x[i + k].re = a.re + tr;        x[i + k].im = a.im + ti;
                x[i + k + half].re = a.re - tr; x[i + k + half].im = a.im - ti;
            }
        }
    }
    if (inverse) {
        const REAL s = (REAL)1 / (REAL)n;
        for (int i = 0; i < n; i++) { x[i].re *= s; x[i].im *= s; }
    }
}

/* ------------------------------------------------------------ direct (conv.go) */

/* directToScalar, dsp/conv/conv.go:117-123: dst[i+j] += a[i]*b[j] (mul, then add). */
static void FN(direct_scalar)(REAL *dst, const REAL *a, const REAL *b, int64_t n, int64_t m) {
    for (int64_t i = 0; i < n; i++)
        for (int64_t j = 0; j < m; j++)
            dst[i + j] += a[i] * b[j];
}

/* directToSIMD, dsp/conv/conv.go:127-154: per input sample temp = b*a[i]
 * (vecmath.ScaleBlock) then dst[i:i+m] += temp (vecmath.AddBlockInPlace). */
static void FN(direct_simd)(REAL *dst, const REAL *a, const REAL *b, int64_t n, int64_t m) {
    REAL *temp = (REAL *)malloc(sizeof(REAL) * (size_t)m);
    for (int64_t i = 0; i < n; i++) {
        const REAL s = a[i];
        for (int64_t j = 0; j < m; j++) temp[j] = b[j] * s;
        REAL *d = dst + i;
        for (int64_t j = 0; j < m; j++) d[j] += temp[j];
    }
    free(temp);
}

/* DirectTo, dsp/conv/conv.go:97-114 (clear, then SIMD path iff m >= 16). */
static void FN(direct_to)(REAL *dst, const REAL *a, int64_t n, const REAL *b, int64_t m) {
    memset(dst, 0, sizeof(REAL) * (size_t)(n + m - 1));
    if (m >= 16) FN(direct_simd)(dst, a, b, n, m);
    else FN(direct_scalar)(dst, a, b, n, m);
}

/* Direct, dsp/conv/conv.go:76-93. */
int FN(orc_direct)(const REAL *a, int64_t n, const REAL *b, int64_t m, REAL *dst) {
    if (n <= 0) return ORC_ERR_EMPTY_INPUT;
    if (m <= 0) return ORC_ERR_EMPTY_KERNEL;
    FN(direct_to)(dst, a, n, b, m);
    return ORC_OK;
}

/* DirectCircular(To), dsp/conv/conv.go:158-189. */
int FN(orc_direct_circular)(const REAL *a, int64_t n, const REAL *b, int64_t m, REAL *dst) {
    if (n <= 0 || m <= 0) return ORC_ERR_EMPTY_INPUT;
    if (n != m) return ORC_ERR_LENGTH_MISMATCH;
    memset(dst, 0, sizeof(REAL) * (size_t)n);
    for (int64_t i = 0; i < n; i++)
        for (int64_t j = 0; j < n; j++)
            dst[(i + j) % n] += a[i] * b[j];
    return ORC_OK;
}

/* ------------------------------------------------- overlap-add (overlap_add.go) */

/* OverlapAdd.Process, dsp/conv/overlap_add.go:108-164, with the sizing of
 * NewOverlapAdd :44-59 (blockSize<=0 -> max(nextPow2(K),256); fft=nextPow2(block+K-1)). */
int FN(orc_ola_process)(const REAL *kernel, int64_t K, int64_t blockSize,
                        const REAL *in, int64_t n, REAL *out) {
    if (K <= 0) return ORC_ERR_EMPTY_KERNEL;
    if (blockSize <= 0) {
        blockSize = orc_next_pow2(K);
        if (blockSize < 256) blockSize = 256;
    }
    const int64_t fftSize = orc_next_pow2(blockSize + K - 1);
    if (n <= 0) return ORC_ERR_EMPTY_INPUT;

    FN(fftplan) *plan = FN(fftplan_new)((int)fftSize);
    FN(cpx) *kfft = (FN(cpx) *)calloc((size_t)fftSize, sizeof(FN(cpx)));
    FN(cpx) *buf = (FN(cpx) *)malloc(sizeof(FN(cpx)) * (size_t)fftSize);
    for (int64_t i = 0; i < K; i++) kfft[i].re = kernel[i];
    FN(fft_exec)(plan, kfft, 0);                       /* :76-83 kernel FFT */

    const int64_t outLen = n + K - 1;
    memset(out, 0, sizeof(REAL) * (size_t)outLen);
    const int64_t numBlocks = (n + blockSize - 1) / blockSize;
    for (int64_t blk = 0; blk < numBlocks; blk++) {
        const int64_t start = blk * blockSize;
        const int64_t end = (start + blockSize < n) ? start + blockSize : n;
        const int64_t blockLen = end - start;
        for (int64_t i = 0; i < fftSize; i++) { buf[i].re = 0; buf[i].im = 0; }   /* :129-131 */
        for (int64_t i = 0; i < blockLen; i++) buf[i].re = in[start + i];        /* :133-135 */
        FN(fft_exec)(plan, buf, 0);                                                /* :138 */
        for (int64_t i = 0; i < fftSize; i++) {                                    /* :144-146 */
            REAL re = buf[i].re * kfft[i].re - buf[i].im * kfft[i].im;
            REAL im = buf[i].re * kfft[i].im + buf[i].im * kfft[i].re;
            buf[i].re = re; buf[i].im = im;
        }
        FN(fft_exec)(plan, buf, 1);                                                /* :149 */
        const int64_t resultLen = blockLen + K - 1;
        for (int64_t i = 0; i < resultLen && start + i < outLen; i++)              /* :157-160 */
            out[start + i] += buf[i].re;
    }
    FN(fftplan_free)(plan);
    free(kfft);
    free(buf);
    return ORC_OK;
}

/* ----------------------------------------------- overlap-save (overlap_save.go) */

/* One OLS block: [history | new samples | zeros] -> FFT -> *H -> IFFT.
 * dsp/conv/overlap_save.go:146-177 (main loop) and :225-245 (tail block). */
static void FN(ols_block)(const FN(fftplan) *plan, const FN(cpx) *kfft, FN(cpx) *buf,
                          const REAL *history, int64_t K, const REAL *newData, int64_t newSamples) {
    const int64_t N = plan->n;
    for (int64_t i = 0; i < N; i++) { buf[i].re = 0; buf[i].im = 0; }
    for (int64_t i = 0; i < K - 1; i++) buf[i].re = history[i];
    for (int64_t i = 0; i < newSamples; i++) buf[K - 1 + i].re = newData[i];
    FN(fft_exec)(plan, buf, 0);
    for (int64_t i = 0; i < N; i++) {
        REAL re = buf[i].re * kfft[i].re - buf[i].im * kfft[i].im;
        REAL im = buf[i].re * kfft[i].im + buf[i].im * kfft[i].re;
        buf[i].re = re; buf[i].im = im;
    }
    FN(fft_exec)(plan, buf, 1);
}

/* OverlapSave.Process, dsp/conv/overlap_save.go:126-254, with the sizing of
 * NewOverlapSave :53-76.  History is cleared at the start of every call (:136-138). */
int FN(orc_ols_process)(const REAL *kernel, int64_t K, int64_t fftSize,
                        const REAL *in, int64_t n, REAL *out) {
    int64_t step;
    int st = orc_ols_sizes(K, fftSize, &fftSize, &step);
    if (st != ORC_OK) return st;
    if (n <= 0) return ORC_ERR_EMPTY_INPUT;

    FN(fftplan) *plan = FN(fftplan_new)((int)fftSize);
    FN(cpx) *kfft = (FN(cpx) *)calloc((size_t)fftSize, sizeof(FN(cpx)));
    FN(cpx) *buf = (FN(cpx) *)malloc(sizeof(FN(cpx)) * (size_t)fftSize);
    REAL *history = (REAL *)calloc((size_t)(K > 1 ? K - 1 : 1), sizeof(REAL));
    for (int64_t i = 0; i < K; i++) kfft[i].re = kernel[i];
    FN(fft_exec)(plan, kfft, 0);                                         /* :96-101 */

    const int64_t outLen = n + K - 1;
    memset(out, 0, sizeof(REAL) * (size_t)outLen);
    int64_t inputPos = 0, outputPos = 0;
    while (inputPos < n) {                                                /* :144 */
        int64_t newSamples = step;
        if (inputPos + newSamples > n) newSamples = n - inputPos;         /* :156-159 */
        FN(ols_block)(plan, kfft, buf, history, K, in + inputPos, newSamples);
        const int64_t validStart = K - 1;                                 /* :182-186 */
        for (int64_t i = 0; i < newSamples && outputPos + i < outLen; i++)
            out[outputPos + i] = buf[validStart + i].re;
        /* :203-215 history = last K-1 input samples seen (zeros before t=0).  The
         * first history loop at :190-201 is fully overwritten and is omitted. */
        const int64_t hs = inputPos + newSamples - (K - 1);
        for (int64_t i = 0; i < K - 1; i++) {
            int64_t idx = hs + i;
            history[i] = (idx >= 0 && idx < n) ? in[idx] : (REAL)0;
        }
        inputPos += newSamples;
        outputPos += newSamples;
    }
    if (outputPos < outLen) {                                             /* :224-251 tail */
        FN(ols_block)(plan, kfft, buf, history, K, NULL, 0);
        const int64_t validStart = K - 1;
        for (int64_t i = 0; outputPos + i < outLen && validStart + i < fftSize; i++)
            out[outputPos + i] = buf[validStart + i].re;
    }
    FN(fftplan_free)(plan);
    free(kfft);
    free(buf);
    free(history);
    return ORC_OK;
}

/* ------------------------------------------------------- Convolve (conv.go) */

/* Convolve, dsp/conv/conv.go:194-216: swap so a is longer; len(b)<=64 -> Direct,
 * else OverlapAddConvolve (overlap_add.go:221-252: block=max(nextPow2(K),256)). */
int FN(orc_convolve)(const REAL *a, int64_t n, const REAL *b, int64_t m, REAL *dst) {
    if (n <= 0) return ORC_ERR_EMPTY_INPUT;
    if (m <= 0) return ORC_ERR_EMPTY_KERNEL;
    if (m > n) { const REAL *t = a; a = b; b = t; int64_t tl = n; n = m; m = tl; }
    if (m <= 64) return FN(orc_direct)(a, n, b, m, dst);
    return FN(orc_ola_process)(b, m, 0, a, n, dst);
}

/* OverlapSaveConvolve, dsp/conv/overlap_save.go:313-342. */
int FN(orc_ols_convolve)(const REAL *sig, int64_t n, const REAL *kernel, int64_t K, REAL *dst) {
    if (K <= 0) return ORC_ERR_EMPTY_KERNEL;
    return FN(orc_ols_process)(kernel, K, 0, sig, n, dst);
}

/* --------------------------------------------------- correlate (correlate.go) */

static REAL *FN(reversed)(const REAL *b, int64_t m) {
    REAL *r = (REAL *)malloc(sizeof(REAL) * (size_t)m);
    for (int64_t i = 0; i < m; i++) r[i] = b[m - 1 - i];
    return r;
}

/* Correlate, dsp/conv/correlate.go:16-28 (either empty -> ErrEmptyInput). */
int FN(orc_correlate)(const REAL *a, int64_t n, const REAL *b, int64_t m, REAL *dst) {
    if (n <= 0 || m <= 0) return ORC_ERR_EMPTY_INPUT;
    REAL *br = FN(reversed)(b, m);
    int st = FN(orc_convolve)(a, n, br, m, dst);
    free(br);
    return st;
}

/* CorrelateDirect, dsp/conv/correlate.go:31-42. */
int FN(orc_correlate_direct)(const REAL *a, int64_t n, const REAL *b, int64_t m, REAL *dst) {
    if (n <= 0 || m <= 0) return ORC_ERR_EMPTY_INPUT;
    REAL *br = FN(reversed)(b, m);
    int st = FN(orc_direct)(a, n, br, m, dst);
    free(br);
    return st;
}

/* l2Norm, dsp/conv/correlate.go:189-196. */
static REAL FN(l2norm)(const REAL *x, int64_t n) {
    REAL sum = 0;
    for (int64_t i = 0; i < n; i++) sum += x[i] * x[i];
    return (REAL)sqrt((double)sum);
}

/* AutoCorrelateNormalized, dsp/conv/correlate.go:63-81 (divide by zero-lag unless 0). */
int FN(orc_autocorrelate_normalized)(const REAL *a, int64_t n, REAL *dst) {
    int st = FN(orc_correlate)(a, n, a, n, dst);
    if (st != ORC_OK) return st;
    const REAL z = dst[n - 1];
    if (z == 0) return ORC_OK;
    for (int64_t i = 0; i < 2 * n - 1; i++) dst[i] /= z;
    return ORC_OK;
}

/* CorrelateNormalized, dsp/conv/correlate.go:86-107. */
int FN(orc_correlate_normalized)(const REAL *a, int64_t n, const REAL *b, int64_t m, REAL *dst) {
    int st = FN(orc_correlate)(a, n, b, m, dst);
    if (st != ORC_OK) return st;
    const REAL np = FN(l2norm)(a, n) * FN(l2norm)(b, m);
    if (np == 0) return ORC_OK;
    for (int64_t i = 0; i < n + m - 1; i++) dst[i] /= np;
    return ORC_OK;
}

/* CorrelateFFT, dsp/conv/correlate.go:111-186. */
int FN(orc_correlate_fft)(const REAL *a, int64_t n, const REAL *b, int64_t m, REAL *dst) {
    if (n <= 0 || m <= 0) return ORC_ERR_EMPTY_INPUT;
    const int64_t N = orc_next_pow2(n + m - 1);
    FN(fftplan) *plan = FN(fftplan_new)((int)N);
    FN(cpx) *fa = (FN(cpx) *)calloc((size_t)N, sizeof(FN(cpx)));
    FN(cpx) *fb = (FN(cpx) *)calloc((size_t)N, sizeof(FN(cpx)));
    for (int64_t i = 0; i < n; i++) fa[i].re = a[i];
    for (int64_t i = 0; i < m; i++) fb[i].re = b[i];
    FN(fft_exec)(plan, fa, 0);
    FN(fft_exec)(plan, fb, 0);
    for (int64_t i = 0; i < N; i++) {                /* :153-159 a * conj(b) */
        REAL br = fb[i].re, bi = -fb[i].im;
        REAL re = fa[i].re * br - fa[i].im * bi;
        REAL im = fa[i].re * bi + fa[i].im * br;
        fa[i].re = re; fa[i].im = im;
    }
    FN(fft_exec)(plan, fa, 1);
    for (int64_t i = 0; i < n; i++) dst[m - 1 + i] = fa[i].re;            /* :177-179 */
    for (int64_t i = 0; i < m - 1; i++) dst[i] = fa[N - m + 1 + i].re;    /* :181-183 */
    FN(fftplan_free)(plan);
    free(fa);
    free(fb);
    return ORC_OK;
}

/* FindPeak, dsp/conv/correlate.go:200-216 (signed, strict >, first max; (-1,0) if empty). */
void FN(orc_find_peak)(const REAL *corr, int64_t len, int64_t *index, REAL *value) {
    if (len <= 0) { *index = -1; *value = 0; return; }
    int64_t idx = 0;
    REAL v = corr[0];
    for (int64_t i = 0; i < len; i++)
        if (corr[i] > v) { idx = i; v = corr[i]; }
    *index = idx;
    *value = v;
}

/* ------------------------------- streaming OLA / OLS (streaming_overlap_*.go) */

typedef struct {
    int64_t K, blockSize, fftSize;
    FN(fftplan) *plan;
    FN(cpx) *kfft, *buf;
    REAL *state;       /* OLA: tail (K-1); OLS: history (K-1) */
    REAL *convResult;  /* OLA only */
    int is_ols;
} FN(stream);

/* NewStreamingOverlapAddT / NewStreamingOverlapSaveT,
 * dsp/conv/streaming_overlap_add.go:41-85, streaming_overlap_save.go:44-84. */
void *FN(orc_stream_create)(const REAL *kernel, int64_t K, int64_t blockSize, int is_ols, int *status) {
    if (K <= 0) { *status = ORC_ERR_EMPTY_KERNEL; return NULL; }
    if (blockSize <= 0) { *status = ORC_ERR_INVALID_ARG; return NULL; }
    FN(stream) *s = (FN(stream) *)calloc(1, sizeof(*s));
    s->K = K; s->blockSize = blockSize; s->is_ols = is_ols;
    s->fftSize = orc_next_pow2(blockSize + K - 1);
    s->plan = FN(fftplan_new)((int)s->fftSize);
    s->kfft = (FN(cpx) *)calloc((size_t)s->fftSize, sizeof(FN(cpx)));
    s->buf = (FN(cpx) *)calloc((size_t)s->fftSize, sizeof(FN(cpx)));
    s->state = (REAL *)calloc((size_t)(K > 1 ? K - 1 : 1), sizeof(REAL));
    s->convResult = (REAL *)calloc((size_t)s->fftSize, sizeof(REAL));
    for (int64_t i = 0; i < K; i++) s->kfft[i].re = kernel[i];
    FN(fft_exec)(s->plan, s->kfft, 0);
    *status = ORC_OK;
    return s;
}

void FN(orc_stream_destroy)(void *h) {
    FN(stream) *s = (FN(stream) *)h;
    if (!s) return;
    FN(fftplan_free)(s->plan);
    free(s->kfft); free(s->buf); free(s->state); free(s->convResult);
    free(s);
}

void FN(orc_stream_reset)(void *h) {
    FN(stream) *s = (FN(stream) *)h;
    for (int64_t i = 0; i < s->K - 1; i++) s->state[i] = 0;
}

int64_t FN(orc_stream_fft_size)(void *h) { return ((FN(stream) *)h)->fftSize; }

/* processBlockCore: streaming_overlap_add.go:98-133 / streaming_overlap_save.go:100-133. */
int FN(orc_stream_process_block)(void *h, const REAL *in, int64_t n, REAL *out) {
    FN(stream) *s = (FN(stream) *)h;
    if (n != s->blockSize) return ORC_ERR_LENGTH_MISMATCH;
    const int64_t N = s->fftSize, K = s->K, B = s->blockSize;
    for (int64_t i = 0; i < N; i++) { s->buf[i].re = 0; s->buf[i].im = 0; }
    if (s->is_ols) {
        for (int64_t i = 0; i < K - 1; i++) s->buf[i].re = s->state[i];
        for (int64_t i = 0; i < B; i++) s->buf[K - 1 + i].re = in[i];
    } else {
        for (int64_t i = 0; i < B; i++) s->buf[i].re = in[i];
    }
    FN(fft_exec)(s->plan, s->buf, 0);
    for (int64_t i = 0; i < N; i++) {
        REAL re = s->buf[i].re * s->kfft[i].re - s->buf[i].im * s->kfft[i].im;
        REAL im = s->buf[i].re * s->kfft[i].im + s->buf[i].im * s->kfft[i].re;
        s->buf[i].re = re; s->buf[i].im = im;
    }
    FN(fft_exec)(s->plan, s->buf, 1);
    if (s->is_ols) {
        for (int64_t i = 0; i < B; i++) out[i] = s->buf[K - 1 + i].re;
        if (B >= K - 1) {                                 /* :127-132 history update */
            for (int64_t i = 0; i < K - 1; i++) s->state[i] = in[B - K + 1 + i];
        } else {
            memmove(s->state, s->state + B, sizeof(REAL) * (size_t)(K - 1 - B));
            for (int64_t i = 0; i < B; i++) s->state[K - 1 - B + i] = in[i];
        }
    } else {
        const int64_t resultLen = B + K - 1;
        for (int64_t i = 0; i < resultLen; i++) s->convResult[i] = s->buf[i].re;
        const int64_t tailLen = K - 1;
        for (int64_t i = 0; i < tailLen && i < resultLen; i++) s->convResult[i] += s->state[i];
        const int64_t newTailLen = resultLen - B;
        for (int64_t i = 0; i < newTailLen; i++) s->state[i] = s->convResult[B + i];
        for (int64_t i = newTailLen; i < tailLen; i++) s->state[i] = 0;
        for (int64_t i = 0; i < B; i++) out[i] = s->convResult[i];
    }
    return ORC_OK;
}

/* ------------------------------------ partitioned convolution (partitioned.go) */

typedef struct {
    int fftOrder, fftSize, partSize, outputPos, latency, mod, modAnd, count;
    FN(cpx) **irSpectra;
    FN(fftplan) *plan;
    FN(cpx) *signalBuf, *signalFreq;
    REAL *convTime;
} FN(pstage);

typedef struct {
    int kernelLen, kernelLenPadded, minBlockOrder, maxBlockOrder, latency;
    REAL *inputBuffer, *outputBuffer;
    int inputBufSize, outputBufLen, blockPos;
    int nstages;
    FN(pstage) **stages;
} FN(pconv);

/* newPartStage + calculateIRSpectra, dsp/conv/partitioned.go:77-130.  IR block is
 * placed in the UPPER half of the 2*partSize buffer (:121-124). */
static FN(pstage) *FN(pstage_new)(const REAL *kernel, int kernelLen, int irOrder, int startPos,
                                  int latency, int count) {
    FN(pstage) *s = (FN(pstage) *)calloc(1, sizeof(*s));
    s->fftOrder = irOrder;
    s->partSize = 1 << irOrder;
    s->fftSize = 1 << (irOrder + 1);
    s->outputPos = startPos;
    s->latency = latency;
    s->mod = 0;
    s->modAnd = s->partSize / latency - 1;
    s->count = count;
    s->plan = FN(fftplan_new)(s->fftSize);
    s->signalBuf = (FN(cpx) *)calloc((size_t)s->fftSize, sizeof(FN(cpx)));
    s->signalFreq = (FN(cpx) *)calloc((size_t)s->fftSize, sizeof(FN(cpx)));
    s->convTime = (REAL *)calloc((size_t)s->fftSize, sizeof(REAL));
    s->irSpectra = (FN(cpx) **)calloc((size_t)count, sizeof(FN(cpx) *));
    for (int b = 0; b < count; b++) {
        FN(cpx) *spec = (FN(cpx) *)calloc((size_t)s->fftSize, sizeof(FN(cpx)));
        int ks = s->outputPos + b * s->partSize;
        int ke = ks + s->partSize;
        if (ke > kernelLen) ke = kernelLen;
        for (int i = ks; i < ke; i++) spec[s->partSize + (i - ks)].re = kernel[i];
        FN(fft_exec)(s->plan, spec, 0);
        s->irSpectra[b] = spec;
    }
    return s;
}

static void FN(pstage_free)(FN(pstage) *s) {
    for (int b = 0; b < s->count; b++) free(s->irSpectra[b]);
    free(s->irSpectra);
    FN(fftplan_free)(s->plan);
    free(s->signalBuf); free(s->signalFreq); free(s->convTime);
    free(s);
}

/* (*partStageT).process, dsp/conv/partitioned.go:134-183.  The single-block and
 * multi-block branches compute the same thing (one multiply + IFFT + add per block). */
static void FN(pstage_process)(FN(pstage) *s, const REAL *inputBuf, int inputLen,
                               REAL *outputBuf, int outputLen) {
    if (s->mod != 0) { s->mod = (s->mod + 1) & s->modAnd; return; }
    const int N = s->fftSize, P = s->partSize;
    const int inputStart = inputLen - N;
    for (int i = 0; i < N; i++) { s->signalFreq[i].re = inputBuf[inputStart + i]; s->signalFreq[i].im = 0; }
    FN(fft_exec)(s->plan, s->signalFreq, 0);
    for (int b = 0; b < s->count; b++) {
        const FN(cpx) *ir = s->irSpectra[b];
        for (int i = 0; i < N; i++) {
            s->signalBuf[i].re = s->signalFreq[i].re * ir[i].re - s->signalFreq[i].im * ir[i].im;
            s->signalBuf[i].im = s->signalFreq[i].re * ir[i].im + s->signalFreq[i].im * ir[i].re;
        }
        FN(fft_exec)(s->plan, s->signalBuf, 1);
        for (int i = 0; i < N; i++) s->convTime[i] = s->signalBuf[i].re;
        const int outPos = s->outputPos + s->latency - P + b * P;         /* :157,:173 */
        if (outPos >= 0 && outPos + P <= outputLen)
            for (int i = 0; i < P; i++) outputBuf[outPos + i] += s->convTime[i];
    }
    s->mod = (s->mod + 1) & s->modAnd;
}

/* NewPartitionedConvolutionT + partitionIR, dsp/conv/partitioned.go:212-332. */
void *FN(orc_part_create)(const REAL *kernel, int64_t K, int minBlockOrder, int maxBlockOrder, int *status) {
    if (K <= 0) { *status = ORC_ERR_EMPTY_IR; return NULL; }
    if (minBlockOrder < 1) { *status = ORC_ERR_INVALID_BLOCK_ORDER; return NULL; }
    if (maxBlockOrder < minBlockOrder) { *status = ORC_ERR_INVALID_BLOCK_ORDER; return NULL; }
    FN(pconv) *p = (FN(pconv) *)calloc(1, sizeof(*p));
    const int latency = 1 << minBlockOrder;
    const int minBlockSize = latency;
    const int kernelLen = (int)K;
    const int kernelLenPadded = ((kernelLen + minBlockSize - 1) / minBlockSize) * minBlockSize;

    int maxIROrd = orc_trunc_log2(kernelLenPadded + minBlockSize) - 1;                   /* :275 */
    int resIRSize = kernelLenPadded - (orc_bits(maxIROrd) - orc_bits(minBlockOrder - 1)); /* :278 */
    if (resIRSize > 0 && ((resIRSize >> maxIROrd) & 1) == 0 && maxIROrd > minBlockOrder) maxIROrd--;
    if (maxIROrd > maxBlockOrder) maxIROrd = maxBlockOrder;
    resIRSize = kernelLenPadded - (orc_bits(maxIROrd) - orc_bits(minBlockOrder - 1));    /* :289 */

    p->stages = (FN(pstage) **)calloc(64, sizeof(FN(pstage) *));
    int startPos = 0, ns = 0;
    for (int order = minBlockOrder; order < maxIROrd; order++) {                         /* :295-312 */
        int count = 1 + ((resIRSize >> order) & 1);
        p->stages[ns++] = FN(pstage_new)(kernel, kernelLen, order, startPos, latency, count);
        startPos += count * (1 << order);
        resIRSize -= (count - 1) * (1 << order);
    }
    int count = 1;                                                                       /* :315-318 */
    if (maxIROrd > 0) { count = 1 + resIRSize / (1 << maxIROrd); if (count < 1) count = 1; }
    p->stages[ns++] = FN(pstage_new)(kernel, kernelLen, maxIROrd, startPos, latency, count);
    p->nstages = ns;

    const int lastOrd = p->stages[ns - 1]->fftOrder;                                     /* :235-244 */
    p->inputBufSize = 2 << lastOrd;
    int outputHistSize = kernelLenPadded - latency;
    if (outputHistSize < 0) outputHistSize = 0;
    p->outputBufLen = outputHistSize + latency;
    p->kernelLen = kernelLen; p->kernelLenPadded = kernelLenPadded;
    p->minBlockOrder = minBlockOrder; p->maxBlockOrder = maxBlockOrder; p->latency = latency;
    p->inputBuffer = (REAL *)calloc((size_t)p->inputBufSize, sizeof(REAL));
    p->outputBuffer = (REAL *)calloc((size_t)p->outputBufLen, sizeof(REAL));
    p->blockPos = 0;
    *status = ORC_OK;
    return p;
}

void FN(orc_part_destroy)(void *h) {
    FN(pconv) *p = (FN(pconv) *)h;
    if (!p) return;
    for (int i = 0; i < p->nstages; i++) FN(pstage_free)(p->stages[i]);
    free(p->stages); free(p->inputBuffer); free(p->outputBuffer);
    free(p);
}

/* ProcessBlock, dsp/conv/partitioned.go:348-396. */
int FN(orc_part_process_block)(void *h, const REAL *in, int64_t n, REAL *out, int64_t nout) {
    FN(pconv) *p = (FN(pconv) *)h;
    if (n != nout) return ORC_ERR_LENGTH_MISMATCH;
    int64_t inPos = 0, remaining = n;
    const int latency = p->latency;
    while (remaining > 0) {
        int chunk = latency - p->blockPos;
        if (chunk > remaining) chunk = (int)remaining;
        memcpy(p->inputBuffer + p->inputBufSize - latency + p->blockPos, in + inPos, sizeof(REAL) * (size_t)chunk);
        memcpy(out + inPos, p->outputBuffer + p->blockPos, sizeof(REAL) * (size_t)chunk);
        p->blockPos += chunk; inPos += chunk; remaining -= chunk;
        if (p->blockPos == latency) {
            const int outLen = p->outputBufLen;
            memmove(p->outputBuffer, p->outputBuffer + latency, sizeof(REAL) * (size_t)(outLen - latency));
            memset(p->outputBuffer + outLen - latency, 0, sizeof(REAL) * (size_t)latency);
            for (int s = 0; s < p->nstages; s++)
                FN(pstage_process)(p->stages[s], p->inputBuffer, p->inputBufSize, p->outputBuffer, outLen);
            memmove(p->inputBuffer, p->inputBuffer + latency, sizeof(REAL) * (size_t)(p->inputBufSize - latency));
            memset(p->inputBuffer + p->inputBufSize - latency, 0, sizeof(REAL) * (size_t)latency);
            p->blockPos = 0;
        }
    }
    return ORC_OK;
}

/* Reset, dsp/conv/partitioned.go:399-407. */
void FN(orc_part_reset)(void *h) {
    FN(pconv) *p = (FN(pconv) *)h;
    memset(p->inputBuffer, 0, sizeof(REAL) * (size_t)p->inputBufSize);
    memset(p->outputBuffer, 0, sizeof(REAL) * (size_t)p->outputBufLen);
    p->blockPos = 0;
    for (int s = 0; s < p->nstages; s++) p->stages[s]->mod = 0;
}

int FN(orc_part_latency)(void *h) { return ((FN(pconv) *)h)->latency; }
int FN(orc_part_kernel_len)(void *h) { return ((FN(pconv) *)h)->kernelLen; }
int FN(orc_part_stage_count)(void *h) { return ((FN(pconv) *)h)->nstages; }
/* StageInfo, dsp/conv/partitioned.go:426-436. */
int FN(orc_part_stage_info)(void *h, int index, int *partSize, int *blockCount, int *startPos) {
    FN(pconv) *p = (FN(pconv) *)h;
    if (index < 0 || index >= p->nstages) return ORC_ERR_STAGE_INDEX;
    *partSize = p->stages[index]->partSize;
    *blockCount = p->stages[index]->count;
    if (startPos) *startPos = p->stages[index]->outputPos;
    return ORC_OK;
}

#undef CAT_
#undef CAT
#undef FN
