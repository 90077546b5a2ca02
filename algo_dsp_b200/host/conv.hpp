// conv.hpp -- header-only C++ mirror of the reference's Go package `conv` (dsp/conv/*.go) on top of
// the C ABI (include/algodsp_cuda.h).  Same names, argument meaning and error behaviour as the Go
// API; Go's (value, error) returns become exceptions carrying the sentinel.  The Go toolchain is
// absent in this image, so this is the compiled-language host side of the drop-in boundary; the cgo
// shim with the literal Go signatures is go/conv_cuda.go.
//
//   auto y  = conv::Convolve(signal, kernel);                 // conv.go:194
//   auto os = conv::NewOverlapSave(kernel, 0);  auto y2 = os.Process(x);   // overlap_save.go:53,126
//   auto [idx, val] = conv::FindPeak(conv::Correlate(a, b));  // correlate.go:16,200
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/algodsp_cuda.h"

namespace conv {

// sentinel errors (conv.go:41-46, partitioned.go:11-15); compare with errors_is()
struct Error : std::runtime_error {
    adsp_status status;
    Error(adsp_status st, const std::string &detail)
        : std::runtime_error(std::string(adsp_status_string(st)) + (detail.empty() ? "" : ": " + detail)), status(st) {}
};
constexpr adsp_status ErrEmptyInput = ADSP_ERR_EMPTY_INPUT;
constexpr adsp_status ErrEmptyKernel = ADSP_ERR_EMPTY_KERNEL;
constexpr adsp_status ErrLengthMismatch = ADSP_ERR_LENGTH_MISMATCH;
constexpr adsp_status ErrInvalidBlockSize = ADSP_ERR_INVALID_BLOCK_SIZE;
constexpr adsp_status ErrInvalidBlockOrder = ADSP_ERR_INVALID_BLOCK_ORDER;
constexpr adsp_status ErrEmptyImpulseResponse = ADSP_ERR_EMPTY_IR;
constexpr adsp_status ErrStageIndexOutOfRange = ADSP_ERR_STAGE_INDEX;
constexpr adsp_status ErrDivisionByZero = ADSP_ERR_DIVISION_BY_ZERO;       // deconvolve.go:14
inline bool errors_is(const Error &e, adsp_status sentinel) { return e.status == sentinel; }

inline void check(adsp_status st) {
    if (st == ADSP_OK) return;
    char buf[512];
    buf[0] = 0;
    if (st >= ADSP_ERR_INVALID_ARG) adsp_last_error(buf, sizeof buf);
    throw Error(st, buf);
}

enum Mode { ModeFull = ADSP_MODE_FULL, ModeSame = ADSP_MODE_SAME, ModeValid = ADSP_MODE_VALID };  // conv.go:57-69

using Vec = std::vector<double>;

class Context {
public:
    explicit Context(int device = 0) { check(adsp_ctx_create(device, &h_)); }
    ~Context() { adsp_ctx_destroy(h_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    adsp_ctx *handle() const { return h_; }
    static Context &Default() { static Context c(0); return c; }
private:
    adsp_ctx *h_ = nullptr;
};

namespace detail {
using Fn = adsp_status (*)(adsp_ctx *, const double *, int64_t, const double *, int64_t, double *);
inline Vec binary(Fn fn, const Vec &a, const Vec &b, Context &ctx) {
    Vec out(a.size() + b.size() > 0 ? a.size() + b.size() - (a.empty() || b.empty() ? 0 : 1) : 1);
    check(fn(ctx.handle(), a.data(), (int64_t)a.size(), b.data(), (int64_t)b.size(), out.data()));
    out.resize(a.size() + b.size() - 1);
    return out;
}
}  // namespace detail

inline Vec Direct(const Vec &a, const Vec &b, Context &c = Context::Default()) { return detail::binary(adsp_direct, a, b, c); }
inline Vec DirectCircular(const Vec &a, const Vec &b, Context &c = Context::Default()) {
    Vec out(a.size() ? a.size() : 1);
    check(adsp_direct_circular(c.handle(), a.data(), (int64_t)a.size(), b.data(), (int64_t)b.size(), out.data()));
    out.resize(a.size());
    return out;
}
inline Vec Convolve(const Vec &a, const Vec &b, Context &c = Context::Default()) { return detail::binary(adsp_convolve, a, b, c); }
inline Vec OverlapAddConvolve(const Vec &s, const Vec &k, Context &c = Context::Default()) { return detail::binary(adsp_overlap_add_convolve, s, k, c); }
inline Vec OverlapSaveConvolve(const Vec &s, const Vec &k, Context &c = Context::Default()) { return detail::binary(adsp_overlap_save_convolve, s, k, c); }
inline Vec Correlate(const Vec &a, const Vec &b, Context &c = Context::Default()) { return detail::binary(adsp_correlate, a, b, c); }
inline Vec CorrelateDirect(const Vec &a, const Vec &b, Context &c = Context::Default()) { return detail::binary(adsp_correlate_direct, a, b, c); }
inline Vec CorrelateFFT(const Vec &a, const Vec &b, Context &c = Context::Default()) { return detail::binary(adsp_correlate_fft, a, b, c); }
inline Vec CorrelateNormalized(const Vec &a, const Vec &b, Context &c = Context::Default()) { return detail::binary(adsp_correlate_normalized, a, b, c); }
inline Vec AutoCorrelate(const Vec &a, Context &c = Context::Default()) { return Correlate(a, a, c); }
inline Vec AutoCorrelateNormalized(const Vec &a, Context &c = Context::Default()) {
    Vec out(a.size() ? 2 * a.size() - 1 : 1);
    check(adsp_autocorrelate_normalized(c.handle(), a.data(), (int64_t)a.size(), out.data()));
    return out;
}
inline Vec trimToMode(const Vec &full, size_t lenA, size_t lenB, Mode mode) {  // conv.go:229
    int64_t s = 0, l = 0;
    adsp_trim_mode((int64_t)lenA, (int64_t)lenB, (adsp_mode)mode, &s, &l);
    return Vec(full.begin() + s, full.begin() + s + l);
}
inline Vec ConvolveMode(const Vec &a, const Vec &b, Mode m, Context &c = Context::Default()) { return trimToMode(Convolve(a, b, c), a.size(), b.size(), m); }
inline Vec CorrelateMode(const Vec &a, const Vec &b, Mode m, Context &c = Context::Default()) { return trimToMode(Correlate(a, b, c), a.size(), b.size(), m); }
inline std::pair<int64_t, double> FindPeak(const Vec &corr, Context &c = Context::Default()) {  // correlate.go:200
    int64_t idx = -1;
    double val = 0;
    check(adsp_find_peak(c.handle(), corr.data(), (int64_t)corr.size(), &idx, &val));
    return {idx, val};
}
inline int64_t LagFromIndex(int64_t index, int64_t lenB) { return adsp_lag_from_index(index, lenB); }
inline int64_t IndexFromLag(int64_t lag, int64_t lenB) { return adsp_index_from_lag(lag, lenB); }

// Reusable convolvers ------------------------------------------------------------------------
class Plan {
public:
    ~Plan() { adsp_plan_destroy(h_); }
    Plan(Plan &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    Plan(const Plan &) = delete;
    int64_t KernelLen() const { return adsp_plan_kernel_len(h_); }
    int64_t FFTSize() const { return adsp_plan_fft_size(h_); }
    void Reset() { adsp_plan_reset(h_); }
    Vec Process(const Vec &input) {                       // overlap_save.go:126 / overlap_add.go:108
        Vec out(input.size() + (size_t)KernelLen());
        check(adsp_plan_process(h_, input.data(), (int64_t)input.size(), out.data(), (int64_t)input.size() + KernelLen() - 1));
        out.resize(input.size() + (size_t)KernelLen() - 1);
        return out;
    }
    void ProcessTo(Vec &output, const Vec &input) {      // overlap_save.go:258: ErrLengthMismatch unless len(output)==len(input)+K-1
        check(adsp_plan_process(h_, input.data(), (int64_t)input.size(), output.data(), (int64_t)output.size()));
    }
    adsp_plan *handle() const { return h_; }
protected:
    Plan() = default;
    adsp_plan *h_ = nullptr;
};
class OverlapSave : public Plan {
public:
    OverlapSave(const Vec &kernel, int64_t fftSize, Context &c = Context::Default()) {
        check(adsp_overlap_save_create(c.handle(), kernel.data(), (int64_t)kernel.size(), fftSize, ADSP_F64, &h_));
    }
    int64_t StepSize() const { return adsp_plan_step_size(h_); }
};
class OverlapAdd : public Plan {
public:
    OverlapAdd(const Vec &kernel, int64_t blockSize, Context &c = Context::Default()) {
        check(adsp_overlap_add_create(c.handle(), kernel.data(), (int64_t)kernel.size(), blockSize, ADSP_F64, &h_));
    }
    int64_t BlockSize() const { return adsp_plan_block_size(h_); }
};
inline OverlapSave NewOverlapSave(const Vec &kernel, int64_t fftSize, Context &c = Context::Default()) { return OverlapSave(kernel, fftSize, c); }
inline OverlapAdd NewOverlapAdd(const Vec &kernel, int64_t blockSize, Context &c = Context::Default()) { return OverlapAdd(kernel, blockSize, c); }

class PartitionedConvolution : public Plan {                // partitioned.go:27
public:
    PartitionedConvolution(const Vec &kernel, int minBlockOrder, int maxBlockOrder, Context &c = Context::Default()) {
        check(adsp_partitioned_create(c.handle(), kernel.data(), (int64_t)kernel.size(), minBlockOrder, maxBlockOrder, ADSP_F64, &h_));
    }
    void ProcessBlock(const Vec &input, Vec &output) {       // partitioned.go:348
        check(adsp_partitioned_process_block(h_, input.data(), (int64_t)input.size(), output.data(), (int64_t)output.size()));
    }
    int Latency() const { return adsp_partitioned_latency(h_); }
    int StageCount() const { return adsp_partitioned_stage_count(h_); }
    std::pair<int, int> StageInfo(int index) const {
        int ps = 0, bc = 0;
        check(adsp_partitioned_stage_info(h_, index, &ps, &bc));
        return {ps, bc};
    }
};
inline PartitionedConvolution NewPartitionedConvolution(const Vec &k, int mn, int mx, Context &c = Context::Default()) { return PartitionedConvolution(k, mn, mx, c); }

// Deconvolution -- deconvolve.go:20-434 (float64 like the reference)
enum DeconvMethod { DeconvNaive = 0, DeconvRegularized = 1, DeconvWiener = 2 };                                   // :20-35
struct DeconvOptions { DeconvMethod Method = DeconvRegularized; double Epsilon = 0, NoiseVariance = 0, SignalVariance = 0; };   // :37-54
inline DeconvOptions DefaultDeconvOptions() { DeconvOptions o; o.Epsilon = 1e-6; return o; }                    // :56
inline Vec Deconvolve(const Vec &signal, const Vec &kernel, const DeconvOptions &opts = DefaultDeconvOptions(),
                      Context &c = Context::Default()) {                                                         // :72
    if (signal.empty()) check(ADSP_ERR_EMPTY_INPUT);
    if (kernel.empty()) check(ADSP_ERR_EMPTY_KERNEL);
    Vec out((size_t)adsp_deconv_out_len((int64_t)signal.size(), (int64_t)kernel.size()));
    check(adsp_deconvolve(c.handle(), signal.data(), (int64_t)signal.size(), kernel.data(), (int64_t)kernel.size(), (int)opts.Method,
                          opts.Epsilon, opts.NoiseVariance, opts.SignalVariance, out.data(), (int64_t)out.size()));
    return out;
}
inline Vec InverseFilter(const Vec &kernel, int64_t length, double epsilon, Context &c = Context::Default()) {   // :359
    if (kernel.empty()) check(ADSP_ERR_EMPTY_KERNEL);
    Vec out((size_t)(length > 0 ? length : 0));
    if (length > 0) check(adsp_inverse_filter(c.handle(), kernel.data(), (int64_t)kernel.size(), length, epsilon, out.data()));
    return out;
}
inline double SNR(const Vec &original, const Vec &recovered) {                                                  // :417
    return adsp_snr(original.data(), (int64_t)original.size(), recovered.data(), (int64_t)recovered.size());
}

// ConvolutionReverb -- dsp/effects/reverb/convolution.go:17-101, extended to `channels` independent streams per call
// (rows of a [channels][n] block with the given stride); channels = 1 is the reference's mono effect.
class ConvolutionReverb : public Plan {
public:
    ConvolutionReverb(const Vec &kernel, int minBlockOrder, int channels = 1, Context &c = Context::Default()) {   // :27 (maxBlockOrder 13)
        check(adsp_partitioned_create_batch(c.handle(), kernel.data(), (int64_t)kernel.size(), minBlockOrder, 13, channels, ADSP_F64, &h_));
    }
    void SetWetDry(double wet, double dry) { check(adsp_partitioned_set_wet_dry(h_, wet, dry)); }                    // :51
    void ProcessInPlace(Vec &block) {                                                                              // :60
        const int64_t ch = adsp_partitioned_channels(h_);
        check(adsp_partitioned_process_in_place_batch(h_, block.data(), (int64_t)block.size() / ch, (int64_t)block.size() / ch));
    }
    int Latency() const { return adsp_partitioned_latency(h_); }                                                   // :98
};
inline ConvolutionReverb NewConvolutionReverb(const Vec &k, int minBlockOrder, int channels = 1, Context &c = Context::Default()) {
    return ConvolutionReverb(k, minBlockOrder, channels, c);
}

}  // namespace conv
