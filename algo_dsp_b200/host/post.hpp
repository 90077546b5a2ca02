// post.hpp -- header-only C++ mirrors of the reference packages right after the dsp/conv path and of the
// dsp/signal generators the path is fed with, on top of the C ABI (include/algodsp_cuda.h, csrc/post.cu, csrc/siggen.cu):
//   namespace ir       measure/ir      Analyzer.SchroederIntegral / FindImpulseStart          (ir.go:94-130, 381-404)
//   namespace sweep    measure/sweep   LogSweep.Generate / InverseFilter / Deconvolve         (sweep.go:73-239)
//   namespace fir      dsp/filter/fir  New, Filter.ProcessBlock / Reset / Order               (filter.go:18-103)
//   namespace resample dsp/resample    NewRational, NewForRates, Resampler.Process, Resample  (resample.go:153-314)
//   namespace signal   dsp/signal      Generator noise / sweeps, Normalize, RemoveDC on the library's index-hash PRNG
// Same names and argument meaning as the Go API; (value, error) returns become exceptions (conv::Error).
#pragma once
#include "conv.hpp"

namespace ir {
using conv::Vec;
constexpr adsp_status ErrEmptyIR = ADSP_ERR_EMPTY_IR;
class Analyzer {
public:
    explicit Analyzer(double sampleRate = 48000.0, conv::Context &c = conv::Context::Default()) : SampleRate(sampleRate), c_(&c) {}
    double SampleRate;
    Vec SchroederIntegral(const Vec &ir) const {                                       // ir.go:94
        if (ir.empty()) conv::check(ErrEmptyIR);
        Vec out(ir.size());
        conv::check(adsp_ir_schroeder(c_->handle(), ir.data(), (int64_t)ir.size(), out.data()));
        return out;
    }
    int64_t FindImpulseStart(const Vec &ir, double thresholdRatio = 0.1) const {       // ir.go:381
        if (ir.empty()) conv::check(ErrEmptyIR);
        int64_t idx = 0;
        conv::check(adsp_ir_find_impulse_start(c_->handle(), ir.data(), (int64_t)ir.size(), thresholdRatio, &idx));
        return idx;
    }
private:
    conv::Context *c_;
};
}  // namespace ir

namespace sweep {
using conv::Vec;
struct LogSweep {                                                                      // sweep.go:28
    double StartFreq, EndFreq, Duration, SampleRate;
    int64_t samples() const { return adsp_logsweep_samples(Duration, SampleRate); }
    Vec Generate() const {                                                             // sweep.go:73
        Vec out((size_t)std::max<int64_t>(samples(), 1));
        conv::check(adsp_logsweep_generate_host(out.data(), StartFreq, EndFreq, Duration, SampleRate));
        out.resize((size_t)samples());
        return out;
    }
    Vec InverseFilter() const {                                                        // sweep.go:104
        Vec out((size_t)std::max<int64_t>(samples(), 1));
        conv::check(adsp_logsweep_inverse_filter_host(out.data(), StartFreq, EndFreq, Duration, SampleRate));
        out.resize((size_t)samples());
        return out;
    }
    Vec Deconvolve(const Vec &response, conv::Context &c = conv::Context::Default()) const {   // sweep.go:164
        if (response.empty()) conv::check(ADSP_ERR_EMPTY_INPUT);                               // sweep.ErrEmptyResponse
        const int64_t n_out = (int64_t)response.size() + samples() - 1;
        Vec out((size_t)std::max<int64_t>(n_out, 1));
        conv::check(adsp_logsweep_deconvolve(c.handle(), response.data(), (int64_t)response.size(), StartFreq, EndFreq, Duration, SampleRate, out.data(),
                                             n_out));
        out.resize((size_t)n_out);
        return out;
    }
};
}  // namespace sweep

namespace fir {
using conv::Vec;
class Filter {                                                                         // filter.go:11
public:
    Filter(const Vec &coeffs, int channels = 1, conv::Context &c = conv::Context::Default()) : channels_(channels) {
        conv::check(adsp_fir_create(c.handle(), coeffs.data(), (int64_t)coeffs.size(), channels, &h_));
    }
    ~Filter() { adsp_fir_destroy(h_); }
    Filter(Filter &&o) noexcept : h_(o.h_), channels_(o.channels_) { o.h_ = nullptr; }
    Filter(const Filter &) = delete;
    int64_t Order() const { return adsp_fir_order(h_); }
    void Reset() { adsp_fir_reset(h_); }
    void ProcessBlock(Vec &buf) {                                                      // filter.go:61: in place, rows of buf.size()/channels
        const int64_t n = (int64_t)buf.size() / channels_;
        conv::check(adsp_fir_process_block(h_, buf.data(), n, n));
    }
    void ProcessBlockDevice(double *buf_dev, int64_t n, int64_t stride) { conv::check(adsp_fir_process_block_device(h_, buf_dev, n, stride)); }
private:
    adsp_fir *h_ = nullptr;
    int channels_;
};
inline Filter New(const Vec &coeffs, int channels = 1, conv::Context &c = conv::Context::Default()) { return Filter(coeffs, channels, c); }   // filter.go:18
}  // namespace fir

namespace resample {
using conv::Vec;
enum Quality { QualityFast = 0, QualityBalanced = 1, QualityBest = 2 };                // resample.go:24-33
struct Options { Quality quality = QualityBalanced; int tapsPerPhase = 0; double cutoffScale = 0, kaiserBeta = 0; int channels = 1; };
inline std::pair<int, int> approximateRatio(double v, int maxDen = 4096) {             // resample_design.go:74
    int n = 1, d = 1;
    adsp_resample_approximate_ratio(v, maxDen, &n, &d);
    return {n, d};
}
class Resampler {                                                                      // resample.go:138
public:
    Resampler(int up, int down, const Options &o = Options(), conv::Context &c = conv::Context::Default()) : channels_(o.channels) {
        conv::check(adsp_resampler_create(c.handle(), up, down, (int)o.quality, o.tapsPerPhase, o.cutoffScale, o.kaiserBeta, o.channels, &h_));
    }
    Resampler(double inRate, double outRate, Quality q, int maxDen, int channels, conv::Context &c) : channels_(channels) {
        conv::check(adsp_resampler_create_for_rates(c.handle(), inRate, outRate, (int)q, maxDen, channels, &h_));
    }
    ~Resampler() { adsp_resampler_destroy(h_); }
    Resampler(Resampler &&o) noexcept : h_(o.h_), channels_(o.channels_) { o.h_ = nullptr; }
    Resampler(const Resampler &) = delete;
    std::pair<int, int> Ratio() const { int u = 1, d = 1; adsp_resampler_ratio(h_, &u, &d); return {u, d}; }
    int TapsPerPhase() const { return adsp_resampler_taps_per_phase(h_); }
    int64_t PredictOutputLen(int64_t inputLen) const { return adsp_resampler_predict_output_len(h_, inputLen); }   // resample.go:295
    void Reset() { adsp_resampler_reset(h_); }                                                                      // resample.go:241
    Vec Process(const Vec &input) {                                                    // resample.go:249; rows of input.size()/channels
        if (input.empty()) return Vec();
        const int64_t n = (int64_t)input.size() / channels_, cap = std::max<int64_t>(PredictOutputLen(n), 1);
        Vec out((size_t)(cap * channels_));
        int64_t got = 0;
        conv::check(adsp_resampler_process(h_, input.data(), n, n, out.data(), cap, cap, &got));
        if (got != cap) {                                                              // rows are `cap` apart: close the gaps
            for (int c = 1; c < channels_; c++) std::copy(out.begin() + c * cap, out.begin() + c * cap + got, out.begin() + c * got);
            out.resize((size_t)(got * channels_));
        }
        return out;
    }
private:
    adsp_resampler *h_ = nullptr;
    int channels_;
};
inline Resampler NewRational(int up, int down, const Options &o = Options(), conv::Context &c = conv::Context::Default()) { return Resampler(up, down, o, c); }
inline Resampler NewForRates(double inRate, double outRate, Quality q = QualityBalanced, int maxDen = 4096, int channels = 1,
                             conv::Context &c = conv::Context::Default()) {            // resample.go:194
    return Resampler(inRate, outRate, q, maxDen, channels, c);
}
inline Vec Resample(const Vec &input, int up, int down, const Options &o = Options()) { return NewRational(up, down, o).Process(input); }   // resample.go:316
}  // namespace resample

namespace signal {
using conv::Vec;
// Generators on the library's stateless index-hash PRNG (sample i of stream `seed` depends on (seed, i) alone, so a shard
// generates exactly its own samples); these are the host twins, bit-identical to the adsp_gen_*_device kernels.
inline Vec Uniform(int64_t n, int64_t seed, int64_t index0 = 0) { Vec o((size_t)n); adsp_gen_uniform_host(o.data(), n, seed, index0); return o; }
inline Vec WhiteNoise(int64_t n, double amplitude, int64_t seed, int64_t index0 = 0) {
    Vec o((size_t)n); adsp_gen_white_host(o.data(), n, amplitude, seed, index0); return o;
}
inline Vec PinkNoise(int64_t n, double amplitude, int64_t seed, int64_t index0 = 0) {
    Vec o((size_t)n); adsp_gen_pink_host(o.data(), n, amplitude, seed, index0); return o;
}
inline Vec DecayingIR(int64_t taps, double decades, int64_t seed) { Vec o((size_t)taps); adsp_gen_decaying_ir_host(o.data(), taps, decades, seed); return o; }
}  // namespace signal
