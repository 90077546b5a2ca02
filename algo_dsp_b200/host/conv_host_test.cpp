// Reads like the reference's conv_test.go / example_test.go, against the C++ host mirror.
// Build: g++ -std=c++17 conv_host_test.cpp -o conv_host_test -L.. -lalgodsp_cuda -Wl,-rpath,$PWD/..
#include <cmath>
#include <cstdio>
#include <string>

#include "conv.hpp"

#define EXPECT(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main() {
    using conv::Vec;
    {   // TestDirect, conv_test.go:9-69
        Vec r = conv::Direct({1, 2, 3}, {1, 1, 1});
        Vec want{1, 3, 6, 5, 3};
        EXPECT(r.size() == want.size());
        for (size_t i = 0; i < r.size(); i++) EXPECT(std::fabs(r[i] - want[i]) <= 1e-10);
    }
    {   // TestDirectErrors, conv_test.go:71-81
        try { conv::Direct({}, {1, 2}); EXPECT(false); } catch (const conv::Error &e) { EXPECT(conv::errors_is(e, conv::ErrEmptyInput)); }
        try { conv::Direct({1, 2}, {}); EXPECT(false); } catch (const conv::Error &e) { EXPECT(conv::errors_is(e, conv::ErrEmptyKernel)); }
    }
    {   // TestOverlapSaveConvolve, conv_test.go:132-169
        Vec signal(500);
        for (size_t i = 0; i < signal.size(); i++) signal[i] = std::sin(2 * M_PI * (double)i / 50);
        Vec kernel{0.2, 0.3, 0.3, 0.2};
        Vec d = conv::Direct(signal, kernel), o = conv::OverlapSaveConvolve(signal, kernel);
        EXPECT(d.size() == o.size());
        double mx = 0;
        for (size_t i = 0; i < d.size(); i++) mx = std::fmax(mx, std::fabs(d[i] - o[i]));
        EXPECT(mx <= 1e-8);
    }
    {   // ExampleOverlapAdd, example_test.go:55-85
        Vec kernel(64);
        for (size_t i = 0; i < 64; i++) kernel[i] = std::exp(-(double)i / 10);
        auto oa = conv::NewOverlapAdd(kernel, 256);
        EXPECT(oa.BlockSize() == 256 && oa.FFTSize() == 512);
        Vec signal(500, 0.5);
        EXPECT(oa.Process(signal).size() == 563);
    }
    {   // ExampleCorrelate, example_test.go:87-102
        auto r = conv::Correlate({0, 0, 0, 1, 2, 3, 2, 1, 0, 0, 0}, {1, 2, 3, 2, 1});
        auto [idx, val] = conv::FindPeak(r);
        EXPECT(idx == 7 && conv::LagFromIndex(idx, 5) == 3 && std::fabs(val - 19.0) < 1e-9);
    }
    {   // TestOverlapSaveInvalidFFTSize, conv_test.go:675-682
        try { conv::NewOverlapSave({0.25, 0.5, 0.25}, 100); EXPECT(false); }
        catch (const conv::Error &e) { EXPECT(conv::errors_is(e, conv::ErrInvalidBlockSize)); }
    }
    {   // ExampleDeconvolve, example_test.go:129-156: "Recovery SNR: 39.6 dB"
        Vec original(50);
        for (size_t i = 0; i < 50; i++) original[i] = std::sin(2 * M_PI * (double)i / 10);
        Vec kernel{0.25, 0.5, 0.25};
        auto opts = conv::DefaultDeconvOptions();
        opts.Epsilon = 1e-3;
        Vec recovered = conv::Deconvolve(conv::Direct(original, kernel), kernel, opts);
        EXPECT(recovered.size() == 50);
        char buf[32];
        std::snprintf(buf, sizeof buf, "%.1f", conv::SNR(original, recovered));
        EXPECT(std::string(buf) == "39.6");
    }
    {   // ConvolutionReverb: wet/dry in place equals dry*x + wet*(delayed convolution), convolution.go:60-83
        Vec kernel(3000);
        for (size_t i = 0; i < kernel.size(); i++) kernel[i] = std::pow(0.999, (double)i) * ((i % 7) ? 0.1 : -0.2);
        Vec x(8000);
        for (size_t i = 0; i < x.size(); i++) x[i] = std::sin(0.01 * (double)i) + ((i * 2654435761u) % 1000) / 1000.0 - 0.5;
        auto rv = conv::NewConvolutionReverb(kernel, 7);
        rv.SetWetDry(0.25, 0.5);
        Vec blk = x;
        rv.ProcessInPlace(blk);
        Vec full = conv::OverlapSaveConvolve(x, kernel);
        const int L = rv.Latency();
        EXPECT(L == 128);
        double mx = 0;
        for (size_t i = 0; i < x.size(); i++) {
            const double wet = (i >= (size_t)L) ? full[i - L] : 0.0;
            mx = std::fmax(mx, std::fabs(blk[i] - (0.5 * x[i] + 0.25 * wet)));
        }
        EXPECT(mx <= 1e-10);
    }
    std::printf("conv_host_test: ok\n");
    return 0;
}
