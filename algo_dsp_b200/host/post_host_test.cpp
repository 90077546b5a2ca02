// Reads like the reference's filter_test.go / resample_test.go / ir_test.go / sweep_test.go, against the C++ host mirror.
// Build: g++ -std=c++17 post_host_test.cpp -o post_host_test -L.. -lalgodsp_cuda -Wl,-rpath,$PWD/..
#include <algorithm>
#include <cmath>
#include <cstdio>

#include "post.hpp"

#define EXPECT(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main() {
    using conv::Vec;
    {   // fir: impulse response equals the coefficients, Order, Reset (filter_test.go:60-112)
        Vec h{0.25, 0.5, 0.25};
        auto f = fir::New(h);
        EXPECT(f.Order() == 2);
        Vec blk{1, 0, 0, 0, 0};
        f.ProcessBlock(blk);
        EXPECT(std::fabs(blk[0] - 0.25) < 1e-15 && std::fabs(blk[1] - 0.5) < 1e-15 && std::fabs(blk[2] - 0.25) < 1e-15 && blk[3] == 0 && blk[4] == 0);
        Vec a{1, 1}, b{1, 1, 1};
        f.Reset();
        f.ProcessBlock(a);                                  // the delay line carries over: a step through two blocks
        f.ProcessBlock(b);
        EXPECT(std::fabs(a[0] - 0.25) < 1e-15 && std::fabs(a[1] - 0.75) < 1e-15 && std::fabs(b[0] - 1.0) < 1e-15 && std::fabs(b[2] - 1.0) < 1e-15);
        // a 257-tap symmetric low-pass over three 8192-sample segments equals the direct convolution
        Vec lp(257);
        double s = 0;
        for (size_t i = 0; i < lp.size(); i++) { lp[i] = 0.5 - 0.5 * std::cos(2 * M_PI * (double)i / 256.0); s += lp[i]; }
        for (auto &v : lp) v /= s;
        Vec x = signal::WhiteNoise(20000, 1.0, 3);
        Vec want = conv::Direct(x, lp);
        auto g = fir::New(lp);
        Vec y = x;
        g.ProcessBlock(y);
        double mx = 0;
        for (size_t i = 0; i < y.size(); i++) mx = std::fmax(mx, std::fabs(y[i] - want[i]));
        EXPECT(mx <= 1e-13);
    }
    {   // resample: ratio reduction, rate constructor, output length, chunked == whole (resample_test.go:20-111)
        auto r = resample::NewRational(320, 294);
        EXPECT(r.Ratio() == std::make_pair(160, 147));
        EXPECT(resample::NewForRates(44100, 48000).Ratio() == std::make_pair(160, 147));
        EXPECT(resample::approximateRatio(48000.0 / 44100.0) == std::make_pair(160, 147));
        Vec x(6000);
        for (size_t i = 0; i < x.size(); i++) x[i] = std::sin(2 * M_PI * 1000.0 * (double)i / 44100.0);
        Vec whole = r.Process(x);
        EXPECT(std::llabs((long long)whole.size() - std::llround(6000.0 * 160 / 147)) <= 1);
        auto rc = resample::NewRational(160, 147);
        Vec parts;
        for (size_t i = 0; i < x.size(); i += 257) {
            Vec p = rc.Process(Vec(x.begin() + i, x.begin() + std::min(x.size(), i + 257)));
            parts.insert(parts.end(), p.begin(), p.end());
        }
        EXPECT(parts == whole);
        // the tone survives: RMS after the transient within 1 % (44.1k -> 48k, resample_test.go:60-88)
        double e = 0;
        for (size_t i = 1000; i < whole.size(); i++) e += whole[i] * whole[i];
        EXPECT(std::fabs(std::sqrt(e / (double)(whole.size() - 1000)) - std::sqrt(0.5)) < 0.01);
        try { resample::NewRational(0, 1); EXPECT(false); } catch (const conv::Error &) {}
    }
    {   // ir: Schroeder integral of an exponential decay starts at 0 dB and never rises; onset of a delayed impulse (ir_test.go:62-105, 344-410)
        Vec h(4800);
        for (size_t i = 0; i < h.size(); i++) h[i] = std::exp(-6.9 * (double)i / 4800.0) * ((i % 2) ? 1.0 : -1.0);
        ir::Analyzer an(48000);
        Vec sch = an.SchroederIntegral(h);
        EXPECT(sch.size() == h.size() && std::fabs(sch[0]) < 1e-9);
        for (size_t i = 1; i < sch.size(); i++) EXPECT(sch[i] <= sch[i - 1] + 1e-12);
        Vec d(1000, 0.0);
        d[300] = 1.0; d[120] = 0.05;
        EXPECT(an.FindImpulseStart(d) == 300);
        try { an.SchroederIntegral({}); EXPECT(false); } catch (const conv::Error &e) { EXPECT(conv::errors_is(e, ir::ErrEmptyIR)); }
    }
    {   // sweep: deconvolving the sweep itself gives an impulse near samples - 1 (sweep_test.go:150-238)
        sweep::LogSweep sw{20, 20000, 0.5, 48000};
        EXPECT(sw.samples() == 24000);
        Vec s = sw.Generate();
        Vec irr = sw.Deconvolve(s);
        EXPECT(irr.size() == 2 * 24000 - 1);
        size_t pk = 0;
        for (size_t i = 0; i < irr.size(); i++) if (std::fabs(irr[i]) > std::fabs(irr[pk])) pk = i;
        EXPECT(std::llabs((long long)pk - 23999) <= 2);
    }
    {   // signal: sample i of a stream depends on (seed, i) alone -- a shard generates exactly its own samples
        Vec a = signal::WhiteNoise(150, 1.0, 7), b = signal::WhiteNoise(100, 1.0, 7, 50);
        EXPECT(std::equal(b.begin(), b.end(), a.begin() + 50));
        Vec p = signal::PinkNoise(300, 1.0, 9), q = signal::PinkNoise(100, 1.0, 9, 200);
        EXPECT(std::equal(q.begin(), q.end(), p.begin() + 200));
        EXPECT(signal::DecayingIR(64, 3.0, 1).size() == 64);
    }
    std::printf("post_host_test: ok\n");
    return 0;
}
