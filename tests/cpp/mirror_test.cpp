// The reference's unit tests (src/lib.rs:903-1406) transliterated against the C++ mirror header.
// Usage: mirror_test <goldens.txt> <mode>   mode = "errors" (no GPU needed) | "gpu"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>

#include "ndrustfft_b200.hpp"

using namespace ndrustfft_b200;
using cd = std::complex<double>;

static std::map<std::string, std::vector<double>> load(const char* path) {
    std::map<std::string, std::vector<double>> m;
    std::ifstream f(path);
    std::string name; size_t count;
    while (f >> name >> count) { std::vector<double> v(count); for (auto& x : v) f >> x; m[name] = v; }
    return m;
}
static int fails = 0;
static void check(bool ok, const char* what) { if (!ok) { std::printf("FAIL %s\n", what); ++fails; } }
static bool close_to(const std::vector<double>& a, const std::vector<double>& b, double tol) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); ++i) if (std::fabs(a[i] - b[i]) > tol) return false;
    return true;
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    auto g = load(argv[1]);
    const bool gpu = std::strcmp(argv[2], "gpu") == 0;
    // error behaviour needs no device: size mismatch text, axis range
    {
        FftHandler<double> h(6);
        std::vector<cd> a(10), b(10);
        auto va = ndview<cd>::c_order(a.data(), {2, 5}); auto vb = ndview<cd>::c_order(b.data(), {2, 5});
        try { ndfft(va, vb, h, 1); check(false, "size mismatch not raised"); }
        catch (const std::runtime_error& e) { check(std::string(e.what()) == "Size mismatch in fft, got 5 expected 6", e.what()); }
        DctHandler<double> hd(6);
        std::vector<double> ra(10), rb(10);
        auto ra_v = ndview<double>::c_order(ra.data(), {5, 2}); auto rb_v = ndview<double>::c_order(rb.data(), {5, 2});
        try { nddct2(ra_v, rb_v, hd, 0); check(false, "dct size mismatch not raised"); }
        catch (const std::runtime_error& e) { check(std::string(e.what()) == "Size mismatch in dct, got 5 expected 6", e.what()); }
    }
    if (gpu) {
        const auto& tm = g["test_matrix"];
        // test_fft (src/lib.rs:903-947)
        std::vector<cd> v(36), vhat(36), back(36);
        for (int i = 0; i < 36; ++i) v[i] = cd(tm[i], tm[i]);
        FftHandler<double> h(6);
        auto vv = ndview<cd>::c_order(v.data(), {6, 6}); auto vh = ndview<cd>::c_order(vhat.data(), {6, 6}); auto vb = ndview<cd>::c_order(back.data(), {6, 6});
        ndfft(vv, vh, h, 1); ndifft(vh, vb, h, 1);
        std::vector<double> re(36), im(36), bre(36);
        for (int i = 0; i < 36; ++i) { re[i] = vhat[i].real(); im[i] = vhat[i].imag(); bre[i] = back[i].real(); }
        check(close_to(re, g["fft_re"], 1e-3) && close_to(im, g["fft_im"], 1e-3), "test_fft");
        check(close_to(bre, tm, 1e-3), "test_fft roundtrip");
        // test_dct1..4 (src/lib.rs:1204-1406)
        for (int k = 1; k <= 4; ++k) {
            std::vector<double> x(tm), y(36);
            DctHandler<double> hd(6);
            auto xv = ndview<double>::c_order(x.data(), {6, 6}); auto yv = ndview<double>::c_order(y.data(), {6, 6});
            if (k == 1) nddct1(xv, yv, hd, 1); else if (k == 2) nddct2(xv, yv, hd, 1); else if (k == 3) nddct3_par(xv, yv, hd, 1); else nddct4(xv, yv, hd, 1);
            check(close_to(y, g["dct" + std::to_string(k)], 1e-3), "test_dct");
        }
        // test_fft_r2c + custom normalisation on the c2r path
        std::vector<double> x(tm), xb(36);
        std::vector<cd> sp(24);
        R2cFftHandler<double> hr(6);
        auto xv = ndview<double>::c_order(x.data(), {6, 6}); auto sv = ndview<cd>::c_order(sp.data(), {6, 4}); auto bv = ndview<double>::c_order(xb.data(), {6, 6});
        ndfft_r2c(xv, sv, hr, 1);
        std::vector<double> sre(24), sim(24);
        for (int i = 0; i < 24; ++i) { sre[i] = sp[i].real(); sim[i] = sp[i].imag(); }
        check(close_to(sre, g["r2c_re"], 1e-3) && close_to(sim, g["r2c_im"], 1e-3), "test_fft_r2c");
        auto hc = hr.normalization(Normalization<cd>::custom([](cd* d, size_t n) { for (size_t i = 0; i < n; ++i) d[i] *= 1.0 / 6.0; }));
        ndifft_r2c(sv, bv, hc, 1);
        check(close_to(xb, tm, 1e-3), "c2r custom norm roundtrip");
    }
    if (gpu) {
        // examples/fft2.rs and examples/rfft2.rs through the one-call compositions
        const double xin[9] = {1, 2, 3, 4, 5, 6, 7, 8, 9};
        std::vector<cd> v(9), vhat(9), back(9), work(9), want(9);
        for (int i = 0; i < 9; ++i) v[i] = cd(xin[i], xin[i]);
        FftHandler<double> h0(3), h1(3);
        auto vv = ndview<cd>::c_order(v.data(), {3, 3}); auto vh = ndview<cd>::c_order(vhat.data(), {3, 3});
        auto vb = ndview<cd>::c_order(back.data(), {3, 3}); auto vw = ndview<cd>::c_order(work.data(), {3, 3}); auto vt = ndview<cd>::c_order(want.data(), {3, 3});
        fft2(vv, vh, h0, h1);
        ndfft(vv, vw, h1, 1); ndfft(vw, vt, h0, 0);
        bool same = true;
        for (int i = 0; i < 9; ++i) same = same && std::abs(vhat[i] - want[i]) < 1e-12;
        check(same && std::abs(vhat[0] - cd(45, 45)) < 1e-9, "fft2 == two ndfft calls (examples/fft2.rs)");
        ifft2(vh, vb, h0, h1);
        for (int i = 0; i < 9; ++i) same = same && std::abs(back[i] - v[i]) < 1e-9;
        check(same, "ifft2 roundtrip");
        std::vector<double> x(xin, xin + 9), xb(9);
        std::vector<cd> sp(6);
        R2cFftHandler<double> hr(3);
        auto xv = ndview<double>::c_order(x.data(), {3, 3}); auto sv = ndview<cd>::c_order(sp.data(), {3, 2}); auto bv = ndview<double>::c_order(xb.data(), {3, 3});
        rfft2(xv, sv, h0, hr);
        check(std::abs(sp[0] - cd(45, 0)) < 1e-9 && std::abs(sp[1] - cd(-4.5, 2.59808)) < 1e-4 && std::abs(sp[2] - cd(-13.5, 7.79423)) < 1e-4, "rfft2 (examples/rfft2.rs)");
        irfft2(sv, bv, h0, hr);
        bool rt = true;
        for (int i = 0; i < 9; ++i) rt = rt && std::fabs(xb[i] - xin[i]) < 1e-9;
        check(rt, "irfft2 roundtrip");
    }
    std::printf(fails ? "MIRROR_FAIL %d\n" : "MIRROR_OK\n", fails);
    return fails ? 1 : 0;
}
