// plan.h — host-side plan builder: factorisation into radix passes, Bluestein length choice, twiddle /
// chirp / permutation tables (computed in long double, rounded once to the working precision).
//
// Replaces: rustfft::FftPlanner::plan_fft_forward/inverse, realfft::RealFftPlanner, rustdct::DctPlanner as
// called from FftHandler::new / R2cFftHandler::new / DctHandler::new (src/lib.rs:294-304, 477-488, 665-679).
#pragma once
#include <cmath>
#include <complex>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "common.h"

namespace ndfb {

using cld = std::complex<long double>;

inline const long double kPiL = 3.14159265358979323846264338327950288419716939937510L;

// exp(-2 pi i num/den) with exact octant reduction of the integer phase
inline cld unit_root(long long num, long long den) {
    num %= den;
    if (num < 0) num += den;
    // reduce to first octant for accuracy
    long double ang = 2.0L * kPiL * (long double)num / (long double)den;
    // use symmetry around multiples of pi/2 : exact cases
    if (num == 0) return cld(1.0L, 0.0L);
    if (4 * num == den) return cld(0.0L, -1.0L);
    if (2 * num == den) return cld(-1.0L, 0.0L);
    if (4 * num == 3 * den) return cld(0.0L, 1.0L);
    return cld(cosl(ang), -sinl(ang));
}

inline bool factorize(int N, std::vector<int>& out) {
    out.clear();
    int rem = N;
    while (rem % 16 == 0) { out.push_back(16); rem /= 16; }
    while (rem % 8 == 0) { out.push_back(8); rem /= 8; }
    while (rem % 4 == 0) { out.push_back(4); rem /= 4; }
    while (rem % 2 == 0) { out.push_back(2); rem /= 2; }
    const int odd[] = {3, 5, 7, 11, 13};
    for (int r : odd)
        while (rem % r == 0) { out.push_back(r); rem /= r; }
    return rem == 1;
}

inline bool is_smooth(long long N) {
    const int p[] = {2, 3, 5, 7, 11, 13};
    for (int r : p)
        while (N % r == 0) N /= r;
    return N == 1;
}

// smallest 2^a 3^b 5^c >= v (a >= 1)
inline int bluestein_length(int v) {
    long long best = 0;
    for (long long p5 = 1; p5 < 4LL * v; p5 *= 5)
        for (long long p3 = p5; p3 < 4LL * v; p3 *= 3) {
            long long m = p3 * 2;
            while (m < v) m *= 2;
            if (best == 0 || m < best) best = m;
        }
    return (int)best;
}

// In-place DIF over radices (naive DFT_r inside), long double; leaves digit-reversed order.
inline void host_dif(std::vector<cld>& x, const std::vector<int>& radices) {
    const int N = (int)x.size();
    int ns = N;
    std::vector<cld> u(16), v(16);
    for (int r : radices) {
        const int s = ns / r;
        for (int b = 0; b < N / r; ++b) {
            int blk = b / s, j = b % s, base = blk * ns + j;
            for (int q = 0; q < r; ++q) u[q] = x[base + q * s];
            for (int k = 0; k < r; ++k) {
                cld acc(0, 0);
                for (int q = 0; q < r; ++q) acc += u[q] * unit_root((long long)q * k, r);
                v[k] = acc * unit_root((long long)k * j, ns);
            }
            for (int q = 0; q < r; ++q) x[base + q * s] = v[q];
        }
        ns = s;
    }
}

inline std::vector<uint32_t> dif_perm(int N, const std::vector<int>& radices) {
    std::vector<uint32_t> p(N);
    for (int k = 0; k < N; ++k) {
        int kk = k, pos = 0, div = N;
        for (int r : radices) {
            div /= r;
            pos += (kk % r) * div;
            kk /= r;
        }
        p[k] = (uint32_t)pos;
    }
    return p;
}

// Tables of one (tile kind, n) schedule, still in long double.
struct CoreTables {
    int kind = 0, n = 0, N = 0, M = 0, B = 0;
    int n_in = 0, n_out = 0, Bl = 0;
    bool in_complex = false, out_complex = false;
    std::vector<int> radix;
    std::vector<cld> tw, tabA, tabB, blu_c, blu_bhat;
    std::vector<uint32_t> perm;
};

inline void build_core(CoreTables& t, int kind, int n, bool tables_only = false) {
    t.kind = kind;
    t.n = n;
    const int m = n / 2 + 1;
    switch (kind) {
        case TK_C2C: t.N = n; t.n_in = n; t.n_out = n; t.in_complex = t.out_complex = true; break;
        case TK_R2C_EVEN: t.N = n / 2; t.n_in = n; t.n_out = m; t.out_complex = true; break;
        case TK_R2C_ODD: t.N = n; t.n_in = n; t.n_out = m; t.out_complex = true; break;
        case TK_C2R_EVEN: t.N = n / 2; t.n_in = m; t.n_out = n; t.in_complex = true; break;
        case TK_C2R_ODD: t.N = n; t.n_in = m; t.n_out = n; t.in_complex = true; break;
        case TK_DCT1: t.N = n - 1; t.n_in = n; t.n_out = n; break;
        case TK_DCT2_EVEN: case TK_DCT3_EVEN: case TK_DCT4_EVEN: t.N = n / 2; t.n_in = n; t.n_out = n; break;
        case TK_DCT2_ODD: case TK_DCT3_ODD: t.N = n; t.n_in = n; t.n_out = n; break;
        case TK_DCT4_ODD: t.N = 2 * n; t.n_in = n; t.n_out = n; break;
    }
    const int N = t.N;
    if (tables_only) {
        // staged path: only the kind's pre/post tables are needed (the core runs as separate launches)
        t.M = 0; t.B = 0; t.radix.clear();
    } else if (N <= 1 || factorize(N, t.radix)) {
        if (N <= 1) t.radix.clear();
        t.M = 0;
        t.B = N;
        t.perm = dif_perm(N, t.radix);
    } else {
        t.M = bluestein_length(2 * N - 1);
        t.B = t.M;
        factorize(t.M, t.radix);
        t.blu_c.resize(N);
        for (long long j = 0; j < N; ++j) t.blu_c[j] = unit_root((j * j) % (2LL * N), 2LL * N);
        std::vector<cld> b(t.M, cld(0, 0));
        for (int j = 0; j < N; ++j) {
            b[j] = std::conj(t.blu_c[j]);
            if (j > 0) b[t.M - j] = std::conj(t.blu_c[j]);
        }
        host_dif(b, t.radix);
        for (auto& v : b) v /= (long double)t.M;
        t.blu_bhat = std::move(b);
    }
    if (!tables_only) {
        t.tw.resize(t.B > 0 ? t.B : 1);
        for (int k = 0; k < t.B; ++k) t.tw[k] = unit_root(k, t.B);
    }
    // lane slots: the core buffer, plus the one extra bin the "zip" kinds read
    t.Bl = t.B;
    if (kind == TK_C2R_EVEN || kind == TK_DCT3_EVEN) t.Bl = std::max(t.Bl, N + 1);
    if (t.Bl < 1) t.Bl = 1;
    // kind tables
    switch (kind) {
        case TK_R2C_EVEN: case TK_C2R_EVEN: case TK_DCT1: case TK_DCT2_EVEN: case TK_DCT3_EVEN:
            t.tabA.resize(N + 1);
            for (int k = 0; k <= N; ++k) t.tabA[k] = unit_root(k, 2LL * N);  // exp(-2 pi i k / (2N))
            break;
        case TK_DCT4_EVEN:
            t.tabA.resize(N);
            for (int j = 0; j < N; ++j) t.tabA[j] = unit_root(j, 2LL * n);   // exp(-i pi j / n)
            break;
        case TK_DCT4_ODD:
            t.tabA.resize(n);
            for (int j = 0; j < n; ++j) t.tabA[j] = unit_root(j, 4LL * n);   // exp(-i pi j / (2n))
            break;
        default: break;
    }
    switch (kind) {
        case TK_DCT2_EVEN: case TK_DCT2_ODD: case TK_DCT3_EVEN: case TK_DCT3_ODD:
            t.tabB.resize(n);
            for (int k = 0; k < n; ++k) t.tabB[k] = unit_root(k, 4LL * n);   // exp(-i pi k / (2n))
            break;
        case TK_DCT4_EVEN:
            t.tabB.resize(N);
            for (int j = 0; j < N; ++j) t.tabB[j] = unit_root(4LL * j + 1, 8LL * n);  // exp(-i pi (4j+1)/(4n))
            break;
        case TK_DCT4_ODD:
            t.tabB.resize(n);
            for (int k = 0; k < n; ++k) t.tabB[k] = unit_root(2LL * k + 1, 8LL * n);  // exp(-i pi (2k+1)/(4n))
            break;
        default: break;
    }
}

}  // namespace ndfb
