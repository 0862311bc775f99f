// butterflies.cuh — register-resident forward DFTs of length 2..16 (sign e^{-2 pi i jk/r}), natural order in and out.
#pragma once
#include "common.h"

namespace ndfb {

// cos/sin(2 pi m / r) for the odd-prime butterflies (40-digit mpmath values rounded to double).
template <int R> NDFB_HD constexpr double odd_cos(int m);
template <int R> NDFB_HD constexpr double odd_sin(int m);

template <> NDFB_HD constexpr double odd_cos<3>(int m) {
    constexpr double t[3] = {1.0, -0.5, -0.5};
    return t[m];
}
template <> NDFB_HD constexpr double odd_sin<3>(int m) {
    constexpr double t[3] = {0.0, 0.8660254037844386467637232, -0.8660254037844386467637232};
    return t[m];
}
template <> NDFB_HD constexpr double odd_cos<5>(int m) {
    constexpr double t[5] = {1.0, 0.3090169943749474241022934, -0.8090169943749474241022934,
                             -0.8090169943749474241022934, 0.3090169943749474241022934};
    return t[m];
}
template <> NDFB_HD constexpr double odd_sin<5>(int m) {
    constexpr double t[5] = {0.0, 0.9510565162951535721164393, 0.587785252292473129168706,
                             -0.587785252292473129168706, -0.9510565162951535721164393};
    return t[m];
}
template <> NDFB_HD constexpr double odd_cos<7>(int m) {
    constexpr double t[7] = {1.0, 0.6234898018587335305250049, -0.2225209339563144042889026,
                             -0.9009688679024191262361023, -0.9009688679024191262361023,
                             -0.2225209339563144042889026, 0.6234898018587335305250049};
    return t[m];
}
template <> NDFB_HD constexpr double odd_sin<7>(int m) {
    constexpr double t[7] = {0.0, 0.7818314824680298087084445, 0.9749279121818236070181317,
                             0.4338837391175581204757683, -0.4338837391175581204757683,
                             -0.9749279121818236070181317, -0.7818314824680298087084445};
    return t[m];
}
template <> NDFB_HD constexpr double odd_cos<11>(int m) {
    constexpr double t[11] = {1.0, 0.8412535328311811688618116, 0.4154150130018864255292741,
                              -0.1423148382732851404437927, -0.6548607339452850640569251,
                              -0.9594929736144973898903681, -0.9594929736144973898903681,
                              -0.6548607339452850640569251, -0.1423148382732851404437927,
                              0.4154150130018864255292741, 0.8412535328311811688618116};
    return t[m];
}
template <> NDFB_HD constexpr double odd_sin<11>(int m) {
    constexpr double t[11] = {0.0, 0.540640817455597582107636, 0.9096319953545183714117154,
                              0.989821441880932732376092, 0.7557495743542582837740358,
                              0.2817325568414296977114179, -0.2817325568414296977114179,
                              -0.7557495743542582837740358, -0.989821441880932732376092,
                              -0.9096319953545183714117154, -0.540640817455597582107636};
    return t[m];
}
template <> NDFB_HD constexpr double odd_cos<13>(int m) {
    constexpr double t[13] = {1.0, 0.8854560256532098959003755, 0.5680647467311558025118076,
                              0.1205366802553230533490677, -0.3546048870425356259696379,
                              -0.7485107481711010986346306, -0.9709418174260520271569823,
                              -0.9709418174260520271569823, -0.7485107481711010986346306,
                              -0.3546048870425356259696379, 0.1205366802553230533490677,
                              0.5680647467311558025118076, 0.8854560256532098959003755};
    return t[m];
}
template <> NDFB_HD constexpr double odd_sin<13>(int m) {
    constexpr double t[13] = {0.0, 0.4647231720437685456560153, 0.8229838658936563945796174,
                              0.9927088740980539928007516, 0.9350162426854148234397846,
                              0.6631226582407952023767855, 0.2393156642875577671487537,
                              -0.2393156642875577671487537, -0.6631226582407952023767855,
                              -0.9350162426854148234397846, -0.9927088740980539928007516,
                              -0.8229838658936563945796174, -0.4647231720437685456560153};
    return t[m];
}

template <typename R>
NDFB_DEV void dft2(Cx<R>& a, Cx<R>& b) {
    Cx<R> t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

// v[0..3] -> DFT4 in natural order
template <typename R>
NDFB_DEV void dft4(Cx<R>& a0, Cx<R>& a1, Cx<R>& a2, Cx<R>& a3) {
    Cx<R> t0 = cadd(a0, a2), t1 = csub(a0, a2);
    Cx<R> t2 = cadd(a1, a3), t3 = cmul_ni(csub(a1, a3));
    a0 = cadd(t0, t2);
    a1 = cadd(t1, t3);
    a2 = csub(t0, t2);
    a3 = csub(t1, t3);
}

template <typename R, int RADIX>
struct Dft;

template <typename R>
struct Dft<R, 2> {
    static NDFB_DEV void run(Cx<R>* v) { dft2(v[0], v[1]); }
};

template <typename R>
struct Dft<R, 4> {
    static NDFB_DEV void run(Cx<R>* v) { dft4(v[0], v[1], v[2], v[3]); }
};

template <typename R>
struct Dft<R, 8> {
    static NDFB_DEV void run(Cx<R>* v) {
        const R h = (R)0.7071067811865475244008444;
        // DFT8 = radix-2 split: evens / odds each a DFT4, then twiddle W8^k on the odd half.
        Cx<R> e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
        Cx<R> o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
        dft4(e0, e1, e2, e3);
        dft4(o0, o1, o2, o3);
        o1 = cmake<R>((o1.x + o1.y) * h, (o1.y - o1.x) * h);    // * (1 - i)/sqrt2
        o2 = cmul_ni(o2);                                        // * -i
        o3 = cmake<R>((o3.y - o3.x) * h, -(o3.x + o3.y) * h);   // * (-1 - i)/sqrt2
        v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
        v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
        v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
        v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
    }
};

template <typename R>
struct Dft<R, 16> {
    static NDFB_DEV void run(Cx<R>* v) {
        // 4 x 4 Cooley-Tukey: j = 4*j1 + j0, k = k0 + 4*k1 ... columns j0 fixed: DFT4 over j1, twiddle W16^{j0*k0}, DFT4 over j0.
        const R c1 = (R)0.9238795325112867561281832, s1 = (R)0.38268343236508977172846;
        const R h = (R)0.7071067811865475244008444;
        Cx<R> a[4][4];  // a[j0][k0]
#pragma unroll
        for (int j0 = 0; j0 < 4; ++j0) {
            Cx<R> x0 = v[j0], x1 = v[j0 + 4], x2 = v[j0 + 8], x3 = v[j0 + 12];
            dft4(x0, x1, x2, x3);
            a[j0][0] = x0; a[j0][1] = x1; a[j0][2] = x2; a[j0][3] = x3;
        }
        // twiddles W16^{j0*k0}: W16^1 = (c1,-s1), W16^2 = (h,-h), W16^3 = (s1,-c1), W16^4 = -i, W16^6 = (-h,-h), W16^9 = (-c1, s1)
        a[1][1] = cmul(a[1][1], cmake<R>(c1, -s1));
        a[1][2] = cmul(a[1][2], cmake<R>(h, -h));
        a[1][3] = cmul(a[1][3], cmake<R>(s1, -c1));
        a[2][1] = cmul(a[2][1], cmake<R>(h, -h));
        a[2][2] = cmul_ni(a[2][2]);
        a[2][3] = cmul(a[2][3], cmake<R>(-h, -h));
        a[3][1] = cmul(a[3][1], cmake<R>(s1, -c1));
        a[3][2] = cmul(a[3][2], cmake<R>(-h, -h));
        a[3][3] = cmul(a[3][3], cmake<R>(-c1, s1));
#pragma unroll
        for (int k0 = 0; k0 < 4; ++k0) {
            Cx<R> x0 = a[0][k0], x1 = a[1][k0], x2 = a[2][k0], x3 = a[3][k0];
            dft4(x0, x1, x2, x3);
            v[k0] = x0; v[k0 + 4] = x1; v[k0 + 8] = x2; v[k0 + 12] = x3;
        }
    }
};

// Odd prime radix by the symmetric O(r^2/2) form:  X[k], X[r-k] = (x0 + sum_j c_jk a_j) -/+ i (sum_j s_jk b_j).
template <typename R, int RADIX>
struct DftOdd {
    static NDFB_DEV void run(Cx<R>* v) {
        constexpr int H = (RADIX - 1) / 2;
        Cx<R> a[H], b[H];
#pragma unroll
        for (int j = 1; j <= H; ++j) {
            a[j - 1] = cadd(v[j], v[RADIX - j]);
            b[j - 1] = csub(v[j], v[RADIX - j]);
        }
        Cx<R> x0 = v[0];
        Cx<R> sum = x0;
#pragma unroll
        for (int j = 0; j < H; ++j) sum = cadd(sum, a[j]);
        v[0] = sum;
#pragma unroll
        for (int k = 1; k <= H; ++k) {
            Cx<R> C = x0, S = cmake<R>((R)0, (R)0);
#pragma unroll
            for (int j = 1; j <= H; ++j) {
                const R c = (R)odd_cos<RADIX>((j * k) % RADIX);
                const R s = (R)odd_sin<RADIX>((j * k) % RADIX);
                C.x += c * a[j - 1].x; C.y += c * a[j - 1].y;
                S.x += s * b[j - 1].x; S.y += s * b[j - 1].y;
            }
            v[k] = cmake<R>(C.x + S.y, C.y - S.x);          // C - i S
            v[RADIX - k] = cmake<R>(C.x - S.y, C.y + S.x);  // C + i S
        }
    }
};

template <typename R> struct Dft<R, 3> : DftOdd<R, 3> {};
template <typename R> struct Dft<R, 5> : DftOdd<R, 5> {};
template <typename R> struct Dft<R, 7> : DftOdd<R, 7> {};
template <typename R> struct Dft<R, 11> : DftOdd<R, 11> {};
template <typename R> struct Dft<R, 13> : DftOdd<R, 13> {};

// ---- radix 32 = 2 x 16 ----
template <typename R>
struct Dft<R, 32> {
    static NDFB_DEV void run(Cx<R>* v) {
        constexpr double wc[16] = {1.0, 0.9807852804032304491262, 0.9238795325112867561282, 0.8314696123025452370788,
                                   0.7071067811865475244008, 0.5555702330196022247428, 0.3826834323650897717285,
                                   0.1950903220161282678483, 0.0, -0.1950903220161282678483, -0.3826834323650897717285,
                                   -0.5555702330196022247428, -0.7071067811865475244008, -0.8314696123025452370788,
                                   -0.9238795325112867561282, -0.9807852804032304491262};
        constexpr double ws[16] = {0.0, 0.1950903220161282678483, 0.3826834323650897717285, 0.5555702330196022247428,
                                   0.7071067811865475244008, 0.8314696123025452370788, 0.9238795325112867561282,
                                   0.9807852804032304491262, 1.0, 0.9807852804032304491262, 0.9238795325112867561282,
                                   0.8314696123025452370788, 0.7071067811865475244008, 0.5555702330196022247428,
                                   0.3826834323650897717285, 0.1950903220161282678483};
        Cx<R> e[16], o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { e[j] = v[2 * j]; o[j] = v[2 * j + 1]; }
        Dft<R, 16>::run(e);
        Dft<R, 16>::run(o);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            Cx<R> t;
            if (k == 0) t = o[0];
            else if (k == 8) t = cmul_ni(o[8]);
            else t = cmul(o[k], cmake<R>((R)wc[k], (R)-ws[k]));
            v[k] = cadd(e[k], t);
            v[k + 16] = csub(e[k], t);
        }
    }
};

// ---- radix 2B (B odd prime) = 2 x B Cooley-Tukey: used for 6 and 10 ----
template <typename R, int B>
struct DftTwiceOdd {
    static NDFB_DEV void run(Cx<R>* v) {
        Cx<R> e[B], o[B];
#pragma unroll
        for (int j = 0; j < B; ++j) { e[j] = v[2 * j]; o[j] = v[2 * j + 1]; }
        Dft<R, B>::run(e);
        Dft<R, B>::run(o);
        v[0] = cadd(e[0], o[0]);
        v[B] = csub(e[0], o[0]);
#pragma unroll
        for (int k = 1; k < B; ++k) {
            // W_{2B}^k = exp(-i pi k / B) = -exp(-2 pi i ((k + B)/2) / B) for odd k, exp(-2 pi i (k/2) / B) for even k
            R c, s;
            if (k % 2 == 0) { c = (R)odd_cos<B>(k / 2); s = (R)odd_sin<B>(k / 2); }
            else { c = -(R)odd_cos<B>((k + B) / 2); s = -(R)odd_sin<B>((k + B) / 2); }
            Cx<R> t = cmul(o[k], cmake<R>(c, -s));
            v[k] = cadd(e[k], t);
            v[k + B] = csub(e[k], t);
        }
    }
};
template <typename R> struct Dft<R, 6> : DftTwiceOdd<R, 3> {};
template <typename R> struct Dft<R, 10> : DftTwiceOdd<R, 5> {};

// ---- composite radix N = A x B by one Cooley-Tukey step with constant twiddles (used for 9 = 3 x 3) ----
template <int N> NDFB_HD constexpr double ct_cos(int m);
template <int N> NDFB_HD constexpr double ct_sin(int m);
template <> NDFB_HD constexpr double ct_cos<9>(int m) {
    constexpr double t[9] = {1.0, 0.7660444431189780352023927, 0.1736481776669303488517166, -0.5, -0.9396926207859083840541093, -0.9396926207859083840541093, -0.5, 0.1736481776669303488517166, 0.7660444431189780352023927};
    return t[m];
}
template <> NDFB_HD constexpr double ct_sin<9>(int m) {
    constexpr double t[9] = {0.0, 0.6427876096865393263226434, 0.984807753012208059366743, 0.8660254037844386467637232, 0.3420201433256687330440996, -0.3420201433256687330440996, -0.8660254037844386467637232, -0.984807753012208059366743, -0.6427876096865393263226434};
    return t[m];
}

template <> NDFB_HD constexpr double ct_cos<12>(int m) {
    constexpr double t[12] = {1.0, 0.8660254037844386467637231708, 0.5, 2.067032109826398823649690305e-43, -0.5, -0.8660254037844386467637231708, -1.0, -0.8660254037844386467637231708, -0.5, 2.23387644065498832429194784e-41, 0.5, 0.8660254037844386467637231708};
    return t[m];
}
template <> NDFB_HD constexpr double ct_sin<12>(int m) {
    constexpr double t[12] = {0.0, 0.5, 0.8660254037844386467637231708, 1.0, 0.8660254037844386467637231708, 0.5, 4.13406421965279764729938061e-43, -0.5, -0.8660254037844386467637231708, -1.0, -0.8660254037844386467637231708, -0.5};
    return t[m];
}
template <> NDFB_HD constexpr double ct_cos<15>(int m) {
    constexpr double t[15] = {1.0, 0.913545457642600895502127572, 0.6691306063588582138262733307, 0.3090169943749474241022934172, -0.1045284632676534713998341548, -0.5, -0.8090169943749474241022934172, -0.9781476007338056379285667479, -0.9781476007338056379285667479, -0.8090169943749474241022934172, -0.5, -0.1045284632676534713998341548, 0.3090169943749474241022934172, 0.6691306063588582138262733307, 0.913545457642600895502127572};
    return t[m];
}
template <> NDFB_HD constexpr double ct_sin<15>(int m) {
    constexpr double t[15] = {0.0, 0.4067366430758002077539859903, 0.743144825477394235014697049, 0.9510565162951535721164393334, 0.994521895368273336922691945, 0.8660254037844386467637231708, 0.5877852522924731291687059546, 0.2079116908177593371017422844, -0.2079116908177593371017422844, -0.5877852522924731291687059546, -0.8660254037844386467637231708, -0.994521895368273336922691945, -0.9510565162951535721164393334, -0.743144825477394235014697049, -0.4067366430758002077539859903};
    return t[m];
}

template <typename R, int A, int B>
struct DftCT {
    static NDFB_DEV void run(Cx<R>* v) {
        constexpr int N = A * B;
        Cx<R> a[A][B];
#pragma unroll
        for (int j0 = 0; j0 < A; ++j0) {
            Cx<R> u[B];
#pragma unroll
            for (int j1 = 0; j1 < B; ++j1) u[j1] = v[j0 + A * j1];
            Dft<R, B>::run(u);
#pragma unroll
            for (int k0 = 0; k0 < B; ++k0) {
                if (j0 * k0 == 0) a[j0][k0] = u[k0];
                else a[j0][k0] = cmul(u[k0], cmake<R>((R)ct_cos<N>((j0 * k0) % N), (R)-ct_sin<N>((j0 * k0) % N)));
            }
        }
#pragma unroll
        for (int k0 = 0; k0 < B; ++k0) {
            Cx<R> w[A];
#pragma unroll
            for (int j0 = 0; j0 < A; ++j0) w[j0] = a[j0][k0];
            Dft<R, A>::run(w);
#pragma unroll
            for (int k1 = 0; k1 < A; ++k1) v[k0 + B * k1] = w[k1];
        }
    }
};
template <typename R> struct Dft<R, 9> : DftCT<R, 3, 3> {};
// 12 = 3 x 4, 15 = 3 x 5, 14 = 2 x 7: wider choice of balanced schedules for the run-time compiled lengths (jit.h)
template <typename R> struct Dft<R, 12> : DftCT<R, 3, 4> {};
template <typename R> struct Dft<R, 15> : DftCT<R, 3, 5> {};
template <typename R> struct Dft<R, 14> : DftTwiceOdd<R, 7> {};

}  // namespace ndfb
