// big_kernels.cuh — element-wise prologue / epilogue kernels of the STAGED path.
//
// Lanes whose transform does not fit one CTA's shared memory (complex rows are handled by the four-step C2C path; this
// is for the real kinds and for lengths with large prime factors) are run as
//     prologue kernel  : strided user array -> contiguous workspace rows z[lane][j]   (same algebra as rsfft_kernel's loads)
//     complex core     : four-step Stockham FFT of the workspace rows, or Bluestein around two of them
//     epilogue kernel  : workspace rows -> strided user array                          (same algebra as tile_kernel's gather)
// i.e. three HBM round trips instead of one: a completeness path, not a fast path.
#pragma once
#include "common.h"
#include "sfft_kernel.cuh"

namespace ndfb {

// BK_C2C, the RKind values of rsfft_kernel (even lengths, packed half-length core), and the odd-length real kinds, which run
// a full-length complex core (DCT-IV: zero-padded 2n-point core) exactly like tile_kernel.cuh does
enum BigKind : int { BK_C2C = 100, BK_R2C_ODD = 201, BK_C2R_ODD = 202, BK_DCT2_ODD = 203, BK_DCT3_ODD = 204, BK_DCT4_ODD = 205 };

struct BigArgs {
    const void* in;       // prologue: user input;        epilogue: unused
    void* out;            // prologue: unused;            epilogue: user output
    void* ws;             // workspace rows (complex), row pitch ldw
    long long ldw;
    long long nlanes;     // total lanes of the call
    long long lane0;      // first lane of this chunk
    long long nchunk;     // lanes in this chunk
    int nbd;
    long long bsz[kMaxBatchDims], bis[kMaxBatchDims], bos[kMaxBatchDims];
    long long is_axis, os_axis;
    int kind;             // BK_C2C or RKind
    int n, N;             // logical length, core length
    int M;                // Bluestein length (0: none); prologue pads rows to M and multiplies by the chirp
    int n_out;            // output lane length
    int conj_in, conj_out;
    double scale;
    const void* tabA;
    const void* tabB;
    const void* chirp;    // c[j] = exp(-i pi j^2 / N), j < N
};

template <typename R>
__global__ void __launch_bounds__(256) big_pro_kernel(const __grid_constant__ BigArgs a) {
    const long long total = a.nchunk * (long long)(a.M ? a.M : a.N);
    const int rowlen = a.M ? a.M : a.N;
    const Cx<R>* __restrict__ tabA = reinterpret_cast<const Cx<R>*>(a.tabA);
    const Cx<R>* __restrict__ tabB = reinterpret_cast<const Cx<R>*>(a.tabB);
    const Cx<R>* __restrict__ chirp = reinterpret_cast<const Cx<R>*>(a.chirp);
    Cx<R>* __restrict__ ws = reinterpret_cast<Cx<R>*>(a.ws);
    const int n = a.n, N = a.N;
    const R zero = (R)0;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long lane = idx / rowlen;
        const int j = (int)(idx - lane * rowlen);
        Cx<R> z = cmake<R>(zero, zero);
        if (j < N) {
            const LaneBase lb = lane_base(a, a.lane0 + lane, true);
            const R* __restrict__ in_r = reinterpret_cast<const R*>(a.in) + lb.bi;   // (only read by the real-input kinds)
            const Cx<R>* __restrict__ in_c = reinterpret_cast<const Cx<R>*>(a.in) + lb.bi;
            auto gin = [&](int t) -> R { return in_r[(long long)t * a.is_axis]; };
            switch (a.kind) {
                case BK_C2C: {
                    z = in_c[(long long)j * a.is_axis];
                    if (a.conj_in) z.y = -z.y;
                } break;
                case RK_R2C: z = cmake<R>(gin(2 * j), gin(2 * j + 1)); break;
                case RK_C2R:
                case RK_DCT3: {
                    Cx<R> xk, xn;
                    const int k2 = N - j;
                    if (a.kind == RK_C2R) {
                        xk = in_c[(long long)j * a.is_axis];
                        xn = in_c[(long long)k2 * a.is_axis];
                        if (j == 0) { xk.y = zero; xn.y = zero; }
                    } else {
                        Cx<R> pk = cmake<R>(gin(j), j == 0 ? zero : -gin(n - j));
                        Cx<R> pn = cmake<R>(gin(k2), -gin(n - k2));
                        xk = cmul(pk, cconj(ldg(&tabB[j])));
                        xn = cmul(pn, cconj(ldg(&tabB[k2])));
                    }
                    const Cx<R> wc = cconj(ldg(&tabA[j]));
                    const Cx<R> E = cadd(xk, cconj(xn)), O = csub(xk, cconj(xn));
                    z = cconj(cadd(E, cmul_i(cmul(wc, O))));
                } break;
                case RK_DCT1: {
                    const int t0 = 2 * j, t1 = 2 * j + 1;
                    z = cmake<R>(gin(t0 <= N ? t0 : 2 * N - t0), gin(t1 <= N ? t1 : 2 * N - t1));
                } break;
                case RK_DCT2: {
                    const int t0 = 2 * j, t1 = 2 * j + 1;
                    z = cmake<R>(gin(t0 < N ? 2 * t0 : 2 * (n - 1 - t0) + 1), gin(t1 < N ? 2 * t1 : 2 * (n - 1 - t1) + 1));
                } break;
                case RK_DCT4: z = cmul(cmake<R>(gin(2 * j), gin(n - 1 - 2 * j)), ldg(&tabA[j])); break;
                case BK_R2C_ODD: z = cmake<R>(gin(j), zero); break;
                case BK_C2R_ODD: {   // conj of the Hermitian completion of the half spectrum (m = n/2 + 1 bins given)
                    const int m = n / 2 + 1;
                    if (j < m) { z = in_c[(long long)j * a.is_axis]; if (j == 0) z.y = zero; z.y = -z.y; }
                    else z = in_c[(long long)(n - j) * a.is_axis];
                } break;
                case BK_DCT2_ODD: {
                    const int h = (n + 1) / 2;
                    z = cmake<R>(gin(j < h ? 2 * j : 2 * (n - 1 - j) + 1), zero);
                } break;
                case BK_DCT3_ODD: {
                    const Cx<R> pk = cmake<R>(gin(j), j == 0 ? zero : -gin(n - j));
                    z = cconj(cmul(pk, cconj(ldg(&tabB[j]))));
                } break;
                case BK_DCT4_ODD: {
                    if (j < n) { const Cx<R> w = ldg(&tabA[j]); const R x = gin(j); z = cmake<R>(x * w.x, x * w.y); }
                } break;
                default: break;
            }
            if (a.M) z = cmul(z, ldg(&chirp[j]));
        }
        ws[lane * a.ldw + j] = z;
    }
}

template <typename R>
__global__ void __launch_bounds__(256) big_epi_kernel(const __grid_constant__ BigArgs a) {
    const long long total = a.nchunk * (long long)a.n_out;
    const Cx<R>* __restrict__ tabA = reinterpret_cast<const Cx<R>*>(a.tabA);
    const Cx<R>* __restrict__ tabB = reinterpret_cast<const Cx<R>*>(a.tabB);
    const Cx<R>* __restrict__ chirp = reinterpret_cast<const Cx<R>*>(a.chirp);
    const Cx<R>* __restrict__ ws = reinterpret_cast<const Cx<R>*>(a.ws);
    const int n = a.n, N = a.N;
    const R sc = (R)a.scale;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long lane = idx / a.n_out;
        const int k = (int)(idx - lane * a.n_out);
        const Cx<R>* __restrict__ row = ws + lane * a.ldw;
        // core output bin (after Bluestein: conj of the second transform times the chirp)
        auto Z = [&](int q) -> Cx<R> { return a.M ? cmul(cconj(row[q]), ldg(&chirp[q])) : row[q]; };
        auto post = [&](int q) -> Cx<R> {
            const Cx<R> zk = Z(q == N ? 0 : q);
            const Cx<R> zc = cconj(Z(q == 0 ? 0 : N - q));
            const Cx<R> w = ldg(&tabA[q]);
            const Cx<R> s = cadd(zk, zc), d = cmul(w, csub(zk, zc));
            return cmake<R>((R)0.5 * (s.x + d.y), (R)0.5 * (s.y - d.x));
        };
        const LaneBase lb = lane_base(a, a.lane0 + lane, true);
        R* __restrict__ out_r = reinterpret_cast<R*>(a.out) + lb.bo;             // (only written by the real-output kinds)
        Cx<R>* __restrict__ out_c = reinterpret_cast<Cx<R>*>(a.out) + lb.bo;
        const long long g = (long long)k * a.os_axis;
        switch (a.kind) {
            case BK_C2C: {
                Cx<R> y = Z(k);
                if (a.conj_out) y.y = -y.y;
                out_c[g] = cscale(y, sc);
            } break;
            case RK_R2C: out_c[g] = cscale(post(k), sc); break;
            case RK_C2R: {
                const Cx<R> y = Z(k >> 1);
                out_r[g] = sc * ((k & 1) ? -y.y : y.x);
            } break;
            case RK_DCT1: out_r[g] = sc * (R)0.5 * post(k).x; break;
            case RK_DCT2: {
                const int kk = k <= N ? k : n - k;
                const Cx<R> A = cmul(post(kk), ldg(&tabB[kk]));
                out_r[g] = sc * (k <= N ? A.x : -A.y);
            } break;
            case RK_DCT3: {
                const int vi = (k & 1) ? (n - 1 - (k >> 1)) : (k >> 1);
                const Cx<R> y = Z(vi >> 1);
                out_r[g] = sc * (R)0.5 * ((vi & 1) ? -y.y : y.x);
            } break;
            case RK_DCT4: {
                const int j = (k & 1) ? ((n - 1 - k) >> 1) : (k >> 1);
                const Cx<R> C = cmul(Z(j), ldg(&tabB[j]));
                out_r[g] = sc * ((k & 1) ? -C.y : C.x);
            } break;
            case BK_R2C_ODD: out_c[g] = cscale(Z(k), sc); break;
            case BK_C2R_ODD: out_r[g] = sc * Z(k).x; break;
            case BK_DCT2_ODD: out_r[g] = sc * cmul(Z(k), ldg(&tabB[k])).x; break;
            case BK_DCT3_ODD: {
                const int vi = (k & 1) ? (n - 1 - (k >> 1)) : (k >> 1);
                out_r[g] = sc * (R)0.5 * Z(vi).x;
            } break;
            case BK_DCT4_ODD: out_r[g] = sc * cmul(Z(k), ldg(&tabB[k])).x; break;
            default: break;
        }
    }
}

}  // namespace ndfb
