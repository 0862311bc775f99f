// sfft_kernel.cuh — the fast path: compile-time-scheduled Stockham FFT, register resident between passes.
//
// For lengths with an instantiated schedule (see sfft_registry.h) one CTA transforms L lanes:
//   * every thread owns E points of one lane, X[i + TL*e] (TL = N/E threads per lane), so the FIRST radix pass reads
//     straight from global memory and the LAST pass writes straight to global memory, both coalesced
//     (along the axis for contiguous rows, across L adjacent lanes for strided columns);
//   * between passes the points are exchanged through shared memory in Stockham autosort order
//     (write Y[(b-k) r + k + q P], read X[i + TL e]); with P passes that is P-1 exchanges, each
//     bank-conflict free thanks to one pad element per r0 points;
//   * twiddles come from one table W_N^t in global memory (L1/L2 resident);
//   * inverse transforms are conj-in / conj-out of the forward schedule; 1/n (Normalization::Default,
//     src/lib.rs:333-338) and the four-step inter-pass twiddle are fused into the store.
// Replaces fft_lane / ifft_lane (src/lib.rs:313-331) plus the lane loop and copies of src/lib.rs:119-163.
#pragma once
#include "butterflies.cuh"
#include "common.h"

namespace ndfb {

struct SfftArgs {
    const void* in;
    void* out;
    long long nlanes;
    int nbd;
    long long bsz[kMaxBatchDims], bis[kMaxBatchDims], bos[kMaxBatchDims];
    long long is_axis, os_axis;
    int conj_in, conj_out;
    double scale;
    const void* tw;  // W_N^t, t < N
    int fs_twiddle, fs_shift;
    const void* fs_lo;
    const void* fs_hi;
};

template <int N_, int TL_, int R0_, int R1_ = 1, int R2_ = 1, int R3_ = 1>
struct Sched {
    static constexpr int N = N_;
    static constexpr int TL = TL_;  // threads cooperating on one lane
    static constexpr int R0 = R0_, R1 = R1_, R2 = R2_, R3 = R3_;
    static constexpr int NP = R1_ == 1 ? 1 : (R2_ == 1 ? 2 : (R3_ == 1 ? 3 : 4));
    static_assert(R0_ * R1_ * R2_ * R3_ == N_, "radices must multiply to N");
    static constexpr int cmax(int a, int b) { return a > b ? a : b; }
    static constexpr int radix(int p) { return p == 0 ? R0_ : p == 1 ? R1_ : p == 2 ? R2_ : R3_; }
    static constexpr int before(int p) { return p == 0 ? 1 : p == 1 ? R0_ : p == 2 ? R0_ * R1_ : R0_ * R1_ * R2_; }
    static constexpr int nbf(int p) { return N_ / radix(p); }                   // butterflies per lane in pass p
    static constexpr int G(int p) { return (nbf(p) + TL_ - 1) / TL_; }          // butterflies per thread
    static constexpr int EP(int p) { return p < NP ? G(p) * radix(p) : 0; }
    static constexpr int E = cmax(cmax(EP(0), EP(1)), cmax(EP(2), EP(3)));      // register slots per thread
    // one pad element per R0 points keeps the stride-R0 writes of pass 0 off a single bank
    static constexpr int pad(int a) { return a + a / R0_; }
    static constexpr int NPAD = N_ + N_ / R0_;
};

template <typename R, class S, int L, bool COLS>
struct SfftCtx {
    Cx<R>* smem;
    const Cx<R>* in;
    Cx<R>* out;
    long long is_axis, os_axis;
    int i, l;  // position within the lane group, lane within the tile
    bool valid;
    NDFB_DEV int addr(int a) const {
        const int p = S::pad(a);
        return COLS ? p * L + l : l * S::NPAD + p;
    }
};

// One Stockham pass.  Butterfly b of this pass (b < N/r) reads X[b + q N/r], multiplies by W_{P r}^{q k}
// (k = b mod P, P = product of the earlier radices) and writes Y[(b-k) r + k + q P].
// Pass 0 reads global memory, the last pass writes global memory (both coalesced in b).
template <typename R, class S, int L, bool COLS, int PASS>
struct SfftPass {
    static constexpr int r = S::radix(PASS);
    static constexpr int P = S::before(PASS);
    static constexpr int G = S::G(PASS);
    static constexpr int NB = S::nbf(PASS);
    static constexpr bool FIRST = PASS == 0;
    static constexpr bool LAST = PASS == S::NP - 1;
    static constexpr bool FULL = (NB % S::TL) == 0;  // no idle threads in this pass

    template <typename StoreF>
    static NDFB_DEV void run(const SfftCtx<R, S, L, COLS>& c, Cx<R> (&v)[S::E], const Cx<R>* __restrict__ tw,
                             const SfftArgs& a, StoreF store) {
#pragma unroll
        for (int m = 0; m < G; ++m) {
            const int b = c.i + S::TL * m;
            if (FULL || b < NB) {
#pragma unroll
                for (int q = 0; q < r; ++q) {
                    if (FIRST) {
                        Cx<R> x = c.valid ? c.in[(long long)(b + q * NB) * c.is_axis] : cmake<R>((R)0, (R)0);
                        if (a.conj_in) x.y = -x.y;
                        v[m * r + q] = x;
                    } else {
                        v[m * r + q] = c.smem[c.addr(b + q * NB)];
                    }
                }
            }
        }
        if (!FIRST && !LAST) __syncthreads();  // every thread has read the previous layout before it is overwritten
#pragma unroll
        for (int m = 0; m < G; ++m) {
            const int b = c.i + S::TL * m;
            if (FULL || b < NB) {
                const int k = b % P;
                if (!FIRST) {
                    constexpr int step = S::N / (P * r);
#pragma unroll
                    for (int q = 1; q < r; ++q) v[m * r + q] = cmul(v[m * r + q], ldg(&tw[q * k * step]));
                }
                Dft<R, r>::run(&v[m * r]);
                if (LAST) {
#pragma unroll
                    for (int q = 0; q < r; ++q) store(b + q * NB, v[m * r + q]);   // (b-k) r + k + q P with P = N/r, k = b
                } else {
#pragma unroll
                    for (int q = 0; q < r; ++q) c.smem[c.addr((b - k) * r + k + q * P)] = v[m * r + q];
                }
            }
        }
        if (!LAST) __syncthreads();
    }
};

template <typename R, class S, int L, bool COLS, int PASS>
struct SfftAll {
    template <typename StoreF>
    static NDFB_DEV void run(const SfftCtx<R, S, L, COLS>& c, Cx<R> (&v)[S::E], const Cx<R>* tw, const SfftArgs& a, StoreF store) {
        SfftPass<R, S, L, COLS, PASS>::run(c, v, tw, a, store);
        if constexpr (PASS + 1 < S::NP) SfftAll<R, S, L, COLS, PASS + 1>::run(c, v, tw, a, store);
    }
};

template <typename R, class S, int L, bool COLS, int MINB>
__global__ void __launch_bounds__(S::TL* L, MINB) sfft_kernel(const __grid_constant__ SfftArgs a) {
    NDFB_DYN_SMEM(smem_raw);
    SfftCtx<R, S, L, COLS> c;
    c.smem = reinterpret_cast<Cx<R>*>(smem_raw);
    const int tid = threadIdx.x;
    if (COLS) { c.l = tid % L; c.i = tid / L; }
    else { c.i = tid % S::TL; c.l = tid / S::TL; }
    // this thread's lane in the global arrays
    long long g = (long long)blockIdx.x * L + c.l;
    c.valid = g < a.nlanes;
    long long bi = 0, bo = 0;
    int j2 = 0;
    if (c.valid) {
#pragma unroll
        for (int d = 0; d < kMaxBatchDims; ++d) {
            if (d < a.nbd) {
                const long long q = g / a.bsz[d];
                const long long rr = g - q * a.bsz[d];
                if (d == 0) j2 = (int)rr;
                bi += rr * a.bis[d];
                bo += rr * a.bos[d];
                g = q;
            }
        }
    }
    c.in = reinterpret_cast<const Cx<R>*>(a.in) + bi;
    c.out = reinterpret_cast<Cx<R>*>(a.out) + bo;
    c.is_axis = a.is_axis;
    c.os_axis = a.os_axis;
    const Cx<R>* __restrict__ tw = reinterpret_cast<const Cx<R>*>(a.tw);
    const R sc = (R)a.scale;
    const R sy = a.conj_out ? -sc : sc;
    Cx<R> v[S::E];
    auto store = [&](int k, Cx<R> val) {
        if (!c.valid) return;
        Cx<R> y = cmake<R>(val.x * sc, val.y * sy);
        if (a.fs_twiddle) {
            const Cx<R>* lo = reinterpret_cast<const Cx<R>*>(a.fs_lo);
            const Cx<R>* hi = reinterpret_cast<const Cx<R>*>(a.fs_hi);
            const unsigned long long ee = (unsigned long long)k * (unsigned long long)j2;
            y = cmul(y, cmul(ldg(&hi[ee >> a.fs_shift]), ldg(&lo[ee & ((1ull << a.fs_shift) - 1)])));
        }
        c.out[(long long)k * c.os_axis] = y;
    };
    SfftAll<R, S, L, COLS, 0>::run(c, v, tw, a, store);
}

}  // namespace ndfb
