// tile_kernel.cuh — the general single-pass axis-transform kernel.
//
// One CTA owns a TILE of L lanes (L a power of two) of one nd* call and keeps it in shared memory from the
// first global read to the last global write, so HBM traffic is exactly input-once + output-once:
//
//   stage-in    global -> shared, coalesced along whichever index is contiguous in the INPUT array
//               (axis index for path A rows, lane index for path B/C strided columns; src/lib.rs:119-163),
//               scattering each element straight to the slot its transform kind needs (r2c packing,
//               Makhoul reorder, even extension, Hermitian completion ...);
//   prologue    pointwise / pairwise in-place fix-up (c2r "zip", DCT-III/IV pre-twiddles, Bluestein chirp);
//   core        in-place decimation-in-frequency radix passes (2,3,4,5,7,8,11,13,16), one __syncthreads each;
//               Bluestein lengths run DIF -> pointwise multiply -> DIT so no reordering pass is needed;
//   epilogue    gather from shared (through the DIF digit-reversal table) with the kind's post-twiddle and
//               normalisation fused, coalesced along whichever index is contiguous in the OUTPUT array.
//
// Replaces: the lane loop + per-lane copies of create_transform!/create_transform_par! (src/lib.rs:100-238) and
// the engine calls inside fft_lane/ifft_lane (:313-331), fft_r2c_lane/ifft_r2c_lane (:497-523), dct1..4_lane (:688-734).
#pragma once
#include "butterflies.cuh"
#include "common.h"

namespace ndfb {

template <typename R>
struct TileCtx {
    const TileArgs& a;
    Cx<R>* buf;          // L * Bl complex slots (+padding)
    long long* lb_in;    // per-lane global base offsets (elements)
    long long* lb_out;
    int* lane_j2;        // four-step: index of the lane along the fastest batch dim
    int tid, T;
    int nl;              // valid lanes in this tile

    NDFB_DEV int phys(int l, int p) const {
        int pp = p + (p >> a.pad_shift);
        return l * a.LP + pp * a.EP;
    }
    NDFB_DEV Cx<R>& slot(int l, int p) const { return buf[phys(l, p)]; }
    NDFB_DEV R& re(int l, int p) const { return reinterpret_cast<R*>(buf)[2 * phys(l, p)]; }
    NDFB_DEV R& im(int l, int p) const { return reinterpret_cast<R*>(buf)[2 * phys(l, p) + 1]; }
    // packed real view: real index i lives in (slot i>>1, part i&1)
    NDFB_DEV R& part(int l, int i) const { return reinterpret_cast<R*>(buf)[2 * phys(l, i >> 1) + (i & 1)]; }
};

NDFB_DEV int pow2ceil_i(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// Run f(lane, idx) over [0,L) x [0,count): lane index fastest across threads (strided-column access) ...
template <typename R, typename F>
NDFB_DEV void for_lane_fast(const TileCtx<R>& c, int count, F f) {
    const int L = c.a.L;
    if (L <= c.T) {
        int tx = c.tid & (L - 1), ty = c.tid >> c.a.log2L, TY = c.T >> c.a.log2L;
        for (int i = ty; i < count; i += TY) f(tx, i);
    } else {
        for (int i = 0; i < count; ++i)
            for (int l = c.tid; l < L; l += c.T) f(l, i);
    }
}
// ... or element index fastest (contiguous rows).
template <typename R, typename F>
NDFB_DEV void for_elem_fast(const TileCtx<R>& c, int count, F f) {
    int TX = pow2ceil_i(count);
    if (TX > c.T) TX = c.T;
    int tx = c.tid & (TX - 1), ty = c.tid / TX, TY = c.T / TX;
    for (int l = ty; l < c.a.L; l += TY)
        for (int i = tx; i < count; i += TX) f(l, i);
}
template <typename R, typename F>
NDFB_DEV void for_tile(const TileCtx<R>& c, int count, bool lane_fast, F f) {
    if (lane_fast) for_lane_fast(c, count, f);
    else for_elem_fast(c, count, f);
}

// ------------------------------------------------------------------------------------------------------
// radix passes
// ------------------------------------------------------------------------------------------------------
template <typename R, int RADIX, bool DIT>
NDFB_DEV void radix_pass(const TileCtx<R>& c, int B, int ns, bool lane_fast) {
    // sub-transform length ns, butterfly stride s = ns / RADIX, B = whole buffer length
    const int s = ns / RADIX;
    const int nb = B / RADIX;
    const int tstep = B / ns;  // twiddle index step: W_ns^{q j} = tw[q*j*tstep]
    const Cx<R>* tw = reinterpret_cast<const Cx<R>*>(c.a.tw);
    const bool s_pow2 = (s & (s - 1)) == 0;
    int s_shift = 0;
    while ((1 << s_shift) < s) ++s_shift;
    for_tile(c, nb, lane_fast, [&](int l, int b) {
        int blk, j;
        if (s_pow2) { blk = b >> s_shift; j = b & (s - 1); }
        else { blk = b / s; j = b - blk * s; }
        const int base = blk * ns + j;
        Cx<R> v[RADIX];
#pragma unroll
        for (int q = 0; q < RADIX; ++q) v[q] = c.slot(l, base + q * s);
        if (DIT && s > 1) {
#pragma unroll
            for (int q = 1; q < RADIX; ++q) v[q] = cmul(v[q], ldg(&tw[q * j * tstep]));
        }
        Dft<R, RADIX>::run(v);
        if (!DIT && s > 1) {
#pragma unroll
            for (int q = 1; q < RADIX; ++q) v[q] = cmul(v[q], ldg(&tw[q * j * tstep]));
        }
#pragma unroll
        for (int q = 0; q < RADIX; ++q) c.slot(l, base + q * s) = v[q];
    });
}

template <typename R, bool DIT>
NDFB_DEV void radix_dispatch(const TileCtx<R>& c, int radix, int B, int ns, bool lane_fast) {
    switch (radix) {
        case 2: radix_pass<R, 2, DIT>(c, B, ns, lane_fast); break;
        case 3: radix_pass<R, 3, DIT>(c, B, ns, lane_fast); break;
        case 4: radix_pass<R, 4, DIT>(c, B, ns, lane_fast); break;
        case 5: radix_pass<R, 5, DIT>(c, B, ns, lane_fast); break;
        case 7: radix_pass<R, 7, DIT>(c, B, ns, lane_fast); break;
        case 8: radix_pass<R, 8, DIT>(c, B, ns, lane_fast); break;
        case 9: radix_pass<R, 9, DIT>(c, B, ns, lane_fast); break;
        case 11: radix_pass<R, 11, DIT>(c, B, ns, lane_fast); break;
        case 13: radix_pass<R, 13, DIT>(c, B, ns, lane_fast); break;
        case 16: radix_pass<R, 16, DIT>(c, B, ns, lane_fast); break;
        default: break;
    }
}

// forward DIF over the whole buffer of length B: natural order in, digit-reversed out
template <typename R>
NDFB_DEV void dif_passes(const TileCtx<R>& c, int B, bool lane_fast) {
    int ns = B;
    for (int p = 0; p < c.a.npass; ++p) {
        radix_dispatch<R, false>(c, c.a.radix[p], B, ns, lane_fast);
        ns /= c.a.radix[p];
        __syncthreads();
    }
}
// transpose of the above: digit-reversed in, natural order out
template <typename R>
NDFB_DEV void dit_passes(const TileCtx<R>& c, int B, bool lane_fast) {
    int ns = 1;
    for (int p = c.a.npass - 1; p >= 0; --p) {
        ns *= c.a.radix[p];
        radix_dispatch<R, true>(c, c.a.radix[p], B, ns, lane_fast);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------------
template <typename R, bool BLU>
__global__ void __launch_bounds__(512) tile_kernel(const __grid_constant__ TileArgs a) {
    NDFB_DYN_SMEM(smem_raw);
    TileCtx<R> c{a, nullptr, nullptr, nullptr, nullptr, 0, 0, 0};
    c.lb_in = reinterpret_cast<long long*>(smem_raw);
    c.lb_out = c.lb_in + a.L;
    c.lane_j2 = reinterpret_cast<int*>(c.lb_out + a.L);
    {
        size_t off = (size_t)a.L * (2 * sizeof(long long) + sizeof(int));
        off = (off + 15) & ~(size_t)15;
        c.buf = reinterpret_cast<Cx<R>*>(smem_raw + off);
    }
    c.tid = threadIdx.x;
    c.T = blockDim.x;
    const long long lane0 = (long long)blockIdx.x * a.L;
    {
        long long rem = a.nlanes - lane0;
        c.nl = rem < a.L ? (int)rem : a.L;
    }
    // per-lane base offsets
    for (int l = c.tid; l < a.L; l += c.T) {
        long long g = lane0 + l, bi = 0, bo = 0;
        int j2 = 0;
        if (l < c.nl) {
            for (int d = 0; d < a.nbd; ++d) {
                long long q = g / a.bsz[d];
                long long r = g - q * a.bsz[d];
                if (d == a.fs_dim) j2 = (int)r;
                bi += r * a.bis[d];
                bo += r * a.bos[d];
                g = q;
            }
        }
        c.lb_in[l] = bi;
        c.lb_out[l] = bo;
        c.lane_j2[l] = j2;
    }
    __syncthreads();

    const int n = a.n, N = a.N;
    const bool smem_lane_fast = (a.EP != 1);  // interleaved layout <=> lane index is the unit-stride one
    const R* in_r = reinterpret_cast<const R*>(a.in);
    const Cx<R>* in_c = reinterpret_cast<const Cx<R>*>(a.in);
    const Cx<R>* tabA = reinterpret_cast<const Cx<R>*>(a.tabA);
    const Cx<R>* tabB = reinterpret_cast<const Cx<R>*>(a.tabB);
    const R zero = (R)0;

    // ---------------- stage-in (scatter form) ----------------
    for_tile(c, a.n_in, a.in_lane_fast != 0, [&](int l, int i) {
        const bool ok = l < c.nl;
        const long long g = c.lb_in[l] + (long long)i * a.is_axis;
        switch (a.kind) {
            case TK_C2C: {
                Cx<R> x = ok ? in_c[g] : cmake<R>(zero, zero);
                if (a.conj_in) x.y = -x.y;
                c.slot(l, i) = x;
            } break;
            case TK_R2C_EVEN: c.part(l, i) = ok ? in_r[g] : zero; break;
            case TK_R2C_ODD: c.slot(l, i) = cmake<R>(ok ? in_r[g] : zero, zero); break;
            case TK_C2R_EVEN: {
                Cx<R> x = ok ? in_c[g] : cmake<R>(zero, zero);
                if (i == 0 || i == N) x.y = zero;  // src/lib.rs:516-521
                c.slot(l, i) = x;
            } break;
            case TK_C2R_ODD: {
                Cx<R> x = ok ? in_c[g] : cmake<R>(zero, zero);
                if (i == 0) x.y = zero;            // src/lib.rs:517
                c.slot(l, i) = cconj(x);           // conj(full[k])
                if (i >= 1) c.slot(l, n - i) = x;  // conj(full[n-k]) = X[k]
            } break;
            case TK_DCT1: {
                R x = ok ? in_r[g] : zero;
                c.part(l, i) = x;
                if (i > 0 && i < N) c.part(l, 2 * N - i) = x;
            } break;
            case TK_DCT2_EVEN: {
                R x = ok ? in_r[g] : zero;
                int p = (i & 1) ? (n - 1 - (i >> 1)) : (i >> 1);
                c.part(l, p) = x;
            } break;
            case TK_DCT2_ODD: {
                R x = ok ? in_r[g] : zero;
                int p = (i & 1) ? (n - 1 - (i >> 1)) : (i >> 1);
                c.slot(l, p) = cmake<R>(x, zero);
            } break;
            case TK_DCT3_EVEN: {
                R y = ok ? in_r[g] : zero;
                if (i == 0) c.slot(l, 0) = cmake<R>(y, zero);
                else if (i < N) c.re(l, i) = y;
                else if (i == N) c.slot(l, N) = cmake<R>(y, -y);
                else c.im(l, n - i) = -y;
            } break;
            case TK_DCT3_ODD: {
                R y = ok ? in_r[g] : zero;
                if (i == 0) c.slot(l, 0) = cmake<R>(y, zero);
                else { c.re(l, i) = y; c.im(l, n - i) = -y; }
            } break;
            case TK_DCT4_EVEN: {
                R x = ok ? in_r[g] : zero;
                if (i & 1) c.im(l, (n - 1 - i) >> 1) = x;
                else c.re(l, i >> 1) = x;
            } break;
            case TK_DCT4_ODD: {
                R x = ok ? in_r[g] : zero;
                Cx<R> w = ldg(&tabA[i]);
                c.slot(l, i) = cmake<R>(x * w.x, x * w.y);
                c.slot(l, n + i) = cmake<R>(zero, zero);
            } break;
            default: break;
        }
    });
    __syncthreads();

    // ---------------- prologue fix-ups (in place) ----------------
    if (a.kind == TK_C2R_EVEN || a.kind == TK_DCT3_EVEN) {
        const bool d3 = a.kind == TK_DCT3_EVEN;
        for_tile(c, N / 2 + 1, smem_lane_fast, [&](int l, int k) {
            const int k2 = N - k;
            Cx<R> xk = c.slot(l, k), xn = c.slot(l, k2);
            if (d3) {
                xk = cmul(xk, cconj(ldg(&tabB[k])));
                xn = cmul(xn, cconj(ldg(&tabB[k2])));
            }
            Cx<R> wc = cconj(ldg(&tabA[k]));   // exp(+2 pi i k / (2N))
            Cx<R> E = cadd(xk, cconj(xn)), O = csub(xk, cconj(xn));
            Cx<R> Tt = cmul_i(cmul(wc, O));
            c.slot(l, k) = cconj(cadd(E, Tt));
            if (k2 != k && k != 0) c.slot(l, k2) = csub(E, Tt);
        });
        __syncthreads();
    } else if (a.kind == TK_DCT3_ODD) {
        for_tile(c, N, smem_lane_fast, [&](int l, int k) {
            c.slot(l, k) = cconj(cmul(c.slot(l, k), cconj(ldg(&tabB[k]))));
        });
        __syncthreads();
    } else if (a.kind == TK_DCT4_EVEN) {
        for_tile(c, N, smem_lane_fast, [&](int l, int j) {
            c.slot(l, j) = cmul(c.slot(l, j), ldg(&tabA[j]));
        });
        __syncthreads();
    }

    // ---------------- core ----------------
    if (BLU) {
        const int M = a.M;
        const Cx<R>* bc = reinterpret_cast<const Cx<R>*>(a.blu_c);
        const Cx<R>* bh = reinterpret_cast<const Cx<R>*>(a.blu_bhat);
        for_tile(c, M, smem_lane_fast, [&](int l, int j) {
            c.slot(l, j) = j < N ? cmul(c.slot(l, j), ldg(&bc[j])) : cmake<R>(zero, zero);
        });
        __syncthreads();
        dif_passes(c, M, smem_lane_fast);
        for_tile(c, M, smem_lane_fast, [&](int l, int p) {
            c.slot(l, p) = cconj(cmul(c.slot(l, p), ldg(&bh[p])));
        });
        __syncthreads();
        dit_passes(c, M, smem_lane_fast);
        for_tile(c, N, smem_lane_fast, [&](int l, int k) {
            c.slot(l, k) = cmul(cconj(c.slot(l, k)), ldg(&bc[k]));
        });
        __syncthreads();
    } else {
        dif_passes(c, N, smem_lane_fast);
    }

    // ---------------- epilogue (gather form) ----------------
    const uint32_t* perm = a.perm;
    auto Y = [&](int l, int k) -> Cx<R> {
        int p = BLU ? k : (int)ldg(&perm[k]);
        return c.slot(l, p);
    };
    // bins 0..N of the length-2N real DFT packed as N complex points
    auto post = [&](int l, int k) -> Cx<R> {
        Cx<R> zk = Y(l, k == N ? 0 : k);
        Cx<R> zc = cconj(Y(l, k == 0 ? 0 : N - k));
        Cx<R> w = ldg(&tabA[k]);
        Cx<R> s = cadd(zk, zc), d = cmul(w, csub(zk, zc));
        // 0.5*(zk+zc) - 0.5*i*w*(zk-zc)
        return cmake<R>((R)0.5 * (s.x + d.y), (R)0.5 * (s.y - d.x));
    };
    const R scale = (R)a.scale;
    R* out_r = reinterpret_cast<R*>(a.out);
    Cx<R>* out_c = reinterpret_cast<Cx<R>*>(a.out);
    for_tile(c, a.n_out, a.out_lane_fast != 0, [&](int l, int k) {
        if (l >= c.nl) return;
        const long long g = c.lb_out[l] + (long long)k * a.os_axis;
        switch (a.kind) {
            case TK_C2C: {
                Cx<R> y = Y(l, k);
                if (a.conj_out) y.y = -y.y;
                y = cscale(y, scale);
                if (a.fs_twiddle) {
                    const Cx<R>* lo = reinterpret_cast<const Cx<R>*>(a.fs_lo);
                    const Cx<R>* hi = reinterpret_cast<const Cx<R>*>(a.fs_hi);
                    unsigned long long e = (unsigned long long)k * (unsigned long long)c.lane_j2[l];
                    Cx<R> w = a.fs_shift >= 40 ? ldg(&lo[e]) : cmul(ldg(&hi[e >> a.fs_shift]), ldg(&lo[e & ((1ull << a.fs_shift) - 1)]));
                    y = cmul(y, w);
                }
                if (a.os_blk) out_c[c.lb_out[l] + (long long)(k / a.os_blk) * a.os_blk_stride + (long long)(k % a.os_blk) * a.os_axis] = y;
                else out_c[g] = y;
            } break;
            case TK_R2C_EVEN: out_c[g] = cscale(post(l, k), scale); break;
            case TK_R2C_ODD: out_c[g] = cscale(Y(l, k), scale); break;
            case TK_C2R_EVEN: {
                Cx<R> y = Y(l, k >> 1);
                out_r[g] = scale * ((k & 1) ? -y.y : y.x);
            } break;
            case TK_C2R_ODD: out_r[g] = scale * Y(l, k).x; break;
            case TK_DCT1: out_r[g] = scale * (R)0.5 * post(l, k).x; break;
            case TK_DCT2_EVEN: {
                const int kk = k <= N ? k : n - k;
                Cx<R> A = cmul(post(l, kk), ldg(&tabB[kk]));
                out_r[g] = scale * (k <= N ? A.x : -A.y);
            } break;
            case TK_DCT2_ODD: out_r[g] = scale * cmul(Y(l, k), ldg(&tabB[k])).x; break;
            case TK_DCT3_EVEN: {
                const int vi = (k & 1) ? (n - 1 - (k >> 1)) : (k >> 1);
                Cx<R> y = Y(l, vi >> 1);
                out_r[g] = scale * (R)0.5 * ((vi & 1) ? -y.y : y.x);
            } break;
            case TK_DCT3_ODD: {
                const int vi = (k & 1) ? (n - 1 - (k >> 1)) : (k >> 1);
                out_r[g] = scale * (R)0.5 * Y(l, vi).x;
            } break;
            case TK_DCT4_EVEN: {
                const int j = (k & 1) ? ((n - 1 - k) >> 1) : (k >> 1);
                Cx<R> C = cmul(Y(l, j), ldg(&tabB[j]));
                out_r[g] = scale * ((k & 1) ? -C.y : C.x);
            } break;
            case TK_DCT4_ODD: out_r[g] = scale * cmul(Y(l, k), ldg(&tabB[k])).x; break;
            default: break;
        }
    });
}

}  // namespace ndfb
