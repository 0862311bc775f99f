// pipe_kernel.cuh — software-pipelined persistent C2C column kernel for tiles that own a whole SM.
//
// The register-resident Stockham kernels (sfft_kernel.cuh) hide their global loads and stores behind OTHER resident CTAs.
// A tile of L x N points that needs most of an SM's shared memory for its exchanges (4 x 4096 c64 = 139 KB: both passes of
// the 2^24-point rows of BASELINE config c5b) runs one CTA per SM, so load -> passes -> store are serialised: measured
// 47 % of the HBM copy rate per pass (profiles/round2/r2i_c5b_variants.txt).  Here the phases of CONSECUTIVE tiles overlap
// inside one persistent CTA:
//   * the next tile's input streams into a shared-memory staging buffer with cp.async (LDGSTS, 16 bytes per request, no
//     registers, no thread waits on it) while the current tile is being transformed; pass 0 then reads its points with LDS;
//   * room for that staging buffer comes from a SPLIT exchange between passes: the real parts of all points go through the
//     shared exchange buffer, then the imaginary parts through the same buffer — half the exchange footprint for two more
//     barriers per exchange (4 x 4096 c64: 128 KB staging + 68 KB exchange = 196 KB of the 227 KB);
//   * stores of the last pass are posted writes: they drain while the next tile's passes run.
// So HBM reads (next tile), FFT work (this tile) and HBM writes (previous tile) proceed together with ONE tile's registers.
// Same arithmetic, same order of operations as sfft_kernel: results are bit-identical.
// Replaces the lane loop of src/lib.rs:119-163 for rustfft lengths that take two HBM passes here (src/lib.rs:294-304).
#pragma once
#include "sfft_kernel.cuh"

namespace ndfb {

#ifdef NDFB_EMU
NDFB_DEV void cpasync16(void* sdst, const void* gsrc) { std::memcpy(sdst, gsrc, 16); }
NDFB_DEV void cpasync_commit() {}
NDFB_DEV void cpasync_wait_all() {}
#else
NDFB_DEV void cpasync16(void* sdst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
NDFB_DEV void cpasync_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
NDFB_DEV void cpasync_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
#endif

// INMODE 0: the tile's L lanes are adjacent in the input (row j of the tile = L contiguous elements): staging layout [j][l]
// INMODE 1: every lane is a contiguous row of the input (axis stride 1): staging layout [l][j], lane pitch padded so that the
//           pass-0 reads (l fastest across threads) spread over the banks
template <typename R, class S, int L, int INMODE>
struct PipeSmem {
    static constexpr int W = (int)(sizeof(Cx<R>) / 4);           // 32-bit words per element
    static constexpr int pitch1() {
        int p = S::N;
        while ((p * W) % 32 != 8 % 32 || (p * (int)sizeof(Cx<R>)) % 16 != 0) ++p;
        return p;
    }
    static constexpr int kPitch = INMODE == 0 ? S::N : pitch1();  // elements between lanes (INMODE 1)
    static constexpr size_t kStage = sizeof(Cx<R>) * (INMODE == 0 ? (size_t)L * S::N : (size_t)L * kPitch);
    static constexpr size_t kXch = sizeof(R) * (size_t)L * S::NPAD;
    static constexpr size_t kTotal = ((((kStage + 15) / 16) * 16 + kXch) + 15) / 16 * 16;   // + the lane table (kLaunch)
    static constexpr size_t kLaunch = kTotal + (size_t)L * 2 * (8 + 8 + 4) + 16;
};

// twiddles (pass >= 1) and butterflies of pass PASS on the points this thread holds
template <typename R, class S, int PASS>
NDFB_DEV void pipe_compute(int i, Cx<R> (&v)[S::E], const Cx<R>* __restrict__ tw) {
    constexpr int r = S::radix(PASS), P = S::before(PASS), G = S::G(PASS), NB = S::nbf(PASS);
    constexpr bool FULL = (NB % S::TL) == 0;
    constexpr bool KCONST = PASS > 0 && (S::TL % P == 0);
    const int k0 = i % P;
#pragma unroll
    for (int m = 0; m < G; ++m) {
        const int b = i + S::TL * m;
        if (FULL || b < NB) {
            if constexpr (PASS > 0) {
                const int k = KCONST ? k0 : b % P;
                const Cx<R>* __restrict__ twp = tw + S::twoff(PASS) + k;
#if NDFB_TW_POW
                if constexpr ((r >= 8 && (r & (r - 1)) == 0 && (S::N & (S::N - 1)) == 0) || (NDFB_TW_POW_ODD && r >= 5)) {
                    Cx<R> t[r];
#pragma unroll
                    for (int q = 1; q < r; ++q) {
                        int hb = 1;
                        while (hb * 2 <= q) hb *= 2;
                        t[q] = (hb == q) ? ldg(&twp[(q - 1) * P]) : cmul(t[hb], t[q - hb]);
                        v[m * r + q] = cmul(v[m * r + q], t[q]);
                    }
                } else
#endif
                {
#pragma unroll
                    for (int q = 1; q < r; ++q) v[m * r + q] = cmul(v[m * r + q], ldg(&twp[(q - 1) * P]));
                }
            }
            Dft<R, r>::run(&v[m * r]);
        }
    }
}

// Split exchange between pass PASS (writer, autosort positions) and pass PASS + 1 (reader): real parts, then imaginary parts,
// through ONE buffer of L x NPAD reals.  Same positions as SfftPass (fast paths included), element type R instead of Cx<R>.
template <typename R, class S, int L, int PASS>
struct PipeExchange {
    static constexpr int rw = S::radix(PASS), Pw = S::before(PASS), Gw = S::G(PASS), NBw = S::nbf(PASS);
    static constexpr int rr = S::radix(PASS + 1), Gr = S::G(PASS + 1), NBr = S::nbf(PASS + 1);
    static constexpr bool FULLw = (NBw % S::TL) == 0, FULLr = (NBr % S::TL) == 0;
    static constexpr bool FW = S::fast_write(PASS), FR = S::fast_read(PASS + 1);
    static constexpr bool KCONST = PASS > 0 && (S::TL % Pw == 0);
    template <class Ctx>
    static NDFB_DEV void run(const Ctx& c, R* __restrict__ sm, Cx<R> (&v)[S::E]) {
        const int k0 = c.i % Pw;
        const int wbase = !FW ? 0 : (PASS == 0 ? c.slot_of(c.i * S::R0P) : c.slot_of(S::pad((c.i - k0) * rw + k0)));
        const int rbase = FR ? c.slot_of(S::pad(c.i)) : 0;
        auto waddr = [&](int m, int q) -> int {
            if (FW) return wbase + (PASS == 0 ? (S::TL * m * S::R0P + q) : S::pad(S::TL * m * rw + q * Pw)) * c.kscale;
            const int b = c.i + S::TL * m;
            const int k = KCONST ? k0 : b % Pw;
            return c.addr((b - k) * rw + k + q * Pw);
        };
        auto raddr = [&](int m, int q) -> int {
            if (FR) return rbase + S::pad(S::TL * m + q * NBr) * c.kscale;
            return c.addr(c.i + S::TL * m + q * NBr);
        };
        R t[S::E];
#pragma unroll
        for (int m = 0; m < Gw; ++m)
            if (FULLw || c.i + S::TL * m < NBw)
#pragma unroll
                for (int q = 0; q < rw; ++q) sm[waddr(m, q)] = v[m * rw + q].x;
        __syncthreads();
#pragma unroll
        for (int m = 0; m < Gr; ++m)
            if (FULLr || c.i + S::TL * m < NBr)
#pragma unroll
                for (int q = 0; q < rr; ++q) t[m * rr + q] = sm[raddr(m, q)];
        __syncthreads();
#pragma unroll
        for (int m = 0; m < Gw; ++m)
            if (FULLw || c.i + S::TL * m < NBw)
#pragma unroll
                for (int q = 0; q < rw; ++q) sm[waddr(m, q)] = v[m * rw + q].y;
        __syncthreads();
#pragma unroll
        for (int m = 0; m < Gr; ++m)
            if (FULLr || c.i + S::TL * m < NBr)
#pragma unroll
                for (int q = 0; q < rr; ++q) v[m * rr + q] = cmake<R>(t[m * rr + q], sm[raddr(m, q)]);
        // the buffer is free for the next exchange (after the last one, the two barriers at the top of the tile loop do that)
        if (PASS + 2 < S::NP) __syncthreads();
    }
};

template <typename R, class S, int L, int PASS, class Ctx, class StoreF>
NDFB_DEV void pipe_passes(const Ctx& c, R* sm, Cx<R> (&v)[S::E], const Cx<R>* __restrict__ tw, StoreF& store) {
    pipe_compute<R, S, PASS>(c.i, v, tw);
    if constexpr (PASS + 1 < S::NP) {
        PipeExchange<R, S, L, PASS>::run(c, sm, v);
        pipe_passes<R, S, L, PASS + 1>(c, sm, v, tw, store);
    } else {
        constexpr int r = S::radix(PASS), G = S::G(PASS), NB = S::nbf(PASS);
#pragma unroll
        for (int m = 0; m < G; ++m) {
            const int b = c.i + S::TL * m;
            if ((NB % S::TL) == 0 || b < NB) {
                auto cur = store.start(b, NB, r);
#pragma unroll
                for (int q = 0; q < r; ++q) store.next(cur, v[m * r + q]);
            }
        }
    }
}

// last-pass store: scale / conjugation, strided output, and the four-step twiddle W_N^{k j2} of output k = b + q NB of lane j2,
// factored as W_N^{b j2} (ONE lookup per butterfly: the hi/lo product, or the single table when shift >= 40) times
// W_N^{q NB j2} = fsq[j2 r + q] (a table of r entries per lane: every thread of a lane reads the same r values, L1 broadcasts).
// sfft_body's MODE 1 / 2 look every W_N^{k j2} up on its own: 2 scattered table loads per POINT, 11.5 GB of L1 sector traffic
// for a 2.1 GB pass of the 2^24-point rows of c5b (profiles/round2/r2q_ncu_c5b_nopipe_summary.txt).
template <typename R>
struct PipeStore {
    Cx<R>* out; long long os_axis; R sc, sy;
    int fs_twiddle, fs_shift; const Cx<R>* lo; const Cx<R>* hi; const Cx<R>* fsq; unsigned j2;
    struct Cur { Cx<R>* p; long long step; Cx<R> w; const Cx<R>* t; };
    NDFB_DEV Cur start(int b, int nb, int r) const {
        Cur u; u.p = out + (long long)b * os_axis; u.step = (long long)nb * os_axis;
        u.w = cmake<R>((R)1, (R)0); u.t = nullptr;
        if (fs_twiddle) {
            const unsigned long long e = (unsigned long long)b * j2;
            u.w = fs_shift >= 40 ? ldg(&lo[(unsigned)e]) : cmul(ldg(&hi[e >> fs_shift]), ldg(&lo[e & ((1ull << fs_shift) - 1)]));
            u.t = fsq + (size_t)j2 * r;
        }
        return u;
    }
    NDFB_DEV void next(Cur& u, Cx<R> val) const {
        Cx<R> y = cmake<R>(val.x * sc, val.y * sy);
        if (fs_twiddle) { y = cmul(y, cmul(u.w, ldg(u.t))); ++u.t; }
        *u.p = y;
        u.p += u.step;
    }
};

template <typename R, class S, int L, int INMODE, int MINB>
__global__ void __launch_bounds__(S::TL* L, MINB) sfft_pipe_kernel(const __grid_constant__ SfftArgs a) {
    static_assert(S::NP >= 2, "single-pass schedules have no exchange to split");
    using SM = PipeSmem<R, S, L, INMODE>;
    constexpr int T = S::TL * L;
    constexpr int N = S::N;
    NDFB_DYN_SMEM(smem_raw);
    Cx<R>* __restrict__ stage = reinterpret_cast<Cx<R>*>(smem_raw);
    R* __restrict__ xch = reinterpret_cast<R*>(smem_raw + ((SM::kStage + 15) / 16) * 16);
    SfftCtx<R, S, L, true> c;
    c.smem = nullptr;
    const int tid = threadIdx.x;
    c.l = tid % L; c.i = tid / L; c.valid = true;
    const Cx<R>* __restrict__ tw = reinterpret_cast<const Cx<R>*>(a.tw);
    const R sc = (R)a.scale;
    const R sgn_in = a.conj_in ? (R)-1 : (R)1;
    const long long ntiles = a.ntiles;

    // lane -> array offsets cost ~25 instructions per batch dim (runtime divisions): computed ONCE per tile and lane by the
    // first L threads, one tile ahead, and shared through a small double-buffered table behind the exchange buffer
    struct LaneTab { long long bi[2][L], bo[2][L]; int j2[2][L]; };
    LaneTab& lt = *reinterpret_cast<LaneTab*>(smem_raw + SM::kTotal);
    auto fill_lanes = [&](long long tile, int slot) {   // threads < L
        const LaneBase lb = lane_base(a, tile * L + tid, true, a.fs_dim);
        lt.bi[slot][tid] = lb.bi; lt.bo[slot][tid] = lb.bo; lt.j2[slot][tid] = lb.j2;
    };
    // all threads: queue the 16-byte pieces of the tile whose lane offsets are in `slot` (every lane of a launched tile exists:
    // the host checks nlanes % L == 0)
    auto issue = [&](int slot) {
        if constexpr (INMODE == 0) {
            constexpr int CPR = (int)(L * sizeof(Cx<R>) / 16);          // pieces per tile row
            constexpr int TOTAL = N * CPR;
            const char* src0 = reinterpret_cast<const char*>(reinterpret_cast<const Cx<R>*>(a.in) + lt.bi[slot][0]);
            const long long row_bytes = a.is_axis * (long long)sizeof(Cx<R>);
            char* dst0 = reinterpret_cast<char*>(stage);
#pragma unroll
            for (int p = tid; p < TOTAL; p += T) {
                const int j = p / CPR, part = p % CPR;
                cpasync16(dst0 + (size_t)p * 16, src0 + (long long)j * row_bytes + part * 16);
            }
        } else {
            constexpr int CPL = (int)(N * sizeof(Cx<R>) / 16);          // pieces per lane
#pragma unroll
            for (int l = 0; l < L; ++l) {
                const char* src = reinterpret_cast<const char*>(reinterpret_cast<const Cx<R>*>(a.in) + lt.bi[slot][l]);
                char* dst = reinterpret_cast<char*>(stage + (size_t)l * SM::kPitch);
                for (int p = tid; p < CPL; p += T) cpasync16(dst + (size_t)p * 16, src + (size_t)p * 16);
            }
        }
        cpasync_commit();
    };

    long long tile = blockIdx.x;
    int slot = 0;
    if (tile < ntiles && tid < L) fill_lanes(tile, 0);
    __syncthreads();
    if (tile < ntiles) issue(0);
    for (; tile < ntiles; tile += gridDim.x, slot ^= 1) {
        const bool more = tile + gridDim.x < ntiles;
        if (more && tid < L) fill_lanes(tile + gridDim.x, slot ^ 1);
        cpasync_wait_all();
        __syncthreads();                         // this tile's input has landed and is visible to every thread
        Cx<R> v[S::E];
        {
            constexpr int r = S::R0, G = S::G(0), NB = S::nbf(0);
#pragma unroll
            for (int m = 0; m < G; ++m) {
                const int b = c.i + S::TL * m;
                if ((NB % S::TL) == 0 || b < NB) {
#pragma unroll
                    for (int q = 0; q < r; ++q) {
                        const int j = b + q * NB;
                        Cx<R> x = INMODE == 0 ? stage[(size_t)j * L + c.l] : stage[(size_t)c.l * SM::kPitch + j];
                        x.y *= sgn_in;
                        v[m * r + q] = x;
                    }
                }
            }
        }
        __syncthreads();                         // every thread has its points: the staging buffer is free again
        if (more) issue(slot ^ 1);
        PipeStore<R> st;
        st.out = reinterpret_cast<Cx<R>*>(a.out) + lt.bo[slot][c.l]; st.os_axis = a.os_axis; st.sc = sc; st.sy = a.conj_out ? -sc : sc;
        st.fs_twiddle = a.fs_twiddle; st.fs_shift = a.fs_shift;
        st.lo = reinterpret_cast<const Cx<R>*>(a.fs_lo); st.hi = reinterpret_cast<const Cx<R>*>(a.fs_hi);
        st.fsq = reinterpret_cast<const Cx<R>*>(a.fs_q); st.j2 = (unsigned)lt.j2[slot][c.l];
        pipe_passes<R, S, L, 0>(c, xch, v, tw, st);
    }
}

}  // namespace ndfb
