// sfft_kernel.cuh — the fast path: compile-time-scheduled Stockham FFT, register resident between passes.
//
// For lengths with an instantiated schedule (see sfft_registry.h) one CTA transforms L lanes:
//   * every thread owns E points of one lane, X[i + TL*e] (TL = N/E threads per lane), so the FIRST radix pass reads
//     straight from global memory and the LAST pass writes straight to global memory, both coalesced
//     (along the axis for contiguous rows, across L adjacent lanes for strided columns);
//   * between passes the points are exchanged through shared memory in Stockham autosort order
//     (write Y[(b-k) r + k + q P], read X[i + TL e]); with P passes that is P-1 exchanges, each
//     bank-conflict free thanks to one pad element per r0 points;
//   * twiddles come from per-pass tables laid out k-fastest so a warp's loads are coalesced (L1/L2 resident);
//   * inverse transforms are conj-in / conj-out of the forward schedule; 1/n (Normalization::Default,
//     src/lib.rs:333-338) and the four-step inter-pass twiddle are fused into the store.
// Replaces fft_lane / ifft_lane (src/lib.rs:313-331) plus the lane loop and copies of src/lib.rs:119-163.
#pragma once
#include "butterflies.cuh"
#include "common.h"

#ifndef NDFB_TW_POW
#define NDFB_TW_POW 1   // +2..10 % on B200 (profiles/r1r_ab_twiddle_powers.jsonl: A = table loads, B = powers)
#endif
#ifndef NDFB_TW_POW_ODD
#define NDFB_TW_POW_ODD 1   // the same for the mixed radices 5..15 (loads W^k, W^2k, W^4k, W^8k; products of depth <= 3): 384-point c128 rows
                            // +7 %, 360-point columns +2 %, the 4095-point DCT-I core +2-3 %, neutral elsewhere, same accuracy
                            // (profiles/round2/r2r_ab_tw_pow_odd.jsonl: A = table loads, B = powers)
#endif

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
    const void* tw;  // per-pass twiddle tables in Sched::twoff layout
    int fs_twiddle, fs_dim, fs_shift;
    int os_blk;            // != 0: output axis index k is split as (k / os_blk, k % os_blk) ...
    long long os_blk_stride;  // ... with this stride for the block index (packed all-to-all send layout)
    int nblk_ptr;          // != 0: block p is written relative to blk_ptr[p] (peer-mapped receive buffers: the store IS the all-to-all)
    void* blk_ptr[8];
    // Producer / consumer overlap between two launches on different streams (dist.SlabR2cFft3d: the r2c pass feeds the
    // exchange pass plane by plane).  Consumer side: before tile t is loaded, wait until wait_cnt[first lane / wait_group]
    // has reached wait_need (the producer adds the lanes it has finished); the launch is persistent (tiles strided by gridDim).
    const unsigned* wait_cnt;
    long long wait_group;
    unsigned wait_need;
    long long ntiles;      // persistent launches: total number of tiles (0 = one tile per CTA)
    long long l2_prefetch_lanes;   // != 0 (contiguous rows): lane distance of the row this CTA asks the L2 to fetch ahead (one
                                   // cp.async.bulk.prefetch.L2 per row: the CTA that later owns it finds it at L2 latency)
    int trans_store;       // != 0 (row kernels): the tile's lanes are far apart in the input but ADJACENT in the output (last pass of a
                           // three-pass split): rows are read along the axis, the last pass re-maps threads so that stores run across lanes
    int bulk_store;        // != 0 (with nblk_ptr): the tile's block for each destination is one contiguous range there: stage the result
                           // in shared memory and send each block with ONE bulk-async copy (cp.async.bulk, the TMA engine)
    const void* fs_lo;
    const void* fs_hi;
    const void* fs_q;      // column kernels, four-step twiddle factored (SfftGStore MODE 4): W_N^{k l} at [k L + l], k < N1, l < L;
                           // pipelined kernels (pipe_kernel.cuh): W_N^{q NB j2} at [j2 r + q], r = last radix, NB = N1 / r
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
    // per-pass twiddle tables, concatenated: pass p >= 1 holds W_{P r}^{q k} at twoff(p) + (q-1) P + k  (k < P fastest,
    // so the 32 lanes of a warp, which hold consecutive butterflies b and therefore consecutive k = b mod P, read
    // consecutive addresses)
    static constexpr int twsize(int p) { return (p >= 1 && p < NP) ? (radix(p) - 1) * before(p) : 0; }
    static constexpr int twoff(int p) { return p <= 1 ? 0 : twoff(p - 1) + twsize(p - 1); }
    static constexpr int TWTOTAL = twsize(1) + twsize(2) + twsize(3);
    static constexpr int E = cmax(cmax(EP(0), EP(1)), cmax(EP(2), EP(3)));      // register slots per thread
    // one pad element per R0 points keeps the stride-R0 writes of pass 0 off a single bank.  An ODD first radix needs none: its
    // stride is already odd, and the pad made it even (13.9.7.5, the 4095-point DCT-I core of c4: 1.60 wavefronts per ideal one
    // with the pad, 1.01 without; 9.9.9: 1.77 -> 1.00; bank model of every exchange in tools/bank_model.py)
    static constexpr bool PADDED = (R0_ % 2) == 0;
    static constexpr int pad(int a) { return PADDED ? a + a / R0_ : a; }
    static constexpr int NPAD = PADDED ? N_ + N_ / R0_ : N_;
    static constexpr int R0P = PADDED ? R0_ + 1 : R0_;   // padded distance of consecutive pass-0 butterflies
    // Address fast paths: when the per-thread part and the compile-time part of an index are each multiples of R0
    // where it matters, pad(thread + const) = pad(thread) + pad(const) and every shared-memory access becomes
    // "one precomputed register + immediate offset" (without padding the positions are additive as they are).
    static constexpr bool fast_read(int p) { return p >= 1 && p < NP && (!PADDED || (TL_ % R0_ == 0 && nbf(p) % R0_ == 0)); }
    static constexpr bool fast_write(int p) {
        return p == 0 ? true : (p < NP - 1 && TL_ % before(p) == 0 && (!PADDED || before(p) % R0_ == 0));
    }
};

// PX: extra elements in the lane pitch of the row layout (transposing rows kernel: lanes become the fastest thread index in
// the last pass, so the pitch must not be a multiple of the bank count)
template <typename R, class S, int L, bool COLS, int PX = 0>
struct SfftCtx {
    Cx<R>* smem;
    int i, l;  // position within the lane group, lane within the tile
    bool valid;
    static constexpr int kPitch = S::NPAD + PX;
    NDFB_DEV int addr(int a) const {
        const int p = S::pad(a);
        return COLS ? p * L + l : l * kPitch + p;
    }
    // slot of padded position `pp` (already padded)
    NDFB_DEV int slot_of(int pp) const { return COLS ? pp * L + l : l * kPitch + pp; }
    static constexpr int kscale = COLS ? L : 1;   // slot distance of one padded position
};

// One Stockham pass.  Butterfly b of this pass (b < N/r) reads X[b + q N/r], multiplies by W_{P r}^{q k}
// (k = b mod P, P = product of the earlier radices) and writes Y[(b-k) r + k + q P].
// Pass 0 takes its inputs from `load(j)` (global memory, coalesced in b), the last pass hands its outputs to
// `store(k, value)`.  SYNC0 / SYNCL add a barrier after the loads of the first / last pass for callers whose
// load or store functor itself goes through the shared buffer (staged rows, pair epilogues).
// Load / store functors may offer a CURSOR interface (member kStrided): the r inputs / outputs of one butterfly are
// NB points apart, so the functor walks a pointer by a precomputed step instead of forming j * stride (a 64-bit multiply
// plus scaling per access: a third of the instructions of the strided-column kernels before; profiles/r1v opmix).
template <class F, class = void> struct sfft_is_strided { static constexpr bool value = false; };
template <class F> struct sfft_is_strided<F, decltype((void)F::kStrided)> { static constexpr bool value = true; };

// strided complex array in global memory (sfft_body, MODE 0 / 1)
template <typename R, bool CG>
struct SfftGLoad {
    static constexpr bool kStrided = true;
    const Cx<R>* in; long long is_axis; R sgn; bool valid;
    struct Cur { const Cx<R>* p; long long step; };
    NDFB_DEV Cur start(int b, int nb) const { Cur u; u.p = in + (long long)b * is_axis; u.step = (long long)nb * is_axis; return u; }
    NDFB_DEV Cx<R> next(Cur& u) const {
        // CG: the array was written by other CTAs of this same launch (fs2_kernel): read it from L2, never from L1
        Cx<R> x = valid ? (CG ? ld_cg(u.p) : ld_stream(u.p)) : cmake<R>((R)0, (R)0);
        u.p += u.step;
        x.y *= sgn;
        return x;
    }
};
// MODE 4: the four-step twiddle W_N^{k j2} of lane j2 = j20 + l (tile base + lane within the tile) factored as
// W_N^{k j20} (the same for every lane of the tile: N1 values per tile, looked up ONCE by the CTA into shared memory behind the
// exchange buffer and read back as broadcasts) x W_N^{k l} (a plan-owned table [k][l] of N1 x L entries, read coalesced across the
// lanes, L1-resident).  MODE 1 / 2 look W_N^{k j2} up per point with a per-lane index:
// 32 different table lines per warp request (the column passes of the 2^24-point rows of c5b: 4.4-4.6 ms per pass against 2.7 ms
// for the pass without a twiddle; profiles/round2/r2u_c5b_launches_*.csv).
template <typename R, int MODE>
struct SfftGStore {
    static constexpr bool kStrided = true;
    Cx<R>* out; long long os_axis; R sc, sy; bool valid; const Cx<R>* fs; unsigned j2;
    const Cx<R>* tq; const Cx<R>* wk; int lane, tl;   // MODE 4: [k][l] table, the tile's W_N^{k j20} in shared memory
    struct Cur { Cx<R>* p; long long step; const Cx<R>* t; unsigned tstep; const Cx<R>* w; };
    NDFB_DEV Cur start(int b, int nb) const {
        Cur u; u.p = out + (long long)b * os_axis; u.step = (long long)nb * os_axis;
        if (MODE == 4) {
            u.t = tq + (unsigned)b * (unsigned)tl + (unsigned)lane; u.tstep = (unsigned)nb * (unsigned)tl;
            u.w = wk + b;
        } else {
            u.t = fs + (unsigned)b * j2; u.tstep = (unsigned)nb * j2;
            u.w = nullptr;
        }
        return u;
    }
    NDFB_DEV void next(Cur& u, Cx<R> val) const {
        Cx<R> y = cmake<R>(val.x * sc, val.y * sy);
        if (MODE == 1) { y = cmul(y, ldg(u.t)); u.t += u.tstep; }   // four-step twiddle W_N^{k j2}, k = b + q nb
        if (MODE == 4) {
            y = cmul(y, cmul(*u.w, ldg(u.t)));
            u.t += u.tstep; u.w += u.tstep / (unsigned)tl;
        }
        if (valid) *u.p = y;
        u.p += u.step;
    }
};

template <typename R, class S, int L, bool COLS, int PASS, bool SYNC0, bool SYNCL>
struct SfftPass {
    static constexpr int r = S::radix(PASS);
    static constexpr int P = S::before(PASS);
    static constexpr int G = S::G(PASS);
    static constexpr int NB = S::nbf(PASS);
    static constexpr bool FIRST = PASS == 0;
    static constexpr bool LAST = PASS == S::NP - 1;
    static constexpr bool FULL = (NB % S::TL) == 0;  // no idle threads in this pass

    static constexpr bool FR = S::fast_read(PASS);
    static constexpr bool FW = !LAST && S::fast_write(PASS);
    static constexpr bool KCONST = !FIRST && (S::TL % P == 0);   // k = b mod P does not depend on m

    template <class Ctx, typename LoadF, typename StoreF>
    static NDFB_DEV void run(const Ctx& c, Cx<R> (&v)[S::E], const Cx<R>* __restrict__ tw,
                             LoadF load, StoreF store) {
        // ---- gather this thread's butterfly inputs ----
        const int rbase = FR ? c.slot_of(S::pad(c.i)) : 0;
#pragma unroll
        for (int m = 0; m < G; ++m) {
            const int b = c.i + S::TL * m;
            if (FULL || b < NB) {
                if constexpr (FIRST && sfft_is_strided<LoadF>::value) {
                    auto cur = load.start(b, NB);
#pragma unroll
                    for (int q = 0; q < r; ++q) v[m * r + q] = load.next(cur);
                } else {
#pragma unroll
                    for (int q = 0; q < r; ++q) {
                        if constexpr (FIRST) v[m * r + q] = load(b + q * NB);
                        else if (FR) v[m * r + q] = c.smem[rbase + S::pad(S::TL * m + q * NB) * c.kscale];
                        else v[m * r + q] = c.smem[c.addr(b + q * NB)];
                    }
                }
            }
        }
        // every thread has read the previous layout before it is overwritten
        if ((!FIRST && !LAST) || (FIRST && SYNC0 && !LAST) || (LAST && SYNCL)) __syncthreads();
        const int k0 = c.i % P;
        const Cx<R>* __restrict__ twp0 = tw + S::twoff(PASS) + k0;
        // write base for the fast paths:  pass 0: b R0P (= R0 + 1 when padded);  later passes: pad((i - k) r + k)
        const int wbase = !FW ? 0 : (FIRST ? c.slot_of(c.i * S::R0P) : c.slot_of(S::pad((c.i - k0) * r + k0)));
#pragma unroll
        for (int m = 0; m < G; ++m) {
            const int b = c.i + S::TL * m;
            if (FULL || b < NB) {
                const int k = KCONST ? k0 : b % P;
                if (!FIRST) {
                    const Cx<R>* __restrict__ twp = KCONST ? twp0 : tw + S::twoff(PASS) + k;
#if NDFB_TW_POW
                    if constexpr ((r >= 8 && (r & (r - 1)) == 0 && (S::N & (S::N - 1)) == 0) || (NDFB_TW_POW_ODD && r >= 5)) {
                        // load W^k, W^2k, W^4k, W^8k and form the other powers as products (depth <= 3): 4 table loads, not 15
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
                if constexpr (LAST && sfft_is_strided<StoreF>::value) {
                    auto cur = store.start(b, NB);
#pragma unroll
                    for (int q = 0; q < r; ++q) store.next(cur, v[m * r + q]);
                } else if constexpr (LAST) {
#pragma unroll
                    for (int q = 0; q < r; ++q) store(b + q * NB, v[m * r + q]);   // (b-k) r + k + q P with P = N/r, k = b
                } else if (FW) {
#pragma unroll
                    for (int q = 0; q < r; ++q) {
                        // pass 0: pad(b r + q) = b R0P + q;  later: pad(thread part) + pad(TL m r + q P)
                        const int cpart = FIRST ? (S::TL * m * S::R0P + q) : S::pad(S::TL * m * r + q * P);
                        c.smem[wbase + cpart * c.kscale] = v[m * r + q];
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < r; ++q) c.smem[c.addr((b - k) * r + k + q * P)] = v[m * r + q];
                }
            }
        }
        if (!LAST) __syncthreads();
    }
};

template <typename R, class S, int L, bool COLS, int PASS, bool SYNC0, bool SYNCL>
struct SfftAll {
    template <class Ctx, typename LoadF, typename StoreF>
    static NDFB_DEV void run(const Ctx& c, Cx<R> (&v)[S::E], const Cx<R>* tw, LoadF load, StoreF store) {
        SfftPass<R, S, L, COLS, PASS, SYNC0, SYNCL>::run(c, v, tw, load, store);
        if constexpr (PASS + 1 < S::NP) SfftAll<R, S, L, COLS, PASS + 1, SYNC0, SYNCL>::run(c, v, tw, load, store);
    }
};

// passes PASS .. STOP-1 only (none of them the last one): callers that give the last pass a mapping of their own
template <typename R, class S, int L, bool COLS, int PASS, int STOP, bool SYNC0>
struct SfftUntil {
    template <class Ctx, typename LoadF, typename StoreF>
    static NDFB_DEV void run(const Ctx& c, Cx<R> (&v)[S::E], const Cx<R>* tw, LoadF load, StoreF store) {
        if constexpr (PASS < STOP) {
            SfftPass<R, S, L, COLS, PASS, SYNC0, false>::run(c, v, tw, load, store);
            SfftUntil<R, S, L, COLS, PASS + 1, STOP, SYNC0>::run(c, v, tw, load, store);
        }
    }
};

// Mirror-paired last pass for the real kinds whose outputs need Z[k] AND Z[N-k] (R2C, DCT-I, DCT-II): when the last pass has
// two butterflies per thread, thread i takes butterflies i and NB - i (thread 0: 0 and NB/2), whose outputs are exactly each
// other's mirror bins (k = b + q NB  <->  N - k = (NB - b) + (r-1-q) NB).  The pair epilogue then runs from registers: no
// write of the spectrum to shared memory, no barrier, no read-back — one shared-memory round trip less per lane.
// (any EVEN number of butterflies per thread works the same way: G/2 mirror pairs per thread)
template <class S>
constexpr bool kMirrorEpi = S::NP >= 2 && S::G(S::NP - 1) % 2 == 0 && S::nbf(S::NP - 1) == S::G(S::NP - 1) * S::TL &&
                            S::nbf(S::NP - 1) % 2 == 0 && S::radix(S::NP - 1) % 2 == 0;
// One butterfly per thread in the last pass: the mirror butterfly lives in ANOTHER thread.  For contiguous rows the butterflies
// are dealt so that lanes t and t ^ 16 of a warp hold butterflies b and NB - b, and the partner's outputs arrive by
// __shfl_xor_sync: the pair epilogue again runs from registers, with warp shuffles instead of a shared-memory round trip.
// Measured on B200 (profiles/round2/r2n_ab_pair_epilogue.jsonl): the shuffle form LOSES 1.3-1.75x against the shared-memory
// epilogue (32 SHFL.32 per thread for eight c128 values move a quarter of what eight LDS.128 move per instruction, and both
// partners repeat the pair arithmetic), so it is compiled only with -DNDFB_SHUFFLE_EPI=1; the in-thread mirror pairs gain 2-12 %.
#ifndef NDFB_SHUFFLE_EPI
#define NDFB_SHUFFLE_EPI 0
#endif
template <class S>
constexpr bool kShuffleEpi = NDFB_SHUFFLE_EPI && S::NP >= 2 && S::G(S::NP - 1) == 1 && S::nbf(S::NP - 1) == S::TL && S::TL % 32 == 0 &&
                             S::radix(S::NP - 1) % 2 == 0;

// Mirror-paired last pass for the kinds whose OUTPUT rows interleave bins k and N-1-k (DCT-III: inverse Makhoul order, DCT-IV:
// out[2k] / out[n-1-2k]): thread i takes butterflies p and NB-1-p, whose outputs k = p + q NB and N-1-k = (NB-1-p) + (r-1-q) NB are
// each other's mirror, and writes whole aligned 16 / 32-byte pieces of the contiguous output row straight from registers: the
// staging copy through shared memory (one write + one read of the lane, a barrier) disappears.  No self-paired butterflies.
template <class S>
constexpr bool kMirrorOut = S::NP >= 2 && S::G(S::NP - 1) % 2 == 0 && S::nbf(S::NP - 1) == S::G(S::NP - 1) * S::TL &&
                            S::nbf(S::NP - 1) % 2 == 0 && S::radix(S::NP - 1) % 2 == 0;

// The same pairing on the FIRST pass for the kinds whose inputs come in mirror pairs (C2R, DCT-III: slots j and N-j are built
// from the same two spectrum bins): schedules that START with the small radix (4.8.8.8) give pass 0 two butterflies per thread.
template <class S>
constexpr bool kMirrorPro = S::NP >= 2 && S::G(0) == 2 && S::nbf(0) == 2 * S::TL && S::nbf(0) % 2 == 0 && S::R0 % 2 == 0;

// lane -> global base offsets (elements of the respective array); also the index along the fastest batch dim
struct LaneBase {
    long long bi, bo;
    int j2;
};
template <typename A>
NDFB_DEV LaneBase lane_base(const A& a, long long g, bool valid, int fs_dim = 0) {
    LaneBase o;
    o.bi = 0; o.bo = 0; o.j2 = 0;
    if (valid) {
        if (a.nlanes <= 0x7fffffffLL) {   // 32-bit index arithmetic (a 64-bit division costs ~80 instructions)
            unsigned g32 = (unsigned)g;
#pragma unroll
            for (int d = 0; d < kMaxBatchDims; ++d) {
                if (d < a.nbd) {
                    unsigned rr = g32;                 // the slowest dim takes what is left: no division (g < nlanes)
                    if (d + 1 < a.nbd) {
                        const unsigned sz = (unsigned)a.bsz[d];
                        const unsigned q = g32 / sz;
                        rr = g32 - q * sz;
                        g32 = q;
                    }
                    if (d == fs_dim) o.j2 = (int)rr;
                    o.bi += (long long)rr * a.bis[d];
                    o.bo += (long long)rr * a.bos[d];
                }
            }
        } else {
#pragma unroll
            for (int d = 0; d < kMaxBatchDims; ++d) {
                if (d < a.nbd) {
                    const long long q = g / a.bsz[d];
                    const long long rr = g - q * a.bsz[d];
                    if (d == fs_dim) o.j2 = (int)rr;
                    o.bi += rr * a.bis[d];
                    o.bo += rr * a.bos[d];
                    g = q;
                }
            }
        }
    }
    return o;
}

// ------------------------------------------------------------------------------------------------------
// bulk-async (TMA) stores: shared memory -> global / peer memory, one instruction per contiguous block
// ------------------------------------------------------------------------------------------------------
#ifdef NDFB_EMU
NDFB_DEV void bulk_store_fence() {}
NDFB_DEV void bulk_store_issue(void* gdst, const void* ssrc, unsigned bytes) {
    unsigned char* d = reinterpret_cast<unsigned char*>(gdst);
    const unsigned char* q = reinterpret_cast<const unsigned char*>(ssrc);
    for (unsigned i = 0; i < bytes; ++i) d[i] = q[i];
}
NDFB_DEV void bulk_store_commit_and_drain() {}
#else
// generic-proxy writes (the STS of the last pass) must be visible to the async proxy before it reads shared memory
NDFB_DEV void bulk_store_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
NDFB_DEV void bulk_store_issue(void* gdst, const void* ssrc, unsigned bytes) {
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(saddr), "r"(bytes) : "memory");
}
// the CTA may not exit (or reuse the buffer) before the engine has READ the staged tile
NDFB_DEV void bulk_store_commit_and_drain() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
#endif

// ------------------------------------------------------------------------------------------------------
// complex-to-complex
// ------------------------------------------------------------------------------------------------------
// MODE 0: plain store;  1: four-step twiddle from the single table W_N^e (N <= 2^17), nothing else;  2: everything
// (hi/lo twiddle product, split / scattered output blocks).  UNIT: both axis strides are 1 (contiguous rows)
template <typename R, class S, int L, bool COLS, int MODE, bool UNIT, bool CG = false>
NDFB_DEV void sfft_body(const SfftArgs& a, const SfftCtx<R, S, L, COLS>& c, const LaneBase& lb, long long tile = 0) {
    const Cx<R>* __restrict__ in = reinterpret_cast<const Cx<R>*>(a.in) + lb.bi;
    Cx<R>* __restrict__ out = reinterpret_cast<Cx<R>*>(a.out) + lb.bo;
    const long long is_axis = UNIT ? 1 : a.is_axis, os_axis = UNIT ? 1 : a.os_axis;
    const Cx<R>* __restrict__ tw = reinterpret_cast<const Cx<R>*>(a.tw);
    const R sc = (R)a.scale;
    const R sy = a.conj_out ? -sc : sc;
    const R sgn_in = a.conj_in ? (R)-1 : (R)1;
    const bool valid = c.valid;
    const int j2 = lb.j2;
    Cx<R> v[S::E];
    if constexpr (MODE == 0 || MODE == 1 || MODE == 4) {
        SfftGLoad<R, CG> gl; gl.in = in; gl.is_axis = is_axis; gl.sgn = sgn_in; gl.valid = valid;
        SfftGStore<R, MODE> gs; gs.out = out; gs.os_axis = os_axis; gs.sc = sc; gs.sy = sy; gs.valid = valid;
        gs.fs = reinterpret_cast<const Cx<R>*>(a.fs_lo); gs.j2 = (unsigned)j2;
        gs.tq = reinterpret_cast<const Cx<R>*>(a.fs_q); gs.lane = c.l; gs.tl = L; gs.wk = nullptr;
        if constexpr (MODE == 4) {
            // the tile's own factor W_N^{k j20}, k < N: one hi / lo lookup per k and TILE (not per point), kept behind the exchange
            // buffer (the host adds N elements to the launch's shared memory); the barriers of the exchanges order it before the
            // last pass reads it
            Cx<R>* wk = c.smem + (S::NP > 1 ? (size_t)L * S::NPAD : 0);
            const Cx<R>* __restrict__ lo = reinterpret_cast<const Cx<R>*>(a.fs_lo);
            const Cx<R>* __restrict__ hi = reinterpret_cast<const Cx<R>*>(a.fs_hi);
            const unsigned j20 = (unsigned)(j2 - c.l);     // first lane of the tile: the same value in every thread
            for (int k = threadIdx.x; k < S::N; k += S::TL * L) {
                const unsigned long long e = (unsigned long long)k * j20;
                wk[k] = a.fs_shift >= 40 ? ldg(&lo[(unsigned)e]) : cmul(ldg(&hi[e >> a.fs_shift]), ldg(&lo[e & ((1ull << a.fs_shift) - 1)]));
            }
            if (S::NP == 1) __syncthreads();
            gs.wk = wk;
        }
        SfftAll<R, S, L, COLS, 0, false, false>::run(c, v, tw, gl, gs);
        return;
    }
    auto load = [&](int j) -> Cx<R> {
        // CG: the array was written by other CTAs of this same launch (fs2_kernel): read it from L2, never from L1
        Cx<R> x = valid ? (CG ? ld_cg(&in[(long long)j * is_axis]) : ld_stream(&in[(long long)j * is_axis])) : cmake<R>((R)0, (R)0);
        x.y *= sgn_in;
        return x;
    };
    if constexpr (MODE == 3) {
        // Scattered output blocks whose per-destination part of this tile is contiguous (blocked exchange layout, column
        // tiles): the last pass stores the tile in natural order [k][lane] into the shared buffer, then ONE bulk-async copy
        // per destination moves os_blk * L elements straight into that GPU's memory.  The store IS the all-to-all, issued by
        // the TMA engine in large transactions instead of 16-byte st.global per thread.
        static_assert(COLS, "bulk block stores are for column tiles");
        auto stage = [&](int k, Cx<R> val) { c.smem[k * L + c.l] = cmake<R>(val.x * sc, val.y * sy); };
        SfftAll<R, S, L, COLS, 0, false, (S::NP > 1)>::run(c, v, tw, load, stage);
        bulk_store_fence();                         // every writer: its stores become visible to the async proxy ...
        __syncthreads();                            // ... and the issuing threads are ordered after all of them
        if (threadIdx.x < (unsigned)a.nblk_ptr) {   // thread p sends block p (only lane 0's base is used)
            const LaneBase lb0 = lane_base(a, tile * L, true, 0);
            const int p = (int)threadIdx.x;
            bulk_store_issue(reinterpret_cast<Cx<R>*>(a.blk_ptr[p]) + lb0.bo, c.smem + (size_t)p * a.os_blk * L,
                             (unsigned)(sizeof(Cx<R>) * (size_t)a.os_blk * L));
            bulk_store_commit_and_drain();
        }
        return;
    }
    auto store = [&](int k, Cx<R> val) {
        if (!valid) return;
        Cx<R> y = cmake<R>(val.x * sc, val.y * sy);
        if (MODE == 1) {
            y = cmul(y, ldg(&reinterpret_cast<const Cx<R>*>(a.fs_lo)[(unsigned)k * (unsigned)j2]));
        } else if (MODE == 2) {
            if (a.fs_twiddle) {
                const Cx<R>* lo = reinterpret_cast<const Cx<R>*>(a.fs_lo);
                const Cx<R>* hi = reinterpret_cast<const Cx<R>*>(a.fs_hi);
                const unsigned long long ee = (unsigned long long)k * (unsigned long long)j2;
                if (a.fs_shift >= 40) y = cmul(y, ldg(&lo[(unsigned)ee]));
                else y = cmul(y, cmul(ldg(&hi[ee >> a.fs_shift]), ldg(&lo[ee & ((1ull << a.fs_shift) - 1)])));
            }
            if (a.os_blk) {
                const int pblk = k / a.os_blk;
                Cx<R>* base = a.nblk_ptr ? reinterpret_cast<Cx<R>*>(a.blk_ptr[pblk]) + lb.bo : out + (long long)pblk * a.os_blk_stride;
                base[(long long)(k - pblk * a.os_blk) * os_axis] = y;
                return;
            }
        }
        out[(long long)k * os_axis] = y;
    };
    SfftAll<R, S, L, COLS, 0, false, false>::run(c, v, tw, load, store);
}

// ------------------------------------------------------------------------------------------------------
// transposing rows: contiguous lanes in, lane-interleaved out.  out[k1 + N1 k] of the LAST pass of a three-pass split
// (2^24 = 256 x 256 x 256): the lanes k1 of a tile are 2 KiB-long contiguous rows in the workspace, N2 N3 elements apart,
// and adjacent elements of the output.  Passes 0 .. NP-2 use the row mapping (threads along the axis: coalesced reads);
// for the last pass the threads are re-dealt with the LANE as the fastest index, so every store instruction writes L
// adjacent output elements.  The exchange between the two mappings is the ordinary shared-memory exchange of the
// schedule; only the lane pitch is padded (PX) so that lane-strided reads of the last pass are conflict free.
// ------------------------------------------------------------------------------------------------------
template <class S, int W>   // W = 32-bit words per element
constexpr int sfft_trans_px() {
    int px = 0;
    while (((S::NPAD + px) * W) % 32 != W % 32) ++px;
    return px;
}
template <typename R, class S, int L, int PASS, class CtxA, class CtxB, class LoadF, class StoreF>
NDFB_DEV void sfft_trans_passes(const CtxA& c, const CtxB& c2, Cx<R> (&v)[S::E], const Cx<R>* tw, LoadF load, StoreF store) {
    if constexpr (PASS == S::NP - 1) {
        SfftPass<R, S, L, false, PASS, false, false>::run(c2, v, tw, load, store);
    } else {
        SfftPass<R, S, L, false, PASS, false, false>::run(c, v, tw, load, store);
        sfft_trans_passes<R, S, L, PASS + 1>(c, c2, v, tw, load, store);
    }
}
template <typename R, class S, int L>
NDFB_DEV void sfft_body_trans(const SfftArgs& a, long long tile) {
    constexpr int PX = sfft_trans_px<S, (int)(sizeof(Cx<R>) / 4)>();
    NDFB_DYN_SMEM(smem_raw);
    SfftCtx<R, S, L, false, PX> c, c2;
    c.smem = c2.smem = reinterpret_cast<Cx<R>*>(smem_raw);
    const int tid = threadIdx.x;
    c.i = tid % S::TL; c.l = tid / S::TL;      // along the axis (global reads of pass 0)
    c2.i = tid / L; c2.l = tid % L;            // across the lanes (global writes of the last pass)
    const long long g = tile * L + c.l, g2 = tile * L + c2.l;
    c.valid = g < a.nlanes; c2.valid = g2 < a.nlanes;
    const LaneBase lb = lane_base(a, g, c.valid, 0), lb2 = lane_base(a, g2, c2.valid, 0);
    const R sc = (R)a.scale;
    SfftGLoad<R, false> gl;
    gl.in = reinterpret_cast<const Cx<R>*>(a.in) + lb.bi; gl.is_axis = a.is_axis; gl.sgn = a.conj_in ? (R)-1 : (R)1; gl.valid = c.valid;
    SfftGStore<R, 0> gs;
    gs.out = reinterpret_cast<Cx<R>*>(a.out) + lb2.bo; gs.os_axis = a.os_axis; gs.sc = sc; gs.sy = a.conj_out ? -sc : sc; gs.valid = c2.valid;
    gs.fs = nullptr; gs.j2 = 0;
    Cx<R> v[S::E];
    sfft_trans_passes<R, S, L, 0>(c, c2, v, reinterpret_cast<const Cx<R>*>(a.tw), gl, gs);
}

// ------------------------------------------------------------------------------------------------------
// Contiguous rows as a PERSISTENT, bulk-async pipelined kernel (TMA 1-D copies + mbarrier):
//     TMA load of tile t+1 -> [in buffer]      |  passes of tile t on [work buffer]  |   TMA store of tile t-1 <- [out buffer]
// One elected thread issues cp.async.bulk global->shared with an mbarrier transaction count; every thread waits on the
// barrier's phase; pass 0 reads its points from the in buffer (LDS instead of LDG: no global latency on the critical
// path), the buffer is handed back to the TMA engine right after pass 0, the last pass writes natural order into the
// out buffer and one thread sends it with cp.async.bulk shared->global.  The register-resident row kernel needs two CTAs
// per SM to cover its own load / store phases; here the copy engine covers them and the SM only issues FFT work.
// ------------------------------------------------------------------------------------------------------
#ifdef NDFB_EMU
struct BulkBar { unsigned long long v; };
NDFB_DEV void bulk_bar_init(BulkBar* b) { b->v = 0; }
NDFB_DEV void bulk_load_issue(void* sdst, const void* gsrc, unsigned bytes, BulkBar*) {
    unsigned char* d = reinterpret_cast<unsigned char*>(sdst);
    const unsigned char* q = reinterpret_cast<const unsigned char*>(gsrc);
    for (unsigned i = 0; i < bytes; ++i) d[i] = q[i];
}
NDFB_DEV void bulk_bar_expect(BulkBar*, unsigned) {}
NDFB_DEV void bulk_bar_wait(BulkBar*, unsigned) {}
NDFB_DEV void bulk_store_wait_read() {}
#else
struct BulkBar { unsigned long long v; };
NDFB_DEV void bulk_bar_init(BulkBar* b) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(b);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the init must be visible to the async proxy
}
NDFB_DEV void bulk_bar_expect(BulkBar* b, unsigned bytes) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(b);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
NDFB_DEV void bulk_load_issue(void* sdst, const void* gsrc, unsigned bytes, BulkBar* b) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(sdst), a = (unsigned)__cvta_generic_to_shared(b);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(gsrc), "r"(bytes), "r"(a)
                 : "memory");
}
NDFB_DEV void bulk_bar_wait(BulkBar* b, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(b);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(a), "r"(parity)
        : "memory");
}
NDFB_DEV void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
#endif

template <typename R, class S, int L>
struct SfftBulkSmem {
    static constexpr size_t kWork = sizeof(Cx<R>) * (size_t)L * S::NPAD;
    static constexpr size_t kIO = sizeof(Cx<R>) * (size_t)L * S::N;
    static constexpr size_t kTotal = kWork + 2 * kIO + 16;
};

template <typename R, class S, int L, int MINB>
__global__ void __launch_bounds__(S::TL* L, MINB) sfft_rows_bulk_kernel(const __grid_constant__ SfftArgs a) {
    static_assert(S::NP > 1, "single-pass schedules have nothing to overlap");
    NDFB_DYN_SMEM(smem_raw);
    using SM = SfftBulkSmem<R, S, L>;
    Cx<R>* work = reinterpret_cast<Cx<R>*>(smem_raw);
    Cx<R>* inb = reinterpret_cast<Cx<R>*>(smem_raw + SM::kWork);
    Cx<R>* outb = reinterpret_cast<Cx<R>*>(smem_raw + SM::kWork + SM::kIO);
    BulkBar* bar = reinterpret_cast<BulkBar*>(smem_raw + SM::kWork + 2 * SM::kIO);
    SfftCtx<R, S, L, false> c;
    c.smem = work;
    const int tid = threadIdx.x;
    c.i = tid % S::TL; c.l = tid / S::TL;
    const long long ntiles = (a.nlanes + L - 1) / L;
    constexpr unsigned kRowBytes = (unsigned)(sizeof(Cx<R>) * S::N);
    const Cx<R>* __restrict__ tw = reinterpret_cast<const Cx<R>*>(a.tw);
    const R sc = (R)a.scale;
    const R sy = a.conj_out ? -sc : sc;
    const R sgn_in = a.conj_in ? (R)-1 : (R)1;
    // one thread: bring the rows of `tile` into the in buffer (one bulk copy per lane: lanes need not be adjacent in memory)
    auto issue_load = [&](long long tile) {
        unsigned total = 0;
        for (int l = 0; l < L; ++l)
            if (tile * L + l < a.nlanes) total += kRowBytes;
        bulk_bar_expect(bar, total);
        for (int l = 0; l < L; ++l) {
            const long long g = tile * L + l;
            if (g >= a.nlanes) break;
            const LaneBase lb = lane_base(a, g, true, 0);
            bulk_load_issue(inb + (size_t)l * S::N, reinterpret_cast<const Cx<R>*>(a.in) + lb.bi, kRowBytes, bar);
        }
    };
    if (tid == 0) {
        bulk_bar_init(bar);
        if ((long long)blockIdx.x < ntiles) issue_load(blockIdx.x);
    }
    __syncthreads();
    unsigned parity = 0;
    Cx<R> v[S::E];
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (tid == 0) bulk_store_wait_read();      // the previous tile's store has finished READING the out buffer
        bulk_bar_wait(bar, parity);                 // this tile's rows have landed in the in buffer
        parity ^= 1u;
        const bool valid = tile * L + c.l < a.nlanes;
        const Cx<R>* __restrict__ row = inb + (size_t)c.l * S::N;
        auto load = [&](int j) -> Cx<R> {
            Cx<R> x = valid ? row[j] : cmake<R>((R)0, (R)0);
            x.y *= sgn_in;
            return x;
        };
        Cx<R>* __restrict__ orow = outb + (size_t)c.l * S::N;
        auto store = [&](int k, Cx<R> val) { orow[k] = cmake<R>(val.x * sc, val.y * sy); };
        SfftPass<R, S, L, false, 0, false, false>::run(c, v, tw, load, store);    // ends with a barrier: the in buffer is free
        if (tid == 0 && tile + gridDim.x < ntiles) issue_load(tile + gridDim.x);  // next tile streams in behind the passes
        SfftAll<R, S, L, false, 1, false, false>::run(c, v, tw, load, store);
        bulk_store_fence();
        __syncthreads();
        if (tid == 0) {
            for (int l = 0; l < L; ++l) {
                const long long g = tile * L + l;
                if (g >= a.nlanes) break;
                const LaneBase lb = lane_base(a, g, true, 0);
                bulk_store_issue(reinterpret_cast<Cx<R>*>(a.out) + lb.bo, outb + (size_t)l * S::N, kRowBytes);
            }
#ifndef NDFB_EMU
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#endif
        }
    }
    if (tid == 0) bulk_store_wait_read();
}

NDFB_DEV unsigned sync_ld_acquire(const unsigned* p) {
#ifdef NDFB_EMU
    return *p;
#else
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}
// bounded poll by one thread (a logic error or a lost producer must not hang the GPU: after ~1 s the consumer goes on)
NDFB_DEV void sync_wait_ge(const unsigned* counter, unsigned need) {
    unsigned spins = 0;
    while (sync_ld_acquire(counter) < need) {
#ifndef NDFB_EMU
        __nanosleep(128);
#endif
        if (++spins > (1u << 23)) break;
    }
}

// lengths whose column kernels carry the bulk-async scatter epilogue (the axis lengths a slab exchange splits; keeping the
// variant out of the other ~400 column instances keeps the library and its build time down)
template <class S, bool COLS>
constexpr bool kSfftBulkStore = COLS && S::NP > 1 && S::N >= 64 && S::N <= 2048;

template <typename R, class S, int L, bool COLS, int MINB>
__global__ void __launch_bounds__(S::TL* L, MINB) sfft_kernel(const __grid_constant__ SfftArgs a) {
    NDFB_DYN_SMEM(smem_raw);
    SfftCtx<R, S, L, COLS> c;
    c.smem = reinterpret_cast<Cx<R>*>(smem_raw);
    const int tid = threadIdx.x;
    if (COLS) { c.l = tid % L; c.i = tid / L; }
    else { c.i = tid % S::TL; c.l = tid / S::TL; }
    const bool plain = !a.fs_twiddle && !a.os_blk;
    if constexpr (kSfftBulkStore<S, COLS>) {
        if (a.ntiles) {
            // persistent consumer of another launch's output (see SfftArgs::wait_cnt): a few CTAs per SM walk the tiles
            for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
                if (a.wait_cnt) {
                    if (tid == 0) sync_wait_ge(&a.wait_cnt[(tile * L) / a.wait_group], a.wait_need);
                    __syncthreads();
                }
                const long long g = tile * L + c.l;
                c.valid = g < a.nlanes;
                const LaneBase lb = lane_base(a, g, c.valid, a.fs_dim);
                if (plain) sfft_body<R, S, L, COLS, 0, false>(a, c, lb, tile);
                else if (a.bulk_store) sfft_body<R, S, L, COLS, 3, false>(a, c, lb, tile);
                else sfft_body<R, S, L, COLS, 2, false>(a, c, lb, tile);
                __syncthreads();     // the shared buffer is reused by the next tile
            }
            return;
        }
    }
    if constexpr (!COLS && S::NP > 1 && L > 1) {
        if (a.trans_store) { sfft_body_trans<R, S, L>(a, blockIdx.x); return; }
    }
#ifndef NDFB_EMU
    if constexpr (!COLS) {
        if (a.l2_prefetch_lanes && tid < L) {
            const long long gp = (long long)blockIdx.x * L + tid + a.l2_prefetch_lanes;
            if (gp < a.nlanes) {
                const LaneBase lp = lane_base(a, gp, true, 0);
                const Cx<R>* row = reinterpret_cast<const Cx<R>*>(a.in) + lp.bi;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(row), "r"((unsigned)(sizeof(Cx<R>) * S::N)) : "memory");
            }
        }
    }
#endif
    const long long g = (long long)blockIdx.x * L + c.l;
    c.valid = g < a.nlanes;
    const LaneBase lb = lane_base(a, g, c.valid, a.fs_dim);
    if (plain && !COLS && a.is_axis == 1 && a.os_axis == 1) sfft_body<R, S, L, COLS, 0, true>(a, c, lb, blockIdx.x);
    else if (plain) sfft_body<R, S, L, COLS, 0, false>(a, c, lb, blockIdx.x);
    else if (COLS && a.fs_twiddle && a.fs_q && !a.os_blk) sfft_body<R, S, L, COLS, 4, false>(a, c, lb, blockIdx.x);
    else if (a.fs_twiddle && a.fs_shift >= 40 && !a.os_blk) sfft_body<R, S, L, COLS, 1, false>(a, c, lb, blockIdx.x);
    else if (kSfftBulkStore<S, COLS> && a.bulk_store) {
        if constexpr (kSfftBulkStore<S, COLS>) sfft_body<R, S, L, COLS, 3, false>(a, c, lb, blockIdx.x);
    } else sfft_body<R, S, L, COLS, 2, false>(a, c, lb, blockIdx.x);
}

// ------------------------------------------------------------------------------------------------------
// Two-pass (four-step) transform of long STRIDED columns in ONE persistent launch, workspace resident in L2.
//
// The columns are cut into groups of Wg adjacent columns.  Pass 1 of group g (N1-point transforms + W_N^{k1 j2} twiddle)
// writes a small ring buffer; pass 2 of the same group (N2-point transforms) reads it back while it is still in the
// 126 MB L2 and writes the user's output.  CTAs draw tiles from one ticket counter in the order
//     P1(0), P1(1), P2(0), P1(2), P2(1), ... , P2(G-1)
// and a pass-2 tile waits until its group's pass-1 tiles have all signalled (a pass-1 tile waits until the ring slot it
// overwrites has been consumed).  Every tile a ticket can wait for has a smaller ticket, so the oldest unfinished tile
// never waits and the launch cannot deadlock, whatever number of CTAs is resident.  HBM sees the input once and the
// output once; the workspace traffic stays in L2.
// ------------------------------------------------------------------------------------------------------
struct Fs2Args {
    SfftArgs a1, a2;        // pass 1: user input -> ring (four-step twiddle on);  pass 2: ring -> user output
    unsigned* sync;         // [0] ticket, [1] error flag, [2 + g] pass-1 tiles done, [2 + G + g] pass-2 tiles done
    int G, Wg, ring;        // groups, columns per group, ring slots
    long long ring_stride;  // elements per ring slot (N * Wg)
    long long col_is, col_os;      // column strides of the user's arrays
    long long j2_is, k1_os;        // user-array strides of the pass-1 lane index j2 and of the pass-2 lane index k1
    int N1, N2;
    unsigned tiles1, tiles2;       // tiles per group
    int dbg;                       // timing experiments only (NDFB_FS2_DBG): 1 = no release fence, 2 = no waits (results invalid)
};

NDFB_DEV unsigned fs2_ld_acquire(const unsigned* p) {
#ifdef NDFB_EMU
    return *p;
#else
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}

// ticket -> (pass, group, tile) in the order P1(0), P1(1), P2(0), P1(2), P2(1), ..., P2(G-1)
struct Fs2Tile { int pass, g; unsigned tile; };
NDFB_DEV Fs2Tile fs2_decode(const Fs2Args& f, unsigned t) {
    Fs2Tile o;
    const unsigned per = f.tiles1 + f.tiles2;
    if (t < f.tiles1) { o.pass = 0; o.g = 0; o.tile = t; return o; }
    const unsigned tt = t - f.tiles1;
    const int u = 1 + (int)(tt / per);
    const unsigned r = tt - (unsigned)(u - 1) * per;
    if (u < f.G && r < f.tiles1) { o.pass = 0; o.g = u; o.tile = r; }
    else { o.pass = 1; o.g = u - 1; o.tile = u < f.G ? r - f.tiles1 : r; }
    return o;
}

// bounded poll (thread 0 only), so that a logic error cannot hang the GPU: it raises the error flag instead
NDFB_DEV void fs2_poll(unsigned* sync, const unsigned* counter, unsigned need) {
    unsigned spins = 0;
    while (fs2_ld_acquire(counter) < need) {
#ifndef NDFB_EMU
        __nanosleep(64);
#endif
        if (++spins > (1u << 22)) { atomicExch(&sync[1], 1u); break; }
    }
}

template <typename R, class S1, int L1, class S2, int L2, int MINB>
__global__ void __launch_bounds__(S1::TL* L1, MINB) fs2_kernel(const __grid_constant__ Fs2Args f) {
    static_assert(S1::TL * L1 == S2::TL * L2, "both passes run with the same CTA size");
    NDFB_DYN_SMEM(smem_raw);
    constexpr size_t kTile1 = sizeof(Cx<R>) * (size_t)L1 * S1::NPAD, kTile2 = sizeof(Cx<R>) * (size_t)L2 * S2::NPAD;
    unsigned& s_ticket = *reinterpret_cast<unsigned*>(smem_raw + (kTile1 > kTile2 ? kTile1 : kTile2));   // behind the tile buffer
    const int tid = threadIdx.x;
    const unsigned total = (unsigned)f.G * (f.tiles1 + f.tiles2);
    unsigned* pending = nullptr;   // (thread 0) counter of the tile this CTA has just finished
    bool pending_fence = false;
    for (;;) {
        __syncthreads();           // every thread has issued the previous tile's stores and is done with the shared buffer
        if (tid == 0) {
            if (pending) {
                if (pending_fence) __threadfence();   // cumulative over the CTA's stores (ordered before by the barrier)
                atomicAdd(pending, 1u);
            }
            const unsigned t = fs2_ld_acquire(&f.sync[1]) ? 0xffffffffu : atomicAdd(&f.sync[0], 1u);   // (error flag: drain)
            s_ticket = t;
            if (t < total && !(f.dbg & 2)) {
                const Fs2Tile w = fs2_decode(f, t);
                if (w.pass == 0) { if (w.g >= f.ring) fs2_poll(f.sync, &f.sync[2 + f.G + w.g - f.ring], f.tiles2); }   // ring slot consumed
                else fs2_poll(f.sync, &f.sync[2 + w.g], f.tiles1);                                                     // group produced
            }
        }
        __syncthreads();
        const unsigned t = s_ticket;
        if (t >= total) return;
        const Fs2Tile w = fs2_decode(f, t);
        const int g = w.g;
        const long long slot = (long long)(g % f.ring) * f.ring_stride;
        if (w.pass == 0) {
            SfftCtx<R, S1, L1, true> c;
            c.smem = reinterpret_cast<Cx<R>*>(smem_raw);
            c.l = tid % L1; c.i = tid / L1; c.valid = true;
            const int cbn = f.Wg / L1;
            const int cb = (int)(w.tile % (unsigned)cbn), j2 = (int)(w.tile / (unsigned)cbn);
            const int wl = cb * L1 + c.l;
            LaneBase lb;
            lb.j2 = j2;
            lb.bi = (long long)j2 * f.j2_is + ((long long)g * f.Wg + wl) * f.col_is;
            lb.bo = slot + (long long)j2 * f.Wg + wl;
            sfft_body<R, S1, L1, true, 1, false>(f.a1, c, lb);
            pending = &f.sync[2 + g];
            pending_fence = !(f.dbg & 1);
        } else {
            SfftCtx<R, S2, L2, true> c;
            c.smem = reinterpret_cast<Cx<R>*>(smem_raw);
            c.l = tid % L2; c.i = tid / L2; c.valid = true;
            const int cbn = f.Wg / L2;
            const int cb = (int)(w.tile % (unsigned)cbn), k1 = (int)(w.tile / (unsigned)cbn);
            const int wl = cb * L2 + c.l;
            LaneBase lb;
            lb.j2 = 0;
            lb.bi = slot + (long long)k1 * f.N2 * f.Wg + wl;
            lb.bo = (long long)k1 * f.k1_os + ((long long)g * f.Wg + wl) * f.col_os;
            sfft_body<R, S2, L2, true, 0, false, true>(f.a2, c, lb);
            pending = &f.sync[2 + f.G + g];
            pending_fence = false;   // a consumer only reports that it has READ its ring slot
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// Bluestein (chirp-z) C2C for lengths without an instantiated schedule, fused in ONE kernel around two Stockham
// transforms of a power-of-two length M >= 2N-1 that does have one:
//     a[j] = x[j] c[j] (zero padded to M)  ->  A = FFT_M(a)  ->  conj(A * bhat)  ->  FFT_M  ->  X[k] = conj(.) c[k]
// The tile never leaves the chip between the two transforms (the first one stores into the shared buffer, the second
// loads from it), so HBM traffic stays input-once + output-once.  Replaces rustfft's BluesteinsAlgorithm / RadersAlgorithm
// behind plan_fft_forward/inverse (src/lib.rs:295-297).
// ------------------------------------------------------------------------------------------------------
struct BsfftArgs {
    SfftArgs base;       // lanes, strides, conj flags, scale, per-pass twiddles of the M-point schedule
    int n;               // actual transform length N (<= (M+1)/2)
    const void* chirp;   // c[j] = exp(-i pi j^2 / N), j < N
    const void* bhat;    // FFT_M(conj-chirp kernel) / M, natural order
};

template <typename R, class S, int L, bool COLS, int MINB>
__global__ void __launch_bounds__(S::TL* L, MINB) bsfft_kernel(const __grid_constant__ BsfftArgs ba) {
    const SfftArgs& a = ba.base;
    NDFB_DYN_SMEM(smem_raw);
    SfftCtx<R, S, L, COLS> c;
    c.smem = reinterpret_cast<Cx<R>*>(smem_raw);
    const int tid = threadIdx.x;
    if (COLS) { c.l = tid % L; c.i = tid / L; }
    else { c.i = tid % S::TL; c.l = tid / S::TL; }
    const long long g = (long long)blockIdx.x * L + c.l;
    c.valid = g < a.nlanes;
    const bool valid = c.valid;
    const LaneBase lb = lane_base(a, g, valid, 0);
    const Cx<R>* __restrict__ in = reinterpret_cast<const Cx<R>*>(a.in) + lb.bi;
    Cx<R>* __restrict__ out = reinterpret_cast<Cx<R>*>(a.out) + lb.bo;
    const Cx<R>* __restrict__ tw = reinterpret_cast<const Cx<R>*>(a.tw);
    const Cx<R>* __restrict__ chirp = reinterpret_cast<const Cx<R>*>(ba.chirp);
    const Cx<R>* __restrict__ bhat = reinterpret_cast<const Cx<R>*>(ba.bhat);
    const long long is_axis = a.is_axis, os_axis = a.os_axis;
    const int N = ba.n;
    const R sc = (R)a.scale;
    const R sy = a.conj_out ? -sc : sc;
    const R sgn_in = a.conj_in ? (R)-1 : (R)1;
    const R zero = (R)0;
    Cx<R> v[S::E];
    auto load1 = [&](int j) -> Cx<R> {
        if (j >= N || !valid) return cmake<R>(zero, zero);
        Cx<R> x = ld_stream(&in[(long long)j * is_axis]);
        x.y *= sgn_in;
        return cmul(x, ldg(&chirp[j]));
    };
    auto store1 = [&](int k, Cx<R> val) { c.smem[c.addr(k)] = cconj(cmul(val, ldg(&bhat[k]))); };
    SfftAll<R, S, L, COLS, 0, false, (S::NP > 1)>::run(c, v, tw, load1, store1);
    __syncthreads();
    auto load2 = [&](int j) -> Cx<R> { return c.smem[c.addr(j)]; };
    auto store2 = [&](int k, Cx<R> val) {
        if (k >= N || !valid) return;
        const Cx<R> y = cmul(cconj(val), ldg(&chirp[k]));
        out[(long long)k * os_axis] = cmake<R>(y.x * sc, y.y * sy);
    };
    SfftAll<R, S, L, COLS, 0, true, false>::run(c, v, tw, load2, store2);
}

// ------------------------------------------------------------------------------------------------------
// real transforms of even length around the same Stockham core of N = n/2 (DCT-I: N = n-1) complex points:
// R2C, C2R, DCT-I..IV.  The algebra is the one verified in tests/kernel_math_model.py and used by
// tile_kernel.cuh; here it is fused into the first-pass loads and the last-pass stores:
//   * strided columns (COLS): every real load/store is coalesced across the L adjacent lanes whatever the
//     row order, so reorders (Makhoul, even extension, zip) cost nothing;
//   * contiguous rows: kinds whose access along the row is strided (DCT-II in, DCT-III out, DCT-IV in/out)
//     are staged through the shared buffer with coalesced copies.
// Replaces fft_r2c_lane / ifft_r2c_lane (src/lib.rs:497-523) and dct1..4_lane (src/lib.rs:688-734).
// ------------------------------------------------------------------------------------------------------
enum RKind : int { RK_R2C = 0, RK_C2R = 1, RK_DCT1 = 2, RK_DCT2 = 3, RK_DCT3 = 4, RK_DCT4 = 5 };

struct RsfftArgs {
    const void* in;
    void* out;
    long long nlanes;
    int nbd;
    long long bsz[kMaxBatchDims], bis[kMaxBatchDims], bos[kMaxBatchDims];
    long long is_axis, os_axis;
    int n;            // logical (real) length
    double scale;
    const void* tw;   // per-pass twiddle tables
    unsigned* done_cnt;     // producer side of SfftArgs::wait_cnt: done_cnt[first lane / done_group] += lanes of this tile, after its stores
    long long done_group;
    const void* tabA; // exp(-2 pi i k / (2N)), k <= N   (DCT-IV: exp(-i pi j / n))
    const void* tabB; // DCT-II/III: exp(-i pi k / (2n));  DCT-IV: exp(-i pi (4j+1) / (4n))
    int vec_out;      // contiguous output rows start on 2-real (1) / 4-real (2) boundaries: DCT-III / DCT-IV may store pairs / quads of
                      // reals straight from registers (kMirrorOut)
};

// UNIT: contiguous rows on both sides (axis strides 1): address arithmetic folds to constants
// VEC: (UNIT rows) output rows are aligned for 2-real vector stores: DCT-III / DCT-IV run the mirror-paired last pass (kMirrorOut)
template <typename R, class S, int L, bool COLS, int KIND, bool UNIT, bool VEC = false>
NDFB_DEV void rsfft_body(const RsfftArgs& a) {
    constexpr int N = S::N;
    constexpr int n = KIND == RK_DCT1 ? N + 1 : 2 * N;   // logical real length (== a.n; the host only launches matching plans)
    constexpr bool IN_CX = KIND == RK_C2R;
    constexpr bool OUT_CX = KIND == RK_R2C;
    // which sides need the coalesced staging copy when the lane is a contiguous row
    constexpr bool STAGE_IN = !COLS && (KIND == RK_DCT2 || KIND == RK_DCT4);
    constexpr bool MIRROR_OUT = VEC && UNIT && !COLS && (KIND == RK_DCT3 || KIND == RK_DCT4) && kMirrorOut<S>;
    constexpr bool STAGE_OUT = !COLS && (KIND == RK_DCT3 || KIND == RK_DCT4) && !MIRROR_OUT;
    constexpr bool PAIR_EPI = KIND == RK_R2C || KIND == RK_DCT1 || KIND == RK_DCT2;   // outputs need Z[k] and Z[N-k]
    NDFB_DYN_SMEM(smem_raw);
    SfftCtx<R, S, L, COLS> c;
    c.smem = reinterpret_cast<Cx<R>*>(smem_raw);
    R* smem_r = reinterpret_cast<R*>(smem_raw);
    const int tid = threadIdx.x;
    if (COLS) { c.l = tid % L; c.i = tid / L; }
    else { c.i = tid % S::TL; c.l = tid / S::TL; }
    const long long g = (long long)blockIdx.x * L + c.l;
    c.valid = g < a.nlanes;
    const bool valid = c.valid;
    const LaneBase lb = lane_base(a, g, valid);
    const R* __restrict__ in_r = reinterpret_cast<const R*>(a.in) + (IN_CX ? 2 * lb.bi : lb.bi);
    const Cx<R>* __restrict__ in_c = reinterpret_cast<const Cx<R>*>(a.in) + lb.bi;
    R* __restrict__ out_r = reinterpret_cast<R*>(a.out) + (OUT_CX ? 2 * lb.bo : lb.bo);
    Cx<R>* __restrict__ out_c = reinterpret_cast<Cx<R>*>(a.out) + lb.bo;
    const long long is_axis = UNIT ? 1 : a.is_axis, os_axis = UNIT ? 1 : a.os_axis;
    const Cx<R>* __restrict__ tw = reinterpret_cast<const Cx<R>*>(a.tw);
    const Cx<R>* __restrict__ tabA = reinterpret_cast<const Cx<R>*>(a.tabA);
    const Cx<R>* __restrict__ tabB = reinterpret_cast<const Cx<R>*>(a.tabB);
    const R sc = (R)a.scale;
    const R zero = (R)0;
    // staged real lane in shared memory (rows only): real index t of lane l
    constexpr int RPITCH = 2 * S::NPAD;
    auto sreal = [&](int t) -> R& { return smem_r[c.l * RPITCH + t]; };
    auto gin = [&](int t) -> R { return valid ? in_r[(long long)t * is_axis] : zero; };          // real input element t

    // staged complex slot j of this lane (two adjacent reals)
    auto scx = [&](int j) -> Cx<R>& { return reinterpret_cast<Cx<R>*>(smem_r + c.l * RPITCH)[j]; };
    if (STAGE_IN) {
        // coalesced global read; the kind's reorder is applied on the shared-memory side so that the first pass reads
        // plain complex slots (conflict free):  DCT-II: Makhoul order v[p(t)];  DCT-IV: u[j] = (x[2j], x[n-1-2j])
        // (compile-time trip count: the copies unroll and all of a thread's loads are in flight together)
        constexpr int IT = (n + S::TL - 1) / S::TL;
#pragma unroll
        for (int m = 0; m < IT; ++m) {
            const int t = c.i + m * S::TL;
            if (n % S::TL != 0 && t >= n) break;
            const int pos = KIND == RK_DCT2 ? ((t & 1) ? n - 1 - (t >> 1) : (t >> 1)) : ((t & 1) ? n - t : t);
            sreal(pos) = valid ? ld_stream(&in_r[(long long)t * is_axis]) : zero;
        }
        __syncthreads();
    }

    // ---- C2R / DCT-III "zip" prologue, pairwise: slots j and N-j are built from the same two spectrum bins
    // (E = X[j] + conj(X[N-j]), O = X[j] - conj(X[N-j]), w = conj(tabA[j]) O:  z[j] = conj(E + i w),  z[N-j] = E - i w,
    // because tabA[N-j] = -conj(tabA[j])), so one thread loads them once and writes both slots to the shared buffer;
    // the first pass then reads plain complex slots.  Halves the global/table loads and the prologue arithmetic. ----
    constexpr bool PAIR_PRO = KIND == RK_C2R || KIND == RK_DCT3;
    constexpr bool MIRROR_PRO = PAIR_PRO && kMirrorPro<S>;
    if (PAIR_PRO && !MIRROR_PRO) {
        constexpr int ITQ = (N / 2 + 1 + S::TL - 1) / S::TL;
#pragma unroll
        for (int m = 0; m < ITQ; ++m) {
            const int j = c.i + m * S::TL;
            if (j > N / 2) break;
            const int k2 = N - j;
            Cx<R> xk, xn;
            if (KIND == RK_C2R) {
                xk = valid ? in_c[(long long)j * is_axis] : cmake<R>(zero, zero);
                xn = valid ? in_c[(long long)k2 * is_axis] : cmake<R>(zero, zero);
                if (j == 0) { xk.y = zero; xn.y = zero; }        // Im X[0], Im X[N] dropped (src/lib.rs:516-521)
            } else {
                // P[k] = (y[k], -y[n-k]) with y[n] = 0, then V[k] = conj(t_k) P[k]
                Cx<R> pk = cmake<R>(gin(j), j == 0 ? zero : -gin(n - j));
                Cx<R> pn = cmake<R>(gin(k2), -gin(n - k2));
                xk = cmul(pk, cconj(ldg(&tabB[j])));
                xn = cmul(pn, cconj(ldg(&tabB[k2])));
            }
            const Cx<R> wc = cconj(ldg(&tabA[j]));
            const Cx<R> E = cadd(xk, cconj(xn)), O = csub(xk, cconj(xn));
            const Cx<R> iw = cmul_i(cmul(wc, O));
            c.smem[c.addr(j)] = cconj(cadd(E, iw));
            if (j != 0 && k2 != j) c.smem[c.addr(k2)] = csub(E, iw);
        }
        __syncthreads();
    }

    // ---- first-pass input z[j], j < N ----
    auto load = [&](int j) -> Cx<R> {
        if (PAIR_PRO) return c.smem[c.addr(j)];
        if (KIND == RK_R2C) {
            if (!COLS && is_axis == 1) {   // (x[2j], x[2j+1]) is one aligned complex load
                return valid ? reinterpret_cast<const Cx<R>*>(in_r)[j] : cmake<R>(zero, zero);
            }
            return cmake<R>(gin(2 * j), gin(2 * j + 1));
        } else if (KIND == RK_C2R || KIND == RK_DCT3) {
            Cx<R> xk, xn;
            const int k2 = N - j;
            if (KIND == RK_C2R) {
                xk = valid ? in_c[(long long)j * is_axis] : cmake<R>(zero, zero);
                xn = valid ? in_c[(long long)k2 * is_axis] : cmake<R>(zero, zero);
                if (j == 0) { xk.y = zero; xn.y = zero; }        // Im X[0], Im X[N] dropped (src/lib.rs:516-521)
            } else {
                // P[k] = (y[k], -y[n-k]) with y[n] = 0, then V[k] = conj(t_k) P[k]
                Cx<R> pk = cmake<R>(gin(j), j == 0 ? zero : -gin(n - j));
                Cx<R> pn = cmake<R>(gin(k2), -gin(n - k2));
                xk = cmul(pk, cconj(ldg(&tabB[j])));
                xn = cmul(pn, cconj(ldg(&tabB[k2])));
            }
            const Cx<R> wc = cconj(ldg(&tabA[j]));
            const Cx<R> E = cadd(xk, cconj(xn)), O = csub(xk, cconj(xn));
            return cconj(cadd(E, cmul_i(cmul(wc, O))));
        } else if (KIND == RK_DCT1) {
            // even extension e[t] = x[t] (t <= N), x[2N - t] otherwise; z[j] = (e[2j], e[2j+1])
            const int t0 = 2 * j, t1 = 2 * j + 1;
            return cmake<R>(gin(t0 <= N ? t0 : 2 * N - t0), gin(t1 <= N ? t1 : 2 * N - t1));
        } else if (KIND == RK_DCT2) {
            // Makhoul: v[t] = x[2t] (t < N) else x[2(n-1-t)+1]; z[j] = (v[2j], v[2j+1])
            if (STAGE_IN) return scx(j);
            const int t0 = 2 * j, t1 = 2 * j + 1;
            return cmake<R>(gin(t0 < N ? 2 * t0 : 2 * (n - 1 - t0) + 1), gin(t1 < N ? 2 * t1 : 2 * (n - 1 - t1) + 1));
        } else {  // RK_DCT4
            if (STAGE_IN) return cmul(scx(j), ldg(&tabA[j]));
            return cmul(cmake<R>(gin(2 * j), gin(n - 1 - 2 * j)), ldg(&tabA[j]));
        }
    };

    // ---- last-pass output ----
    auto put = [&](int t, R val) {   // real output element t
        if (STAGE_OUT) sreal(t) = val;
        else if (valid) out_r[(long long)t * os_axis] = val;
    };
    auto store = [&](int k, Cx<R> y) {
        if (PAIR_EPI) {
            c.smem[c.addr(k)] = y;   // natural order; combined below
        } else if (KIND == RK_C2R) {
            if (!COLS && os_axis == 1) {
                if (valid) reinterpret_cast<Cx<R>*>(out_r)[k] = cmake<R>(sc * y.x, -sc * y.y);
            } else {
                put(2 * k, sc * y.x);
                put(2 * k + 1, -sc * y.y);
            }
        } else if (KIND == RK_DCT3) {
            // v[2k] = y.x, v[2k+1] = -y.y;  x[o(t)] = v[t]/2, o(t) = 2t (t < N) else 2(n-1-t)+1
            const int t0 = 2 * k, t1 = 2 * k + 1;
            const R h = (R)0.5 * sc;
            if (STAGE_OUT) { scx(k) = cmake<R>(h * y.x, -h * y.y); return; }   // reorder applied by the copy-out loop
            put(t0 < N ? 2 * t0 : 2 * (n - 1 - t0) + 1, h * y.x);
            put(t1 < N ? 2 * t1 : 2 * (n - 1 - t1) + 1, -h * y.y);
        } else {  // RK_DCT4
            const Cx<R> C = cmul(y, ldg(&tabB[k]));
            if (STAGE_OUT) { scx(k) = cmake<R>(sc * C.x, -sc * C.y); return; }
            put(2 * k, sc * C.x);
            put(n - 1 - 2 * k, -sc * C.y);
        }
    };

    Cx<R> v[S::E];
    // bins k and k2 = N - k of the length-2N real DFT from Z[k] = zk and Z[N-k] = zm (k = 0: zm = Z[0], k2 = N);
    // both == false: only bin k is produced (its mirror is another thread's)
    auto emit = [&](int k, Cx<R> zk, Cx<R> zm, bool both) {
        const int k2 = N - k;
        const Cx<R> zc = cconj(zm);
        const Cx<R> w = ldg(&tabA[k]);
        const Cx<R> s = cadd(zk, zc), d = cmul(w, csub(zk, zc));
        const Cx<R> X = cmake<R>((R)0.5 * (s.x + d.y), (R)0.5 * (s.y - d.x));
        const Cx<R> X2 = cmake<R>((R)0.5 * (s.x - d.y), (R)-0.5 * (s.y + d.x));
        if (!valid) return;
        const bool two = both && k2 != k;
        if (KIND == RK_R2C) {
            out_c[(long long)k * os_axis] = cmake<R>(sc * X.x, sc * X.y);
            if (two) out_c[(long long)k2 * os_axis] = cmake<R>(sc * X2.x, sc * X2.y);
        } else if (KIND == RK_DCT1) {
            out_r[(long long)k * os_axis] = (R)0.5 * sc * X.x;
            if (two) out_r[(long long)k2 * os_axis] = (R)0.5 * sc * X2.x;
        } else if (KIND == RK_DCT2) {
            const Cx<R> A = cmul(X, ldg(&tabB[k]));
            out_r[(long long)k * os_axis] = sc * A.x;
            if (k > 0 && k < N) out_r[(long long)(n - k) * os_axis] = -sc * A.y;
            if (two) {
                const Cx<R> A2 = cmul(X2, ldg(&tabB[k2]));
                out_r[(long long)k2 * os_axis] = sc * A2.x;
                if (k2 > 0 && k2 < N) out_r[(long long)(n - k2) * os_axis] = -sc * A2.y;
            }
        }
    };
    if constexpr (PAIR_EPI && kMirrorEpi<S>) {
        constexpr int LP = S::NP - 1, r = S::radix(LP), P = S::before(LP), NB = S::nbf(LP), HG = S::G(LP) / 2;
        SfftUntil<R, S, L, COLS, 0, LP, STAGE_IN || PAIR_PRO>::run(c, v, tw, load, store);
        // pair m of this thread: butterflies p = i + TL m and NB - p (p = 0: butterflies 0 and NB/2, both their own mirror)
#pragma unroll
        for (int m = 0; m < HG; ++m) {
            const int p = c.i + S::TL * m;
            const int b0 = p, b1 = p == 0 ? NB / 2 : NB - p;
            Cx<R>* u = &v[2 * m * r];
#pragma unroll
            for (int q = 0; q < r; ++q) {
                u[q] = c.smem[c.addr(b0 + q * NB)];
                u[r + q] = c.smem[c.addr(b1 + q * NB)];
            }
            const Cx<R>* __restrict__ t0 = tw + S::twoff(LP) + (b0 % P);
            const Cx<R>* __restrict__ t1 = tw + S::twoff(LP) + (b1 % P);
#pragma unroll
            for (int q = 1; q < r; ++q) {
                u[q] = cmul(u[q], ldg(&t0[(q - 1) * P]));
                u[r + q] = cmul(u[r + q], ldg(&t1[(q - 1) * P]));
            }
            Dft<R, r>::run(&u[0]);
            Dft<R, r>::run(&u[r]);
        }
#pragma unroll
        for (int m = 0; m < HG; ++m) {
            const int p = c.i + S::TL * m;
            Cx<R>* u = &v[2 * m * r];
            if (p != 0) {
#pragma unroll
                for (int q = 0; q < r; ++q) emit(p + q * NB, u[q], u[r + (r - 1 - q)], true);
            } else {
                emit(0, u[0], u[0], true);                                         // bins 0 and N, both from Z[0]
#pragma unroll
                for (int q = 1; q < r / 2; ++q) emit(q * NB, u[q], u[r - q], true);
                emit((r / 2) * NB, u[r / 2], u[r / 2], true);                      // k = N/2 pairs with itself
#pragma unroll
                for (int q = 0; q < r / 2; ++q) emit(NB / 2 + q * NB, u[r + q], u[r + (r - 1 - q)], true);
            }
        }
        return;
    }
    if constexpr (PAIR_EPI && !COLS && !kMirrorEpi<S> && kShuffleEpi<S>) {
        constexpr int LP = S::NP - 1, r = S::radix(LP), P = S::before(LP), NB = S::nbf(LP);
        SfftUntil<R, S, L, COLS, 0, LP, STAGE_IN || PAIR_PRO>::run(c, v, tw, load, store);
        // deal the butterflies: warp w of the lane, lanes 0..15 take 16 w + t, lanes 16..31 take NB - 16 w - (t - 16)
        // (NB itself does not exist: lane 16 of warp 0 takes NB/2, which like 0 is its own mirror)
        const int w = c.i >> 5, t = c.i & 31;
        const bool self = (w == 0) && (t == 0 || t == 16);
        const int b = t < 16 ? 16 * w + t : (self ? NB / 2 : NB - 16 * w - (t - 16));
#pragma unroll
        for (int q = 0; q < r; ++q) v[q] = c.smem[c.addr(b + q * NB)];
        {
            const Cx<R>* __restrict__ t0 = tw + S::twoff(LP) + (b % P);
#pragma unroll
            for (int q = 1; q < r; ++q) v[q] = cmul(v[q], ldg(&t0[(q - 1) * P]));
        }
        Dft<R, r>::run(&v[0]);
        // Z[N - k] for k = b + q NB is the partner's output r-1-q (own output (r-q) % r for b = 0, r-1-q for b = NB/2)
        Cx<R> zm[r];
#pragma unroll
        for (int q = 0; q < r; ++q) {
            Cx<R> got;
            got.x = __shfl_xor_sync(0xffffffffu, v[r - 1 - q].x, 16);
            got.y = __shfl_xor_sync(0xffffffffu, v[r - 1 - q].y, 16);
            zm[q] = got;
        }
        if (self) {
#pragma unroll
            for (int q = 0; q < r; ++q) zm[q] = (b == 0) ? v[(r - q) % r] : v[r - 1 - q];
        }
#pragma unroll
        for (int q = 0; q < r; ++q) emit(b + q * NB, v[q], zm[q], false);
        if (b == 0) emit(0, v[0], v[0], true);        // also bin N (bin 0 is written twice with the same value)
        return;
    }
    if constexpr (MIRROR_OUT && !MIRROR_PRO) {
        constexpr int LP = S::NP - 1, r = S::radix(LP), P = S::before(LP), NB = S::nbf(LP), HG = S::G(LP) / 2;
        SfftUntil<R, S, L, COLS, 0, LP, STAGE_IN || PAIR_PRO>::run(c, v, tw, load, store);
#pragma unroll
        for (int m = 0; m < HG; ++m) {
            const int b0 = c.i + S::TL * m, b1 = NB - 1 - b0;
            Cx<R>* u = &v[2 * m * r];
#pragma unroll
            for (int q = 0; q < r; ++q) {
                u[q] = c.smem[c.addr(b0 + q * NB)];
                u[r + q] = c.smem[c.addr(b1 + q * NB)];
            }
            const Cx<R>* __restrict__ t0 = tw + S::twoff(LP) + (b0 % P);
            const Cx<R>* __restrict__ t1 = tw + S::twoff(LP) + (b1 % P);
#pragma unroll
            for (int q = 1; q < r; ++q) {
                u[q] = cmul(u[q], ldg(&t0[(q - 1) * P]));
                u[r + q] = cmul(u[r + q], ldg(&t1[(q - 1) * P]));
            }
            Dft<R, r>::run(&u[0]);
            Dft<R, r>::run(&u[r]);
        }
        if (!valid) return;
        auto put2 = [&](int t, R x0, R x1) { *reinterpret_cast<Cx<R>*>(out_r + t) = cmake<R>(x0, x1); };   // t even: aligned pair
        // four adjacent reals as ONE store when the rows are aligned for it (a.vec_out == 2): a whole 32-byte sector per thread in
        // f64 (st.global.v4.f64 = STG.E.ENL2.256 on sm_100a) instead of two half-sector stores from the same thread
        const bool wide = a.vec_out >= 2;
        auto put4 = [&](int t, R x0, R x1, R x2, R x3) {
#ifndef NDFB_EMU
            if (wide) {
                if constexpr (sizeof(R) == 8) {
                    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(out_r + t), "d"((double)x0), "d"((double)x1), "d"((double)x2), "d"((double)x3) : "memory");
                } else {
                    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out_r + t), "f"((float)x0), "f"((float)x1), "f"((float)x2), "f"((float)x3) : "memory");
                }
                return;
            }
#endif
            put2(t, x0, x1);
            put2(t + 2, x2, x3);
        };
#pragma unroll
        for (int m = 0; m < HG; ++m) {
            const int b0 = c.i + S::TL * m;
            const Cx<R>* u = &v[2 * m * r];
#pragma unroll
            for (int q = 0; q < r; ++q) {
                const int k = b0 + q * NB;                    // its mirror N-1-k is output r-1-q of butterfly NB-1-b0
                const Cx<R> yk = u[q], ym = u[r + (r - 1 - q)];
                if (KIND == RK_DCT3) {
                    // x[4 lo .. 4 lo + 3] = h (Re y_lo, -Im y_hi, -Im y_lo, Re y_hi), lo = min(k, N-1-k)  (k < N/2 <=> q < r/2)
                    const R h = (R)0.5 * sc;
                    const bool klo = q < r / 2;
                    const Cx<R> ylo = klo ? yk : ym, yhi = klo ? ym : yk;
                    const int t = 4 * (klo ? k : N - 1 - k);
                    put4(t, h * ylo.x, -h * yhi.y, -h * ylo.y, h * yhi.x);
                } else {
                    const Cx<R> Ck = cmul(yk, ldg(&tabB[k])), Cm = cmul(ym, ldg(&tabB[N - 1 - k]));
                    put2(2 * k, sc * Ck.x, -sc * Cm.y);                  // out[2k], out[2k+1] = out[n-1-2k']
                    put2(n - 2 - 2 * k, sc * Cm.x, -sc * Ck.y);          // out[2k'] = out[n-2-2k], out[n-1-2k]
                }
            }
        }
        return;
    }
    if constexpr (MIRROR_PRO) {
        // pass 0 with butterflies i and NB - i in one thread: the zip of bins j and N - j feeds both, straight from global memory
        constexpr int r = S::R0, NB = S::nbf(0);
        auto zip = [&](int j, Cx<R>& zj, Cx<R>& zm) {      // z[j] and z[N - j] (j = 0: zm unused; j = N/2: both equal)
            const int k2 = N - j;
            Cx<R> xk, xn;
            if (KIND == RK_C2R) {
                xk = valid ? in_c[(long long)j * is_axis] : cmake<R>(zero, zero);
                xn = valid ? in_c[(long long)k2 * is_axis] : cmake<R>(zero, zero);
                if (j == 0) { xk.y = zero; xn.y = zero; }        // Im X[0], Im X[N] dropped (src/lib.rs:516-521)
            } else {
                Cx<R> pk = cmake<R>(gin(j), j == 0 ? zero : -gin(n - j));
                Cx<R> pn = cmake<R>(gin(k2), -gin(n - k2));
                xk = cmul(pk, cconj(ldg(&tabB[j])));
                xn = cmul(pn, cconj(ldg(&tabB[k2])));
            }
            const Cx<R> wc = cconj(ldg(&tabA[j]));
            const Cx<R> E = cadd(xk, cconj(xn)), O = csub(xk, cconj(xn));
            const Cx<R> iw = cmul_i(cmul(wc, O));
            zj = cconj(cadd(E, iw));
            zm = csub(E, iw);
        };
        const int i = c.i;
        const int b0 = i, b1 = i == 0 ? NB / 2 : NB - i;
        Cx<R> dump;
        if (i != 0) {
#pragma unroll
            for (int q = 0; q < r; ++q) zip(b0 + q * NB, v[q], v[r + (r - 1 - q)]);
        } else {
            zip(0, v[0], dump);
#pragma unroll
            for (int q = 1; q < r / 2; ++q) zip(q * NB, v[q], v[r - q]);
            zip((r / 2) * NB, v[r / 2], dump);
#pragma unroll
            for (int q = 0; q < r / 2; ++q) zip(NB / 2 + q * NB, v[r + q], v[r + (r - 1 - q)]);
        }
        Dft<R, r>::run(&v[0]);
        Dft<R, r>::run(&v[r]);
#pragma unroll
        for (int q = 0; q < r; ++q) {
            c.smem[c.addr(b0 * r + q)] = v[q];          // pass 0 of the autosort: Y[b r + q]
            c.smem[c.addr(b1 * r + q)] = v[r + q];
        }
        __syncthreads();
        SfftAll<R, S, L, COLS, 1, false, (PAIR_EPI || STAGE_OUT) && (S::NP > 1)>::run(c, v, tw, load, store);
    } else {
        SfftAll<R, S, L, COLS, 0, STAGE_IN || PAIR_PRO, (PAIR_EPI || STAGE_OUT) && (S::NP > 1)>::run(c, v, tw, load, store);
    }

    if (PAIR_EPI) {
        __syncthreads();
        // bins 0..N of the length-2N real DFT from the packed N-point result.  Bins k and N-k need the same two slots
        // Z[k], Z[N-k] and conjugate-related factors (tabA[N-k] = -conj(tabA[k])), so one thread produces both: with
        // s = Z[k] + conj(Z[N-k]), d = w (Z[k] - conj(Z[N-k])):  X[k] = (s - i d)/2,  X[N-k] = (conj(s) - i conj(d))/2.
        // (k = 0 pairs with N, both from Z[0]; k = N/2 pairs with itself.)
        constexpr int ITP = (N / 2 + 1 + S::TL - 1) / S::TL;
#pragma unroll
        for (int m = 0; m < ITP; ++m) {
            const int k = c.i + m * S::TL;
            if (k > N / 2) break;
            const int k2 = N - k;
            const Cx<R> zk = c.smem[c.addr(k)];
            const Cx<R> zc = cconj(c.smem[c.addr(k == 0 ? 0 : k2)]);
            const Cx<R> w = ldg(&tabA[k]);
            const Cx<R> s = cadd(zk, zc), d = cmul(w, csub(zk, zc));
            const Cx<R> X = cmake<R>((R)0.5 * (s.x + d.y), (R)0.5 * (s.y - d.x));
            const Cx<R> X2 = cmake<R>((R)0.5 * (s.x - d.y), (R)-0.5 * (s.y + d.x));
            if (!valid) continue;
            const bool two = k2 != k;
            if (KIND == RK_R2C) {
                out_c[(long long)k * os_axis] = cmake<R>(sc * X.x, sc * X.y);
                if (two) out_c[(long long)k2 * os_axis] = cmake<R>(sc * X2.x, sc * X2.y);
            } else if (KIND == RK_DCT1) {
                out_r[(long long)k * os_axis] = (R)0.5 * sc * X.x;
                if (two) out_r[(long long)k2 * os_axis] = (R)0.5 * sc * X2.x;
            } else {  // RK_DCT2
                const Cx<R> A = cmul(X, ldg(&tabB[k]));
                out_r[(long long)k * os_axis] = sc * A.x;
                if (k > 0) out_r[(long long)(n - k) * os_axis] = -sc * A.y;      // (k <= N/2 < N)
                if (two) {
                    const Cx<R> A2 = cmul(X2, ldg(&tabB[k2]));
                    out_r[(long long)k2 * os_axis] = sc * A2.x;
                    if (k2 < N) out_r[(long long)(n - k2) * os_axis] = -sc * A2.y;
                }
            }
        }
    }
    if (STAGE_OUT) {
        __syncthreads();
        constexpr int ITO = (n + S::TL - 1) / S::TL;
        if (valid)
#pragma unroll
            for (int m = 0; m < ITO; ++m) {
                const int t = c.i + m * S::TL;
                if (n % S::TL != 0 && t >= n) break;
                // DCT-III: out[t] = v[p(t)] (inverse Makhoul);  DCT-IV: even t from the real parts, odd t from the imaginary ones
                const int pos = KIND == RK_DCT3 ? ((t & 1) ? n - 1 - (t >> 1) : (t >> 1)) : ((t & 1) ? n - t : t);
                out_r[(long long)t * os_axis] = sreal(pos);
            }
    }
}

// release of a finished tile to a consumer launch: every thread's stores are ordered before the counter update
NDFB_DEV void rsfft_signal(const RsfftArgs& a, int L) {
    if (!a.done_cnt) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const long long g0 = (long long)blockIdx.x * L;
        const long long left = a.nlanes - g0;
        atomicAdd(&a.done_cnt[g0 / a.done_group], (unsigned)(left < L ? left : L));
    }
}

template <typename R, class S, int L, bool COLS, int KIND, int MINB>
__global__ void __launch_bounds__(S::TL* L, MINB) rsfft_kernel(const __grid_constant__ RsfftArgs a) {
    if constexpr (!COLS) {
        if (a.is_axis == 1 && a.os_axis == 1) {
            if constexpr ((KIND == RK_DCT3 || KIND == RK_DCT4) && kMirrorOut<S>) {
                if (a.vec_out) { rsfft_body<R, S, L, COLS, KIND, true, true>(a); rsfft_signal(a, L); return; }
            }
            rsfft_body<R, S, L, COLS, KIND, true>(a); rsfft_signal(a, L); return;
        }
    }
    rsfft_body<R, S, L, COLS, KIND, false>(a);
    rsfft_signal(a, L);
}

}  // namespace ndfb
