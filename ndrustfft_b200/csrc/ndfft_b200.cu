// ndfft_b200.cu — C ABI (include/ndfft_b200.h), launch geometry and device plumbing.
//
// Built two ways from the same sources:
//   nvcc -gencode arch=compute_100a,code=sm_100a  -> ndrustfft_b200/lib/libndfft_b200.so   (THE product)
//   g++ -x c++ -DNDFB_EMU                         -> tests/emu/libndfft_b200_emu.so        (test-only SIMT emulation,
//                                                    never loaded by the package; see tests/emu/simt_emu.h)
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ndfft_b200.h"
#include "common.h"
#include "devapi.h"
#include "hostio.h"
#include "jit.h"
#include "plan.h"
#include "big_kernels.cuh"
#include "sfft_inst.h"
#include "pipe_inst.h"
#include "tile_kernel.cuh"

namespace ndfb {

static thread_local std::string g_err;
std::string& err_slot() { return g_err; }
static std::atomic<uint64_t> g_launches{0};
std::atomic<uint64_t>& launch_counter() { return g_launches; }
static thread_local size_t g_smem_floor = 0;
size_t& launch_smem_floor() { return g_smem_floor; }
// producer / consumer hints for the NEXT fast-path launch of this thread (ndfb_hint_next_launch_signal / _wait)
struct SyncHint {
    unsigned* signal_cnt = nullptr; long long signal_group = 0;
    const unsigned* wait_cnt = nullptr; long long wait_group = 0; unsigned wait_need = 0; int ctas_per_sm = 0;
    void clear() { *this = SyncHint(); }
};
static thread_local SyncHint g_sync_hint;
// a hint is for the NEXT call only: whatever way that call ends (argument error included), the hint must not leak into a later one
struct SyncHintScope { ~SyncHintScope() { g_sync_hint.clear(); } };
static thread_local bool g_trans_next = false;   // launch_c2c -> launch_sfft: run the entry as a transposing rows kernel

// ------------------------------------------------------------------------------------------------------
// plans
// ------------------------------------------------------------------------------------------------------
struct DeviceTables {
    void *tw = nullptr, *tabA = nullptr, *tabB = nullptr, *blu_c = nullptr, *blu_bhat = nullptr;
    uint32_t* perm = nullptr;
    void *fs_lo = nullptr, *fs_hi = nullptr;
    bool ready = false;
    std::map<uint32_t, void*> sfft_tw;  // per-pass Stockham twiddle tables, keyed by the radix schedule
};

struct Core {
    CoreTables t;
    DeviceTables d;
    std::mutex mu;
};

template <typename R>
static int upload_cx(void** dst, const std::vector<cld>& v) {
    if (v.empty()) { *dst = nullptr; return 0; }
    std::vector<Cx<R>> h(v.size());
    for (size_t i = 0; i < v.size(); ++i) { h[i].x = (R)v[i].real(); h[i].y = (R)v[i].imag(); }
    int rc = dev_malloc(dst, h.size() * sizeof(Cx<R>));
    if (rc) return rc;
    rc = dev_h2d(*dst, h.data(), h.size() * sizeof(Cx<R>), 0);
    if (rc) return rc;
    return dev_sync(0);
}

}  // namespace ndfb

struct ndfb_plan {
    int kind, dtype, device;
    size_t n;
    std::mutex mu;
    std::map<std::pair<int, int>, std::unique_ptr<ndfb::Core>> cores;  // (tile kind, n) -> schedule
    // four-step inter-pass twiddles, keyed by total length
    struct FsTw { void *lo = nullptr, *hi = nullptr; int shift = 0; long long ntot = 0; };
    std::map<long long, FsTw> fs;
    // pipelined kernels: W_N^{q NB j2} at [j2 r + q], keyed by (N, first factor, last radix of its schedule)
    std::map<std::pair<long long, std::pair<int, int>>, void*> fsq;
    // staged Bluestein (lengths with a large prime factor that do not fit one CTA): chirp c[j] and FFT_M(kernel)/M
    struct BigBlu { void *chirp = nullptr, *bhat = nullptr; int M = 0; };
    std::map<long long, BigBlu> bigblu;
};

namespace ndfb {

static Core* get_core(ndfb_plan* p, int tk, int n) {
    std::lock_guard<std::mutex> g(p->mu);
    auto key = std::make_pair(tk, n);
    auto it = p->cores.find(key);
    if (it != p->cores.end()) return it->second.get();
    std::unique_ptr<Core> c(new Core());
    build_core(c->t, tk, n);
    Core* raw = c.get();
    p->cores[key] = std::move(c);
    return raw;
}

template <typename R>
static int ensure_device(ndfb_plan* p, Core* c) {
    std::lock_guard<std::mutex> g(c->mu);
    if (c->d.ready) return 0;
    int rc = dev_set(p->device);
    if (rc) return rc;
    if ((rc = upload_cx<R>(&c->d.tw, c->t.tw))) return rc;
    if ((rc = upload_cx<R>(&c->d.tabA, c->t.tabA))) return rc;
    if ((rc = upload_cx<R>(&c->d.tabB, c->t.tabB))) return rc;
    if ((rc = upload_cx<R>(&c->d.blu_c, c->t.blu_c))) return rc;
    if ((rc = upload_cx<R>(&c->d.blu_bhat, c->t.blu_bhat))) return rc;
    if (!c->t.perm.empty()) {
        void* q = nullptr;
        if ((rc = dev_malloc(&q, c->t.perm.size() * 4))) return rc;
        if ((rc = dev_h2d(q, c->t.perm.data(), c->t.perm.size() * 4, 0))) return rc;
        if ((rc = dev_sync(0))) return rc;
        c->d.perm = (uint32_t*)q;
    }
    c->d.ready = true;
    return 0;
}

template <typename R>
static int get_fs_twiddles(ndfb_plan* p, long long Ntot, ndfb_plan::FsTw* out) {
    std::lock_guard<std::mutex> g(p->mu);
    auto it = p->fs.find(Ntot);
    if (it != p->fs.end()) { *out = it->second; return 0; }
    int shift = 0;
    while ((1LL << (2 * shift)) < Ntot) ++shift;  // lo table ~ sqrt(N)
    if (shift < 1) shift = 1;
    if (Ntot <= (1LL << 17)) shift = 40;          // short enough for ONE table W_N^e, e = k1*j2 < N: no hi/lo product
    const long long nlo = shift >= 40 ? Ntot : (1LL << shift), nhi = shift >= 40 ? 1 : ((Ntot - 1) >> shift) + 1;
    std::vector<cld> lo(nlo), hi(nhi);
    for (long long a = 0; a < nlo; ++a) lo[a] = unit_root(a, Ntot);
    for (long long b = 0; b < nhi; ++b) hi[b] = unit_root(shift >= 40 ? 0 : (b << shift), Ntot);
    ndfb_plan::FsTw t;
    t.shift = shift;
    t.ntot = Ntot;
    int rc;
    if ((rc = dev_set(p->device))) return rc;
    if ((rc = upload_cx<R>(&t.lo, lo))) return rc;
    if ((rc = upload_cx<R>(&t.hi, hi))) return rc;
    p->fs[Ntot] = t;
    *out = t;
    return 0;
}

// pipelined kernels: the q-dependent factor of the four-step twiddle, W_N^{q NB j2} for the last radix r of the N1-point
// schedule (NB = N1 / r) and every lane j2 < nj2, laid out [j2][q] (pipe_kernel.cuh: PipeStore)
template <typename R>
static int get_fs_q(ndfb_plan* p, long long Ntot, int N1, int r, long long nj2, void** out) {
    std::lock_guard<std::mutex> g(p->mu);
    const auto key = std::make_pair(Ntot, std::make_pair(N1, r));
    auto it = p->fsq.find(key);
    if (it != p->fsq.end()) { *out = it->second; return 0; }
    const long long NB = N1 / r;
    std::vector<cld> t((size_t)nj2 * r);
    for (long long j2 = 0; j2 < nj2; ++j2)
        for (int q = 0; q < r; ++q) t[(size_t)j2 * r + q] = unit_root((long long)(((unsigned long long)q * NB % Ntot) * j2 % Ntot), Ntot);
    int rc;
    void* d = nullptr;
    if ((rc = dev_set(p->device))) return rc;
    if ((rc = upload_cx<R>(&d, t))) return rc;
    p->fsq[key] = d;
    *out = d;
    return 0;
}

// column kernels: the lane-dependent factor of the four-step twiddle, W_N^{k l} for output index k < N1 and lane l < L of a
// tile, laid out [k][l] (sfft_kernel.cuh: SfftGStore MODE 4); keyed like the pipelined kernels' table with -L as third key
template <typename R>
static int get_fs_kl(ndfb_plan* p, long long Ntot, int N1, int L, void** out) {
    std::lock_guard<std::mutex> g(p->mu);
    const auto key = std::make_pair(Ntot, std::make_pair(N1, -L));
    auto it = p->fsq.find(key);
    if (it != p->fsq.end()) { *out = it->second; return 0; }
    std::vector<cld> t((size_t)N1 * L);
    for (long long k = 0; k < N1; ++k)
        for (int l = 0; l < L; ++l) t[(size_t)k * L + l] = unit_root(k * l % Ntot, Ntot);
    int rc;
    void* d = nullptr;
    if ((rc = dev_set(p->device))) return rc;
    if ((rc = upload_cx<R>(&d, t))) return rc;
    p->fsq[key] = d;
    *out = d;
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// workspaces (per host thread)
// ------------------------------------------------------------------------------------------------------
struct Pool {
    struct Slot { void* p = nullptr; size_t bytes = 0; int device = -1; };
    Slot slots[9];   // 0/1: host staging in/out, 2: four-step workspace, 3/4: staged-path rows, 5: nested four-step workspace,
                     // 6: tile counters of the fused two-pass kernel,
                     // 7/8: intermediates of ndfb_exec_chain
    // The workspaces belong to the host thread, not to a stream.  A call on stream B right after a call on stream A
    // would reuse a workspace A's kernels may still be using, so each device-memory call that took a workspace records
    // an event when it is done and the next call on a DIFFERENT stream waits for it first (same stream: stream order).
    stream_t cur = nullptr, last = nullptr;
    bool in_call = false, touched = false, have_event = false, waited = false;
#ifndef NDFB_EMU
    cudaEvent_t ev = nullptr;
#endif
    ~Pool() {}  // device memory is reclaimed at process exit; explicit release via ndfb_release_workspaces
    void begin_call(stream_t s) {
        cur = s; in_call = true; touched = false; waited = false;
#ifndef NDFB_EMU
        // inside a stream capture the caller owns the ordering (an event from outside may not be waited on there)
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(s, &cs) != cudaSuccess) cudaGetLastError();
        if (cs != cudaStreamCaptureStatusNone) in_call = false;
#endif
    }
    void end_call() {
#ifndef NDFB_EMU
        if (in_call && touched) {
            if (!ev && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { ev = nullptr; cudaGetLastError(); }
            if (ev && cudaEventRecord(ev, cur) == cudaSuccess) { have_event = true; last = cur; }
            else { cudaGetLastError(); have_event = false; }
        }
#endif
        in_call = false;
    }
    int get(int which, int device, size_t bytes, void** out) {
#ifndef NDFB_EMU
        if (in_call && !waited && have_event && last != cur) {
            if (cudaStreamWaitEvent(cur, ev, 0) != cudaSuccess) { cudaGetLastError(); cudaDeviceSynchronize(); }
            waited = true;
        }
#endif
        touched = true;
        Slot& s = slots[which];
        if (s.p && (s.bytes < bytes || s.device != device)) { dev_free(s.p); s.p = nullptr; s.bytes = 0; }
        if (!s.p) {
            int rc = dev_malloc(&s.p, bytes);
            if (rc) { s.p = nullptr; return rc; }
            s.bytes = bytes;
            s.device = device;
        }
        *out = s.p;
        return 0;
    }
    void release() {
        for (auto& s : slots) { if (s.p) dev_free(s.p); s.p = nullptr; s.bytes = 0; }
        have_event = false;
    }
};
static thread_local Pool g_pool;

// ------------------------------------------------------------------------------------------------------
// launch geometry
// ------------------------------------------------------------------------------------------------------
struct BDim { long long size, is, os; };

static long long llabs_(long long v) { return v < 0 ? -v : v; }
static int pow2floor(long long v) { int p = 1; while ((long long)p * 2 <= v) p *= 2; return p; }
static int pow2ceil(long long v) { int p = 1; while (p < v) p *= 2; return p; }
static int ilog2(int v) { int s = 0; while ((1 << s) < v) ++s; return s; }

struct LaunchSpec {
    // one tile-kernel launch over `dims` batch dims
    Core* core;
    const void* in;
    void* out;
    std::vector<BDim> dims;  // batch dims, fastest first (already ordered/merged)
    long long is_axis, os_axis;
    int conj_in = 0, conj_out = 0;
    double scale = 1.0;
    int fs_twiddle = 0;
    int fs_dim = 0;   // which batch dim carries the four-step index j2
    int os_blk = 0;   // split output axis (see TileArgs::os_blk)
    long long os_blk_stride = 0;
    int nblk_ptr = 0;
    void* blk_ptr[8] = {nullptr};
    ndfb_plan::FsTw fs;
    bool keep_dim_order = false;
    bool trans = false;   // last pass of a three-pass split: contiguous lanes in, lane-interleaved out (sfft_body_trans)
    int max_L = 0;    // != 0: widest tile allowed (transposing passes want 32/L points of a lane per warp >= one sector)
};

// ---- fast path lookup ----
// Which schedule family to prefer: 'A' (few passes, many registers) or 'B' (more CTAs per SM).  The default table
// below comes from same-run A/B measurements on B200 (tools/tune_variants.py, profiles/r1f_tune*.jsonl);
// NDFB_SFFT_FAMILY=A|B overrides it for measurements.
static int preferred_family(bool f64, int N, bool cols, bool real_kind) {
    if (const char* f = std::getenv("NDFB_SFFT_FAMILY")) return (f[0] == 'B' || f[0] == 'b') ? 1 : 0;
    // B won every same-run comparison except where its tile is narrower than a 32-byte sector per row (handled by
    // better_entry's first rule): c3 0.46 ms vs 0.52-0.56 ms, c2 rows 0.308 vs 0.329 ms, c4 rows 0.11-0.125 vs 0.126-0.145 ms.
    (void)f64; (void)N; (void)cols; (void)real_kind;
    return 1;
}

// resident threads per SM this entry can expect (registers estimated from the points per thread)
template <typename E>
static int resident_threads(const E* e) {
    const int regs = e->f64 ? e->E * 4 + (e->fam >= 1 ? 32 : 56) : e->E * 2 + (e->fam >= 1 ? 32 : 44);
    int ctas = 65536 / (e->threads * (regs < 32 ? 32 : regs));
    if (e->smem) ctas = std::min<int>(ctas, (int)((227 * 1024) / e->smem));
    ctas = std::max(ctas, e->minb);     // __launch_bounds__ guarantees at least this many (register-capped variants)
    if (e->smem) ctas = std::min<int>(ctas, (int)((227 * 1024) / e->smem));
    ctas = std::min(ctas, 2048 / e->threads);
    ctas = std::max(1, std::min(ctas, 32));
    return ctas * e->threads;
}

template <typename E>
static bool better_entry(const E* e, const E* best, long long nlanes, int pref_fam, size_t lane_elem_bytes, long long axis_stride_bytes) {
    if (!best) return true;
    // 0. strided tiles must cover at least one 32-byte sector per row
    if (e->cols) {
        const bool e_sec = (size_t)e->L * lane_elem_bytes >= 32, b_sec = (size_t)best->L * lane_elem_bytes >= 32;
        if (e_sec != b_sec) return e_sec;
    }
    // 1. a tile not much wider than the batch, 2. the preferred family
    const bool e_fit = e->L <= 2 * nlanes, b_fit = best->L <= 2 * nlanes;
    if (e_fit != b_fit) return e_fit;
    const bool e_f = e->fam == pref_fam, b_f = best->fam == pref_fam;
    if (e_f != b_f) return e_f;
    // 3. rows more than ~1 MiB apart (one 2 MiB page per few rows): the widest tile amortises the TLB misses
    //    (c3 axis 0, L = 4 vs 8: 0.63 vs 0.47 ms); otherwise occupancy decides (1000-point f64 columns, L = 2 vs 4:
    //    0.528 vs 0.453 of the roofline), ties go to the wider tile
    if (e->cols && axis_stride_bytes >= (1 << 20)) {
        if (e->L != best->L) return e->L > best->L;
    }
    const int er = resident_threads(e), br = resident_threads(best);
    if (er != br) return er > br;
    return e->L > best->L;
}

static const SfftEntry* find_sfft(bool f64, int N, bool cols, long long nlanes, long long axis_stride_bytes = 0, int max_L = 0) {
    static const bool disabled = std::getenv("NDFB_DISABLE_SFFT") != nullptr;
    if (disabled) return nullptr;
    const SfftEntry* tabs[4] = {kSfft_f32_rows, kSfft_f32_cols, kSfft_f64_rows, kSfft_f64_cols};
    const int counts[4] = {kSfft_f32_rows_count, kSfft_f32_cols_count, kSfft_f64_rows_count, kSfft_f64_cols_count};
    const int which = (f64 ? 2 : 0) + (cols ? 1 : 0);
    // measurement hook: NDFB_SFFT_PICK=<N>:<index> takes the index-th registry entry of that length
    if (const char* pick = std::getenv("NDFB_SFFT_PICK")) {
        int pn = 0, pi = 0;
        if (sscanf(pick, "%d:%d", &pn, &pi) == 2 && pn == N) {
            int seen = 0;
            for (int i = 0; i < counts[which]; ++i)
                if (tabs[which][i].N == N && seen++ == pi) return &tabs[which][i];
        }
    }
    const int pref = preferred_family(f64, N, cols, false);
    const SfftEntry* best = nullptr;
    for (int i = 0; i < counts[which]; ++i) {
        const SfftEntry* e = &tabs[which][i];
        if (e->N != N) continue;
        if (max_L && e->L > max_L) continue;
        if (better_entry(e, best, nlanes, pref, f64 ? (size_t)16 : (size_t)8, axis_stride_bytes)) best = e;
    }
    return best;
}

template <typename R>
static int launch_sfft(ndfb_plan* p, const SfftEntry* e, const LaunchSpec& s, stream_t stream);

// ---- run-time compiled schedules (csrc/jit.h): every smooth length without an instantiated schedule ----
#ifndef NDFB_EMU
static bool jit_enabled() {
    static const bool off = std::getenv("NDFB_NO_JIT") != nullptr || std::getenv("NDFB_DISABLE_SFFT") != nullptr;
    return !off;
}
static bool jit_warned_once(const char* what) {
    static std::mutex mu;
    static std::map<std::string, int> seen;
    std::lock_guard<std::mutex> g(mu);
    return seen[what]++ > 0;
}
static void jit_set_kind(SfftEntry*, int) {}
static void jit_set_kind(RsfftEntry* e, int kind) { e->kind = kind; }
// C2C entry (kind < 0) or real-kind entry; the returned pointers live for the process
template <class Entry>
static const Entry* jit_entry(bool f64, int kind, int N, bool cols, long long nlanes) {
    if (!jit_enabled()) return nullptr;
    static std::mutex mu;
    static std::map<std::string, std::unique_ptr<Entry>> reg;
    JitSched js;
    if (!jit_plan(N, f64, cols, kind >= 0, nlanes, &js)) return nullptr;
    char key[160];
    int dev = 0;
    cudaGetDevice(&dev);
    snprintf(key, sizeof key, "%d:%d:%d:%d:%d:%d:%d:%d:%d:%d:%d:%d", dev, (int)f64, kind, N, js.TL, js.r[0], js.r[1], js.r[2], js.r[3], js.L, js.cols, js.minb);
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = reg.find(key);
        if (it != reg.end()) return it->second.get();
    }
    void* func = nullptr;
    const int rc = jit_get_kernel(f64, kind, js, &func);
    if (rc) {
        if (!jit_warned_once("jit")) fprintf(stderr, "[ndfb] note: run-time schedule compilation unavailable (%s); such lengths use the general kernel\n", err_slot().c_str());
        err_slot().clear();
        return nullptr;
    }
    std::unique_ptr<Entry> e(new Entry());
    std::memset((void*)e.get(), 0, sizeof(Entry));
    e->f64 = f64 ? 1 : 0; e->N = N; e->cols = js.cols; e->L = js.L; e->threads = js.threads; e->E = js.E; e->minb = js.minb; e->fam = 1;
    for (int i = 0; i < 4; ++i) e->r[i] = js.r[i];
    e->twtotal = js.twtotal; e->smem = js.smem; e->launch = nullptr; e->jit_func = func;
    jit_set_kind(e.get(), kind);
    std::lock_guard<std::mutex> g(mu);
    auto& slot = reg[key];
    if (!slot) slot = std::move(e);
    return slot.get();
}
#else
template <class Entry>
static const Entry* jit_entry(bool, int, int, bool, long long) { return nullptr; }
template <typename A>
static int jit_launch(void*, const A&, unsigned, unsigned, size_t, stream_t) { return fail(NDFB_E_UNSUPPORTED, "no run-time compilation in the emulation build"); }
#endif


template <typename R>
static int launch_tile(ndfb_plan* p, const LaunchSpec& s, stream_t stream, std::string* describe = nullptr) {
    Core* c = s.core;
    const CoreTables& t = c->t;
    TileArgs a;
    std::memset(&a, 0, sizeof a);
    a.in = s.in; a.out = s.out;
    a.kind = t.kind; a.n = t.n; a.n_in = t.n_in; a.n_out = t.n_out; a.N = t.N; a.M = t.M;
    a.is_axis = s.is_axis; a.os_axis = s.os_axis;
    a.conj_in = s.conj_in; a.conj_out = s.conj_out;
    a.scale = s.scale;
    a.npass = (int)t.radix.size();
    if (a.npass > kMaxPass) return fail(NDFB_E_UNSUPPORTED, "too many radix passes (%d)", a.npass);
    for (int i = 0; i < a.npass; ++i) a.radix[i] = t.radix[i];
    a.pad_shift = 31;

    long long nlanes = 1;
    for (auto& d : s.dims) nlanes *= d.size;
    if (nlanes == 0 || t.n_out == 0) return 0;
    a.nlanes = nlanes;
    if ((int)s.dims.size() > kMaxBatchDims) return fail(NDFB_E_UNSUPPORTED, "internal: too many batch dims");
    a.nbd = (int)s.dims.size();
    for (int d = 0; d < a.nbd; ++d) { a.bsz[d] = s.dims[d].size; a.bis[d] = s.dims[d].is; a.bos[d] = s.dims[d].os; }

    const size_t in_elem = (t.in_complex ? 2 : 1) * sizeof(R), out_elem = (t.out_complex ? 2 : 1) * sizeof(R);
    const size_t cs = sizeof(Cx<R>);
    const bool have_batch = a.nbd > 0 && nlanes > 1;
    a.in_lane_fast = have_batch && t.n_in > 1 && llabs_(a.bis[0]) < llabs_(s.is_axis);
    a.out_lane_fast = have_batch && t.n_out > 1 && llabs_(a.bos[0]) < llabs_(s.os_axis);
    const bool strided = a.in_lane_fast || a.out_lane_fast;

    const size_t cap = dev_smem_cap(p->device) - 1024;
    const size_t lane_bytes = (size_t)t.Bl * cs;
    const size_t per_lane_over = 2 * sizeof(long long) + sizeof(int);
    long long Lmax = (long long)(cap / (lane_bytes + per_lane_over));
    if (Lmax < 1) return fail(NDFB_E_UNSUPPORTED, "lane of %d complex points does not fit in shared memory", t.Bl);
    Lmax = pow2floor(Lmax);
    long long Lwant;
    if (strided) {
        size_t minelem = std::min(in_elem, out_elem);
        Lwant = std::max<long long>(1, 128 / (long long)minelem);
        // keep the tile modest when the lanes are long, but never below one 32-byte sector per row
        while (Lwant * lane_bytes > 64 * 1024 && Lwant * minelem > 32) Lwant /= 2;
    } else {
        Lwant = pow2floor(std::max<long long>(1, (long long)(32 * 1024 / cs) / t.Bl));
        const int sms = dev_sm_count(p->device);
        while (Lwant > 1 && (nlanes + Lwant - 1) / Lwant < 2LL * sms) Lwant /= 2;
    }
    long long L = std::min<long long>(std::min(Lwant, Lmax), pow2ceil(nlanes));
    if (L > 512) L = 512;
    a.L = (int)L;
    a.log2L = ilog2(a.L);
    a.Bl = t.Bl;
    if (strided && a.L > 1) { a.LP = 1; a.EP = a.L; }
    else { a.LP = t.Bl; a.EP = 1; }
    long long work = (long long)a.L * std::max(t.B, 1);
    int T = pow2ceil((work + 7) / 8);
    T = std::max(64, std::min(512, T));
    if (T < a.L) T = std::min(512, a.L);

    a.tw = c->d.tw; a.tabA = c->d.tabA; a.tabB = c->d.tabB; a.perm = c->d.perm;
    a.blu_c = c->d.blu_c; a.blu_bhat = c->d.blu_bhat;
    a.fs_twiddle = s.fs_twiddle; a.fs_dim = s.fs_dim;
    a.fs_shift = s.fs.shift; a.fs_lo = s.fs.lo; a.fs_hi = s.fs.hi;
    a.os_blk = s.os_blk; a.os_blk_stride = s.os_blk_stride;

    size_t smem = (((size_t)a.L * per_lane_over + 15) & ~(size_t)15) + (size_t)a.L * lane_bytes;
    const long long grid = (nlanes + a.L - 1) / a.L;
    if (grid > 0x7fffffffLL) return fail(NDFB_E_UNSUPPORTED, "too many tiles");
    if (describe) {
        char buf[256];
        snprintf(buf, sizeof buf, "{\"kernel\":\"tile_kernel<%s,%s>\",\"L\":%d,\"threads\":%d,\"smem\":%zu,\"grid\":%lld,\"layout\":\"%s\"}",
                 sizeof(R) == 4 ? "f32" : "f64", t.M ? "bluestein" : "direct", a.L, T, smem, grid, a.EP == 1 ? "lane-major" : "interleaved");
        *describe = buf;
        return 0;
    }
    if (t.M) return dev_launch(tile_kernel<R, true>, (unsigned)grid, (unsigned)T, smem, stream, a);
    return dev_launch(tile_kernel<R, false>, (unsigned)grid, (unsigned)T, smem, stream, a);
}

// W_{P r}^{q k} for every pass p >= 1 of the schedule, in Sched::twoff layout
template <typename R>
static int get_sfft_twiddles(ndfb_plan* p, Core* c, const SfftEntry* e, void** out) {
    std::lock_guard<std::mutex> g(c->mu);
    const uint32_t key = (uint32_t)e->r[0] | ((uint32_t)e->r[1] << 8) | ((uint32_t)e->r[2] << 16) | ((uint32_t)e->r[3] << 24);
    auto it = c->d.sfft_tw.find(key);
    if (it != c->d.sfft_tw.end()) { *out = it->second; return 0; }
    std::vector<cld> t;
    long long P = e->r[0];
    for (int ps = 1; ps < 4 && e->r[ps] > 1; ++ps) {
        const int r = e->r[ps];
        for (int q = 1; q < r; ++q)
            for (long long k = 0; k < P; ++k) t.push_back(unit_root((long long)q * k, P * r));
        P *= r;
    }
    if ((int)t.size() != e->twtotal) return fail(NDFB_E_INVALID, "internal: twiddle table size mismatch (%zu vs %d)", t.size(), e->twtotal);
    if (t.empty()) t.push_back(cld(1, 0));
    void* d = nullptr;
    int rc = dev_set(p->device);
    if (rc) return rc;
    if ((rc = upload_cx<R>(&d, t))) return rc;
    c->d.sfft_tw[key] = d;
    *out = d;
    return 0;
}

template <typename R>
static int launch_sfft(ndfb_plan* p, const SfftEntry* e, const LaunchSpec& s, stream_t stream) {
    const bool trans_next = g_trans_next;
    g_trans_next = false;
    SfftArgs a;
    std::memset(&a, 0, sizeof a);
    a.in = s.in; a.out = s.out;
    long long nlanes = 1;
    for (auto& d : s.dims) nlanes *= d.size;
    if (nlanes == 0) return 0;
    a.nlanes = nlanes;
    a.nbd = (int)s.dims.size();
    for (int d = 0; d < a.nbd; ++d) { a.bsz[d] = s.dims[d].size; a.bis[d] = s.dims[d].is; a.bos[d] = s.dims[d].os; }
    a.is_axis = s.is_axis; a.os_axis = s.os_axis;
    a.conj_in = s.conj_in; a.conj_out = s.conj_out; a.scale = s.scale;
    {
        void* twd = nullptr;
        int rc = get_sfft_twiddles<R>(p, s.core, e, &twd);
        if (rc) return rc;
        a.tw = twd;
    }
    a.fs_twiddle = s.fs_twiddle; a.fs_dim = s.fs_dim; a.fs_shift = s.fs.shift; a.fs_lo = s.fs.lo; a.fs_hi = s.fs.hi;
    a.os_blk = s.os_blk; a.os_blk_stride = s.os_blk_stride;
    a.nblk_ptr = s.nblk_ptr;
    for (int i = 0; i < 8; ++i) a.blk_ptr[i] = s.blk_ptr[i];
    a.trans_store = trans_next ? 1 : 0;
    // Column tiles whose lanes are consecutive j2 (the lane dim IS the twiddle dim and holds whole tiles): the four-step twiddle is
    // formed as (tile-uniform hi/lo lookup) x (coalesced [k][l] table) instead of one scattered lookup per point (MODE 4).
    // (whole tiles only: every thread of the CTA derives the same tile base j20 from its own lane, and the tile's factor needs N more
    // elements of shared memory behind the exchange buffer, which must still fit the SM)
    if (s.fs_twiddle && e->cols && s.fs_dim == 0 && !s.os_blk && !s.nblk_ptr && !s.dims.empty() && s.dims[0].size % e->L == 0 && s.fs.ntot &&
        e->smem + (size_t)e->N * sizeof(Cx<R>) <= dev_smem_cap(p->device) && !std::getenv("NDFB_NO_FS_FACTORED")) {
        void* tq = nullptr;
        int rc = get_fs_kl<R>(p, s.fs.ntot, e->N, e->L, &tq);
        if (rc) return rc;
        a.fs_q = tq;
        // room behind the exchange buffer for the tile's own factor W_N^{k j20}, k < N1 (looked up once per tile)
        launch_smem_floor() = std::max(launch_smem_floor(), e->smem + (size_t)e->N * sizeof(Cx<R>));
        if (std::getenv("NDFB_TRACE")) fprintf(stderr, "[ndfb] fs twiddle factored: tile-uniform lookup x [k][l] table (%d x %d)\n", e->N, e->L);
    }
    // long contiguous rows: every CTA asks the L2 for the row that will be started when it retires (NDFB_L2_PREFETCH=<waves>, 0 = off)
    // Measured on B200 (tools/ab_l2_prefetch.py, profiles/round2/r2k_ab_l2_prefetch.jsonl): half a wave ahead gains 2-4 % on 32-64 KiB
    // rows (8192-point c64: 0.255 -> 0.245 ms), one or two waves ahead lose (the lines are evicted or fight the demand loads).
    if (!e->cols && s.is_axis == 1 && !trans_next && (size_t)e->N * sizeof(Cx<R>) >= 32768 && ((uintptr_t)s.in % 16) == 0) {
        static const double waves = std::getenv("NDFB_L2_PREFETCH") ? atof(std::getenv("NDFB_L2_PREFETCH")) : 0.5;
        bool even = true;
        if (sizeof(R) == 4) for (auto& d : s.dims) if (d.is % 2) even = false;
        if (waves > 0 && even) a.l2_prefetch_lanes = (long long)(waves * dev_sm_count(p->device) * std::max(1, e->minb) * e->L);
    }
    // Scattered blocks that are contiguous per (tile, destination) go out as bulk-async copies (sfft_body MODE 3): column tile
    // over the fastest batch dim, that dim contiguous in the output and a whole number of tiles long, rows of a block L apart.
    if (s.nblk_ptr && e->cols && e->r[1] > 1 && e->N >= 64 && e->N <= 2048 && !s.fs_twiddle && !s.dims.empty() && s.dims[0].os == 1 && s.dims[0].size % e->L == 0 &&
        s.os_axis == e->L && (size_t)s.os_blk * e->L * sizeof(Cx<R>) % 16 == 0 && !std::getenv("NDFB_NO_BULK_STORE")) {
        bool aligned = true;
        for (int i = 0; i < s.nblk_ptr; ++i) if ((uintptr_t)s.blk_ptr[i] % 16) aligned = false;
        for (size_t d = 1; d < s.dims.size(); ++d) if ((s.dims[d].os * (long long)sizeof(Cx<R>)) % 16) aligned = false;
        a.bulk_store = aligned ? 1 : 0;
    }
    if (a.bulk_store && std::getenv("NDFB_TRACE")) fprintf(stderr, "[ndfb] scatter blocks as bulk-async copies: %d x %zu bytes per tile\n", s.nblk_ptr, (size_t)s.os_blk * e->L * sizeof(Cx<R>));
    long long grid = (nlanes + e->L - 1) / e->L;
    if (grid > 0x7fffffffLL) return fail(NDFB_E_UNSUPPORTED, "too many tiles");
    if (g_sync_hint.wait_cnt) {
        // consumer of another launch's output: persistent CTAs that wait for their tile's producer group
        const SyncHint h = g_sync_hint;
        g_sync_hint.clear();
        if (!(e->cols && e->r[1] > 1 && e->N >= 64 && e->N <= 2048) || s.fs_twiddle || h.wait_group % e->L)
            return fail(NDFB_E_UNSUPPORTED, "launch wait hint: this transform has no persistent consumer kernel");
        a.wait_cnt = h.wait_cnt; a.wait_group = h.wait_group; a.wait_need = h.wait_need;
        a.ntiles = grid;
        grid = std::min<long long>(grid, (long long)dev_sm_count(p->device) * std::max(1, h.ctas_per_sm));
        if (std::getenv("NDFB_TRACE")) fprintf(stderr, "[ndfb] persistent consumer launch: %lld tiles on %lld CTAs\n", a.ntiles, grid);
    }
    if (e->jit_func) return jit_launch(e->jit_func, a, (unsigned)grid, (unsigned)e->threads, e->smem, stream);
    return e->launch(a, (unsigned)grid, stream);
}

template <typename R>
static int get_big_bluestein(ndfb_plan* p, long long N, ndfb_plan::BigBlu* out);

static const BsfftEntry* find_bsfft(bool f64, int M, bool cols, long long nlanes, long long axis_stride_bytes) {
    static const bool disabled = std::getenv("NDFB_DISABLE_SFFT") != nullptr;
    if (disabled || std::getenv("NDFB_DISABLE_BSFFT")) return nullptr;
    const BsfftEntry* tab = f64 ? kBsfft_f64 : kBsfft_f32;
    const int count = f64 ? kBsfft_f64_count : kBsfft_f32_count;
    const BsfftEntry* best = nullptr;
    for (int i = 0; i < count; ++i) {
        const BsfftEntry* e = &tab[i];
        if (e->M != M || e->cols != (cols ? 1 : 0)) continue;
        if (better_entry(e, best, nlanes, 1, f64 ? (size_t)16 : (size_t)8, axis_stride_bytes)) best = e;
    }
    return best;
}

// NDFB_ROWS_BULK=1 selects the persistent bulk-async row kernels (A/B against the register-resident ones; default below)
static int rows_bulk_enabled() {   // 0 off, 1 on for batches that fill the GPU, 2 forced (tests)
    const char* e = std::getenv("NDFB_ROWS_BULK");
    return e ? atoi(e) : 0;
}

// Software-pipelined persistent column kernel (pipe_kernel.cuh) for tiles that own a whole SM: NDFB_PIPE=0 (default) off, 1 where
// the register-resident kernel would run one CTA per SM and the batch gives every SM several tiles, 2 wherever an instance exists
// (tests).  *done = 1 when it has launched.  Measured on B200 (profiles/round2/r2r_ab_pipe.jsonl): bit-identical results, the same
// speed within +-6 % (c5b 11.28 vs 11.30 ms; 2048-point c128 columns 0.337 vs 0.359 ms; 4096-point c64 columns 0.448 vs 0.421 ms),
// so the register-resident kernels stay the default: what bounds these tiles is the barrier-separated phases of the ONE CTA an
// SM can hold (a 128 KB tile is half the register file), not exposed load latency.
static int pipe_mode() {
    const char* e = std::getenv("NDFB_PIPE");
    return e ? atoi(e) : 0;
}
template <typename R>
static int try_pipe(ndfb_plan* p, const SfftEntry* e, const LaunchSpec& s, long long nlanes, stream_t stream, int* done) {
    *done = 0;
    const int mode = pipe_mode();
    if (!mode || s.os_blk || s.nblk_ptr || s.trans || g_sync_hint.wait_cnt || s.dims.empty() || (int)s.dims.size() > kMaxBatchDims) return 0;
    if (mode < 2 && !(e && e->cols && e->smem * 2 > (size_t)(227 * 1024))) return 0;   // rows: only on request (A/B runs, tests)
    const size_t cs = sizeof(Cx<R>);
    if ((uintptr_t)s.in % 16) return 0;
    const int N = s.core->t.N;
    const int sms = dev_sm_count(p->device);
    // contiguous rows in AND out take the one-lane tiles (the tile is one row: no lane interleaving to pay for), every other
    // layout the widest tile (full sectors per tile row)
    const bool rows_io = s.is_axis == 1 && s.os_axis == 1;
    const SfftPipeEntry* pe = nullptr;
    for (int i = 0; i < kSfftPipe_count; ++i) {
        const SfftPipeEntry* c = &kSfftPipe[i];
        if (c->f64 != (sizeof(R) == 8 ? 1 : 0) || c->N != N) continue;
        if ((c->L == 1) != rows_io) continue;
        if (const char* v = std::getenv("NDFB_PIPE_L")) { if (c->L != atoi(v)) continue; }   // measurement hook: tile width
        if (nlanes % c->L || s.dims[0].size % c->L) continue;              // every tile is full and lies in one row of the fastest batch dim
        bool ok = true;
        if (c->inmode == 0) {
            // L adjacent lanes = one 16-byte-aligned piece of every tile row
            if (s.dims[0].is != 1 || (c->L * cs) % 16 || (size_t)llabs_(s.is_axis) * cs % 16) ok = false;
            for (size_t d = 1; d < s.dims.size(); ++d) if ((size_t)llabs_(s.dims[d].is) * cs % 16) ok = false;
        } else {
            if (s.is_axis != 1 || ((size_t)N * cs) % 16) ok = false;
            for (auto& d : s.dims) if ((size_t)llabs_(d.is) * cs % 16) ok = false;
        }
        if (!ok) continue;
        if (!pe || c->L > pe->L) pe = c;
    }
    if (!pe) return 0;
    const long long ntiles = nlanes / pe->L;
    if (ntiles > 0x7fffffffLL) return 0;
    if (mode < 2 && ntiles < 4LL * sms * std::max(1, pe->minb)) return 0;
    SfftArgs a;
    std::memset(&a, 0, sizeof a);
    a.in = s.in; a.out = s.out; a.nlanes = nlanes; a.nbd = (int)s.dims.size();
    for (int d = 0; d < a.nbd; ++d) { a.bsz[d] = s.dims[d].size; a.bis[d] = s.dims[d].is; a.bos[d] = s.dims[d].os; }
    a.is_axis = s.is_axis; a.os_axis = s.os_axis; a.conj_in = s.conj_in; a.conj_out = s.conj_out; a.scale = s.scale;
    a.fs_twiddle = s.fs_twiddle; a.fs_dim = s.fs_dim; a.fs_shift = s.fs.shift; a.fs_lo = s.fs.lo; a.fs_hi = s.fs.hi;
    a.ntiles = ntiles;
    if (s.fs_twiddle) {
        int rl = pe->r[0];
        for (int k = 1; k < 4; ++k) if (pe->r[k] > 1) rl = pe->r[k];
        if (!s.fs.ntot || s.fs_dim < 0 || s.fs_dim >= (int)s.dims.size()) return 0;
        void* fq = nullptr;
        int rcq = get_fs_q<R>(p, s.fs.ntot, pe->N, rl, s.dims[s.fs_dim].size, &fq);
        if (rcq) return rcq;
        a.fs_q = fq;
    }
    SfftEntry proxy;
    std::memset((void*)&proxy, 0, sizeof proxy);
    for (int k = 0; k < 4; ++k) proxy.r[k] = pe->r[k];
    proxy.twtotal = pe->twtotal;
    void* twd = nullptr;
    int rc = get_sfft_twiddles<R>(p, s.core, &proxy, &twd);
    if (rc) return rc;
    a.tw = twd;
    const long long grid = std::min<long long>(ntiles, (long long)sms * std::max(1, pe->minb));
    if (std::getenv("NDFB_TRACE"))
        fprintf(stderr, "[ndfb] sfft %s N=%d %s pipelined (cp.async staging, split exchange) L=%d T=%d smem=%zu in=%s tiles=%lld grid=%lld radix=%d.%d.%d.%d\n",
                sizeof(R) == 8 ? "f64" : "f32", pe->N, rows_io ? "rows" : "cols", pe->L, pe->threads, pe->smem, pe->inmode ? "rows" : "lane-adjacent", ntiles, grid, pe->r[0], pe->r[1], pe->r[2], pe->r[3]);
    *done = 1;
    return pe->launch(a, (unsigned)grid, stream);
}

// C2C launch: the instantiated Stockham schedule when there is one, else the general tile kernel
template <typename R>
static int launch_c2c(ndfb_plan* p, const LaunchSpec& s, stream_t stream) {
    const CoreTables& t = s.core->t;
    if (t.kind == TK_C2C && t.M == 0 && (int)s.dims.size() <= kMaxBatchDims) {
        long long nlanes = 1;
        for (auto& d : s.dims) nlanes *= d.size;
        const bool have_batch = !s.dims.empty() && nlanes > 1;
        const bool cols = have_batch && t.N > 1 &&
                          (llabs_(s.dims[0].is) < llabs_(s.is_axis) || llabs_(s.dims[0].os) < llabs_(s.os_axis));
        if (s.trans && s.is_axis == 1 && !s.dims.empty() && s.dims[0].os == 1 && !s.fs_twiddle && !s.os_blk) {
            const SfftEntry* et = find_sfft(sizeof(R) == 8, t.N, false, nlanes);
            if (!et && nlanes > 0) et = jit_entry<SfftEntry>(sizeof(R) == 8, -1, t.N, false, nlanes);
            if (et && et->r[1] > 1 && et->L > 1) {
                if (std::getenv("NDFB_TRACE")) fprintf(stderr, "[ndfb] sfft %s N=%d rows->lanes (transposing) L=%d T=%d lanes=%lld\n", sizeof(R) == 8 ? "f64" : "f32", et->N, et->L, et->threads, nlanes);
                LaunchSpec st = s;
                st.trans = false;
                launch_smem_floor() = et->smem + (size_t)et->L * sizeof(Cx<R>) * 16;   // room for the padded lane pitch
                g_trans_next = true;
                return launch_sfft<R>(p, et, st, stream);
            }
        }
        const SfftEntry* e = find_sfft(sizeof(R) == 8, t.N, cols, nlanes, std::max(llabs_(s.is_axis), llabs_(s.os_axis)) * (long long)sizeof(Cx<R>), s.max_L);
        if (!e && s.max_L) e = find_sfft(sizeof(R) == 8, t.N, cols, nlanes, std::max(llabs_(s.is_axis), llabs_(s.os_axis)) * (long long)sizeof(Cx<R>));
        if (!e && nlanes > 0) e = jit_entry<SfftEntry>(sizeof(R) == 8, -1, t.N, cols, nlanes);
        // contiguous rows: the persistent bulk-async (TMA + mbarrier) kernel when there is one for this length
        if (!cols && s.is_axis == 1 && s.os_axis == 1 && !s.fs_twiddle && !s.os_blk && !s.nblk_ptr && rows_bulk_enabled()) {
            const SfftBulkEntry* be = nullptr;
            for (int i = 0; i < kSfftBulk_count; ++i)
                if (kSfftBulk[i].f64 == (sizeof(R) == 8 ? 1 : 0) && kSfftBulk[i].N == t.N) be = &kSfftBulk[i];
            bool ok = be && (nlanes >= 2LL * be->L * dev_sm_count(p->device) || rows_bulk_enabled() >= 2) && ((uintptr_t)s.in % 16) == 0 &&
                      ((uintptr_t)s.out % 16) == 0;
            if (ok && sizeof(R) == 4)
                for (auto& d : s.dims) if ((d.is % 2) || (d.os % 2)) ok = false;    // every row starts on a 16-byte boundary
            if (ok) {
                SfftArgs a;
                std::memset(&a, 0, sizeof a);
                a.in = s.in; a.out = s.out; a.nlanes = nlanes; a.nbd = (int)s.dims.size();
                for (int d = 0; d < a.nbd; ++d) { a.bsz[d] = s.dims[d].size; a.bis[d] = s.dims[d].is; a.bos[d] = s.dims[d].os; }
                a.is_axis = 1; a.os_axis = 1; a.conj_in = s.conj_in; a.conj_out = s.conj_out; a.scale = s.scale;
                SfftEntry proxy;
                std::memset(&proxy, 0, sizeof proxy);
                for (int i = 0; i < 4; ++i) proxy.r[i] = be->r[i];
                proxy.twtotal = be->twtotal;
                void* twd = nullptr;
                int rc = get_sfft_twiddles<R>(p, s.core, &proxy, &twd);
                if (rc) return rc;
                a.tw = twd;
                const long long ntiles = (nlanes + be->L - 1) / be->L;
                const long long grid = std::min<long long>(ntiles, (long long)dev_sm_count(p->device) * be->minb);
                if (std::getenv("NDFB_TRACE")) fprintf(stderr, "[ndfb] sfft %s N=%d rows bulk-async persistent L=%d T=%d smem=%zu tiles=%lld grid=%lld\n", sizeof(R) == 8 ? "f64" : "f32", be->N, be->L, be->threads, be->smem, ntiles, grid);
                return be->launch(a, (unsigned)grid, stream);
            }
        }
        if ((cols || (s.is_axis == 1 && s.os_axis == 1)) && nlanes > 0 && !s.dims.empty()) {
            int done = 0;
            const int rc = try_pipe<R>(p, e, s, nlanes, stream, &done);
            if (rc || done) return rc;
        }
        if (e) {
            const bool trace = std::getenv("NDFB_TRACE") != nullptr;
            if (trace) fprintf(stderr, "[ndfb] sfft %s N=%d %s L=%d T=%d smem=%zu lanes=%lld minb=%d fam=%c%s radix=%d.%d.%d.%d\n", sizeof(R) == 8 ? "f64" : "f32", e->N, e->cols ? "cols" : "rows", e->L, e->threads, e->smem, nlanes, e->minb, e->fam ? 'B' : 'A', e->jit_func ? " jit" : "", e->r[0], e->r[1], e->r[2], e->r[3]);
            return launch_sfft<R>(p, e, s, stream);
        }
    }
    if (s.nblk_ptr) return fail(NDFB_E_UNSUPPORTED, "scattered output blocks need a length with an instantiated Stockham schedule");
    // no schedule for this length: Bluestein fused around two power-of-two Stockham transforms (bsfft_kernel), if M fits
    if (t.kind == TK_C2C && t.N >= 2 && !s.fs_twiddle && !s.os_blk && (int)s.dims.size() <= kMaxBatchDims) {
        long long M = 1;
        while (M < 2LL * t.N - 1) M <<= 1;
        long long nlanes = 1;
        for (auto& d : s.dims) nlanes *= d.size;
        const bool have_batch = !s.dims.empty() && nlanes > 1;
        const bool cols = have_batch && (llabs_(s.dims[0].is) < llabs_(s.is_axis) || llabs_(s.dims[0].os) < llabs_(s.os_axis));
        const BsfftEntry* e = M <= 8192 ? find_bsfft(sizeof(R) == 8, (int)M, cols, nlanes, std::max(llabs_(s.is_axis), llabs_(s.os_axis)) * (long long)sizeof(Cx<R>)) : nullptr;
        if (e && nlanes > 0) {
            ndfb_plan::BigBlu blu;
            int rc = get_big_bluestein<R>(p, t.N, &blu);
            if (rc) return rc;
            BsfftArgs ba;
            std::memset(&ba, 0, sizeof ba);
            SfftArgs& a = ba.base;
            a.in = s.in; a.out = s.out; a.nlanes = nlanes;
            a.nbd = (int)s.dims.size();
            for (int d = 0; d < a.nbd; ++d) { a.bsz[d] = s.dims[d].size; a.bis[d] = s.dims[d].is; a.bos[d] = s.dims[d].os; }
            a.is_axis = s.is_axis; a.os_axis = s.os_axis;
            a.conj_in = s.conj_in; a.conj_out = s.conj_out; a.scale = s.scale;
            SfftEntry proxy;
            std::memset(&proxy, 0, sizeof proxy);
            for (int i = 0; i < 4; ++i) proxy.r[i] = e->r[i];
            proxy.twtotal = e->twtotal;
            void* twd = nullptr;
            if ((rc = get_sfft_twiddles<R>(p, s.core, &proxy, &twd))) return rc;
            a.tw = twd;
            ba.n = t.N; ba.chirp = blu.chirp; ba.bhat = blu.bhat;
            const bool trace = std::getenv("NDFB_TRACE") != nullptr;
            if (trace) fprintf(stderr, "[ndfb] bsfft %s N=%d M=%d %s L=%d T=%d smem=%zu lanes=%lld\n", sizeof(R) == 8 ? "f64" : "f32", t.N, e->M, e->cols ? "cols" : "rows", e->L, e->threads, e->smem, nlanes);
            const long long grid = (nlanes + e->L - 1) / e->L;
            if (grid > 0x7fffffffLL) return fail(NDFB_E_UNSUPPORTED, "too many tiles");
            return e->launch(ba, (unsigned)grid, stream);
        }
    }
    return launch_tile<R>(p, s, stream);
}

// ---- real-transform fast path ----
static const RsfftEntry* find_rsfft(bool f64, int rkind, int N, bool cols, long long nlanes, long long axis_stride_bytes = 0) {
    static const bool disabled = std::getenv("NDFB_DISABLE_SFFT") != nullptr;
    if (disabled) return nullptr;
    struct Tab { const RsfftEntry* e; int n; };
#define RSFFT_TABLE(name) {kRsfft_##name, kRsfft_##name##_count},
    static const Tab tabs[] = {
#include "rsfft_tables.inc"
    };
#undef RSFFT_TABLE
    if (const char* pick = std::getenv("NDFB_RSFFT_PICK")) {   // measurement hook, as NDFB_SFFT_PICK
        int pn = 0, pi = 0;
        if (sscanf(pick, "%d:%d", &pn, &pi) == 2 && pn == N) {
            int seen = 0;
            for (const Tab& t : tabs)
                for (int i = 0; i < t.n; ++i) {
                    const RsfftEntry* e = &t.e[i];
                    if (e->f64 != (f64 ? 1 : 0) || e->kind != rkind || e->N != N || e->cols != (cols ? 1 : 0)) continue;
                    if (seen++ == pi) return e;
                }
        }
    }
    // family 2 = family B with the small radix FIRST: only for the kinds whose prologue pairs bins j and N-j (C2R, DCT-III),
    // where it lets pass 0 run mirror-paired from registers (kMirrorPro); preferred there, ignored elsewhere
    // Measured on B200 (profiles/round2/r2m_ab_mirror_first_pass.txt): the saved prologue round trip does not pay for the
    // radix-4-first schedule (c3 ndifft_r2c 0.426 -> 0.460 ms, DCT-III rows 0.109 = 0.109 ms, columns 0.146 -> 0.141 ms),
    // so it is opt-in: NDFB_MIRROR_PRO=1.
    // Measured choices the generic rules below cannot see (same-box A/B, profiles/round2/r3i_ab_dct1_schedules.jsonl): the 4095-point core
    // as 15.13.7.3 on 320 threads keeps 88 % of its thread slots busy against 71 % for 13.9.7.5 on 512 (DCT-I of 4096 points, c4):
    // f64 rows 0.164 -> 0.146 ms with two CTAs per SM (three spill: 0.218 ms), f32 rows 0.173 -> 0.150 ms with three, f64 columns +4 %.
    if (N == 4095 && !std::getenv("NDFB_NO_TUNED_PICKS")) {
        const int want_minb = cols ? 1 : (f64 ? 2 : 3);
        for (const Tab& t : tabs)
            for (int i = 0; i < t.n; ++i) {
                const RsfftEntry* e = &t.e[i];
                if (e->f64 == (f64 ? 1 : 0) && e->kind == rkind && e->N == N && e->cols == (cols ? 1 : 0) && e->r[0] == 15 && e->minb == want_minb) return e;
            }
    }
    const bool wants_rev = (rkind == RK_C2R || rkind == RK_DCT3) && std::getenv("NDFB_MIRROR_PRO") != nullptr;
    const int pref = wants_rev ? 2 : preferred_family(f64, N, cols, true);
    const RsfftEntry* best = nullptr;
    for (const Tab& t : tabs)
        for (int i = 0; i < t.n; ++i) {
            const RsfftEntry* e = &t.e[i];
            if (e->f64 != (f64 ? 1 : 0) || e->kind != rkind || e->N != N || e->cols != (cols ? 1 : 0)) continue;
            if (e->fam == 2 && !wants_rev) continue;
            if (better_entry(e, best, nlanes, pref, f64 ? (size_t)8 : (size_t)4, axis_stride_bytes)) best = e;
        }
    return best;
}

// The even-length real kinds: TileKind -> RKind of rsfft_kernel (odd lengths stay on the general kernel)
static int rkind_of(int tk) {
    switch (tk) {
        case TK_R2C_EVEN: return RK_R2C;
        case TK_C2R_EVEN: return RK_C2R;
        case TK_DCT1: return RK_DCT1;
        case TK_DCT2_EVEN: return RK_DCT2;
        case TK_DCT3_EVEN: return RK_DCT3;
        case TK_DCT4_EVEN: return RK_DCT4;
        default: return -1;
    }
}

template <typename R>
static int launch_real(ndfb_plan* p, const LaunchSpec& s, stream_t stream) {
    const CoreTables& t = s.core->t;
    const int rk = rkind_of(t.kind);
    if (rk >= 0 && t.M == 0 && t.N >= 2 && (int)s.dims.size() <= kMaxBatchDims) {
        long long nlanes = 1;
        for (auto& d : s.dims) nlanes *= d.size;
        const bool have_batch = !s.dims.empty() && nlanes > 1;
        const bool cols = have_batch && (llabs_(s.dims[0].is) < llabs_(s.is_axis) || llabs_(s.dims[0].os) < llabs_(s.os_axis));
        // the row kernels read / write (re, im)-style pairs of reals with one vector access: needs pair alignment
        bool ok = true;
        if (!cols) {
            if (rk == RK_R2C && s.is_axis == 1) {
                ok = ((uintptr_t)s.in % (2 * sizeof(R))) == 0;
                for (auto& d : s.dims) if (d.is % 2) ok = false;
            }
            if (rk == RK_C2R && s.os_axis == 1) {
                ok = ((uintptr_t)s.out % (2 * sizeof(R))) == 0;
                for (auto& d : s.dims) if (d.os % 2) ok = false;
            }
        }
        const RsfftEntry* e = ok ? find_rsfft(sizeof(R) == 8, rk, t.N, cols, nlanes, std::max(llabs_(s.is_axis), llabs_(s.os_axis)) * (long long)sizeof(R)) : nullptr;
        if (!e && ok && nlanes > 0) {
            e = jit_entry<RsfftEntry>(sizeof(R) == 8, rk, t.N, cols, nlanes);
        }
        if (e) {
            const bool trace = std::getenv("NDFB_TRACE") != nullptr;
            if (trace) fprintf(stderr, "[ndfb] rsfft kind=%d %s N=%d %s L=%d T=%d smem=%zu lanes=%lld minb=%d fam=%c%s\n", rk, sizeof(R) == 8 ? "f64" : "f32", e->N, e->cols ? "cols" : "rows", e->L, e->threads, e->smem, nlanes, e->minb, e->fam == 2 ? 'R' : e->fam ? 'B' : 'A', e->jit_func ? " jit" : "");
            RsfftArgs a;
            std::memset(&a, 0, sizeof a);
            a.in = s.in; a.out = s.out; a.nlanes = nlanes;
            if (nlanes == 0) return 0;
            a.nbd = (int)s.dims.size();
            for (int d = 0; d < a.nbd; ++d) { a.bsz[d] = s.dims[d].size; a.bis[d] = s.dims[d].is; a.bos[d] = s.dims[d].os; }
            a.is_axis = s.is_axis; a.os_axis = s.os_axis;
            a.n = t.n; a.scale = s.scale;
            a.tabA = s.core->d.tabA; a.tabB = s.core->d.tabB;
            // contiguous DCT-III / DCT-IV output rows that start on 2-real boundaries: mirror-paired last pass, pairs of reals stored
            // straight from registers (sfft_kernel.cuh: kMirrorOut); NDFB_NO_MIRROR_OUT keeps the staged copy-out (A/B runs)
            // Measured on B200: DCT-IV +10-11 % (4096^2 f64 rows 0.084 -> 0.076 ms), DCT-III f32 +13 %, f64 +6 % once its four reals per
            // sector go out as ONE 256-bit store (two 16-byte stores per thread, i.e. half sectors per request, lost 3 %:
            // profiles/round2/r2s_ab_mirror_out.jsonl, r3g_ab_mirror_out_wide.jsonl).  NDFB_NO_MIRROR_OUT keeps the staged copy-out.
            const bool want_mirror = rk == RK_DCT4 || rk == RK_DCT3;
            if (!cols && s.os_axis == 1 && want_mirror && !std::getenv("NDFB_NO_MIRROR_OUT")) {
                bool al = ((uintptr_t)s.out % (2 * sizeof(R))) == 0, al4 = ((uintptr_t)s.out % (4 * sizeof(R))) == 0;
                for (auto& d : s.dims) { if (d.os % 2) al = false; if (d.os % 4) al4 = false; }
                a.vec_out = al ? (al4 ? 2 : 1) : 0;
                if (a.vec_out && trace) fprintf(stderr, "[ndfb] real kind=%d rows: mirror-paired output pass where the schedule allows it (aligned contiguous rows)\n", rk);
            }
            SfftEntry proxy;
            std::memset(&proxy, 0, sizeof proxy);
            for (int i = 0; i < 4; ++i) proxy.r[i] = e->r[i];
            proxy.twtotal = e->twtotal;
            void* twd = nullptr;
            int rc = get_sfft_twiddles<R>(p, s.core, &proxy, &twd);
            if (rc) return rc;
            a.tw = twd;
            const long long grid = (nlanes + e->L - 1) / e->L;
            if (grid > 0x7fffffffLL) return fail(NDFB_E_UNSUPPORTED, "too many tiles");
            if (g_sync_hint.signal_cnt) {
                const SyncHint h = g_sync_hint;
                g_sync_hint.clear();
                if (h.signal_group % e->L) return fail(NDFB_E_UNSUPPORTED, "launch signal hint: group size is not a multiple of the tile width");
                a.done_cnt = h.signal_cnt; a.done_group = h.signal_group;
            }
            if (e->jit_func) return jit_launch(e->jit_func, a, (unsigned)grid, (unsigned)e->threads, e->smem, stream);
            return e->launch(a, (unsigned)grid, stream);
        }
    }
    return launch_tile<R>(p, s, stream);
}

// order batch dims by input stride and merge the ones that are contiguous in both arrays
static void normalize_dims(std::vector<BDim>& dims) {
    std::vector<BDim> v;
    for (auto& d : dims) if (d.size != 1) v.push_back(d);
    std::stable_sort(v.begin(), v.end(), [](const BDim& x, const BDim& y) {
        if (llabs_(x.is) != llabs_(y.is)) return llabs_(x.is) < llabs_(y.is);
        return llabs_(x.os) < llabs_(y.os);
    });
    std::vector<BDim> m;
    for (auto& d : v) {
        if (!m.empty() && d.is == m.back().is * m.back().size && d.os == m.back().os * m.back().size) m.back().size *= d.size;
        else m.push_back(d);
    }
    dims.swap(m);
}

static bool fits_one_tile(ndfb_plan* p, const CoreTables& t, size_t cs) {
    const size_t cap = dev_smem_cap(p->device) - 1024;
    return (size_t)t.Bl * cs + 64 <= cap;
}

// Long strided columns, N = N1*N2 <= 2^17: both passes in ONE persistent launch with the workspace ring resident in L2
// (fs2_kernel).  *done = 0 when the case does not qualify (the caller then runs the two-launch path).
template <typename R>
static int exec_fs2(ndfb_plan* p, long long N, bool inverse, double scale, const void* in, void* out, const BDim& cols,
                    long long is_axis, long long os_axis, stream_t stream, int* done) {
    *done = 0;
    // Opt-in (NDFB_FS2=1).  Measured on B200 in the bench's c2 step (profiles/r1u_fused_two_pass.jsonl, r1z notes in
    // DESIGN.md 4.4b): the fused launch halves the DRAM traffic (1.13 GB instead of 2.15 GB per call) but is bound by
    // instruction issue / latency on the SMs (0.344 ms), while the two separate pass kernels, after the cursor addressing,
    // each run at 98 % of the HBM copy peak (0.334 ms for both) -- so the two-launch path is the default again.
    if (!std::getenv("NDFB_FS2") || std::getenv("NDFB_NO_FS2") || N > (1LL << 17)) return 0;
    const bool f64 = sizeof(R) == 8;
    if (f64 && !std::getenv("NDFB_FS2_F64")) return 0;
    if (N != 8192 && !std::getenv("NDFB_FS2_ALL")) return 0;
    const size_t cs = sizeof(Cx<R>);
    const Fs2Entry* e = nullptr;
    const int want_t = std::getenv("NDFB_FS2_T") ? atoi(std::getenv("NDFB_FS2_T")) : 0;
    for (int i = 0; i < kFs2_count; ++i)
        if (kFs2[i].f64 == (f64 ? 1 : 0) && (long long)kFs2[i].N1 * kFs2[i].N2 == N && (!want_t || kFs2[i].threads == want_t)) { e = &kFs2[i]; break; }
    if (!e) return 0;
    // group width: the widest divisor of the column count that keeps one group's intermediate within the budget
    size_t budget = (size_t)8 << 20;
    if (const char* v = std::getenv("NDFB_FS2_KB")) budget = (size_t)atoll(v) << 10;
    const long long W0 = cols.size, lm = std::max(e->L1, e->L2);
    long long Wg = 0;
    for (long long w = lm; w <= W0; w += lm)
        if (W0 % w == 0 && (size_t)w * (size_t)N * cs <= budget) Wg = w;
    if (!Wg) return 0;
    const long long G = W0 / Wg;
    if (G < 4 || G > (1 << 20)) return 0;
    int ring = 3;
    if (const char* v = std::getenv("NDFB_FS2_RING")) ring = std::max(2, atoi(v));
    int rc;
    void *ws = nullptr, *sync = nullptr;
    if ((rc = g_pool.get(2, p->device, (size_t)ring * (size_t)N * (size_t)Wg * cs, &ws))) return rc;
    const size_t sync_bytes = (size_t)(2 + 2 * G) * sizeof(unsigned);
    if ((rc = g_pool.get(6, p->device, sync_bytes, &sync))) return rc;
    Core* c1 = get_core(p, TK_C2C, e->N1);
    Core* c2 = get_core(p, TK_C2C, e->N2);
    if ((rc = ensure_device<R>(p, c1)) || (rc = ensure_device<R>(p, c2))) return rc;
    ndfb_plan::FsTw fs;
    if ((rc = get_fs_twiddles<R>(p, N, &fs))) return rc;
    if (fs.shift < 40) return 0;
    SfftEntry px1, px2;
    std::memset(&px1, 0, sizeof px1); std::memset(&px2, 0, sizeof px2);
    for (int i = 0; i < 4; ++i) { px1.r[i] = e->r1[i]; px2.r[i] = e->r2[i]; }
    px1.twtotal = e->tw1; px2.twtotal = e->tw2;
    void *tw1 = nullptr, *tw2 = nullptr;
    if ((rc = get_sfft_twiddles<R>(p, c1, &px1, &tw1)) || (rc = get_sfft_twiddles<R>(p, c2, &px2, &tw2))) return rc;
    Fs2Args f;
    std::memset(&f, 0, sizeof f);
    f.a1.in = in; f.a1.out = ws; f.a1.is_axis = e->N2 * is_axis; f.a1.os_axis = (long long)e->N2 * Wg;
    f.a1.conj_in = inverse ? 1 : 0; f.a1.scale = 1.0; f.a1.tw = tw1;
    f.a1.fs_twiddle = 1; f.a1.fs_shift = fs.shift; f.a1.fs_lo = fs.lo; f.a1.fs_hi = fs.hi;
    f.a2.in = ws; f.a2.out = out; f.a2.is_axis = Wg; f.a2.os_axis = e->N1 * os_axis;
    f.a2.conj_out = inverse ? 1 : 0; f.a2.scale = scale; f.a2.tw = tw2;
    f.sync = (unsigned*)sync;
    f.G = (int)G; f.Wg = (int)Wg; f.ring = (int)std::min<long long>(ring, G);
    f.ring_stride = N * Wg;
    f.col_is = cols.is; f.col_os = cols.os; f.j2_is = is_axis; f.k1_os = os_axis;
    f.N1 = e->N1; f.N2 = e->N2;
    if (const char* v = std::getenv("NDFB_FS2_DBG")) f.dbg = atoi(v);
    f.tiles1 = (unsigned)((Wg / e->L1) * e->N2);
    f.tiles2 = (unsigned)((Wg / e->L2) * e->N1);
    const long long total = G * (long long)(f.tiles1 + f.tiles2);
    if (total > 0x7fffffffLL) return 0;
    if ((rc = dev_memset0(sync, sync_bytes, stream))) return rc;
    const long long grid = std::min<long long>(total, (long long)dev_sm_count(p->device) * e->minb);
    if (std::getenv("NDFB_TRACE"))
        fprintf(stderr, "[ndfb] fs2 %s N=%lld = %d x %d fused two-pass, %lld groups of %lld columns, ring %d, grid %lld\n", f64 ? "f64" : "f32", N,
                e->N1, e->N2, G, Wg, f.ring, grid);
    *done = 1;
    return e->launch(f, (unsigned)grid, stream);
}

// C2C of a length too long for one CTA's shared memory: four-step N = N1*N2 through a device workspace.
template <typename R>
static int exec_four_step(ndfb_plan* p, long long N, bool inverse, double scale, const void* in, void* out,
                          std::vector<BDim> dims, long long is_axis, long long os_axis, stream_t stream,
                          int depth = 0, int conj_in_override = -1) {
    // conj_in_override: nested calls transform the workspace (already conjugated by the outer pass 1) and only need the
    // output conjugation + scale of an inverse transform
    const bool conj_in = conj_in_override >= 0 ? conj_in_override != 0 : inverse;
    const size_t cs = sizeof(Cx<R>);
    const size_t cap = dev_smem_cap(p->device) - 1024;
    long long cap1 = (long long)((cap - 256) / (4 * cs));  // pass 1 wants >= 4 lanes per tile (32-byte rows)
    long long cap2 = (long long)((cap - 256) / cs);
    if (const char* f = std::getenv("NDFB_FS_CAP")) cap1 = cap2 = atoll(f);   // test hook: pretend the chip is tiny
    // Soft cap of the second factor.  Measured on B200 (profiles/round2/r2x_fs_medium.txt, 4.3 GB arrays): with both factors >= 1024 the
    // passes run one 139 KB tile per SM at ~0.5 of the copy rate, three passes over small tiles run at ~0.9 each:
    // 2^20 c64 5.45 -> 4.02 ms, 2^22 c64 5.10 -> 4.09 ms, 2^20 c128 5.37 -> 4.13 ms, 2^22 c128 5.50 -> 4.44 ms.
    const long long cap2s = std::getenv("NDFB_FS_CAP") || std::getenv("NDFB_FS_TWO_PASS") ? cap2 : std::min<long long>(cap2, 512);
    if (!is_smooth(N)) return fail(NDFB_E_UNSUPPORTED, "length %lld has a prime factor > 13 and is too long for the single-pass Bluestein kernel", N);
    // N1 * N2 = N with both factors on chip; prefer factors that have an instantiated Stockham schedule (and, for the
    // column pass, a tile at least one 32-byte sector wide), then the most square split
    normalize_dims(dims);
    const bool strided_lanes = !dims.empty() && llabs_(dims[0].is) < llabs_(is_axis) && llabs_(dims[0].os) < llabs_(os_axis);
    if (depth == 0 && strided_lanes && dims.size() == 1 && conj_in_override < 0 && dims[0].is > 0 && dims[0].os > 0) {
        int done = 0;
        int rc = exec_fs2<R>(p, N, inverse, scale, in, out, dims[0], is_axis, os_axis, stream, &done);
        if (rc || done) return rc;
    }
    // Optional L2-resident workspace (NDFB_FS_L2_KB=<group size>; off by default): run the two passes group by group over
    // a slice of the lanes and reuse ONE workspace, so that pass 2 reads what pass 1 just wrote from L2 and the two-pass
    // transform costs one HBM round trip.  Measured on B200 (profiles/r1t_l2_groups.jsonl): the DRAM traffic does halve,
    // but 32-64 MB groups are 15-25 us kernels of 2-4 waves whose ramp-up and tail cost more than the saved pass
    // (c2 axis 0: 0.37 ms ungrouped, 0.42 ms in 64 MB groups even replayed from a CUDA graph) -- it needs a persistent
    // two-pass kernel with device-side group dependencies to pay off.
    if (depth == 0 && !dims.empty()) {
        size_t target = 0;
        if (const char* e = std::getenv("NDFB_FS_L2_KB")) target = (size_t)atoll(e) << 10;   // 0 disables; tests shrink it
        long long nb_all = 1;
        for (auto& d : dims) nb_all *= d.size;
        const size_t lane_bytes = (size_t)N * cs;
        if (target && (size_t)nb_all * lane_bytes > target + target / 2) {
            // slice the outermost batch dim; a single index of it may still be too big (then the recursion slices the next)
            const int cd = (int)dims.size() - 1;
            long long inner = nb_all / dims[cd].size;
            long long per = (long long)(target / ((size_t)inner * lane_bytes));     // indices of dim cd per group
            const bool innermost_cols = strided_lanes && cd == 0;
            if (innermost_cols) per = per / 64 * 64;                                // keep whole 512-byte column groups
            if (per < 1 && cd > 0) per = 1;
            if (per >= 1 && per < dims[cd].size && (!innermost_cols || per >= 64)) {
                for (long long c0 = 0; c0 < dims[cd].size; c0 += per) {
                    std::vector<BDim> sub = dims;
                    sub[cd].size = std::min(per, dims[cd].size - c0);
                    int rc = exec_four_step<R>(p, N, inverse, scale, (const char*)in + c0 * dims[cd].is * (long long)cs,
                                               (char*)out + c0 * dims[cd].os * (long long)cs, sub, is_axis, os_axis, stream, 0, conj_in_override);
                    if (rc) return rc;
                }
                return 0;
            }
        }
    }
    long long best1 = 0, nested1 = 0;
    int best_score = -1, nested_score = -1;
    for (long long d = 1; d * d <= N; ++d) {
        if (N % d) continue;
        long long cands[2] = {d, N / d};
        for (long long n1 : cands) {
            long long n2 = N / n1;
            // second factors beyond 512 points run one CTA per SM (128 KB tiles): where a third pass is possible it is preferred
            bool nested = n2 > cap2s;
            if (n1 > cap1 || n1 < 2 || n2 < 2) continue;
            if (nested && (depth > 0 || strided_lanes || n2 > cap1 * cap2 || (int)dims.size() + 2 > kMaxBatchDims)) {
                if (n2 > cap2) continue;
                nested = false;         // no third pass here: the big second-factor tile it is
            }
            if (const char* f = std::getenv("NDFB_FS_N1")) { if (depth == 0 && atoll(f) != n1) continue; }
            else if (nested && n1 != 256 && n1 != 64) continue;    // three-pass split: 64- or 256-point first pass (256-byte rows)
            int score = nested ? 1 : 0;
            if (nested) {
                const SfftEntry* e1 = find_sfft(sizeof(R) == 8, (int)n1, true, 1 << 20);
                if (e1 && (size_t)e1->L * cs >= 128) score += 2;
                // two-pass splits whose tiles are at least 64 bytes wide outrank the three-pass ones (decided after the loop)
                // ties: f32 the 256-point first pass (64 x 2^24 c64: 256^3 8.55 ms, 128 x .. 8.60, 64 x (512 x 512) 8.89, 512 x .. 8.99),
                // f64 the 64-point one (2^20 c128: 4.13 vs 4.33 ms, 2^21: 4.21 vs 4.36 ms; profiles/round2/r2y_fs_medium.txt)
                const bool prefer = sizeof(R) == 4 ? n1 > nested1 : (nested1 == 0 || n1 < nested1);
                if (score > nested_score || (score == nested_score && prefer)) { nested_score = score; nested1 = n1; }
                continue;
            }
            const SfftEntry* e1 = find_sfft(sizeof(R) == 8, (int)n1, true, 1 << 20);
            // pass 2 stores out[k1 + N1*k2]: its OUTPUT is lane-interleaved in both variants, so it runs the column
            // layout and wants a tile at least one sector wide too (2^24 f32 x 64: 2048 x 8192 14.6 ms, 4096 x 4096 12.4 ms)
            const SfftEntry* e2 = find_sfft(sizeof(R) == 8, (int)n2, true, 1 << 20);
            if (e1 && (size_t)e1->L * cs >= 32) score += 2;
            if (e1 && (size_t)e1->L * cs >= 64) score += 1;
            if (e2 && (size_t)e2->L * cs >= 32) score += 2;
            if (e2 && (size_t)e2->L * cs >= 64) score += 1;
            // f64: a first factor beyond 512 points (147 KB tiles) loses to three passes (2^19: 4.10 vs 3.90 ms, 2^20: 4.48 vs 4.16 ms);
            // f32 it wins while its tile rows stay >= 64 bytes (2^19: 3.58 vs 4.38 ms, 2^20: 3.92 vs 4.31 ms; profiles/round2/r3b, r3c_fs_medium.txt)
            if (sizeof(R) == 8 && std::max(n1, n2) > 512 && score > 0) score -= 1;
            if (score > best_score || (score == best_score && llabs_(n1 - n2) < llabs_(best1 - N / best1))) { best_score = score; best1 = n1; }
        }
    }
    // Two passes whose tile rows are only 32 bytes wide (2^24 = 4096 x 4096 in f32) are bounded by what HBM gives such tiles (~0.5 of
    // the copy rate per pass): three passes over 256-byte rows win there.  Measured on B200, 64 x 2^24 c64
    // (profiles/round2/r2v_c5b_variants.txt): 64 x (512 x 512) 10.28 ms, 256^3 10.54 ms, 4096 x 4096 11.33 ms.
    bool pick_nested = false;
    if (nested1 && (best_score < 6 || !best1) && nested_score >= (best1 ? 3 : 0)) { best1 = nested1; best_score = nested_score; pick_nested = true; }
    if (!best1) return fail(NDFB_E_UNSUPPORTED, "length %lld is too long for the two-pass decomposition (max about %lld)", N, cap1 * cap2);
    const long long N1 = best1, N2 = N / N1;
    if ((int)dims.size() + 1 > kMaxBatchDims) return fail(NDFB_E_UNSUPPORTED, "four-step transform with more than %d batch dims", kMaxBatchDims - 1);
    long long nb = 1;
    for (auto& d : dims) nb *= d.size;
    void* ws = nullptr;
    int rc = g_pool.get(depth == 0 ? 2 : 5, p->device, (size_t)nb * (size_t)N * cs, &ws);
    if (rc) return rc;
    const bool nested2 = pick_nested;
    Core* c1 = get_core(p, TK_C2C, (int)N1);
    Core* c2 = nested2 ? nullptr : get_core(p, TK_C2C, (int)N2);
    if ((rc = ensure_device<R>(p, c1))) return rc;
    if (c2 && (rc = ensure_device<R>(p, c2))) return rc;
    ndfb_plan::FsTw fs;
    if ((rc = get_fs_twiddles<R>(p, N, &fs))) return rc;
    const bool strided = strided_lanes;
    LaunchSpec s1, s2;
    s1.core = c1; s1.in = in; s1.out = ws;
    s2.core = c2; s2.in = ws; s2.out = out;
    if (!strided) {
        // pass 1: lanes (j2, batch...), transform over j1 (stride N2), twiddle W_N^{k1 j2}; ws[b][k1][j2]
        s1.dims.push_back({N2, is_axis, 1});
        long long wstride = N;
        for (auto& d : dims) { s1.dims.push_back({d.size, d.is, wstride}); wstride *= d.size; }
        s1.is_axis = N2 * is_axis; s1.os_axis = N2;
        s1.fs_dim = 0;
        // pass 2: lanes (k1, batch...), transform over j2 (contiguous), out[k1 + N1*k2]
        s2.dims.push_back({N1, N2, os_axis});
        wstride = N;
        for (auto& d : dims) { s2.dims.push_back({d.size, wstride, d.os}); wstride *= d.size; }
        s2.is_axis = 1; s2.os_axis = N1 * os_axis;
    } else {
        // strided columns: keep the fastest batch dim (adjacent columns) innermost in the workspace too,
        // ws[outer][j][col], so that both passes read and write full rows of adjacent lanes
        const long long W0 = dims[0].size;
        s1.dims.push_back({W0, dims[0].is, 1});
        s1.dims.push_back({N2, is_axis, W0});
        long long wstride = N * W0;
        for (size_t d = 1; d < dims.size(); ++d) { s1.dims.push_back({dims[d].size, dims[d].is, wstride}); wstride *= dims[d].size; }
        s1.is_axis = N2 * is_axis; s1.os_axis = N2 * W0;
        s1.fs_dim = 1;
        s2.dims.push_back({W0, 1, dims[0].os});
        s2.dims.push_back({N1, N2 * W0, os_axis});
        wstride = N * W0;
        for (size_t d = 1; d < dims.size(); ++d) { s2.dims.push_back({dims[d].size, wstride, dims[d].os}); wstride *= dims[d].size; }
        s2.is_axis = W0; s2.os_axis = N1 * os_axis;
    }
    s1.conj_in = conj_in; s1.fs_twiddle = 1; s1.fs = fs;
    if ((rc = launch_c2c<R>(p, s1, stream))) return rc;
    s2.conj_out = inverse; s2.scale = scale;
    if (!nested2 && !strided && depth == 0 && !std::getenv("NDFB_NO_FS_TRANSPOSE") && !std::getenv("NDFB_NO_TRANS_STORE")) {
        // Last pass of a TWO-pass split: its lanes k1 are rows of the workspace (N2 elements apart) and adjacent elements of the
        // output — the transposing rows kernel's case as it stands (launch_c2c falls back to the column tile when there is no
        // multi-lane row schedule for N2 or the output rows are not contiguous).  Measured on B200, 8192 rows of 2^16 points c64
        // (profiles/round2/r3a_launches_*.csv): the column tile reads 32 lanes x 8 bytes per warp request and takes 2.67 ms where the
        // first pass takes 1.45 ms.
        s2.trans = true;
    }
    if (!nested2 && !strided && depth > 0 && !std::getenv("NDFB_NO_FS_TRANSPOSE")) {
        // Transposing last pass of the three-pass split: tile the lanes along the dim that is contiguous in the OUTPUT (the
        // outer level's k1, output stride 1) instead of the one contiguous in the workspace, and cap the tile at 8 lanes
        // so that a warp still reads 32/L consecutive points (>= one 32-byte sector) of each lane's row.  Measured on
        // B200, 64 x 2^24 c64 as 256 x (256 x 256): 21.96 -> 12.21 ms (profiles/r2b_c5b_three_pass.jsonl); the two-pass
        // 4096 x 4096 split (11.2 ms) still wins where it fits, so this matters for lengths beyond two on-chip factors.
        size_t best = 0;
        for (size_t d = 1; d < s2.dims.size(); ++d)
            if (llabs_(s2.dims[d].os) < llabs_(s2.dims[best].os)) best = d;
        if (best != 0) {
            std::swap(s2.dims[0], s2.dims[best]);
            s2.max_L = sizeof(R) == 8 ? 16 : 8;
            // preferred: the transposing rows kernel (full rows in, L adjacent output elements per store); the capped column
            // tile above remains the fallback when there is no multi-pass row schedule for N2
            s2.trans = !std::getenv("NDFB_NO_TRANS_STORE");
        }
    }
    {
        const bool trace = std::getenv("NDFB_TRACE") != nullptr;
        if (trace) fprintf(stderr, "[ndfb] four-step N=%lld = %lld x %lld (%s lanes%s)\n", N, N1, N2, strided ? "strided" : "contiguous", nested2 ? ", second factor split again" : "");
    }
    if (nested2)   // the second factor is itself too long for one CTA: split it again (three passes in total)
        return exec_four_step<R>(p, N2, inverse, scale, s2.in, s2.out, s2.dims, s2.is_axis, s2.os_axis, stream, depth + 1, /*conj_in=*/0);
    return launch_c2c<R>(p, s2, stream);
}

struct OpInfo {
    int tk;
    bool in_complex, out_complex;
    long long n_in, n_out;
    int conj_in = 0, conj_out = 0;
    double scale = 1.0;
    const char* what = "fft";
    int os_blk = 0;
    long long os_blk_stride = 0;
    int nblk_ptr = 0;
    void* blk_ptr[8] = {nullptr};
};

static int op_info(const ndfb_plan* p, int op, int norm, OpInfo* o) {
    const long long n = (long long)p->n;
    const long long m = n / 2 + 1;
    const bool def = norm == NDFB_NORM_DEFAULT;
    switch (op) {
        case NDFB_OP_FFT:
        case NDFB_OP_IFFT:
            if (p->kind != NDFB_C2C) return fail(NDFB_E_INVALID, "op %d needs a C2C plan", op);
            o->tk = TK_C2C; o->in_complex = o->out_complex = true; o->n_in = o->n_out = n;
            if (op == NDFB_OP_IFFT) { o->conj_in = o->conj_out = 1; o->scale = def && n > 0 ? 1.0 / (double)n : 1.0; }
            return 0;
        case NDFB_OP_R2C:
            if (p->kind != NDFB_R2C) return fail(NDFB_E_INVALID, "op %d needs an R2C plan", op);
            o->tk = (n % 2 == 0 && n >= 2) ? TK_R2C_EVEN : TK_R2C_ODD;
            o->in_complex = false; o->out_complex = true; o->n_in = n; o->n_out = m;
            return 0;
        case NDFB_OP_C2R:
            if (p->kind != NDFB_R2C) return fail(NDFB_E_INVALID, "op %d needs an R2C plan", op);
            o->tk = (n % 2 == 0 && n >= 2) ? TK_C2R_EVEN : TK_C2R_ODD;
            o->in_complex = true; o->out_complex = false; o->n_in = m; o->n_out = n;
            o->scale = def && n > 0 ? 1.0 / (double)n : 1.0;
            return 0;
        case NDFB_OP_DCT1: case NDFB_OP_DCT2: case NDFB_OP_DCT3: case NDFB_OP_DCT4: {
            if (p->kind != NDFB_DCT) return fail(NDFB_E_INVALID, "op %d needs a DCT plan", op);
            const bool even = n % 2 == 0 && n >= 2;
            if (op == NDFB_OP_DCT1) o->tk = TK_DCT1;
            else if (op == NDFB_OP_DCT2) o->tk = even ? TK_DCT2_EVEN : TK_DCT2_ODD;
            else if (op == NDFB_OP_DCT3) o->tk = even ? TK_DCT3_EVEN : TK_DCT3_ODD;
            else o->tk = even ? TK_DCT4_EVEN : TK_DCT4_ODD;
            o->in_complex = o->out_complex = false; o->n_in = o->n_out = n;
            o->scale = def ? 2.0 : 1.0;
            o->what = "dct";
            return 0;
        }
        default: return fail(NDFB_E_INVALID, "unknown op %d", op);
    }
}


// ------------------------------------------------------------------------------------------------------
// staged path: prologue kernel -> complex core on workspace rows -> epilogue kernel (big_kernels.cuh)
// ------------------------------------------------------------------------------------------------------
struct BigMulArgs { void* ws; const void* bhat; long long ldw, nchunk; int M; };
template <typename R>
__global__ void __launch_bounds__(256) big_mul_entry(const __grid_constant__ BigMulArgs a) {
    Cx<R>* __restrict__ ws = reinterpret_cast<Cx<R>*>(a.ws);
    const Cx<R>* __restrict__ bhat = reinterpret_cast<const Cx<R>*>(a.bhat);
    const long long total = a.nchunk * (long long)a.M;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long lane = idx / a.M;
        const int k = (int)(idx - lane * a.M);
        Cx<R>* q = &ws[lane * a.ldw + k];
        *q = cconj(cmul(*q, ldg(&bhat[k])));
    }
}

// forward C2C of `rows` contiguous rows of length N (N smooth): one tile per row if it fits, else four-step
template <typename R>
static int c2c_rows(ndfb_plan* p, long long N, const void* in, void* out, long long rows, long long ld, stream_t stream) {
    const size_t cs = sizeof(Cx<R>);
    const size_t cap = dev_smem_cap(p->device) - 1024;
    if ((size_t)N * cs + 64 <= cap && !std::getenv("NDFB_FORCE_FOUR_STEP")) {
        Core* c = get_core(p, TK_C2C, (int)N);
        int rc = ensure_device<R>(p, c);
        if (rc) return rc;
        LaunchSpec s;
        s.core = c; s.in = in; s.out = out;
        s.dims.push_back({rows, ld, ld});
        s.is_axis = 1; s.os_axis = 1;
        return launch_c2c<R>(p, s, stream);
    }
    std::vector<BDim> dims;
    dims.push_back({rows, ld, ld});
    return exec_four_step<R>(p, N, false, 1.0, in, out, dims, 1, 1, stream);
}

// host FFT (iterative radix 2, double) for the Bluestein kernel spectrum of the staged path
static void host_fft_pow2(std::vector<std::complex<double>>& x) {
    const size_t n = x.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(x[i], x[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        std::vector<std::complex<double>> w(len / 2);
        for (size_t k = 0; k < len / 2; ++k) { cld u = unit_root((long long)k, (long long)len); w[k] = std::complex<double>((double)u.real(), (double)u.imag()); }
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                const std::complex<double> a = x[i + k], b = x[i + k + len / 2] * w[k];
                x[i + k] = a + b;
                x[i + k + len / 2] = a - b;
            }
    }
}

template <typename R>
static int get_big_bluestein(ndfb_plan* p, long long N, ndfb_plan::BigBlu* out) {
    std::lock_guard<std::mutex> g(p->mu);
    auto it = p->bigblu.find(N);
    if (it != p->bigblu.end()) { *out = it->second; return 0; }
    long long M = 1;
    while (M < 2 * N - 1) M <<= 1;
    if (M > (1LL << 26)) return fail(NDFB_E_UNSUPPORTED, "length %lld needs a %lld-point Bluestein convolution: too long", N, M);
    std::vector<cld> chirp(N);
    for (long long j = 0; j < N; ++j) chirp[j] = unit_root((j * j) % (2 * N), 2 * N);
    std::vector<std::complex<double>> b(M, std::complex<double>(0, 0));
    for (long long j = 0; j < N; ++j) {
        std::complex<double> v((double)chirp[j].real(), -(double)chirp[j].imag());
        b[j] = v;
        if (j > 0) b[M - j] = v;
    }
    host_fft_pow2(b);
    std::vector<cld> bh(M);
    for (long long k = 0; k < M; ++k) bh[k] = cld(b[k].real() / (double)M, b[k].imag() / (double)M);
    ndfb_plan::BigBlu t;
    t.M = (int)M;
    int rc;
    if ((rc = dev_set(p->device))) return rc;
    if ((rc = upload_cx<R>(&t.chirp, chirp))) return rc;
    if ((rc = upload_cx<R>(&t.bhat, bh))) return rc;
    p->bigblu[N] = t;
    *out = t;
    return 0;
}

template <typename R>
static int exec_big(ndfb_plan* p, const OpInfo& o, double scale, const void* in, void* out, std::vector<BDim> dims,
                    long long is_axis, long long os_axis, stream_t stream) {
    const size_t cs = sizeof(Cx<R>);
    int bk = o.tk == TK_C2C ? (int)BK_C2C : rkind_of(o.tk);
    const long long n = (long long)p->n;
    long long N = o.tk == TK_C2C ? n : (o.tk == TK_DCT1 ? n - 1 : n / 2);
    switch (o.tk) {   // odd lengths: full-length complex core
        case TK_R2C_ODD: bk = BK_R2C_ODD; N = n; break;
        case TK_C2R_ODD: bk = BK_C2R_ODD; N = n; break;
        case TK_DCT2_ODD: bk = BK_DCT2_ODD; N = n; break;
        case TK_DCT3_ODD: bk = BK_DCT3_ODD; N = n; break;
        case TK_DCT4_ODD: bk = BK_DCT4_ODD; N = 2 * n; break;
        default: break;
    }
    if (bk < 0) return fail(NDFB_E_UNSUPPORTED, "%s of length %zu has no staged schedule", o.what, p->n);
    if (N < 1) return fail(NDFB_E_UNSUPPORTED, "length too short for the staged path");
    normalize_dims(dims);
    if ((int)dims.size() > kMaxBatchDims) return fail(NDFB_E_UNSUPPORTED, "staged transform with more than %d batch dims", kMaxBatchDims);
    long long nlanes = 1;
    for (auto& d : dims) nlanes *= d.size;
    if (nlanes == 0) return 0;
    int rc;
    // kind tables (tabA / tabB) without the single-tile tables
    Core* kc = nullptr;
    if (o.tk != TK_C2C) {
        std::lock_guard<std::mutex> g(p->mu);
        auto key = std::make_pair(o.tk + 1000, (int)n);
        auto it = p->cores.find(key);
        if (it == p->cores.end()) {
            std::unique_ptr<Core> c(new Core());
            build_core(c->t, o.tk, (int)n, /*tables_only=*/true);
            kc = c.get();
            p->cores[key] = std::move(c);
        } else kc = it->second.get();
    }
    if (kc && (rc = ensure_device<R>(p, kc))) return rc;
    ndfb_plan::BigBlu blu;
    const bool bluestein = !is_smooth(N);
    if (bluestein && (rc = get_big_bluestein<R>(p, N, &blu))) return rc;
    const long long rowlen = bluestein ? blu.M : N;
    const long long budget = 1LL << 30;   // bytes per workspace buffer
    long long chunk = std::max<long long>(1, budget / (rowlen * (long long)cs));
    chunk = std::min(chunk, nlanes);
    void *w1 = nullptr, *w2 = nullptr;
    if ((rc = g_pool.get(3, p->device, (size_t)(chunk * rowlen) * cs, &w1))) return rc;
    if ((rc = g_pool.get(4, p->device, (size_t)(chunk * rowlen) * cs, &w2))) return rc;
    {
        const bool trace = std::getenv("NDFB_TRACE") != nullptr;
        if (trace) fprintf(stderr, "[ndfb] staged kind=%d n=%lld core N=%lld %s rowlen=%lld lanes=%lld chunk=%lld\n", bk, n, N, bluestein ? "bluestein" : "direct", rowlen, nlanes, chunk);
    }
    BigArgs a;
    std::memset(&a, 0, sizeof a);
    a.in = in; a.out = out; a.ldw = rowlen; a.nlanes = nlanes;
    a.nbd = (int)dims.size();
    for (int d = 0; d < a.nbd; ++d) { a.bsz[d] = dims[d].size; a.bis[d] = dims[d].is; a.bos[d] = dims[d].os; }
    a.is_axis = is_axis; a.os_axis = os_axis;
    a.kind = bk; a.n = (int)n; a.N = (int)N; a.M = bluestein ? blu.M : 0; a.n_out = (int)o.n_out;
    a.conj_in = o.conj_in; a.conj_out = o.conj_out; a.scale = scale;
    a.tabA = kc ? kc->d.tabA : nullptr; a.tabB = kc ? kc->d.tabB : nullptr;
    a.chirp = blu.chirp;
    const int sms = dev_sm_count(p->device);
    for (long long l0 = 0; l0 < nlanes; l0 += chunk) {
        const long long nc = std::min(chunk, nlanes - l0);
        a.lane0 = l0; a.nchunk = nc;
        a.ws = w1;
        unsigned grid = (unsigned)std::min<long long>((nc * rowlen + 255) / 256, (long long)sms * 16);
        if ((rc = dev_launch(big_pro_kernel<R>, grid, 256, 0, stream, a))) return rc;
        const void* result = nullptr;
        if (!bluestein) {
            if ((rc = c2c_rows<R>(p, N, w1, w2, nc, rowlen, stream))) return rc;
            result = w2;
        } else {
            if ((rc = c2c_rows<R>(p, rowlen, w1, w2, nc, rowlen, stream))) return rc;
            BigMulArgs m{w2, blu.bhat, rowlen, nc, blu.M};
            if ((rc = dev_launch(big_mul_entry<R>, grid, 256, 0, stream, m))) return rc;
            if ((rc = c2c_rows<R>(p, rowlen, w2, w1, nc, rowlen, stream))) return rc;
            result = w1;
        }
        a.ws = const_cast<void*>(result);
        grid = (unsigned)std::min<long long>((nc * (long long)a.n_out + 255) / 256, (long long)sms * 16);
        if ((rc = dev_launch(big_epi_kernel<R>, grid, 256, 0, stream, a))) return rc;
    }
    return 0;
}

// Runs body(in', out', dims', blk') once per index combination of the batch dims beyond the `lim` fastest ones
// (dims are normalised: fastest input stride first), with the pointers advanced accordingly.  The reference accepts any
// ndarray Dimension (src/lib.rs:105-115), so views with many non-mergeable dims must work on every path.
template <typename F>
static int peel_batch_dims(const std::vector<BDim>& dims, int lim, size_t ie, size_t oe, const void* in, void* out,
                           void* const* blk_ptr, int nblk, F&& body) {
    if ((int)dims.size() <= lim) return body(in, out, dims, blk_ptr);
    std::vector<BDim> inner(dims.begin(), dims.begin() + lim);
    std::vector<BDim> outer(dims.begin() + lim, dims.end());
    long long nouter = 1;
    for (auto& d : outer) nouter *= d.size;
    for (long long g = 0; g < nouter; ++g) {
        long long rem = g, io = 0, oo = 0;
        for (auto& d : outer) { long long r = rem % d.size; rem /= d.size; io += r * d.is; oo += r * d.os; }
        void* blk[8] = {nullptr};
        for (int i = 0; i < nblk && i < 8; ++i) blk[i] = (char*)blk_ptr[i] + oo * (long long)oe;
        int rc = body((const char*)in + io * (long long)ie, (char*)out + oo * (long long)oe, inner, nblk ? blk : blk_ptr);
        if (rc) return rc;
    }
    return 0;
}

template <typename R>
static int exec_device(ndfb_plan* p, const OpInfo& o, double extra_scale, const void* in, void* out, int ndim,
                       const size_t* shape_in, const ptrdiff_t* strides_in, const size_t* shape_out,
                       const ptrdiff_t* strides_out, int axis, stream_t stream) {
    int rc = dev_set(p->device);
    if (rc) return rc;
    std::vector<BDim> dims;
    for (int d = 0; d < ndim; ++d)
        if (d != axis) dims.push_back({(long long)shape_in[d], (long long)strides_in[d], (long long)strides_out[d]});
    for (auto& d : dims) if (d.size == 0) return 0;
    if (p->n == 0) return 0;
    const long long is_axis = strides_in[axis], os_axis = strides_out[axis];
    const double scale = o.scale * extra_scale;
    const size_t cs = sizeof(Cx<R>);

    if (o.tk == TK_DCT1 && p->n < 2) return fail(NDFB_E_UNSUPPORTED, "DCT-I needs n >= 2");
    if (p->n > (size_t)(1 << 30)) return fail(NDFB_E_UNSUPPORTED, "length %zu too long", p->n);
    Core* c = nullptr;
    bool single = true;
    {
        // estimate lane slots before building the (possibly huge) tables
        long long N_est = (long long)p->n;
        if (o.tk == TK_R2C_EVEN || o.tk == TK_C2R_EVEN || o.tk == TK_DCT2_EVEN || o.tk == TK_DCT3_EVEN || o.tk == TK_DCT4_EVEN) N_est = p->n / 2;
        if (o.tk == TK_DCT4_ODD) N_est = 2 * (long long)p->n;
        const size_t cap = dev_smem_cap(p->device) - 1024;
        if ((size_t)N_est * cs + 64 > cap) single = false;
    }
    // test hook: exercise the two-pass path at sizes the emulator/tests can afford
    if (single && o.tk == TK_C2C && !o.os_blk && p->n >= 4 && std::getenv("NDFB_FORCE_FOUR_STEP")) {
        bool composite = false;
        for (size_t d = 2; d * d <= p->n; ++d) if (p->n % d == 0) composite = true;
        if (composite && is_smooth((long long)p->n)) single = false;
    }
    // long strided columns whose single-pass tile would be narrower than one 32-byte sector per row: two passes over
    // full-width rows beat one pass over half sectors (c2 axis 0: 8192-point c64 columns)
    if (single && o.tk == TK_C2C && !o.os_blk && is_smooth((long long)p->n) && p->n >= 1024) {
        std::vector<BDim> nd = dims;
        normalize_dims(nd);
        long long nl = 1;
        for (auto& d : nd) nl *= d.size;
        const bool strided = !nd.empty() && nl > 1 && llabs_(nd[0].is) < llabs_(is_axis) && llabs_(nd[0].os) < llabs_(os_axis);
        if (strided) {
            const SfftEntry* e = find_sfft(sizeof(R) == 8, (int)p->n, true, nl);
            const char* ov = std::getenv("NDFB_STRIDED_FOURSTEP");
            bool narrow = e ? (size_t)e->L * cs < 32 : true;
#ifndef NDFB_EMU
            if (!e && jit_enabled()) {   // a run-time compiled single-pass schedule with a wide enough tile beats two passes
                JitSched js;
                if (jit_plan((int)p->n, sizeof(R) == 8, true, false, nl, &js)) narrow = (size_t)js.L * cs < 32;
            }
#endif
            if (ov) narrow = ov[0] == '1';
            if (narrow) single = false;
        }
    }
    if (single) {
        c = get_core(p, o.tk, (int)p->n);
        if (!fits_one_tile(p, c->t, cs)) single = false;
    }
    if (!single && o.os_blk) return fail(NDFB_E_UNSUPPORTED, "split output axis is only available for single-pass complex transforms");
    if (single && !o.os_blk && std::getenv("NDFB_FORCE_STAGED") && p->n >= 3) single = false;
    const size_t ie = (o.in_complex ? 2 : 1) * sizeof(R), oe = (o.out_complex ? 2 : 1) * sizeof(R);
    normalize_dims(dims);
    if (!single) {
        // the multi-pass paths index fewer batch dims in their kernels (the passes add dims of their own): peel the rest
        const bool four_step = o.tk == TK_C2C && is_smooth((long long)p->n) && !std::getenv("NDFB_FORCE_STAGED");
        int lim = kMaxBatchDims;
        if (four_step) lim = ((long long)p->n > (1LL << 18) || std::getenv("NDFB_FS_CAP")) ? kMaxBatchDims - 2 : kMaxBatchDims - 1;   // three passes add two dims
        return peel_batch_dims(dims, lim, ie, oe, in, out, nullptr, 0, [&](const void* pin, void* pout, const std::vector<BDim>& d2, void* const*) -> int {
            if (four_step) return exec_four_step<R>(p, (long long)p->n, o.conj_in != 0, scale, pin, pout, d2, is_axis, os_axis, stream);
            return exec_big<R>(p, o, scale, pin, pout, d2, is_axis, os_axis, stream);
        });
    }
    if ((rc = ensure_device<R>(p, c))) return rc;
    // more batch dims than the kernel indexes: the slowest ones are peeled on the host (scattered block pointers move along)
    return peel_batch_dims(dims, kMaxBatchDims, ie, oe, in, out, o.blk_ptr, o.nblk_ptr, [&](const void* pin, void* pout, const std::vector<BDim>& d2, void* const* blk) -> int {
        LaunchSpec s;
        s.core = c; s.in = pin; s.out = pout; s.dims = d2; s.is_axis = is_axis; s.os_axis = os_axis;
        s.conj_in = o.conj_in; s.conj_out = o.conj_out; s.scale = scale;
        s.os_blk = o.os_blk; s.os_blk_stride = o.os_blk_stride; s.nblk_ptr = o.nblk_ptr;
        for (int i = 0; i < 8; ++i) s.blk_ptr[i] = blk ? blk[i] : nullptr;
        return o.tk == TK_C2C ? launch_c2c<R>(p, s, stream) : launch_real<R>(p, s, stream);
    });
}

// byte span [lo, hi) touched by a strided array, relative to its base pointer
static void span_of(int ndim, const size_t* shape, const ptrdiff_t* strides, size_t elem, long long* lo, long long* hi, bool* dense) {
    long long mn = 0, mx = 0, count = 1;
    for (int d = 0; d < ndim; ++d) {
        if (shape[d] == 0) { *lo = 0; *hi = 0; *dense = true; return; }
        long long ext = (long long)(shape[d] - 1) * (long long)strides[d];
        if (ext < 0) mn += ext; else mx += ext;
        count *= (long long)shape[d];
    }
    *lo = mn * (long long)elem;
    *hi = (mx + 1) * (long long)elem;
    *dense = (mx - mn + 1) == count;
}

static bool is_c_order(int ndim, const size_t* shape, const ptrdiff_t* strides) {
    long long expect = 1;
    for (int d = ndim - 1; d >= 0; --d) {
        if (shape[d] != 1 && strides[d] != expect) return false;
        expect *= (long long)shape[d];
    }
    return true;
}

static std::vector<ptrdiff_t> c_strides_of(int ndim, const size_t* shape) {
    std::vector<ptrdiff_t> st(ndim);
    long long acc = 1;
    for (int d = ndim - 1; d >= 0; --d) { st[d] = acc; acc *= (long long)shape[d]; }
    return st;
}

// A host array as the device path wants it.  Views with gaps between their elements (slices, steps: the siblings of
// ndarray's multi_slice_mut / split_at share the allocation) are packed into a private dense C-order buffer, so the
// library neither reads nor — for outputs — ever WRITES a byte outside the view's logical elements
// (the reference only touches lane elements, src/lib.rs:119-163).  Dense arrays (any order, any stride signs) are used as they are.
struct HostArray {
    const void* user = nullptr;          // caller's pointer (element [0, 0, ...])
    void* base = nullptr;                // pointer the device path is handed for element [0, 0, ...] (user or packed)
    std::vector<ptrdiff_t> strides;      // strides matching `base`
    std::unique_ptr<char[]> packed;      // owns the dense copy when the view has gaps (uninitialised: every element is written)
    bool is_packed = false;
    long long lo = 0, hi = 0;            // byte span relative to base
    void init(const void* ptr, int ndim, const size_t* shape, const ptrdiff_t* st, size_t elem, bool gather) {
        user = ptr;
        bool dense;
        span_of(ndim, shape, st, elem, &lo, &hi, &dense);
        if (dense) { base = const_cast<void*>(ptr); strides.assign(st, st + ndim); return; }
        is_packed = true;
        size_t total = elem;
        for (int d = 0; d < ndim; ++d) total *= shape[d];
        packed.reset(new char[total ? total : 1]);
        strides = c_strides_of(ndim, shape);
        base = packed.get();
        lo = 0; hi = (long long)total;
        if (gather) host_nd_copy(true, packed.get(), const_cast<void*>(ptr), ndim, shape, st, elem);
    }
    void scatter_back(int ndim, const size_t* shape, const ptrdiff_t* user_strides, size_t elem) {
        if (is_packed) host_nd_copy(false, packed.get(), const_cast<void*>(user), ndim, shape, user_strides, elem);
    }
};

#ifndef NDFB_EMU
struct HostPipe {
    cudaStream_t s[3] = {nullptr, nullptr, nullptr};
    static constexpr int kMaxChunks = 16;
    cudaEvent_t ev_in[kMaxChunks], ev_k[kMaxChunks];
    int device = -1;
    StageRing ring;
    int init(int dev) {
        if (device != dev) {
            for (int i = 0; i < 3; ++i) NDFB_CUDA(cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking));
            for (int i = 0; i < kMaxChunks; ++i) {
                NDFB_CUDA(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
                NDFB_CUDA(cudaEventCreateWithFlags(&ev_k[i], cudaEventDisableTiming));
            }
            device = dev;
        }
        return ring.ensure(dev);
    }
};
static thread_local HostPipe g_pipe;

// Large standard-layout host arrays: split along a non-transformed dim and pipeline
//     [host threads: caller memory -> pinned slot] -> H2D(c) | kernel(c-1) | D2H(c-2) -> [pinned slot -> caller memory]
// on three streams, so both PCIe directions, the GPU work and the host-side staging copies overlap.
template <typename R>
static int exec_host_pipelined(ndfb_plan* p, const OpInfo& o, double extra_scale, const void* in, void* out, void* din, void* dout,
                               int ndim, const size_t* shape_in, const ptrdiff_t* strides_in, const size_t* shape_out,
                               const ptrdiff_t* strides_out, int axis, size_t ie, size_t oe, int* done) {
    *done = 0;
    if (ndim < 2 || std::getenv("NDFB_NO_HOST_PIPELINE")) return 0;
    if (!is_c_order(ndim, shape_in, strides_in) || !is_c_order(ndim, shape_out, strides_out)) return 0;
    size_t total_in = ie, total_out = oe;
    for (int d = 0; d < ndim; ++d) { total_in *= shape_in[d]; total_out *= shape_out[d]; }
    if (total_in + total_out < (size_t)(16u << 20)) return 0;
    const int d = axis == 0 ? 1 : 0;               // split dim: outermost non-transformed dim
    const size_t nd = shape_in[d];
    int K = HostPipe::kMaxChunks;   // 16 pieces: fill + drain of the three-stage pipeline cost 2/16 of a transfer
    if (const char* e = std::getenv("NDFB_HOST_CHUNKS")) K = std::max(2, std::min(HostPipe::kMaxChunks, atoi(e)));
    while (K > 1 && (nd / K < 1 || (total_in / K) < (size_t)(2u << 20))) K /= 2;
    // column pieces are 2-D copies: keep every row segment >= 16 KiB or the DMA engines lose bandwidth
    // (8192 x 8192 c64, axis 0: 16 pieces 15.2 ms, 4 pieces 13.6 ms per call)
    if (d != 0 && !std::getenv("NDFB_HOST_CHUNKS"))
        while (K > 2 && (nd / K) * (size_t)strides_in[d] * ie < (size_t)(16u << 10)) K /= 2;
    if (K < 2 || nd < (size_t)K) return 0;
    int rc = g_pipe.init(p->device);
    if (rc) return rc;
    StageRing& ring = g_pipe.ring;
    const bool in_pageable = host_ptr_pageable(in), out_pageable = host_ptr_pageable(out);
    // 2-D copy geometry of one chunk [lo, hi) of dim d:  d == 0: one contiguous range; d == 1 (axis 0): shape[0] rows
    std::vector<size_t> shi(shape_in, shape_in + ndim), sho(shape_out, shape_out + ndim);
    for (int c = 0; c < K; ++c) {
        const size_t lo = nd * c / K, hi = nd * (c + 1) / K;
        if (hi == lo) continue;
        shi[d] = sho[d] = hi - lo;
        const size_t ioff = lo * (size_t)strides_in[d] * ie, ooff = lo * (size_t)strides_out[d] * oe;
        const size_t iw = (hi - lo) * (size_t)strides_in[d] * ie, ow = (hi - lo) * (size_t)strides_out[d] * oe;
        const size_t irows = d == 0 ? 1 : shape_in[0], orows = d == 0 ? 1 : shape_out[0];
        const size_t ipitch = d == 0 ? iw : (size_t)strides_in[0] * ie, opitch = d == 0 ? ow : (size_t)strides_out[0] * oe;
        if ((rc = ring.h2d((char*)din + ioff, ipitch, (const char*)in + ioff, ipitch, iw, irows, in_pageable, g_pipe.s[0]))) { cudaDeviceSynchronize(); g_pipe.ring.pend.clear(); return rc; }
        NDFB_CUDA(cudaEventRecord(g_pipe.ev_in[c], g_pipe.s[0]));
        NDFB_CUDA(cudaStreamWaitEvent(g_pipe.s[1], g_pipe.ev_in[c], 0));
        rc = exec_device<R>(p, o, extra_scale, (const char*)din + ioff, (char*)dout + ooff, ndim, shi.data(), strides_in, sho.data(),
                            strides_out, axis, g_pipe.s[1]);
        if (rc) { cudaDeviceSynchronize(); g_pipe.ring.pend.clear(); return rc; }
        NDFB_CUDA(cudaEventRecord(g_pipe.ev_k[c], g_pipe.s[1]));
        NDFB_CUDA(cudaStreamWaitEvent(g_pipe.s[2], g_pipe.ev_k[c], 0));
        if ((rc = ring.d2h((char*)out + ooff, opitch, (const char*)dout + ooff, opitch, ow, orows, out_pageable, g_pipe.s[2]))) { cudaDeviceSynchronize(); g_pipe.ring.pend.clear(); return rc; }
        if ((rc = ring.drain(false))) { cudaDeviceSynchronize(); g_pipe.ring.pend.clear(); return rc; }
    }
    if ((rc = ring.drain(true))) return rc;
    NDFB_CUDA(cudaStreamSynchronize(g_pipe.s[2]));
    *done = 1;
    return 0;
}
#endif

template <typename R>
static int exec_any(ndfb_plan* p, const OpInfo& o, double extra_scale, const void* in, void* out, int ndim,
                    const size_t* shape_in, const ptrdiff_t* strides_in, const size_t* shape_out,
                    const ptrdiff_t* strides_out, int axis, int mem, void* stream_v) {
    stream_t stream = (stream_t)stream_v;
    if (mem == NDFB_MEM_DEVICE) {
        g_pool.begin_call(stream);
        const int rc = exec_device<R>(p, o, extra_scale, in, out, ndim, shape_in, strides_in, shape_out, strides_out, axis, stream);
        g_pool.end_call();
        return rc;
    }
    // host arrays: dense views move as their byte span (strides unchanged), views with gaps as packed logical elements
    const size_t ie = (o.in_complex ? 2 : 1) * sizeof(R), oe = (o.out_complex ? 2 : 1) * sizeof(R);
    for (int d = 0; d < ndim; ++d) if (shape_in[d] == 0 || shape_out[d] == 0) return 0;
    HostArray hi_, ho_;
    hi_.init(in, ndim, shape_in, strides_in, ie, /*gather=*/true);
    ho_.init(out, ndim, shape_out, strides_out, oe, /*gather=*/false);
    const ptrdiff_t* si = hi_.strides.data();
    const ptrdiff_t* so = ho_.strides.data();
    const long long ilo = hi_.lo, ihi = hi_.hi, olo = ho_.lo, ohi = ho_.hi;
    if (ihi == ilo || ohi == olo) return 0;
    int rc = dev_set(p->device);
    if (rc) return rc;
    void *din = nullptr, *dout = nullptr;
    if ((rc = g_pool.get(0, p->device, (size_t)(ihi - ilo), &din))) return rc;
    if ((rc = g_pool.get(1, p->device, (size_t)(ohi - olo), &dout))) return rc;
#ifndef NDFB_EMU
    {
        int done = 0;
        rc = exec_host_pipelined<R>(p, o, extra_scale, hi_.base, ho_.base, din, dout, ndim, shape_in, si, shape_out, so, axis, ie, oe, &done);
        if (rc) return rc;
        if (done) { ho_.scatter_back(ndim, shape_out, strides_out, oe); return 0; }
    }
#endif
    if ((rc = dev_h2d(din, (const char*)hi_.base + ilo, (size_t)(ihi - ilo), stream))) return rc;
    rc = exec_device<R>(p, o, extra_scale, (const char*)din - ilo, (char*)dout - olo, ndim, shape_in, si, shape_out, so, axis, stream);
    if (rc) return rc;
    if ((rc = dev_d2h((char*)ho_.base + olo, dout, (size_t)(ohi - olo), stream))) return rc;
    if ((rc = dev_sync(stream))) return rc;
    ho_.scatter_back(ndim, shape_out, strides_out, oe);
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// multi-axis chains (ndfb_exec_chain): intermediates stay on the device
// ------------------------------------------------------------------------------------------------------
struct ChainStep { ndfb_plan* p; OpInfo o; int axis; };
struct ChainView {
    void* ptr = nullptr;
    std::vector<size_t> shape;
    std::vector<ptrdiff_t> strides;
    int where = 0;   // 0: caller's input, 1: caller's output, 2/3: workspace slot 7/8
};

static std::vector<ptrdiff_t> c_strides(const std::vector<size_t>& shape) {
    std::vector<ptrdiff_t> st(shape.size());
    long long acc = 1;
    for (int d = (int)shape.size() - 1; d >= 0; --d) { st[d] = acc; acc *= (long long)shape[d]; }
    return st;
}

// Where step i writes: the caller's output as soon as every later step keeps shape and element type (they then run in
// place on it), else a contiguous workspace (in place on the workspace when the step itself keeps shape and type).
template <typename R>
static int chain_dst(const std::vector<ChainStep>& st, int i, const ChainView& src, ChainView* dst, const ChainView& out, int device) {
    const int n = (int)st.size();
    bool rest_keeps = true;
    for (int j = i + 1; j < n; ++j)
        if (st[j].o.in_complex != st[j].o.out_complex || st[j].o.n_in != st[j].o.n_out) rest_keeps = false;
    if (i == n - 1 || rest_keeps) { *dst = out; return 0; }
    const OpInfo& o = st[i].o;
    const bool keeps = o.in_complex == o.out_complex && o.n_in == o.n_out;
    if (keeps && src.where >= 2) { *dst = src; return 0; }
    dst->shape = src.shape;
    dst->shape[st[i].axis] = (size_t)o.n_out;
    dst->strides = c_strides(dst->shape);
    dst->where = src.where == 2 ? 3 : 2;
    size_t bytes = (o.out_complex ? 2 : 1) * sizeof(R);
    for (size_t v : dst->shape) bytes *= v;
    return g_pool.get(dst->where == 2 ? 7 : 8, device, bytes, &dst->ptr);
}

template <typename R>
static int chain_device(const std::vector<ChainStep>& st, const ChainView& in, const ChainView& out, stream_t stream,
                        int first = 0, int last = -1, ChainView* cur_io = nullptr) {
    // runs steps [first, last]; cur_io carries the current array between partial runs (host pipeline)
    const int n = (int)st.size();
    if (last < 0) last = n - 1;
    ChainView cur = cur_io && first > 0 ? *cur_io : in;
    for (int i = first; i <= last; ++i) {
        ChainView dst;
        int rc = chain_dst<R>(st, i, cur, &dst, out, st[i].p->device);
        if (rc) return rc;
        rc = exec_device<R>(st[i].p, st[i].o, 1.0, cur.ptr, dst.ptr, (int)cur.shape.size(), cur.shape.data(), cur.strides.data(),
                            dst.shape.data(), dst.strides.data(), st[i].axis, stream);
        if (rc) return rc;
        cur = dst;
    }
    if (cur_io) *cur_io = cur;
    return 0;
}

#ifndef NDFB_EMU
// One chunk [lo, hi) of dim d of a C-ordered array as a 2-D copy (d == 0: one contiguous range; d == 1: shape[0] rows).
struct ChunkGeom { size_t off, width, rows, pitch; };
static ChunkGeom chunk_geom(const ChainView& v, int d, size_t lo, size_t hi, size_t elem) {
    ChunkGeom g;
    g.off = lo * (size_t)v.strides[d] * elem;
    g.width = (hi - lo) * (size_t)v.strides[d] * elem;
    g.rows = d == 0 ? 1 : v.shape[0];
    g.pitch = d == 0 ? g.width : (size_t)v.strides[0] * elem;
    return g;
}
static int pipeline_pieces(size_t nd, size_t total_bytes, size_t seg_bytes_per_index, bool two_d) {
    int K = HostPipe::kMaxChunks;
    if (const char* e = std::getenv("NDFB_HOST_CHUNKS")) return std::max(1, std::min(HostPipe::kMaxChunks, std::min((int)nd, atoi(e))));
    while (K > 1 && (nd / K < 1 || (total_bytes / K) < (size_t)(2u << 20))) K /= 2;
    if (two_d) while (K > 2 && (nd / K) * seg_bytes_per_index < (size_t)(16u << 10)) K /= 2;
    return K;
}

// Host arrays, >= 2 steps: upload pieces while the first step runs on the pieces already there; middle steps on the
// whole array; the last step runs piece by piece with the download of the finished pieces behind it.
template <typename R>
static int chain_host_pipelined(const std::vector<ChainStep>& st, const void* hin, void* hout, const ChainView& din, const ChainView& dout,
                                size_t ie, size_t oe, int* done) {
    *done = 0;
    const int n = (int)st.size(), ndim = (int)din.shape.size();
    if (n < 2 || ndim < 2 || std::getenv("NDFB_NO_HOST_PIPELINE")) return 0;
    if (!is_c_order(ndim, din.shape.data(), din.strides.data()) || !is_c_order(ndim, dout.shape.data(), dout.strides.data())) return 0;
    size_t total_in = ie, total_out = oe;
    for (int d = 0; d < ndim; ++d) { total_in *= din.shape[d]; total_out *= dout.shape[d]; }
    if (total_in + total_out < (size_t)(16u << 20)) return 0;
    const int d0 = st[0].axis == 0 ? 1 : 0, dl = st[n - 1].axis == 0 ? 1 : 0;
    const int K0 = pipeline_pieces(din.shape[d0], total_in, (size_t)din.strides[d0] * ie, d0 != 0);
    const int KL = pipeline_pieces(dout.shape[dl], total_out, (size_t)dout.strides[dl] * oe, dl != 0);
    if (K0 < 2 || KL < 2) return 0;
    ndfb_plan* p = st[0].p;
    int rc = g_pipe.init(p->device);
    if (rc) return rc;
    StageRing& ring = g_pipe.ring;
    const bool in_pageable = host_ptr_pageable(hin), out_pageable = host_ptr_pageable(hout);
    // first step, piece by piece behind the upload
    ChainView d1;
    if ((rc = chain_dst<R>(st, 0, din, &d1, dout, p->device))) return rc;
    const size_t e1 = (st[0].o.out_complex ? 2 : 1) * sizeof(R);
    {
        const size_t nd = din.shape[d0];
        ChainView a = din, b = d1;
        for (int c = 0; c < K0; ++c) {
            const size_t lo = nd * c / K0, hi = nd * (c + 1) / K0;
            if (hi == lo) continue;
            const ChunkGeom g = chunk_geom(din, d0, lo, hi, ie);
            if ((rc = ring.h2d((char*)din.ptr + g.off, g.pitch, (const char*)hin + g.off, g.pitch, g.width, g.rows, in_pageable, g_pipe.s[0]))) { cudaDeviceSynchronize(); g_pipe.ring.pend.clear(); return rc; }
            NDFB_CUDA(cudaEventRecord(g_pipe.ev_in[c], g_pipe.s[0]));
            NDFB_CUDA(cudaStreamWaitEvent(g_pipe.s[1], g_pipe.ev_in[c], 0));
            a.shape[d0] = b.shape[d0] = hi - lo;
            rc = exec_device<R>(p, st[0].o, 1.0, (const char*)din.ptr + g.off, (char*)d1.ptr + lo * (size_t)d1.strides[d0] * e1, ndim,
                                a.shape.data(), a.strides.data(), b.shape.data(), b.strides.data(), st[0].axis, g_pipe.s[1]);
            if (rc) { cudaDeviceSynchronize(); g_pipe.ring.pend.clear(); return rc; }
        }
    }
    // middle steps on the whole array
    ChainView cur = d1;
    if (n > 2 && (rc = chain_device<R>(st, din, dout, g_pipe.s[1], 1, n - 2, &cur))) { cudaDeviceSynchronize(); g_pipe.ring.pend.clear(); return rc; }
    // last step, piece by piece ahead of the download
    {
        const ChainStep& L = st[n - 1];
        const size_t el = (L.o.in_complex ? 2 : 1) * sizeof(R);
        const size_t nd = dout.shape[dl];
        ChainView a = cur, b = dout;
        for (int c = 0; c < KL; ++c) {
            const size_t lo = nd * c / KL, hi = nd * (c + 1) / KL;
            if (hi == lo) continue;
            const ChunkGeom g = chunk_geom(dout, dl, lo, hi, oe);
            a.shape[dl] = b.shape[dl] = hi - lo;
            rc = exec_device<R>(L.p, L.o, 1.0, (const char*)cur.ptr + lo * (size_t)cur.strides[dl] * el, (char*)dout.ptr + g.off, ndim,
                                a.shape.data(), a.strides.data(), b.shape.data(), b.strides.data(), L.axis, g_pipe.s[1]);
            if (rc) { cudaDeviceSynchronize(); g_pipe.ring.pend.clear(); return rc; }
            NDFB_CUDA(cudaEventRecord(g_pipe.ev_k[c], g_pipe.s[1]));
            NDFB_CUDA(cudaStreamWaitEvent(g_pipe.s[2], g_pipe.ev_k[c], 0));
            if ((rc = ring.d2h((char*)hout + g.off, g.pitch, (const char*)dout.ptr + g.off, g.pitch, g.width, g.rows, out_pageable, g_pipe.s[2]))) { cudaDeviceSynchronize(); g_pipe.ring.pend.clear(); return rc; }
            if ((rc = ring.drain(false))) { cudaDeviceSynchronize(); g_pipe.ring.pend.clear(); return rc; }
        }
    }
    if ((rc = ring.drain(true))) return rc;
    NDFB_CUDA(cudaStreamSynchronize(g_pipe.s[2]));
    NDFB_CUDA(cudaStreamSynchronize(g_pipe.s[1]));
    *done = 1;
    return 0;
}
#endif

template <typename R>
static int chain_any(const std::vector<ChainStep>& st, const void* in, void* out, int ndim, const size_t* shape_in,
                     const ptrdiff_t* strides_in, const size_t* shape_out, const ptrdiff_t* strides_out, int mem, void* stream_v) {
    stream_t stream = (stream_t)stream_v;
    ndfb_plan* p = st[0].p;
    int rc = dev_set(p->device);
    if (rc) return rc;
    ChainView vi, vo;
    vi.shape.assign(shape_in, shape_in + ndim); vi.strides.assign(strides_in, strides_in + ndim); vi.where = 0;
    vo.shape.assign(shape_out, shape_out + ndim); vo.strides.assign(strides_out, strides_out + ndim); vo.where = 1;
    if (mem == NDFB_MEM_DEVICE) {
        vi.ptr = const_cast<void*>(in); vo.ptr = out;
        g_pool.begin_call(stream);
        rc = chain_device<R>(st, vi, vo, stream);
        g_pool.end_call();
        return rc;
    }
    const size_t ie = (st.front().o.in_complex ? 2 : 1) * sizeof(R), oe = (st.back().o.out_complex ? 2 : 1) * sizeof(R);
    for (int d = 0; d < ndim; ++d) if (shape_in[d] == 0 || shape_out[d] == 0) return 0;
    HostArray hi_, ho_;   // views with gaps travel as packed logical elements (see exec_any)
    hi_.init(in, ndim, shape_in, strides_in, ie, /*gather=*/true);
    ho_.init(out, ndim, shape_out, strides_out, oe, /*gather=*/false);
    vi.strides = hi_.strides; vo.strides = ho_.strides;
    const long long ilo = hi_.lo, ihi = hi_.hi, olo = ho_.lo, ohi = ho_.hi;
    if (ihi == ilo || ohi == olo) return 0;
    void *din = nullptr, *dout = nullptr;
    if ((rc = g_pool.get(0, p->device, (size_t)(ihi - ilo), &din))) return rc;
    if ((rc = g_pool.get(1, p->device, (size_t)(ohi - olo), &dout))) return rc;
    vi.ptr = (char*)din - ilo; vo.ptr = (char*)dout - olo;
#ifndef NDFB_EMU
    {
        int done = 0;
        rc = chain_host_pipelined<R>(st, hi_.base, ho_.base, vi, vo, ie, oe, &done);
        if (rc) return rc;
        if (done) { ho_.scatter_back(ndim, shape_out, strides_out, oe); return 0; }
    }
#endif
    if ((rc = dev_h2d(din, (const char*)hi_.base + ilo, (size_t)(ihi - ilo), stream))) return rc;
    if ((rc = chain_device<R>(st, vi, vo, stream))) return rc;
    if ((rc = dev_d2h((char*)ho_.base + olo, dout, (size_t)(ohi - olo), stream))) return rc;
    if ((rc = dev_sync(stream))) return rc;
    ho_.scatter_back(ndim, shape_out, strides_out, oe);
    return 0;
}

}  // namespace ndfb

// ------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------
using namespace ndfb;

extern "C" {

int ndfb_plan_create(ndfb_plan** out, int kind, int dtype, size_t n, int device) {
    if (!out) return fail(NDFB_E_INVALID, "null plan pointer");
    if (kind < NDFB_C2C || kind > NDFB_DCT) return fail(NDFB_E_INVALID, "unknown plan kind %d", kind);
    if (dtype != NDFB_F32 && dtype != NDFB_F64) return fail(NDFB_E_INVALID, "unknown dtype %d", dtype);
    if (device < 0) return fail(NDFB_E_INVALID, "negative device index");
    ndfb_plan* p = new (std::nothrow) ndfb_plan();
    if (!p) return fail(NDFB_E_ALLOC, "out of host memory");
    p->kind = kind; p->dtype = dtype; p->device = device; p->n = n;
    *out = p;
    return NDFB_OK;
}

void ndfb_plan_destroy(ndfb_plan* p) {
    if (!p) return;
    for (auto& kv : p->cores) {
        DeviceTables& d = kv.second->d;
        void* ptrs[] = {d.tw, d.tabA, d.tabB, d.blu_c, d.blu_bhat, d.perm};
        for (void* q : ptrs) if (q) dev_free(q);
        for (auto& kv2 : d.sfft_tw) if (kv2.second) dev_free(kv2.second);
    }
    for (auto& kv : p->fs) { if (kv.second.lo) dev_free(kv.second.lo); if (kv.second.hi) dev_free(kv.second.hi); }
    for (auto& kv : p->fsq) if (kv.second) dev_free(kv.second);
    for (auto& kv : p->bigblu) { if (kv.second.chirp) dev_free(kv.second.chirp); if (kv.second.bhat) dev_free(kv.second.bhat); }
    delete p;
}

static int check_call(const ndfb_plan* p, const OpInfo& o, int ndim, const size_t* shape_in, const size_t* shape_out, int axis) {
    // order of checks mirrors the reference: axis index (src/lib.rs:116), lane sizes (assert_size: input first,
    // then output; src/lib.rs:314-315, 498-499, 507-508, 689-690), then ndarray's Zip shape check.
    if (axis < 0 || axis >= ndim) return fail(NDFB_E_AXIS, "axis %d out of range for %d-dimensional array", axis, ndim);
    if ((long long)shape_in[axis] != o.n_in)
        return fail(NDFB_E_SIZE_MISMATCH, "Size mismatch in %s, got %zu expected %lld", o.what, shape_in[axis], o.n_in);
    if ((long long)shape_out[axis] != o.n_out)
        return fail(NDFB_E_SIZE_MISMATCH, "Size mismatch in %s, got %zu expected %lld", o.what, shape_out[axis], o.n_out);
    for (int d = 0; d < ndim; ++d)
        if (d != axis && shape_in[d] != shape_out[d])
            return fail(NDFB_E_SHAPE, "input and output shapes differ along dimension %d (%zu vs %zu)", d, shape_in[d], shape_out[d]);
    (void)p;
    return 0;
}

static int exec_common(const ndfb_plan* plan, int op, int norm, double extra_scale, size_t out_block, ptrdiff_t out_block_stride,
                       const void* in, void* out, int ndim, const size_t* shape_in, const ptrdiff_t* strides_in,
                       const size_t* shape_out, const ptrdiff_t* strides_out, int axis, int mem, void* stream,
                       int nblk_ptr = 0, void* const* blk_ptrs = nullptr);

int ndfb_exec_scaled(const ndfb_plan* plan, int op, int norm, double extra_scale, const void* in, void* out, int ndim,
                     const size_t* shape_in, const ptrdiff_t* strides_in, const size_t* shape_out,
                     const ptrdiff_t* strides_out, int axis, int mem, void* stream) {
    return exec_common(plan, op, norm, extra_scale, 0, 0, in, out, ndim, shape_in, strides_in, shape_out, strides_out, axis, mem, stream);
}

int ndfb_exec_scatter_out(const ndfb_plan* plan, int op, int norm, double extra_scale, size_t out_block, int nblocks,
                          void* const* block_ptrs, const void* in, int ndim, const size_t* shape_in, const ptrdiff_t* strides_in,
                          const size_t* shape_out, const ptrdiff_t* strides_out, int axis, void* stream) {
    if (op != NDFB_OP_FFT && op != NDFB_OP_IFFT) return fail(NDFB_E_UNSUPPORTED, "scattered output blocks are only available for ndfft / ndifft");
    if (!block_ptrs || nblocks < 1 || nblocks > 8) return fail(NDFB_E_INVALID, "1..8 block pointers expected");
    if (out_block == 0 || !shape_out || axis < 0 || axis >= ndim || shape_out[axis] != out_block * (size_t)nblocks)
        return fail(NDFB_E_INVALID, "out_block * nblocks must equal the output lane length");
    return exec_common(plan, op, norm, extra_scale, out_block, 0, in, block_ptrs[0], ndim, shape_in, strides_in, shape_out, strides_out,
                       axis, NDFB_MEM_DEVICE, stream, nblocks, block_ptrs);
}

static int exec_common(const ndfb_plan* plan, int op, int norm, double extra_scale, size_t out_block, ptrdiff_t out_block_stride,
                       const void* in, void* out, int ndim, const size_t* shape_in, const ptrdiff_t* strides_in,
                       const size_t* shape_out, const ptrdiff_t* strides_out, int axis, int mem, void* stream,
                       int nblk_ptr, void* const* blk_ptrs) {
    SyncHintScope hint_scope;
    (void)hint_scope;
    if (!plan || !shape_in || !strides_in || !shape_out || !strides_out) return fail(NDFB_E_INVALID, "null argument");
    if (ndim < 1 || ndim > NDFB_MAX_DIMS) return fail(NDFB_E_INVALID, "ndim %d outside 1..%d", ndim, NDFB_MAX_DIMS);
    if (norm != NDFB_NORM_NONE && norm != NDFB_NORM_DEFAULT) return fail(NDFB_E_INVALID, "unknown norm %d", norm);
    if (mem != NDFB_MEM_HOST && mem != NDFB_MEM_DEVICE) return fail(NDFB_E_INVALID, "unknown mem %d", mem);
    ndfb_plan* p = const_cast<ndfb_plan*>(plan);  // lazily built caches are guarded by mutexes
    DeviceGuard device_guard;
    (void)device_guard;
    OpInfo o;
    int rc = op_info(p, op, norm, &o);
    if (rc) return rc;
    if ((rc = check_call(p, o, ndim, shape_in, shape_out, axis))) return rc;
    o.os_blk = (int)out_block; o.os_blk_stride = (long long)out_block_stride;
    o.nblk_ptr = nblk_ptr;
    for (int i = 0; i < nblk_ptr && i < 8; ++i) o.blk_ptr[i] = blk_ptrs[i];
    bool empty = false;
    for (int d = 0; d < ndim; ++d) if (shape_in[d] == 0 || shape_out[d] == 0) empty = true;
    if (empty) return NDFB_OK;
    if (!in || !out) return fail(NDFB_E_INVALID, "null data pointer");
    rc = p->dtype == NDFB_F32
             ? exec_any<float>(p, o, extra_scale, in, out, ndim, shape_in, strides_in, shape_out, strides_out, axis, mem, stream)
             : exec_any<double>(p, o, extra_scale, in, out, ndim, shape_in, strides_in, shape_out, strides_out, axis, mem, stream);
    if (g_sync_hint.signal_cnt || g_sync_hint.wait_cnt) {   // the call took a path that cannot honour the hint: say so
        g_sync_hint.clear();
        if (!rc) rc = fail(NDFB_E_UNSUPPORTED, "launch signal / wait hint was not consumed by this call");
    }
    return rc;
}

int ndfb_exec_split_out(const ndfb_plan* plan, int op, int norm, double extra_scale, size_t out_block, ptrdiff_t out_block_stride,
                        const void* in, void* out, int ndim, const size_t* shape_in, const ptrdiff_t* strides_in,
                        const size_t* shape_out, const ptrdiff_t* strides_out, int axis, void* stream) {
    if (op != NDFB_OP_FFT && op != NDFB_OP_IFFT) return fail(NDFB_E_UNSUPPORTED, "split output axis is only available for ndfft / ndifft");
    if (out_block == 0 || (axis >= 0 && axis < ndim && shape_out && shape_out[axis] % out_block))
        return fail(NDFB_E_INVALID, "out_block must divide the output lane length");
    return exec_common(plan, op, norm, extra_scale, out_block, out_block_stride, in, out, ndim, shape_in, strides_in, shape_out, strides_out,
                       axis, NDFB_MEM_DEVICE, stream);
}

int ndfb_exec(const ndfb_plan* plan, int op, int norm, const void* in, void* out, int ndim, const size_t* shape_in,
              const ptrdiff_t* strides_in, const size_t* shape_out, const ptrdiff_t* strides_out, int axis, int mem,
              void* stream) {
    return ndfb_exec_scaled(plan, op, norm, 1.0, in, out, ndim, shape_in, strides_in, shape_out, strides_out, axis, mem, stream);
}

int ndfb_exec_chain(const ndfb_step* steps, int nsteps, const void* in, void* out, int ndim, const size_t* shape_in,
                    const ptrdiff_t* strides_in, const size_t* shape_out, const ptrdiff_t* strides_out, int mem, void* stream) {
    SyncHintScope hint_scope;   // launch hints do not apply to chains
    (void)hint_scope;
    if (g_sync_hint.signal_cnt || g_sync_hint.wait_cnt) return fail(NDFB_E_UNSUPPORTED, "launch signal / wait hints apply to single ndfb_exec calls, not to chains");
    if (!steps || nsteps < 1 || nsteps > 16) return fail(NDFB_E_INVALID, "1..16 steps expected");
    if (!shape_in || !strides_in || !shape_out || !strides_out) return fail(NDFB_E_INVALID, "null argument");
    if (ndim < 1 || ndim > NDFB_MAX_DIMS) return fail(NDFB_E_INVALID, "ndim %d outside 1..%d", ndim, NDFB_MAX_DIMS);
    if (mem != NDFB_MEM_HOST && mem != NDFB_MEM_DEVICE) return fail(NDFB_E_INVALID, "unknown mem %d", mem);
    DeviceGuard device_guard;
    (void)device_guard;
    std::vector<ChainStep> st(nsteps);
    std::vector<size_t> cur(shape_in, shape_in + ndim);
    bool cur_complex = false;
    for (int i = 0; i < nsteps; ++i) {
        const ndfb_step& s = steps[i];
        if (!s.plan) return fail(NDFB_E_INVALID, "step %d: null plan", i);
        if (s.norm != NDFB_NORM_NONE && s.norm != NDFB_NORM_DEFAULT) return fail(NDFB_E_INVALID, "step %d: unknown norm %d", i, s.norm);
        st[i].p = const_cast<ndfb_plan*>(s.plan);
        st[i].axis = s.axis;
        if (st[i].p->dtype != st[0].p->dtype || st[i].p->device != st[0].p->device)
            return fail(NDFB_E_INVALID, "step %d: all plans of a chain must share dtype and device", i);
        int rc = op_info(st[i].p, s.op, s.norm, &st[i].o);
        if (rc) return rc;
        const OpInfo& o = st[i].o;
        if (s.axis < 0 || s.axis >= ndim) return fail(NDFB_E_AXIS, "axis %d out of range for %d-dimensional array", s.axis, ndim);
        if (i > 0 && o.in_complex != cur_complex)
            return fail(NDFB_E_INVALID, "step %d reads %s data but step %d wrote %s data", i, o.in_complex ? "complex" : "real", i - 1, cur_complex ? "complex" : "real");
        if ((long long)cur[s.axis] != o.n_in)
            return fail(NDFB_E_SIZE_MISMATCH, "Size mismatch in %s, got %zu expected %lld", o.what, cur[s.axis], o.n_in);
        cur[s.axis] = (size_t)o.n_out;
        cur_complex = o.out_complex;
    }
    for (int d = 0; d < ndim; ++d) {
        if (cur[d] == shape_out[d]) continue;
        bool transformed = false;
        const char* what = "fft";
        for (auto& c : st) if (c.axis == d) { transformed = true; what = c.o.what; }
        if (transformed) return fail(NDFB_E_SIZE_MISMATCH, "Size mismatch in %s, got %zu expected %zu", what, shape_out[d], cur[d]);
        return fail(NDFB_E_SHAPE, "input and output shapes differ along dimension %d (%zu vs %zu)", d, shape_in[d], shape_out[d]);
    }
    for (int d = 0; d < ndim; ++d) if (shape_in[d] == 0 || shape_out[d] == 0) return NDFB_OK;
    if (!in || !out) return fail(NDFB_E_INVALID, "null data pointer");
    if (st[0].p->dtype == NDFB_F32) return chain_any<float>(st, in, out, ndim, shape_in, strides_in, shape_out, strides_out, mem, stream);
    return chain_any<double>(st, in, out, ndim, shape_in, strides_in, shape_out, strides_out, mem, stream);
}

size_t ndfb_plan_describe(const ndfb_plan* plan, char* buf, size_t cap) {
    if (!plan) return 0;
    ndfb_plan* p = const_cast<ndfb_plan*>(plan);
    std::string s = "{\"kind\":" + std::to_string(p->kind) + ",\"dtype\":\"" + (p->dtype == NDFB_F32 ? "f32" : "f64") +
                    "\",\"n\":" + std::to_string(p->n) + ",\"ops\":[";
    int first_op = p->kind == NDFB_C2C ? NDFB_OP_FFT : p->kind == NDFB_R2C ? NDFB_OP_R2C : NDFB_OP_DCT1;
    int last_op = p->kind == NDFB_C2C ? NDFB_OP_IFFT : p->kind == NDFB_R2C ? NDFB_OP_C2R : NDFB_OP_DCT4;
    bool first = true;
    for (int op = first_op; op <= last_op; ++op) {
        OpInfo o;
        if (op_info(p, op, NDFB_NORM_DEFAULT, &o)) continue;
        if (!first) s += ",";
        first = false;
        s += "{\"op\":" + std::to_string(op) + ",\"tile_kind\":" + std::to_string(o.tk);
        const size_t cs = p->dtype == NDFB_F32 ? 8 : 16;
        long long N_est = (long long)p->n;
        if (o.tk == TK_R2C_EVEN || o.tk == TK_C2R_EVEN || o.tk == TK_DCT2_EVEN || o.tk == TK_DCT3_EVEN || o.tk == TK_DCT4_EVEN) N_est = p->n / 2;
        if (o.tk == TK_DCT4_ODD) N_est = 2 * (long long)p->n;
        if (o.tk == TK_DCT1) N_est = (long long)p->n - 1;
        if (p->n == 0 || (o.tk == TK_DCT1 && p->n < 2)) { s += ",\"family\":\"empty\"}"; continue; }
        if ((size_t)N_est * cs + 64 > 226 * 1024) {
            const bool staged_ok = true;   // every kind has a staged schedule (big_kernels.cuh)
            s += std::string(",\"family\":\"") + (o.tk == TK_C2C && is_smooth((long long)p->n) ? "four-step" : (staged_ok ? "staged" : "unsupported")) + "\",\"N\":" + std::to_string(N_est) + "}";
            continue;
        }
        Core* c = get_core(p, o.tk, (int)p->n);
        s += ",\"family\":\"" + std::string(c->t.M ? "bluestein" : "direct") + "\",\"N\":" + std::to_string(c->t.N) +
             ",\"M\":" + std::to_string(c->t.M) + ",\"radix\":[";
        for (size_t i = 0; i < c->t.radix.size(); ++i) s += (i ? "," : "") + std::to_string(c->t.radix[i]);
        s += "],\"lane_slots\":" + std::to_string(c->t.Bl) + "}";
    }
    s += "]}";
    if (buf && cap) {
        size_t ncopy = std::min(cap - 1, s.size());
        std::memcpy(buf, s.data(), ncopy);
        buf[ncopy] = 0;
    }
    return s.size() + 1;
}

int ndfb_jit_compile_check(int dtype, int rkind, size_t core_n, int cols, char* info, size_t cap) {
#ifdef NDFB_EMU
    (void)dtype; (void)rkind; (void)core_n; (void)cols; (void)info; (void)cap;
    return fail(NDFB_E_UNSUPPORTED, "no run-time compilation in the emulation build");
#else
    JitSched js;
    if (core_n > (size_t)(1 << 20) || !jit_plan((int)core_n, dtype == NDFB_F64, cols != 0, rkind >= 0, 1 << 20, &js))
        return fail(NDFB_E_UNSUPPORTED, "no run-time schedule for a %zu-point core (not 13-smooth, more than 4 passes, or too long for one CTA)", core_n);
    char expr[256];
    const char* R = dtype == NDFB_F64 ? "double" : "float";
    if (rkind < 0)
        snprintf(expr, sizeof expr, "ndfb::sfft_kernel<%s, ndfb::Sched<%d, %d, %d, %d, %d, %d>, %d, %s, %d>", R, js.N, js.TL, js.r[0], js.r[1], js.r[2], js.r[3], js.L, js.cols ? "true" : "false", js.minb);
    else
        snprintf(expr, sizeof expr, "ndfb::rsfft_kernel<%s, ndfb::Sched<%d, %d, %d, %d, %d, %d>, %d, %s, %d, %d>", R, js.N, js.TL, js.r[0], js.r[1], js.r[2], js.r[3], js.L, js.cols ? "true" : "false", rkind, js.minb);
    std::vector<char> cubin;
    std::string lowered, log;
    const int rc = jit_compile(expr, &cubin, &lowered, &log);
    if (info && cap) snprintf(info, cap, "{\"kernel\":\"%s\",\"threads\":%d,\"smem\":%zu,\"E\":%d,\"cubin_bytes\":%zu}", expr, js.threads, js.smem, js.E, cubin.size());
    return rc;
#endif
}

// ---- device memory / stream helpers for hosts without CUDA bindings (the Rust shim's DeviceArray and Stream) ----
int ndfb_device_alloc(void** ptr, size_t bytes, int device) {
    if (!ptr) return fail(NDFB_E_INVALID, "null pointer");
    DeviceGuard guard; (void)guard;
    int rc = dev_set(device);
    if (rc) return rc;
    return dev_malloc(ptr, bytes);
}
void ndfb_device_free(void* ptr) { if (ptr) dev_free(ptr); }
int ndfb_memcpy(void* dst, const void* src, size_t bytes, int kind, int device, void* stream) {
    if ((!dst || !src) && bytes) return fail(NDFB_E_INVALID, "null pointer");
    if (kind != NDFB_COPY_H2D && kind != NDFB_COPY_D2H && kind != NDFB_COPY_D2D) return fail(NDFB_E_INVALID, "unknown copy kind %d", kind);
    if (bytes == 0) return 0;
#ifdef NDFB_EMU
    (void)device; (void)stream;
    std::memcpy(dst, src, bytes);
    return 0;
#else
    DeviceGuard guard; (void)guard;
    int rc = dev_set(device);
    if (rc) return rc;
    const cudaMemcpyKind k = kind == NDFB_COPY_H2D ? cudaMemcpyHostToDevice : kind == NDFB_COPY_D2H ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    // pageable host memory moves through the calling thread's pinned ring in slot-sized pieces (hostio.h)
    if (kind != NDFB_COPY_D2D && bytes >= ((size_t)4 << 20)) {
        const void* host = kind == NDFB_COPY_H2D ? src : dst;
        if (host_ptr_pageable(host)) {
            if ((rc = g_pipe.init(device))) return rc;
            StageRing& ring = g_pipe.ring;
            cudaStream_t s = (cudaStream_t)stream;
            if (kind == NDFB_COPY_H2D) rc = ring.h2d(dst, bytes, src, bytes, bytes, 1, true, s);
            else { rc = ring.d2h(dst, bytes, src, bytes, bytes, 1, true, s); if (!rc) rc = ring.drain(true); }
            return rc;
        }
    }
    NDFB_CUDA(cudaMemcpyAsync(dst, src, bytes, k, (cudaStream_t)stream));
    if (kind == NDFB_COPY_D2H) NDFB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));   // the host may read `dst` on return
    return 0;
#endif
}
int ndfb_stream_create(void** stream, int device) {
    if (!stream) return fail(NDFB_E_INVALID, "null pointer");
#ifdef NDFB_EMU
    (void)device; *stream = nullptr; return 0;
#else
    DeviceGuard guard; (void)guard;
    int rc = dev_set(device);
    if (rc) return rc;
    cudaStream_t s;
    NDFB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (void*)s;
    return 0;
#endif
}
void ndfb_stream_destroy(void* stream) {
#ifndef NDFB_EMU
    if (stream) cudaStreamDestroy((cudaStream_t)stream);
#else
    (void)stream;
#endif
}
int ndfb_stream_sync(void* stream) { return dev_sync((stream_t)stream); }

void ndfb_hint_next_launch_smem(size_t bytes) { launch_smem_floor() = bytes; }
void ndfb_hint_next_launch_signal(void* counters, long long lanes_per_group) {
    g_sync_hint.signal_cnt = (unsigned*)counters; g_sync_hint.signal_group = lanes_per_group > 0 ? lanes_per_group : 1;
}
void ndfb_hint_next_launch_wait(const void* counters, long long lanes_per_group, unsigned need, int ctas_per_sm) {
    g_sync_hint.wait_cnt = (const unsigned*)counters; g_sync_hint.wait_group = lanes_per_group > 0 ? lanes_per_group : 1;
    g_sync_hint.wait_need = need; g_sync_hint.ctas_per_sm = ctas_per_sm;
}
const char* ndfb_last_error(void) { return g_err.c_str(); }
const char* ndfb_version(void) { return version_string(); }
uint64_t ndfb_launch_count(void) { return g_launches.load(); }
void ndfb_release_workspaces(void) {
    g_pool.release();
#ifndef NDFB_EMU
    g_pipe.ring.release();
#endif
}

}  // extern "C"
