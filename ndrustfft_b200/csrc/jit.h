// jit.h — run-time compiled Stockham schedules for lengths without an instantiated one.
//
// rustfft plans ANY length (src/lib.rs:294-304; "performs best on sizes which are multiples of 2 or 3", :245).  The
// library ships ~900 ahead-of-time instances for the lengths of the BASELINE configs; every other
// {2,3,5,7,11,13}-smooth length up to one CTA's shared memory gets the SAME register-resident kernels
// (sfft_kernel / rsfft_kernel, csrc/sfft_kernel.cuh) compiled for its own radix schedule the first time a plan
// needs it: NVRTC -> sm_100a cubin -> cudaLibraryLoadData, cached in memory and on disk
// ($NDFB_JIT_CACHE or ~/.cache/ndfft_b200).  The kernel sources are embedded in the library (jit_sources.inc,
// generated from the headers by tools/embed_src.py), so nothing but libnvrtc is needed at run time.
// Without libnvrtc (or with NDFB_NO_JIT=1) such lengths run the general tile kernel: slower, same results.
#pragma once
#ifndef NDFB_EMU
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "devapi.h"

namespace ndfb {

#include "jit_sources.inc"   // kJitSrc_common_h, kJitSrc_butterflies_cuh, kJitSrc_sfft_kernel_cuh

struct Nvrtc {
    typedef int (*create_t)(void**, const char*, const char*, int, const char* const*, const char* const*);
    typedef int (*destroy_t)(void**);
    typedef int (*compile_t)(void*, int, const char* const*);
    typedef int (*size_t_fn)(void*, size_t*);
    typedef int (*get_t)(void*, char*);
    typedef int (*addname_t)(void*, const char*);
    typedef int (*lowered_t)(void*, const char*, const char**);
    typedef const char* (*errstr_t)(int);
    void* h = nullptr;
    create_t create = nullptr; destroy_t destroy = nullptr; compile_t compile = nullptr;
    size_t_fn cubin_size = nullptr, log_size = nullptr; get_t cubin = nullptr, log = nullptr;
    addname_t add_name = nullptr; lowered_t lowered = nullptr; errstr_t errstr = nullptr;
    bool ok = false;
    static Nvrtc& get() {
        static Nvrtc n;
        static std::once_flag once;
        std::call_once(once, [] {
            const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
            for (const char* nm : names) { n.h = dlopen(nm, RTLD_NOW | RTLD_LOCAL); if (n.h) break; }
            if (!n.h) return;
            n.create = (create_t)dlsym(n.h, "nvrtcCreateProgram");
            n.destroy = (destroy_t)dlsym(n.h, "nvrtcDestroyProgram");
            n.compile = (compile_t)dlsym(n.h, "nvrtcCompileProgram");
            n.cubin_size = (size_t_fn)dlsym(n.h, "nvrtcGetCUBINSize");
            n.cubin = (get_t)dlsym(n.h, "nvrtcGetCUBIN");
            n.log_size = (size_t_fn)dlsym(n.h, "nvrtcGetProgramLogSize");
            n.log = (get_t)dlsym(n.h, "nvrtcGetProgramLog");
            n.add_name = (addname_t)dlsym(n.h, "nvrtcAddNameExpression");
            n.lowered = (lowered_t)dlsym(n.h, "nvrtcGetLoweredName");
            n.errstr = (errstr_t)dlsym(n.h, "nvrtcGetErrorString");
            n.ok = n.create && n.destroy && n.compile && n.cubin_size && n.cubin && n.log_size && n.log && n.add_name && n.lowered;
        });
        return n;
    }
};

inline uint64_t jit_hash(const std::string& s) {
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
    return h;
}

inline std::string jit_cache_dir() {
    std::string d;
    if (const char* e = std::getenv("NDFB_JIT_CACHE")) d = e;
    else if (const char* h = std::getenv("HOME")) d = std::string(h) + "/.cache/ndfft_b200";
    else d = "/tmp/ndfft_b200_cache";
    return d;
}

inline void jit_mkdirs(const std::string& d) {
    std::string cur;
    for (size_t i = 0; i <= d.size(); ++i) {
        if (i == d.size() || d[i] == '/') { if (!cur.empty()) mkdir(cur.c_str(), 0755); }
        if (i < d.size()) cur.push_back(d[i]);
    }
}

// Compiles `name_expr` (a kernel template instantiation from sfft_kernel.cuh) to a cubin; returns the lowered name too.
inline int jit_compile(const std::string& name_expr, std::vector<char>* cubin, std::string* lowered, std::string* log_out) {
    Nvrtc& n = Nvrtc::get();
    if (!n.ok) return fail(NDFB_E_UNSUPPORTED, "libnvrtc not available: cannot compile a schedule at run time");
    // on-disk cache (keyed by the library version + sources + the instantiation)
    static const uint64_t src_hash = jit_hash(std::string(kJitSrc_common_h) + kJitSrc_butterflies_cuh + kJitSrc_sfft_kernel_cuh + version_string());
    char key[64];
    snprintf(key, sizeof key, "%016llx_%016llx", (unsigned long long)src_hash, (unsigned long long)jit_hash(name_expr));
    const std::string dir = jit_cache_dir(), base = dir + "/" + key;
    if (!std::getenv("NDFB_JIT_NO_DISK_CACHE")) {
        std::ifstream fc(base + ".cubin", std::ios::binary), fn(base + ".name");
        if (fc && fn) {
            std::stringstream ss; ss << fc.rdbuf();
            const std::string bytes = ss.str();
            std::string nm; std::getline(fn, nm);
            if (bytes.size() > 64 && !nm.empty()) { cubin->assign(bytes.begin(), bytes.end()); *lowered = nm; return 0; }
        }
    }
    const std::string src = "#include \"sfft_kernel.cuh\"\n";
    const char* hdr_src[] = {kJitSrc_common_h, kJitSrc_butterflies_cuh, kJitSrc_sfft_kernel_cuh};
    const char* hdr_names[] = {"common.h", "butterflies.cuh", "sfft_kernel.cuh"};
    void* prog = nullptr;
    int rc = n.create(&prog, src.c_str(), "ndfb_jit.cu", 3, hdr_src, hdr_names);
    if (rc) return fail(NDFB_E_CUDA, "nvrtcCreateProgram failed (%d)", rc);
    n.add_name(prog, name_expr.c_str());
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "-default-device", "-DNDFB_JIT=1"};
    rc = n.compile(prog, 5, opts);
    if (rc) {
        size_t ls = 0; n.log_size(prog, &ls);
        std::string lg(ls, 0); if (ls) n.log(prog, &lg[0]);
        if (log_out) *log_out = lg;
        n.destroy(&prog);
        return fail(NDFB_E_CUDA, "run-time compilation of %s failed: %.300s", name_expr.c_str(), lg.c_str());
    }
    const char* low = nullptr;
    n.lowered(prog, name_expr.c_str(), &low);
    size_t cs = 0; n.cubin_size(prog, &cs);
    if (!low || cs == 0) { n.destroy(&prog); return fail(NDFB_E_CUDA, "run-time compilation produced no cubin for %s", name_expr.c_str()); }
    cubin->resize(cs);
    n.cubin(prog, cubin->data());
    *lowered = low;
    n.destroy(&prog);
    if (!std::getenv("NDFB_JIT_NO_DISK_CACHE")) {
        jit_mkdirs(dir);
        char tmp[32]; snprintf(tmp, sizeof tmp, ".tmp%d", (int)getpid());
        { std::ofstream f(base + ".cubin" + tmp, std::ios::binary); f.write(cubin->data(), (std::streamsize)cubin->size()); }
        { std::ofstream f(base + ".name" + tmp); f << *lowered << "\n"; }
        rename((base + ".cubin" + tmp).c_str(), (base + ".cubin").c_str());
        rename((base + ".name" + tmp).c_str(), (base + ".name").c_str());
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// schedule choice (the run-time twin of tools/gen_sfft.py)
// ------------------------------------------------------------------------------------------------------
struct JitSched {
    int N = 0, TL = 0, r[4] = {1, 1, 1, 1}, np = 0, L = 1, cols = 0, minb = 1, E = 0, threads = 0, twtotal = 0;
    size_t smem = 0;
};

inline bool jit_is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// cost of a schedule: every pass is a shared-memory exchange; E (points per thread) sets the register footprint:
// f64 wants E <= 8..10 (64-80 registers, 3-4 CTAs/SM: "family B", profiles/r1f_tune_*), f32 E <= 16
inline long jit_cost(int N, bool f64, const int* r, int np, int TL, int* E_out) {
    int E = 0;
    double util = 0;
    for (int i = 0; i < np; ++i) {
        const int nbf = N / r[i], G = (nbf + TL - 1) / TL;
        E = std::max(E, G * r[i]);
        util += (double)nbf / ((double)TL * G);
    }
    util /= np;
    long pen;
    if (f64) pen = E <= 8 ? 0 : E <= 10 ? 150 : E <= 12 ? 600 : E <= 16 ? 1200 : 2500 + 100L * (E - 16);
    else pen = E <= 16 ? 0 : E <= 20 ? 300 : E <= 24 ? 800 : 2000 + 100L * (E - 24);
    if (E_out) *E_out = E;
    return 1000L * np + pen + (long)(400.0 * (1.0 - util));
}

// best factorisation of N into <= 4 radices from the butterflies that exist (butterflies.cuh), with the threads per lane
inline bool jit_radices(int N, bool f64, int out[4], int* npass, int* TL_out) {
    static const int cand[] = {16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2};
    struct Best { long score = -1; int r[4] = {1, 1, 1, 1}; int np = 0, TL = 0; } best;
    int cur[4];
    struct Rec {
        static void go(int rem, int depth, int maxr, int* cur, int N, bool f64, Best* best) {
            if (rem == 1 && depth > 0) {
                for (int i = 0; i < depth; ++i) {          // candidate thread counts: one butterfly per thread in pass i
                    const int TL = N / cur[i];
                    if (TL < 2 || TL > 1024) continue;
                    bool dup = false;
                    for (int j = 0; j < i; ++j) if (cur[j] == cur[i]) dup = true;
                    if (dup) continue;
                    const long sc = jit_cost(N, f64, cur, depth, TL, nullptr);
                    if (best->score < 0 || sc < best->score) {
                        best->score = sc; best->np = depth; best->TL = TL;
                        for (int k = 0; k < 4; ++k) best->r[k] = k < depth ? cur[k] : 1;
                    }
                }
                return;
            }
            if (depth == 4) return;
            for (int c : cand) {
                if (c > maxr || rem % c) continue;
                cur[depth] = c;
                go(rem / c, depth + 1, c, cur, N, f64, best);
            }
        }
    };
    Rec::go(N, 0, 16, cur, N, f64, &best);
    if (best.score < 0) return false;
    for (int i = 0; i < 4; ++i) out[i] = best.r[i];    // non-increasing: the small leftover radix runs last
    *npass = best.np;
    *TL_out = best.TL;
    return true;
}

inline int jit_npad(int N, int r0) { return (r0 % 2) ? N : N + N / r0; }   // Sched::NPAD

inline bool jit_plan(int N, bool f64, bool cols, bool real_kind, long long nlanes, JitSched* s) {
    if (N < 4) return false;
    int np = 0, TL = 0;
    if (!jit_radices(N, f64, s->r, &np, &TL)) return false;
    s->N = N; s->np = np; s->cols = cols ? 1 : 0;
    s->TL = TL;
    int E = 0, tw = 0;
    long P = 1;
    for (int i = 0; i < np; ++i) {
        const int nbf = N / s->r[i], G = (nbf + TL - 1) / TL;
        E = std::max(E, G * s->r[i]);
        if (i >= 1) tw += (s->r[i] - 1) * (int)P;
        P *= s->r[i];
    }
    s->E = E; s->twtotal = tw;
    const size_t cs = f64 ? 16 : 8;
    const size_t lane_smem = (size_t)jit_npad(N, s->r[0]) * cs;
    auto finish = [&](int L) {
        s->L = L; s->threads = TL * L;
        s->smem = (np > 1 || real_kind) ? (size_t)L * lane_smem : 0;
        int regs = E * (f64 ? 4 : 2) + 40;
        regs = std::min(regs, 255);
        long mb = 65536 / ((long)s->threads * regs);
        if (s->smem) mb = std::min<long>(mb, (long)((227 * 1024) / s->smem));
        mb = std::max<long>(1, std::min<long>(mb, 8));
        // mixed-radix butterflies want ~120 (f64) / ~80 (f32) registers; capping them for twice the CTAs per SM wins
        // 1.2-1.5x (profiles/r1z_tune_mixed.jsonl): same rule as tools/gen_sfft.py capped_variants
        if (!jit_is_pow2(N)) {
            const long cap = 65536 / ((long)s->threads * mb * 2);
            if (cap >= (f64 ? 64 : 48) && s->threads * mb * 2 <= 2048 && (!s->smem || (size_t)(mb * 2) * s->smem <= (size_t)227 * 1024)) mb *= 2;
        }
        s->minb = (int)mb;
    };
    if (!cols) {
        int L = std::max(1, 256 / TL);
        while (L > 1 && (size_t)L * lane_smem > (size_t)72 * 1024) L /= 2;
        int p = 1; while (p * 2 <= L) p *= 2;
        L = p;
        while (L > 1 && L > 2 * nlanes) L /= 2;
        if ((size_t)L * lane_smem > (size_t)220 * 1024 || TL * L > 1024) return false;
        finish(L);
        return s->threads >= 32 || nlanes * TL < 32 || true;
    }
    const int widths[] = {32, 16, 8, 4, 2};
    for (int L : widths) {
        if ((size_t)L * (real_kind ? cs / 2 : cs) > (size_t)(N <= 256 ? 256 : 128)) continue;
        const int T = TL * L;
        const int tmax = (E * (f64 ? 4 : 2) <= 32) ? 1024 : 512;
        if (T > tmax || T < 32 || (size_t)L * lane_smem > (size_t)200 * 1024) continue;
        if (L > 2 * nlanes && L > 2) continue;
        finish(L);
        return true;
    }
    return false;
}

struct JitKernel { void* func = nullptr; JitSched s; };

// `kind` < 0: sfft_kernel (C2C);  0..5: rsfft_kernel of that RKind
inline int jit_get_kernel(bool f64, int kind, const JitSched& s, void** func) {
    static std::mutex mu;
    static std::map<std::string, void*> cache;
    char expr[256];
    const char* R = f64 ? "double" : "float";
    if (kind < 0)
        snprintf(expr, sizeof expr, "ndfb::sfft_kernel<%s, ndfb::Sched<%d, %d, %d, %d, %d, %d>, %d, %s, %d>", R, s.N, s.TL, s.r[0], s.r[1], s.r[2], s.r[3], s.L,
                 s.cols ? "true" : "false", s.minb);
    else
        snprintf(expr, sizeof expr, "ndfb::rsfft_kernel<%s, ndfb::Sched<%d, %d, %d, %d, %d, %d>, %d, %s, %d, %d>", R, s.N, s.TL, s.r[0], s.r[1], s.r[2], s.r[3],
                 s.L, s.cols ? "true" : "false", kind, s.minb);
    int dev = 0;
    cudaGetDevice(&dev);
    const std::string key = std::to_string(dev) + ":" + expr;
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *func = it->second; return it->second ? 0 : fail(NDFB_E_UNSUPPORTED, "run-time compilation failed earlier for %s", expr); }
    std::vector<char> cubin;
    std::string lowered, log;
    int rc = jit_compile(expr, &cubin, &lowered, &log);
    if (rc) { cache[key] = nullptr; return rc; }
    cudaLibrary_t lib = nullptr;
    cudaError_t e = cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e != cudaSuccess) { cudaGetLastError(); cache[key] = nullptr; return fail(NDFB_E_CUDA, "cudaLibraryLoadData failed for %s: %s", expr, cudaGetErrorString(e)); }
    cudaKernel_t k = nullptr;
    e = cudaLibraryGetKernel(&k, lib, lowered.c_str());
    if (e != cudaSuccess) { cudaGetLastError(); cache[key] = nullptr; return fail(NDFB_E_CUDA, "cudaLibraryGetKernel(%s) failed: %s", lowered.c_str(), cudaGetErrorString(e)); }
    e = cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    if (e != cudaSuccess) { cudaGetLastError(); cache[key] = nullptr; return fail(NDFB_E_CUDA, "cudaFuncSetAttribute on a run-time compiled kernel failed: %s", cudaGetErrorString(e)); }
    if (std::getenv("NDFB_TRACE")) fprintf(stderr, "[ndfb] jit compiled %s\n", expr);
    cache[key] = (void*)k;
    *func = (void*)k;
    return 0;
}

template <typename A>
inline int jit_launch(void* func, const A& a, unsigned grid, unsigned block, size_t smem, stream_t stream) {
    A copy = a;
    void* args[] = {&copy};
    {
        size_t& floor_ = launch_smem_floor();
        if (floor_ > smem) smem = floor_ < (size_t)(227 * 1024) ? floor_ : (size_t)(227 * 1024);
        floor_ = 0;
    }
    NDFB_CUDA(cudaLaunchKernel((const void*)func, dim3(grid), dim3(block), args, smem, stream));
    launch_counter()++;
    return 0;
}

}  // namespace ndfb
#endif  // !NDFB_EMU
