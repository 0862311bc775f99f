// common.h — types shared by the host plan builder and the device kernels.
#pragma once
#ifdef __CUDACC_RTC__
// run-time compilation of a schedule (jit.h): NVRTC has no host headers; the kernel sources need only these
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef unsigned long uintptr_t;
#else
#include <cstddef>
#include <cstdint>
#endif

#ifdef NDFB_EMU
#include "simt_emu.h"  // tests/emu: g++-compilable stand-ins for the CUDA spellings (test-only build)
#define NDFB_HD inline
#define NDFB_DEV inline
#define NDFB_DYN_SMEM(name) unsigned char* name = simt::block()->smem
#else
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#endif
#define NDFB_HD __host__ __device__ __forceinline__
#define NDFB_DEV __device__ __forceinline__
#define NDFB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace ndfb {

// Interleaved complex, layout-compatible with Rust's #[repr(C)] num_complex::Complex<T> (src/lib.rs:83).
template <typename R>
struct alignas(2 * sizeof(R)) Cx {
    R x, y;
};

template <typename R> NDFB_HD Cx<R> cmake(R x, R y) { Cx<R> c; c.x = x; c.y = y; return c; }
template <typename R> NDFB_HD Cx<R> cadd(Cx<R> a, Cx<R> b) { return cmake<R>(a.x + b.x, a.y + b.y); }
template <typename R> NDFB_HD Cx<R> csub(Cx<R> a, Cx<R> b) { return cmake<R>(a.x - b.x, a.y - b.y); }
template <typename R> NDFB_HD Cx<R> cmul(Cx<R> a, Cx<R> b) { return cmake<R>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
template <typename R> NDFB_HD Cx<R> cconj(Cx<R> a) { return cmake<R>(a.x, -a.y); }
template <typename R> NDFB_HD Cx<R> cscale(Cx<R> a, R s) { return cmake<R>(a.x * s, a.y * s); }
template <typename R> NDFB_HD Cx<R> cmul_i(Cx<R> a) { return cmake<R>(-a.y, a.x); }      // i*a
template <typename R> NDFB_HD Cx<R> cmul_ni(Cx<R> a) { return cmake<R>(a.y, -a.x); }     // -i*a

// read-only (non-coherent) table loads
#ifdef NDFB_EMU
template <typename R> NDFB_DEV Cx<R> ldg(const Cx<R>* p) { return *p; }
NDFB_DEV uint32_t ldg(const uint32_t* p) { return *p; }
#else
NDFB_DEV Cx<float> ldg(const Cx<float>* p) {
    float2 v = __ldg(reinterpret_cast<const float2*>(p));
    return cmake<float>(v.x, v.y);
}
NDFB_DEV Cx<double> ldg(const Cx<double>* p) {
    double2 v = __ldg(reinterpret_cast<const double2*>(p));
    return cmake<double>(v.x, v.y);
}
NDFB_DEV uint32_t ldg(const uint32_t* p) { return __ldg(p); }
#endif

// streaming loads of the user's array: each element is read once, so it should not displace the twiddle tables from L1
// (the 8192-point row kernel re-read 17 M twiddle sectors per launch from L2 before; profiles/r1l_ncu_bench_summary.txt)
#ifndef NDFB_STREAM_LD
#define NDFB_STREAM_LD 0   // measured neutral on B200 (profiles/r1q_ab_stream.jsonl): off
#endif
#if defined(NDFB_EMU) || !NDFB_STREAM_LD
template <typename T> NDFB_DEV T ld_stream(const T* p) { return *p; }
#else
NDFB_DEV Cx<float> ld_stream(const Cx<float>* p) {
    float x, y;
    asm("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "l"(p) : "memory");
    return cmake<float>(x, y);
}
NDFB_DEV Cx<double> ld_stream(const Cx<double>* p) {
    double x, y;
    asm("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "l"(p) : "memory");
    return cmake<double>(x, y);
}
NDFB_DEV float ld_stream(const float* p) {
    float x;
    asm("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(x) : "l"(p) : "memory");
    return x;
}
NDFB_DEV double ld_stream(const double* p) {
    double x;
    asm("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(x) : "l"(p) : "memory");
    return x;
}
#endif

// L2-coherent load (bypasses the non-coherent L1): for data written by other CTAs of the same launch
#ifdef NDFB_EMU
template <typename T> NDFB_DEV T ld_cg(const T* p) { return *p; }
#else
NDFB_DEV Cx<float> ld_cg(const Cx<float>* p) {
    float2 v = __ldcg(reinterpret_cast<const float2*>(p));
    return cmake<float>(v.x, v.y);
}
NDFB_DEV Cx<double> ld_cg(const Cx<double>* p) {
    double2 v = __ldcg(reinterpret_cast<const double2*>(p));
    return cmake<double>(v.x, v.y);
}
#endif

// What one launch of the tile kernel computes around its forward complex FFT core of length N.
// (n = the handler's logical length; see DESIGN.md "kinds" for the algebra, verified in tests/kernel_math_model.py.)
enum TileKind : int {
    TK_C2C = 0,        // N = n            complex -> complex   (flags: conj_in / conj_out / four-step twiddle)
    TK_R2C_EVEN = 1,   // N = n/2          real n -> complex n/2+1 via packed half-length FFT + post-twiddle
    TK_R2C_ODD = 2,    // N = n            real n -> complex n/2+1 via full complex FFT
    TK_C2R_EVEN = 3,   // N = n/2          complex n/2+1 -> real n (pre-twiddle "zip", Im(DC)/Im(Nyq) dropped)
    TK_C2R_ODD = 4,    // N = n            Hermitian completion + full complex FFT
    TK_DCT1 = 5,       // N = n-1          even extension to 2(n-1) reals, packed
    TK_DCT2_EVEN = 6,  // N = n/2          Makhoul reorder + packed real FFT + quarter-wave twiddle
    TK_DCT2_ODD = 7,   // N = n            Makhoul reorder + full complex FFT
    TK_DCT3_EVEN = 8,  // N = n/2          inverse of DCT2_EVEN
    TK_DCT3_ODD = 9,   // N = n
    TK_DCT4_EVEN = 10, // N = n/2          pre/post twiddled half-length complex FFT
    TK_DCT4_ODD = 11   // N = 2n           zero-padded complex FFT with pre/post twiddles
};

constexpr int kMaxPass = 24;
constexpr int kMaxBatchDims = 4;

// Kernel argument block (passed by value, < 1 KB).
struct TileArgs {
    const void* in;
    void* out;
    int kind;
    int n;        // logical length
    int n_in;     // lane length of the input in its own element type
    int n_out;    // lane length of the output
    int N;        // core FFT length seen by prologue/epilogue
    int M;        // Bluestein convolution length (0 = direct)
    int Bl;       // complex slots per lane in shared memory
    int L;        // lanes per tile (power of two)
    int log2L;
    int LP, EP;   // shared-memory pitches (complex elements): slot(l, p) = l*LP + p*EP
    int pad_shift;  // physical position = p + (p >> pad_shift); 31 = no padding
    long long nlanes;
    int nbd;                          // batch dims in use
    long long bsz[kMaxBatchDims];     // extents, fastest first
    long long bis[kMaxBatchDims];     // input strides (elements)
    long long bos[kMaxBatchDims];     // output strides (elements)
    long long is_axis, os_axis;       // strides along the transformed axis (elements)
    int in_lane_fast, out_lane_fast;  // which index the global-memory loops run fastest
    int conj_in, conj_out;            // TK_C2C flags
    int npass;
    int radix[kMaxPass];
    const void* tw;        // W_B^k, k < B (B = M ? M : N), forward sign
    const void* tabA;      // kind-specific table A
    const void* tabB;      // kind-specific table B
    const uint32_t* perm;  // DIF output positions (direct plans)
    const void* blu_c;     // Bluestein chirp c[j], j < N
    const void* blu_bhat;  // FFT_M(conj chirp kernel)/M stored in DIF order
    double scale;          // multiplied into every output element
    // four-step (TK_C2C only): multiply output element k of lane j2 by W_fsN^{k*j2}
    int fs_twiddle;
    int fs_dim;            // batch dim whose index is j2
    int fs_shift;          // W = hi[e >> fs_shift] * lo[e & ((1<<fs_shift)-1)], e = k*j2
    const void* fs_lo;
    const void* fs_hi;
    // TK_C2C only: output axis index k split as (k / os_blk, k % os_blk) with stride os_blk_stride for the block index
    int os_blk;
    long long os_blk_stride;
};

}  // namespace ndfb
