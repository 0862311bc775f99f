// devapi.h — device plumbing shared by every translation unit: error slot, launch counter, CUDA wrappers
// (or plain host memory + the SIMT emulator in the test-only NDFB_EMU build).
#pragma once
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>

#include "../../include/ndfft_b200.h"
#include "common.h"

namespace ndfb {

// ------------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------------
std::string& err_slot();  // thread-local, defined in ndfft_b200.cu

inline int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    err_slot() = buf;
    return code;
}

std::atomic<uint64_t>& launch_counter();  // defined in ndfft_b200.cu
size_t& launch_smem_floor();               // thread-local, consumed by the next launch (occupancy hint); ndfft_b200.cu

// ------------------------------------------------------------------------------------------------------
// device plumbing (CUDA, or plain host memory in the emulation build)
// ------------------------------------------------------------------------------------------------------
#ifdef NDFB_EMU
typedef void* stream_t;
inline int dev_set(int) { return 0; }
struct DeviceGuard {};
inline int dev_malloc(void** p, size_t bytes) { *p = std::malloc(bytes ? bytes : 1); return *p ? 0 : NDFB_E_ALLOC; }
inline void dev_free(void* p) { std::free(p); }
inline int dev_h2d(void* d, const void* h, size_t bytes, stream_t) { std::memcpy(d, h, bytes); return 0; }
inline int dev_d2h(void* h, const void* d, size_t bytes, stream_t) { std::memcpy(h, d, bytes); return 0; }
inline int dev_sync(stream_t) { return 0; }
inline int dev_memset0(void* d, size_t bytes, stream_t) { std::memset(d, 0, bytes); return 0; }
inline size_t dev_smem_cap(int) { return 227 * 1024; }
inline int dev_sm_count(int) { return 148; }
template <typename K, typename A>
inline int dev_launch(K kernel, unsigned grid, unsigned block, size_t smem, stream_t, const A& a) {
    A copy = a;
    {
        size_t& floor_ = launch_smem_floor();   // same occupancy / extra-room hint as the CUDA build
        if (floor_ > smem) smem = floor_;
        floor_ = 0;
    }
    simt::launch(dim3(grid), dim3(block), smem, [&]() { kernel(copy); });
    launch_counter()++;
    return 0;
}
inline const char* version_string() { return "ndfft_b200 0.2 emu (CPU SIMT emulation, tests only)"; }
#else
typedef cudaStream_t stream_t;
inline int cuda_fail(cudaError_t e, const char* what) {
    return fail(NDFB_E_CUDA, "CUDA error in %s: %s", what, cudaGetErrorString(e));
}
#define NDFB_CUDA(call)                                        \
    do {                                                       \
        cudaError_t e_ = (call);                               \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);    \
    } while (0)
inline int dev_set(int dev) { NDFB_CUDA(cudaSetDevice(dev)); return 0; }
// restores the caller's current device when an entry point returns (the host application, e.g. torch, owns it)
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; } }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
inline int dev_malloc(void** p, size_t bytes) {
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 1);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(NDFB_E_ALLOC, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); }
    return 0;
}
inline void dev_free(void* p) { cudaFree(p); }
inline int dev_h2d(void* d, const void* h, size_t bytes, stream_t s) { NDFB_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s)); return 0; }
inline int dev_d2h(void* h, const void* d, size_t bytes, stream_t s) { NDFB_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s)); return 0; }
inline int dev_sync(stream_t s) { NDFB_CUDA(cudaStreamSynchronize(s)); return 0; }
inline int dev_memset0(void* d, size_t bytes, stream_t s) { NDFB_CUDA(cudaMemsetAsync(d, 0, bytes, s)); return 0; }
inline size_t dev_smem_cap(int dev) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || v <= 0) { cudaGetLastError(); return 227 * 1024; }
    return (size_t)v;
}
inline int dev_sm_count(int dev) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) { cudaGetLastError(); return 148; }
    return v;
}
template <typename K, typename A>
inline int dev_launch(K kernel, unsigned grid, unsigned block, size_t smem, stream_t s, const A& a) {
    static thread_local std::map<std::pair<int, const void*>, size_t> attr_set;   // per (device, kernel)
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    const std::pair<int, const void*> key(cur_dev, (const void*)kernel);
    auto it = attr_set.find(key);
    if (it == attr_set.end() || it->second < smem) {
        NDFB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        attr_set[key] = 227 * 1024;
    }
    {
        size_t& floor_ = launch_smem_floor();
        if (floor_ > smem) smem = floor_ < (size_t)(227 * 1024) ? floor_ : (size_t)(227 * 1024);
        floor_ = 0;
    }
    kernel<<<grid, block, smem, s>>>(a);
    NDFB_CUDA(cudaGetLastError());
    launch_counter()++;
    return 0;
}
inline const char* version_string() { return "ndfft_b200 0.2 sm_100a"; }
#endif


}  // namespace ndfb
