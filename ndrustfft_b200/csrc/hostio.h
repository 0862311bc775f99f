// hostio.h — host-array plumbing of the NDFB_MEM_HOST path: a small copy-thread pool, logical-element pack / unpack
// of strided host arrays, and (CUDA build only) the pinned staging ring that lets PAGEABLE caller memory
// (what ndarray's ArrayBase::as_ptr hands the shim, src/lib.rs:105-115) pipeline with the GPU.
//
// Why a ring: cudaMemcpyAsync on pageable memory is staged by the driver through one internal buffer by a single
// thread and blocks the caller, so the H2D | kernel | D2H overlap of exec_host_pipelined cannot happen.  Here the
// library owns 2 x kSlots pinned buffers; worker threads copy caller memory <-> pinned slots while the DMA engines and
// the kernels work on the neighbouring pieces.  Pointers that are already pinned / registered skip the ring.
#pragma once
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#endif

#include "devapi.h"

namespace ndfb {

// ------------------------------------------------------------------------------------------------------
// copy-thread pool (process wide; one parallel job at a time)
// ------------------------------------------------------------------------------------------------------
class CopyPool {
  public:
    static CopyPool& get() {
        static CopyPool* p = new CopyPool();   // never destroyed: workers may be parked on the condvar at process exit
        return *p;
    }
    int threads() const { return nthreads_; }
    // runs f(t, T) for t = 0..T-1 (the caller is worker 0) and returns when all are done
    void run(const std::function<void(int, int)>& f) {
        if (nthreads_ <= 1) { f(0, 1); return; }
        std::lock_guard<std::mutex> job_guard(job_mu_);
        {
            std::lock_guard<std::mutex> g(mu_);
            job_ = &f;
            pending_ = nthreads_ - 1;
            ++gen_;
        }
        cv_.notify_all();
        f(0, nthreads_);
        std::unique_lock<std::mutex> g(mu_);
        cv_done_.wait(g, [&] { return pending_ == 0; });
        job_ = nullptr;
    }

  private:
    CopyPool() {
        int n = 0;
        if (const char* e = std::getenv("NDFB_HOST_THREADS")) n = atoi(e);
        if (n <= 0) {
            const unsigned hw = std::thread::hardware_concurrency();
            n = (int)std::max(1u, std::min(12u, hw * 3 / 4));
        }
        nthreads_ = std::min(n, 64);
        for (int t = 1; t < nthreads_; ++t) std::thread([this, t] { worker(t); }).detach();
    }
    void worker(int t) {
        unsigned seen = 0;
        for (;;) {
            const std::function<void(int, int)>* f;
            {
                std::unique_lock<std::mutex> g(mu_);
                cv_.wait(g, [&] { return gen_ != seen; });
                seen = gen_;
                f = job_;
            }
            if (f) (*f)(t, nthreads_);
            {
                std::lock_guard<std::mutex> g(mu_);
                if (--pending_ == 0) cv_done_.notify_one();
            }
        }
    }
    int nthreads_ = 1;
    std::mutex job_mu_, mu_;
    std::condition_variable cv_, cv_done_;
    const std::function<void(int, int)>* job_ = nullptr;
    unsigned gen_ = 0;
    int pending_ = 0;
};

// Large copies bypass the cache on the store side (non-temporal stores): a staging copy is read once by the DMA engine
// or by the caller much later, and regular stores would first READ every destination line (read-for-ownership),
// i.e. 3 bytes of DRAM traffic per byte copied instead of 2.
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("avx2"))) inline void stream_copy_avx2(char* d, const char* s, size_t n) {
    while (n && ((uintptr_t)d & 31)) { *d++ = *s++; --n; }
    size_t v = n / 128;
    for (size_t i = 0; i < v; ++i) {
        const __m256i a = _mm256_loadu_si256((const __m256i*)(s)), b = _mm256_loadu_si256((const __m256i*)(s + 32));
        const __m256i c = _mm256_loadu_si256((const __m256i*)(s + 64)), e = _mm256_loadu_si256((const __m256i*)(s + 96));
        _mm256_stream_si256((__m256i*)(d), a); _mm256_stream_si256((__m256i*)(d + 32), b);
        _mm256_stream_si256((__m256i*)(d + 64), c); _mm256_stream_si256((__m256i*)(d + 96), e);
        s += 128; d += 128;
    }
    n -= v * 128;
    if (n) std::memcpy(d, s, n);
    _mm_sfence();
}
inline void big_copy(void* d, const void* s, size_t n) {
    static const bool avx2 = __builtin_cpu_supports("avx2") && !std::getenv("NDFB_NO_STREAM_COPY");
    if (avx2 && n >= (size_t)(64 << 10)) stream_copy_avx2((char*)d, (const char*)s, n);
    else std::memcpy(d, s, n);
}
#else
inline void big_copy(void* d, const void* s, size_t n) { std::memcpy(d, s, n); }
#endif

// rows x width bytes, both sides pitched; split evenly by bytes over the pool
inline void parallel_copy_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows) {
    const size_t total = width * rows;
    if (total == 0) return;
    auto body = [&](int t, int T) {
        size_t lo = total * (size_t)t / (size_t)T, hi = total * (size_t)(t + 1) / (size_t)T;
        while (lo < hi) {
            const size_t r = lo / width, c = lo - r * width;
            const size_t n = std::min(width - c, hi - lo);
            big_copy((char*)dst + r * dpitch + c, (const char*)src + r * spitch + c, n);
            lo += n;
        }
    };
    if (total < ((size_t)1 << 20)) body(0, 1);
    else CopyPool::get().run(body);
}

// ------------------------------------------------------------------------------------------------------
// logical-element pack / unpack of a strided host array <-> dense C-order buffer
// (never touches the bytes between the elements of a view: ndarray hands out sibling views of one allocation)
// ------------------------------------------------------------------------------------------------------
inline void host_nd_copy(bool pack, void* dense, void* strided, int ndim, const size_t* shape, const ptrdiff_t* strides, size_t elem) {
    size_t total = 1;
    for (int d = 0; d < ndim; ++d) total *= shape[d];
    if (total == 0) return;
    const size_t inner = shape[ndim - 1];
    const ptrdiff_t inner_st = strides[ndim - 1];
    const size_t outer = total / inner;
    auto body = [&](int t, int T) {
        const size_t lo = outer * (size_t)t / (size_t)T, hi = outer * (size_t)(t + 1) / (size_t)T;
        for (size_t o = lo; o < hi; ++o) {
            size_t rem = o;
            long long off = 0;
            for (int d = ndim - 2; d >= 0; --d) { const size_t i = rem % shape[d]; rem /= shape[d]; off += (long long)i * (long long)strides[d]; }
            char* dn = (char*)dense + o * inner * elem;
            char* st = (char*)strided + off * (long long)elem;
            if (inner_st == 1) {
                if (pack) std::memcpy(dn, st, inner * elem); else std::memcpy(st, dn, inner * elem);
            } else {
                for (size_t i = 0; i < inner; ++i) {
                    char* a = dn + i * elem;
                    char* b = st + (long long)i * inner_st * (long long)elem;
                    if (pack) std::memcpy(a, b, elem); else std::memcpy(b, a, elem);
                }
            }
        }
    };
    if (total * elem < ((size_t)1 << 20)) body(0, 1);
    else CopyPool::get().run(body);
}

#ifndef NDFB_EMU
// ------------------------------------------------------------------------------------------------------
// pinned staging ring (per host thread)
// ------------------------------------------------------------------------------------------------------
inline bool host_ptr_pageable(const void* p) {
    static const bool off = std::getenv("NDFB_NO_STAGING") != nullptr;   // A/B hook: async copies straight from pageable memory
    if (off) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

struct StageRing {
    static constexpr int kSlots = 4;
    int device = -1;
    size_t slot_bytes = 0;
    void* in_slot[kSlots] = {nullptr};
    void* out_slot[kSlots] = {nullptr};
    cudaEvent_t in_free[kSlots], out_ready[kSlots];
    bool in_busy[kSlots] = {false};
    unsigned in_next = 0, out_next = 0;
    struct Pending { int slot; void* h; size_t hpitch, width, rows; };
    std::deque<Pending> pend;

    int ensure(int dev) {
        if (device == dev && slot_bytes) return 0;
        release();
        size_t mb = 16;
        if (const char* e = std::getenv("NDFB_STAGE_MB")) mb = (size_t)std::max(1, atoi(e));
        for (int i = 0; i < kSlots; ++i) {
            NDFB_CUDA(cudaHostAlloc(&in_slot[i], mb << 20, cudaHostAllocDefault));
            NDFB_CUDA(cudaHostAlloc(&out_slot[i], mb << 20, cudaHostAllocDefault));
            NDFB_CUDA(cudaEventCreateWithFlags(&in_free[i], cudaEventDisableTiming));
            NDFB_CUDA(cudaEventCreateWithFlags(&out_ready[i], cudaEventDisableTiming));
            in_busy[i] = false;
        }
        slot_bytes = mb << 20;
        device = dev;
        return 0;
    }
    void release() {
        if (!slot_bytes) return;
        for (int i = 0; i < kSlots; ++i) {
            if (in_slot[i]) cudaFreeHost(in_slot[i]);
            if (out_slot[i]) cudaFreeHost(out_slot[i]);
            cudaEventDestroy(in_free[i]);
            cudaEventDestroy(out_ready[i]);
            in_slot[i] = out_slot[i] = nullptr;
        }
        slot_bytes = 0;
        device = -1;
        pend.clear();
    }
    // finish the host side of completed downloads (all == false: only those whose DMA has already finished)
    int drain(bool all) {
        while (!pend.empty()) {
            const Pending& q = pend.front();
            if (!all) {
                const cudaError_t e = cudaEventQuery(out_ready[q.slot]);
                if (e == cudaErrorNotReady) break;
                if (e != cudaSuccess) return cuda_fail(e, "cudaEventQuery");
            } else {
                NDFB_CUDA(cudaEventSynchronize(out_ready[q.slot]));
            }
            parallel_copy_2d(q.h, q.hpitch, out_slot[q.slot], q.width, q.width, q.rows);
            pend.pop_front();
        }
        return 0;
    }
    // the piece is cut into sub-pieces of whole rows (or row segments when one row exceeds a slot) that fit a slot
    template <typename F>
    int for_subpieces(size_t width, size_t rows, F&& f) {
        if (width == 0 || rows == 0) return 0;
        if (width <= slot_bytes) {
            const size_t per = std::max<size_t>(1, slot_bytes / width);
            for (size_t r = 0; r < rows; r += per) {
                int rc = f(r, (size_t)0, width, std::min(per, rows - r));
                if (rc) return rc;
            }
        } else {
            for (size_t r = 0; r < rows; ++r)
                for (size_t c = 0; c < width; c += slot_bytes) {
                    int rc = f(r, c, std::min(slot_bytes, width - c), (size_t)1);
                    if (rc) return rc;
                }
        }
        return 0;
    }
    int h2d(void* d, size_t dpitch, const void* h, size_t hpitch, size_t width, size_t rows, bool pageable, cudaStream_t s) {
        if (!pageable) {
            NDFB_CUDA(cudaMemcpy2DAsync(d, dpitch, h, hpitch, width, rows, cudaMemcpyHostToDevice, s));
            return 0;
        }
        return for_subpieces(width, rows, [&](size_t r0, size_t c0, size_t w, size_t nr) -> int {
            const int slot = (int)(in_next++ % kSlots);
            if (in_busy[slot]) {
                int rc = drain(false);   // useful work before blocking
                if (rc) return rc;
                NDFB_CUDA(cudaEventSynchronize(in_free[slot]));
            }
            parallel_copy_2d(in_slot[slot], w, (const char*)h + r0 * hpitch + c0, hpitch, w, nr);
            NDFB_CUDA(cudaMemcpy2DAsync((char*)d + r0 * dpitch + c0, dpitch, in_slot[slot], w, w, nr, cudaMemcpyHostToDevice, s));
            NDFB_CUDA(cudaEventRecord(in_free[slot], s));
            in_busy[slot] = true;
            return 0;
        });
    }
    int d2h(void* h, size_t hpitch, const void* d, size_t dpitch, size_t width, size_t rows, bool pageable, cudaStream_t s) {
        if (!pageable) {
            NDFB_CUDA(cudaMemcpy2DAsync(h, hpitch, d, dpitch, width, rows, cudaMemcpyDeviceToHost, s));
            return 0;
        }
        return for_subpieces(width, rows, [&](size_t r0, size_t c0, size_t w, size_t nr) -> int {
            while ((int)pend.size() >= kSlots) {
                const size_t before = pend.size();
                int rc = drain(false);
                if (rc) return rc;
                if (pend.size() == before) {   // nothing finished yet: block on the oldest
                    const Pending q = pend.front();
                    NDFB_CUDA(cudaEventSynchronize(out_ready[q.slot]));
                    parallel_copy_2d(q.h, q.hpitch, out_slot[q.slot], q.width, q.width, q.rows);
                    pend.pop_front();
                }
            }
            const int slot = (int)(out_next++ % kSlots);
            NDFB_CUDA(cudaMemcpy2DAsync(out_slot[slot], w, (const char*)d + r0 * dpitch + c0, dpitch, w, nr, cudaMemcpyDeviceToHost, s));
            NDFB_CUDA(cudaEventRecord(out_ready[slot], s));
            pend.push_back(Pending{slot, (char*)h + r0 * hpitch + c0, hpitch, w, nr});
            return 0;
        });
    }
};
#endif  // !NDFB_EMU

}  // namespace ndfb
