// simt_emu.h — a tiny SIMT emulator so the CUDA kernel SOURCES can be compiled with g++ and stepped on a CPU.
//
// TEST INFRASTRUCTURE ONLY.  It exists so that `pytest -m "not gpu"` can exercise the real kernel logic
// (index maps, twiddles, prologues/epilogues, radix schedules) in a container without a GPU.  The product
// package never loads the library built from it: ndrustfft_b200/_lib.py only ever opens the nvcc-built
// libndfft_b200.so and raises when it is missing.  Nothing here is timed, shipped or used as a fallback.
//
// Model: one CUDA thread = one ucontext fiber; a block's fibers run round-robin on the calling OS thread;
// __syncthreads() yields to the scheduler, which resumes the block once every live fiber has arrived.
// Blocks run one after another.  Warp shuffles are emulated with a per-warp mailbox and warp-level yields.
#pragma once
#include <ucontext.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

namespace simt {

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;
    bool done = false;
    dim3 tid;
    unsigned linear = 0;
};

struct BlockState {
    dim3 blockIdx, blockDim, gridDim;
    unsigned char* smem = nullptr;
    std::vector<Fiber> fibers;
    ucontext_t sched;
    Fiber* cur = nullptr;
    const std::function<void()>* body = nullptr;
    // warp shuffle mailbox: 8 bytes per thread
    std::vector<uint64_t> mailbox;
    std::vector<unsigned> warp_arrivals;
    std::vector<unsigned> warp_generation;
    // block barrier
    unsigned live = 0, bar_arrived = 0, bar_generation = 0;
};

inline BlockState*& block() {
    static thread_local BlockState* b = nullptr;
    return b;
}

inline void yield_to_scheduler() {
    BlockState* b = block();
    Fiber* f = b->cur;
    swapcontext(&f->ctx, &b->sched);
}

inline void syncthreads() {
    BlockState* b = block();
    unsigned gen = b->bar_generation;
    if (++b->bar_arrived >= b->live) {
        b->bar_arrived = 0;
        b->bar_generation++;
    }
    while (b->bar_generation == gen) yield_to_scheduler();
}

inline void fiber_entry() {
    BlockState* b = block();
    Fiber* f = b->cur;
    (*b->body)();
    f->done = true;
    // an exited thread counts as arrived at every later barrier (sm_70+ semantics)
    b->live--;
    if (b->live > 0 && b->bar_arrived >= b->live) {
        b->bar_arrived = 0;
        b->bar_generation++;
    }
    swapcontext(&f->ctx, &b->sched);
}

// Warp-level exchange: every lane of the warp must call with the same sequence.
template <typename T>
inline T shfl_generic(T v, unsigned src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    BlockState* b = block();
    Fiber* f = b->cur;
    unsigned nthreads = b->blockDim.x * b->blockDim.y * b->blockDim.z;
    unsigned warp = f->linear / 32, lane = f->linear % 32;
    unsigned wsize = std::min(32u, nthreads - warp * 32);
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    b->mailbox[f->linear] = raw;
    // arrive, wait for the whole warp
    unsigned gen = b->warp_generation[warp];
    if (++b->warp_arrivals[warp] == wsize) {
        b->warp_arrivals[warp] = 0;
        b->warp_generation[warp]++;
    }
    while (b->warp_generation[warp] == gen) yield_to_scheduler();
    uint64_t got = b->mailbox[warp * 32 + (src_lane % wsize)];
    // second phase so nobody overwrites the mailbox early
    gen = b->warp_generation[warp];
    if (++b->warp_arrivals[warp] == wsize) {
        b->warp_arrivals[warp] = 0;
        b->warp_generation[warp]++;
    }
    while (b->warp_generation[warp] == gen) yield_to_scheduler();
    T out;
    std::memcpy(&out, &got, sizeof(T));
    (void)lane;
    return out;
}

inline void launch(dim3 grid, dim3 blockDim, size_t smem_bytes, const std::function<void()>& body) {
    const size_t kStack = 64 * 1024;
    unsigned nthreads = blockDim.x * blockDim.y * blockDim.z;
    BlockState st;
    st.blockDim = blockDim;
    st.gridDim = grid;
    st.body = &body;
    st.fibers.resize(nthreads);
    st.mailbox.assign(nthreads, 0);
    st.warp_arrivals.assign((nthreads + 31) / 32, 0);
    st.warp_generation.assign((nthreads + 31) / 32, 0);
    static thread_local char* stack_pool = nullptr;
    static thread_local size_t stack_pool_size = 0;
    if (stack_pool_size < kStack * (size_t)nthreads) {
        std::free(stack_pool);
        stack_pool_size = kStack * (size_t)nthreads;
        stack_pool = (char*)std::malloc(stack_pool_size);  // untouched pages stay unmapped
    }
    std::vector<unsigned char> smem(smem_bytes + 64);
    st.smem = smem.data();
    BlockState* prev = block();
    block() = &st;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                st.blockIdx = dim3(bx, by, bz);
                std::memset(smem.data(), 0xCD, smem.size());  // poison: uninitialised reads show up as garbage
                unsigned lin = 0;
                for (unsigned tz = 0; tz < blockDim.z; ++tz)
                    for (unsigned ty = 0; ty < blockDim.y; ++ty)
                        for (unsigned tx = 0; tx < blockDim.x; ++tx, ++lin) {
                            Fiber& f = st.fibers[lin];
                            f.done = false;
                            f.tid = dim3(tx, ty, tz);
                            f.linear = lin;
                            f.stack = stack_pool + kStack * (size_t)lin;
                            getcontext(&f.ctx);
                            f.ctx.uc_stack.ss_sp = f.stack;
                            f.ctx.uc_stack.ss_size = kStack;
                            f.ctx.uc_link = &st.sched;
                            makecontext(&f.ctx, (void (*)())fiber_entry, 0);
                        }
                std::fill(st.warp_arrivals.begin(), st.warp_arrivals.end(), 0u);
                st.live = nthreads;
                st.bar_arrived = 0;
                bool any = true;
                while (any) {
                    any = false;
                    for (unsigned i = 0; i < nthreads; ++i) {
                        Fiber& f = st.fibers[i];
                        if (f.done) continue;
                        st.cur = &f;
                        swapcontext(&st.sched, &f.ctx);
                        if (!f.done) any = true;
                    }
                }
            }
    block() = prev;
}

}  // namespace simt

// ---- CUDA spellings mapped onto the emulator ---------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __grid_constant__
using simt::dim3;
#define threadIdx (simt::block()->cur->tid)
#define blockIdx (simt::block()->blockIdx)
#define blockDim (simt::block()->blockDim)
#define gridDim (simt::block()->gridDim)
#define __syncthreads() simt::syncthreads()
#define __syncwarp(...) ((void)0)
template <typename T>
inline T __ldg(const T* p) { return *p; }
// blocks run one after another, fibers of a block are cooperative: plain read-modify-write is atomic here
template <typename T>
inline T atomicAdd(T* p, T v) { T old = *p; *p = old + v; return old; }
template <typename T>
inline T atomicExch(T* p, T v) { T old = *p; *p = v; return old; }
inline void __threadfence() {}
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int lanemask) {
    return simt::shfl_generic(v, (simt::block()->cur->linear % 32) ^ (unsigned)lanemask);
}
template <typename T>
inline T __shfl_sync(unsigned, T v, int src) { return simt::shfl_generic(v, (unsigned)src); }
