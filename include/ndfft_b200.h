/*
 * ndfft_b200.h — C ABI of the B200-native replacement for ndrustfft's hot path:
 * the batched 1-D transform along one axis of an n-dimensional array.
 *
 * The reference (preiter93/ndrustfft v0.5.0, /root/reference/src/lib.rs) has no FFI of its own; its
 * seam is the public Rust API.  Each entry point below is what a Rust shim (`rust/ndrustfft-b200`)
 * binds with `extern "C"` to replace one piece of that API:
 *
 *   ndfb_plan_create(NDFB_C2C, ..)  <- FftHandler::new        src/lib.rs:294-304 (FftPlanner::plan_fft_forward/inverse)
 *   ndfb_plan_create(NDFB_R2C, ..)  <- R2cFftHandler::new     src/lib.rs:477-488 (RealFftPlanner)
 *   ndfb_plan_create(NDFB_DCT, ..)  <- DctHandler::new        src/lib.rs:665-679 (DctPlanner::plan_dct1..4)
 *   ndfb_plan_destroy               <- Drop of the Arc<dyn ..> plans held by the handlers (src/lib.rs:270-275, 452-458, 641-648)
 *   ndfb_exec                       <- the bodies of create_transform! / create_transform_par!  src/lib.rs:100-238
 *                                      together with the per-lane methods they call:
 *        NDFB_OP_FFT   fft_lane       :313-318     NDFB_OP_IFFT  ifft_lane      :321-331
 *        NDFB_OP_R2C   fft_r2c_lane   :497-503     NDFB_OP_C2R   ifft_r2c_lane  :506-523
 *        NDFB_OP_DCT1..4  dct1..4_lane :688-734
 *   norm = NDFB_NORM_NONE / NDFB_NORM_DEFAULT   <- Normalization::None / ::Default   :89-98, 333-338, 525-531, 736-741
 *        (Normalization::Custom(fn) is a host function pointer; the shim applies it on the host)
 *   NDFB_E_SIZE_MISMATCH + ndfb_last_error()    <- assert_size panics  :340-347, 533-540, 743-750
 *
 * Conventions: plain pointers and sizes only; no exceptions cross the boundary; every function
 * returns 0 or a negative NDFB_E_* code and leaves a thread-local message in ndfb_last_error().
 * A plan is immutable after creation and may be shared between threads and streams (the handlers
 * are `Clone` + shared by `&` across rayon workers in the reference, src/lib.rs:169-238).
 * There is no CPU fallback: without a usable CUDA device every exec returns NDFB_E_CUDA.
 */
#ifndef NDFFT_B200_H
#define NDFFT_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define NDFB_API __attribute__((visibility("default")))
#else
#define NDFB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ndfb_plan ndfb_plan;

enum ndfb_kind  { NDFB_C2C = 0, NDFB_R2C = 1, NDFB_DCT = 2 };
enum ndfb_dtype { NDFB_F32 = 0, NDFB_F64 = 1 };
enum ndfb_op {
    NDFB_OP_FFT = 0,   /* ndfft      : complex n -> complex n, forward, never scaled                 */
    NDFB_OP_IFFT = 1,  /* ndifft     : complex n -> complex n, backward, Default = 1/n after          */
    NDFB_OP_R2C = 2,   /* ndfft_r2c  : real n -> complex n/2+1, never scaled                          */
    NDFB_OP_C2R = 3,   /* ndifft_r2c : complex n/2+1 -> real n, Default = 1/n, Im(DC), Im(Nyq) ignored */
    NDFB_OP_DCT1 = 4,  /* nddct1..4  : real n -> real n, Default = scipy (2x rustdct), None = rustdct  */
    NDFB_OP_DCT2 = 5,
    NDFB_OP_DCT3 = 6,
    NDFB_OP_DCT4 = 7
};
enum ndfb_norm { NDFB_NORM_NONE = 0, NDFB_NORM_DEFAULT = 1 };
enum ndfb_mem  { NDFB_MEM_HOST = 0, NDFB_MEM_DEVICE = 1 };

enum ndfb_status {
    NDFB_OK = 0,
    NDFB_E_INVALID = -1,        /* bad enum / null pointer / ndim out of range                              */
    NDFB_E_SIZE_MISMATCH = -2,  /* lane length != handler length ("Size mismatch in fft|dct, got G expected E") */
    NDFB_E_SHAPE = -3,          /* non-transformed dimensions of input and output differ (ndarray Zip panic)    */
    NDFB_E_AXIS = -4,           /* axis >= ndim (index panic at src/lib.rs:116)                                 */
    NDFB_E_ALLOC = -5,
    NDFB_E_CUDA = -6,           /* CUDA error or no device: the product path never falls back to the CPU        */
    NDFB_E_UNSUPPORTED = -7     /* length outside what this build plans                         */
};

#define NDFB_MAX_DIMS 8

/* Build a plan for transforms of logical length n (the handler's `n`: the REAL length for R2C and DCT).
 * Works without a GPU (tables are uploaded lazily at first exec on `device`). */
NDFB_API int ndfb_plan_create(ndfb_plan** out, int kind, int dtype, size_t n, int device);
NDFB_API void ndfb_plan_destroy(ndfb_plan* plan);

/* Writes a JSON description of the schedule each op of this plan would run (kernel family, radix passes,
 * Bluestein length, tile geometry for a contiguous lane) into buf; returns the length needed. */
NDFB_API size_t ndfb_plan_describe(const ndfb_plan* plan, char* buf, size_t cap);

/* One nd* call.  shape/strides describe `in` and `out` as ndarray does: `ndim` extents and SIGNED strides
 * in ELEMENTS of the respective element type (real scalar, or interleaved {re,im} complex).  `mem` says
 * whether both pointers are host memory (synchronous) or device memory on the plan's device (asynchronous on
 * `stream`, a cudaStream_t; NULL = default stream).
 * Host memory may be PAGEABLE (what ndarray's as_ptr() hands the Rust shim, src/lib.rs:105-115): large C-ordered
 * arrays are cut into pieces and pipelined  caller memory -> pinned ring slot (copy threads) -> H2D | kernel | D2H ->
 * pinned ring slot -> caller memory;  pointers that are already pinned or cudaHostRegister'ed skip the ring.
 * Views with gaps between their elements are packed / unpacked on the host: the library never reads or writes a byte of
 * host memory that is not a logical element of the view (sibling views of one allocation stay intact).
 * Environment: NDFB_HOST_THREADS (copy threads, default min(12, 3/4 of the cores)), NDFB_STAGE_MB (slot size, default 16).
 * In place: `in == out` with identical shape and strides is supported for the ops that keep shape and element type
 * (FFT, IFFT, DCT1..4) — every tile is read completely before it is written (the reference always takes a separate
 * output, src/lib.rs:107).  Partially overlapping arrays are not supported. */
NDFB_API int ndfb_exec(const ndfb_plan* plan, int op, int norm,
              const void* in, void* out, int ndim,
              const size_t* shape_in, const ptrdiff_t* strides_in,
              const size_t* shape_out, const ptrdiff_t* strides_out,
              int axis, int mem, void* stream);

/* Same, with an explicit extra real factor multiplied into the result (1.0 = none).  Lets the shim fold
 * simple custom normalisations into the kernel epilogue instead of a host pass. */
NDFB_API int ndfb_exec_scaled(const ndfb_plan* plan, int op, int norm, double extra_scale,
                     const void* in, void* out, int ndim,
                     const size_t* shape_in, const ptrdiff_t* strides_in,
                     const size_t* shape_out, const ptrdiff_t* strides_out,
                     int axis, int mem, void* stream);

/* ndfft / ndifft on DEVICE memory whose OUTPUT axis is stored in blocks: output element k of a lane goes to
 *   lane_base + (k / out_block) * out_block_stride + (k % out_block) * strides_out[axis]
 * (elements).  This is the packed send layout of a slab all-to-all: the axis pass writes each destination rank's
 * chunk contiguously (or straight into a peer-mapped buffer), so the transpose needs no separate pack kernel.
 * No counterpart in the reference (it has no distributed path); used by ndrustfft_b200.dist. */
NDFB_API int ndfb_exec_split_out(const ndfb_plan* plan, int op, int norm, double extra_scale,
                        size_t out_block, ptrdiff_t out_block_stride,
                        const void* in, void* out, int ndim,
                        const size_t* shape_in, const ptrdiff_t* strides_in,
                        const size_t* shape_out, const ptrdiff_t* strides_out,
                        int axis, void* stream);

/* Same, but block p of every output lane is written relative to block_ptrs[p] (device pointers, 1..8 of them, e.g. the
 * peer-mapped receive buffers of the other GPUs): element k lands at
 *   block_ptrs[k / out_block] + lane_offset + (k % out_block) * strides_out[axis].
 * With NVLink peer mappings the axis pass's store IS the all-to-all of the slab transpose (fused compute + collective). */
NDFB_API int ndfb_exec_scatter_out(const ndfb_plan* plan, int op, int norm, double extra_scale,
                          size_t out_block, int nblocks, void* const* block_ptrs,
                          const void* in, int ndim,
                          const size_t* shape_in, const ptrdiff_t* strides_in,
                          const size_t* shape_out, const ptrdiff_t* strides_out,
                          int axis, void* stream);

/* Thread-local message of the last failing call on this thread ("" if none). */
NDFB_API const char* ndfb_last_error(void);

/* One axis transform of a chain. */
typedef struct ndfb_step {
    const ndfb_plan* plan;   /* all plans of a chain share dtype and device */
    int op;                  /* NDFB_OP_* valid for the plan's kind */
    int norm;                /* NDFB_NORM_NONE / NDFB_NORM_DEFAULT */
    int axis;
} ndfb_step;

/* Multi-axis transform: applies steps[0], steps[1], ... in order, step i reading what step i-1 wrote — the same
 * result as nsteps ndfb_exec calls through intermediate `work` arrays, which is how the reference's users compose
 * fft2 / rfft2 / fft3 (examples/fft2.rs:23-27 and :55-59, examples/rfft2.rs:29-33).  Here the
 * intermediates never leave the device: they live in `out` itself when shape and element type allow (later steps then
 * run in place) or in library workspaces, and with mem == NDFB_MEM_HOST the data crosses PCIe once in each direction
 * (first step overlapped with the upload, last step with the download) instead of once per axis.
 * shape_in / strides_in describe `in` (element type = input type of steps[0]); shape_out / strides_out describe `out`
 * (output type of the last step).  The shape evolves along each r2c / c2r axis exactly as in the single calls, and the
 * same size checks apply per step (NDFB_E_SIZE_MISMATCH carries the reference's message).  `in` is never modified;
 * `in == out` is allowed when strides are identical and every step keeps shape and element type (C2C, DCT). */
NDFB_API int ndfb_exec_chain(const ndfb_step* steps, int nsteps,
                    const void* in, void* out, int ndim,
                    const size_t* shape_in, const ptrdiff_t* strides_in,
                    const size_t* shape_out, const ptrdiff_t* strides_out,
                    int mem, void* stream);

/* Device memory and streams for hosts without CUDA bindings of their own (the Rust shim's `DeviceArray` / `Stream`,
 * SURVEY.md 8f-2): arrays allocated here are passed to ndfb_exec / ndfb_exec_chain with mem = NDFB_MEM_DEVICE, so a chain
 * of transforms pays PCIe once on the way in and once on the way out instead of on every axis (the reference's callers
 * compose axes through `work` arrays, examples/fft2.rs:23-27).  ndfb_memcpy is asynchronous on `stream` for H2D / D2D
 * (pageable host memory goes through the pinned ring) and returns after completion for D2H. */
enum ndfb_copy_kind { NDFB_COPY_H2D = 0, NDFB_COPY_D2H = 1, NDFB_COPY_D2D = 2 };
NDFB_API int ndfb_device_alloc(void** ptr, size_t bytes, int device);
NDFB_API void ndfb_device_free(void* ptr);
NDFB_API int ndfb_memcpy(void* dst, const void* src, size_t bytes, int kind, int device, void* stream);
NDFB_API int ndfb_stream_create(void** stream, int device);
NDFB_API void ndfb_stream_destroy(void* stream);
NDFB_API int ndfb_stream_sync(void* stream);

/* Library identification: "ndfft_b200 <version> sm_100a" for the CUDA build. */
NDFB_API const char* ndfb_version(void);

/* Number of kernel launches issued by this library since load (all threads); used by bench.py to report
 * `gpu_launches`. */
NDFB_API uint64_t ndfb_launch_count(void);

/* Lengths without an ahead-of-time instantiated schedule get the same register-resident Stockham kernels compiled at
 * run time for their own radix schedule (NVRTC -> sm_100a cubin, cached in memory and under $NDFB_JIT_CACHE or
 * ~/.cache/ndfft_b200; rustfft plans any length, src/lib.rs:294-304).  This entry point plans and compiles — without
 * loading, so it works on a machine without a GPU — the schedule for a `core_n`-point complex core
 * (rkind < 0: ndfft / ndifft;  0..5: the real kinds R2C, C2R, DCT-I..IV whose core is n/2 or n-1 points) and writes a
 * JSON description into `info`.  Returns NDFB_E_UNSUPPORTED when the length has no such schedule or libnvrtc is absent. */
NDFB_API int ndfb_jit_compile_check(int dtype, int rkind, size_t core_n, int cols, char* info, size_t cap);

/* Occupancy hint for the NEXT kernel launch of the calling thread: request at least `bytes` of dynamic shared memory
 * (e.g. 116 KiB => one CTA per SM), so that a link-bound kernel leaves room for another stream's HBM-bound kernel. */
NDFB_API void ndfb_hint_next_launch_smem(size_t bytes);

/* Producer / consumer overlap of two calls issued on DIFFERENT streams (no event between them): the next real-kind
 * fast-path launch of the calling thread (ndfft_r2c ...) adds the number of lanes of every finished tile to
 * counters[first_lane / lanes_per_group] (32-bit, zeroed by the caller); the next ndfft / ndifft launch on strided
 * columns runs as `ctas_per_sm` persistent CTAs per SM and loads a tile only after counters[first_lane / lanes_per_group]
 * has reached `need`.  Lane indices count the non-transformed elements in memory order of the INPUT of the respective
 * call.  dist.SlabR2cFft3d uses the pair to start the exchange pass on plane p as soon as plane p's r2c rows are done.
 * A call that cannot honour a pending hint returns NDFB_E_UNSUPPORTED (the consumer waits are bounded: no hang). */
NDFB_API void ndfb_hint_next_launch_signal(void* counters, long long lanes_per_group);
NDFB_API void ndfb_hint_next_launch_wait(const void* counters, long long lanes_per_group, unsigned need, int ctas_per_sm);

/* Release cached workspaces / pinned staging buffers held by the calling thread's pools. */
NDFB_API void ndfb_release_workspaces(void);

#ifdef __cplusplus
}
#endif
#endif /* NDFFT_B200_H */
