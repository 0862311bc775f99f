//! Raw bindings of include/ndfft_b200.h.
use std::os::raw::{c_char, c_double, c_int, c_void};

#[repr(C)]
pub struct NdfbPlan {
    _private: [u8; 0],
}

pub const NDFB_C2C: c_int = 0;
pub const NDFB_R2C: c_int = 1;
pub const NDFB_DCT: c_int = 2;
pub const NDFB_F32: c_int = 0;
pub const NDFB_F64: c_int = 1;
pub const NDFB_OP_FFT: c_int = 0;
pub const NDFB_OP_IFFT: c_int = 1;
pub const NDFB_OP_R2C: c_int = 2;
pub const NDFB_OP_C2R: c_int = 3;
pub const NDFB_OP_DCT1: c_int = 4;
pub const NDFB_OP_DCT2: c_int = 5;
pub const NDFB_OP_DCT3: c_int = 6;
pub const NDFB_OP_DCT4: c_int = 7;
pub const NDFB_NORM_NONE: c_int = 0;
pub const NDFB_NORM_DEFAULT: c_int = 1;
pub const NDFB_MEM_HOST: c_int = 0;
pub const NDFB_MEM_DEVICE: c_int = 1;
pub const NDFB_COPY_H2D: c_int = 0;
pub const NDFB_COPY_D2H: c_int = 1;
pub const NDFB_COPY_D2D: c_int = 2;

/// `struct ndfb_step` of the C header.
#[repr(C)]
#[derive(Clone, Copy)]
pub struct NdfbStep {
    pub plan: *const NdfbPlan,
    pub op: c_int,
    pub norm: c_int,
    pub axis: c_int,
}

extern "C" {
    pub fn ndfb_plan_create(out: *mut *mut NdfbPlan, kind: c_int, dtype: c_int, n: usize, device: c_int) -> c_int;
    pub fn ndfb_plan_destroy(plan: *mut NdfbPlan);
    pub fn ndfb_exec(
        plan: *const NdfbPlan, op: c_int, norm: c_int, input: *const c_void, output: *mut c_void, ndim: c_int,
        shape_in: *const usize, strides_in: *const isize, shape_out: *const usize, strides_out: *const isize,
        axis: c_int, mem: c_int, stream: *mut c_void,
    ) -> c_int;
    pub fn ndfb_exec_scaled(
        plan: *const NdfbPlan, op: c_int, norm: c_int, extra_scale: c_double, input: *const c_void,
        output: *mut c_void, ndim: c_int, shape_in: *const usize, strides_in: *const isize,
        shape_out: *const usize, strides_out: *const isize, axis: c_int, mem: c_int, stream: *mut c_void,
    ) -> c_int;
    /// Multi-axis chain (include/ndfft_b200.h: ndfb_exec_chain); `steps` points at `nsteps` NdfbStep records.
    pub fn ndfb_exec_chain(
        steps: *const NdfbStep, nsteps: c_int, input: *const c_void, output: *mut c_void, ndim: c_int,
        shape_in: *const usize, strides_in: *const isize, shape_out: *const usize, strides_out: *const isize,
        mem: c_int, stream: *mut c_void,
    ) -> c_int;
    /// Device memory / stream helpers (include/ndfft_b200.h) behind `DeviceArray` and `Stream`.
    pub fn ndfb_device_alloc(ptr: *mut *mut c_void, bytes: usize, device: c_int) -> c_int;
    pub fn ndfb_device_free(ptr: *mut c_void);
    pub fn ndfb_memcpy(dst: *mut c_void, src: *const c_void, bytes: usize, kind: c_int, device: c_int, stream: *mut c_void) -> c_int;
    pub fn ndfb_stream_create(stream: *mut *mut c_void, device: c_int) -> c_int;
    pub fn ndfb_stream_destroy(stream: *mut c_void);
    pub fn ndfb_stream_sync(stream: *mut c_void) -> c_int;
    pub fn ndfb_last_error() -> *const c_char;
    pub fn ndfb_version() -> *const c_char;
}
