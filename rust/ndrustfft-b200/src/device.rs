//! Device-resident arrays, streams and multi-axis chains.
//!
//! The reference composes N-D transforms through host `work` arrays (examples/fft2.rs:23-27, examples/rfft2.rs:29-33).
//! With a GPU behind the same API that pattern would cross PCIe twice per axis; these types let a caller upload once,
//! run any number of `nd*_dev` calls (or one `ndchain`) on a stream, and download once.
use crate::{dtype_of, ffi, last_error, norm_code, DctHandler, FftHandler, FftNum, R2cFftHandler};
use ndarray::{Array, ArrayBase, Data, DataMut, Dimension};
use num_complex::Complex;
use num_traits::FloatConst;
use std::marker::PhantomData;
use std::os::raw::{c_int, c_void};

fn device_index() -> c_int {
    std::env::var("NDFB_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0)
}

/// A CUDA stream owned by the library (`ndfb_stream_create`).  `Stream::default()` is the legacy default stream.
pub struct Stream {
    raw: *mut c_void,
    owned: bool,
}
unsafe impl Send for Stream {}
impl Default for Stream {
    fn default() -> Self {
        Stream { raw: std::ptr::null_mut(), owned: false }
    }
}
impl Stream {
    /// A new non-blocking stream on the device selected by `NDFB_DEVICE` (default 0).
    pub fn new() -> Self {
        let mut raw: *mut c_void = std::ptr::null_mut();
        let rc = unsafe { ffi::ndfb_stream_create(&mut raw, device_index()) };
        assert!(rc == 0, "{}", last_error());
        Stream { raw, owned: true }
    }
    /// Blocks until everything queued on the stream has finished.
    pub fn synchronize(&self) {
        let rc = unsafe { ffi::ndfb_stream_sync(self.raw) };
        assert!(rc == 0, "{}", last_error());
    }
    pub(crate) fn raw(&self) -> *mut c_void {
        self.raw
    }
}
impl Drop for Stream {
    fn drop(&mut self) {
        if self.owned {
            unsafe { ffi::ndfb_stream_destroy(self.raw) }
        }
    }
}

/// An owned, C-ordered n-dimensional array in GPU memory with element type `A` (`T` or `Complex<T>`).
pub struct DeviceArray<A, D: Dimension> {
    ptr: *mut c_void,
    dim: D,
    strides: Vec<isize>,
    _elem: PhantomData<A>,
}
unsafe impl<A: Send, D: Dimension> Send for DeviceArray<A, D> {}

impl<A: Copy, D: Dimension> DeviceArray<A, D> {
    /// Uninitialised device array of the given shape.
    pub fn uninit(dim: D) -> Self {
        let n: usize = dim.slice().iter().product();
        let mut ptr: *mut c_void = std::ptr::null_mut();
        let rc = unsafe { ffi::ndfb_device_alloc(&mut ptr, n * std::mem::size_of::<A>(), device_index()) };
        assert!(rc == 0, "{}", last_error());
        let mut strides = vec![0isize; dim.ndim()];
        let mut acc = 1isize;
        for (d, s) in dim.slice().iter().enumerate().rev() {
            strides[d] = acc;
            acc *= *s as isize;
        }
        DeviceArray { ptr, dim, strides, _elem: PhantomData }
    }
    /// Uploads a host array (any layout; non-standard layouts are made contiguous first).  Asynchronous on `stream` for
    /// pinned memory; pageable memory is staged through the library's pinned ring.
    pub fn from_array<S: Data<Elem = A>>(host: &ArrayBase<S, D>, stream: &Stream) -> Self {
        let dev = Self::uninit(host.raw_dim());
        let owned;
        let src = if host.is_standard_layout() {
            host.as_ptr()
        } else {
            owned = host.as_standard_layout().into_owned();
            owned.as_ptr()
        };
        let rc = unsafe {
            ffi::ndfb_memcpy(dev.ptr, src as *const c_void, dev.len() * std::mem::size_of::<A>(), ffi::NDFB_COPY_H2D, device_index(), stream.raw())
        };
        assert!(rc == 0, "{}", last_error());
        if !host.is_standard_layout() {
            stream.synchronize(); // `owned` is dropped at the end of this function
        }
        dev
    }
    /// Downloads into a new host array (returns after the copy has completed).
    pub fn to_array(&self, stream: &Stream) -> Array<A, D>
    where
        A: num_traits::Zero,
    {
        let mut host = Array::<A, D>::zeros(self.dim.clone());
        self.copy_to(&mut host, stream);
        host
    }
    /// Downloads into an existing standard-layout host array of the same shape.
    pub fn copy_to<S: DataMut<Elem = A>>(&self, host: &mut ArrayBase<S, D>, stream: &Stream) {
        assert!(host.shape() == self.dim.slice(), "shape mismatch");
        assert!(host.is_standard_layout(), "copy_to needs a standard-layout host array");
        let rc = unsafe {
            ffi::ndfb_memcpy(host.as_mut_ptr() as *mut c_void, self.ptr, self.len() * std::mem::size_of::<A>(), ffi::NDFB_COPY_D2H, device_index(), stream.raw())
        };
        assert!(rc == 0, "{}", last_error());
    }
    /// Number of elements.
    pub fn len(&self) -> usize {
        self.dim.slice().iter().product()
    }
    /// True when the array has no elements.
    pub fn is_empty(&self) -> bool {
        self.len() == 0
    }
    /// Shape, as `ndarray` reports it.
    pub fn shape(&self) -> &[usize] {
        self.dim.slice()
    }
    /// Raw device pointer (e.g. to hand to other CUDA code).
    pub fn as_device_ptr(&self) -> *mut c_void {
        self.ptr
    }
}
impl<A, D: Dimension> Drop for DeviceArray<A, D> {
    fn drop(&mut self) {
        unsafe { ffi::ndfb_device_free(self.ptr) }
    }
}

#[allow(clippy::too_many_arguments)]
fn exec_dev<A: Copy, B: Copy, D: Dimension>(
    plan: *const ffi::NdfbPlan, op: c_int, norm: c_int, input: &DeviceArray<A, D>, output: &mut DeviceArray<B, D>, axis: usize, stream: &Stream,
) {
    let _ = output.shape()[axis]; // same index panic as the reference (src/lib.rs:116)
    let rc = unsafe {
        ffi::ndfb_exec(
            plan, op, norm, input.ptr as *const c_void, output.ptr, input.dim.ndim() as c_int,
            input.dim.slice().as_ptr(), input.strides.as_ptr(), output.dim.slice().as_ptr(), output.strides.as_ptr(),
            axis as c_int, ffi::NDFB_MEM_DEVICE, stream.raw(),
        )
    };
    assert!(rc == 0, "{}", last_error());
}

fn no_custom<T>(n: &crate::Normalization<T>) {
    if let crate::Normalization::Custom(_) = n {
        panic!("Normalization::Custom is a host callback: use the host-array functions, or None/Default on device arrays");
    }
}

/// `ndfft` on device arrays, asynchronous on `stream`.
pub fn ndfft_dev<T: FftNum + FloatConst, D: Dimension>(
    input: &DeviceArray<Complex<T>, D>, output: &mut DeviceArray<Complex<T>, D>, handler: &FftHandler<T>, axis: usize, stream: &Stream,
) {
    exec_dev(handler.plan.0, ffi::NDFB_OP_FFT, ffi::NDFB_NORM_NONE, input, output, axis, stream);
}
/// `ndifft` on device arrays.
pub fn ndifft_dev<T: FftNum + FloatConst, D: Dimension>(
    input: &DeviceArray<Complex<T>, D>, output: &mut DeviceArray<Complex<T>, D>, handler: &FftHandler<T>, axis: usize, stream: &Stream,
) {
    no_custom(&handler.norm);
    exec_dev(handler.plan.0, ffi::NDFB_OP_IFFT, norm_code(&handler.norm), input, output, axis, stream);
}
/// `ndfft_r2c` on device arrays.
pub fn ndfft_r2c_dev<T: FftNum + FloatConst, D: Dimension>(
    input: &DeviceArray<T, D>, output: &mut DeviceArray<Complex<T>, D>, handler: &R2cFftHandler<T>, axis: usize, stream: &Stream,
) {
    exec_dev(handler.plan.0, ffi::NDFB_OP_R2C, ffi::NDFB_NORM_NONE, input, output, axis, stream);
}
/// `ndifft_r2c` on device arrays.
pub fn ndifft_r2c_dev<T: FftNum + FloatConst, D: Dimension>(
    input: &DeviceArray<Complex<T>, D>, output: &mut DeviceArray<T, D>, handler: &R2cFftHandler<T>, axis: usize, stream: &Stream,
) {
    no_custom(&handler.norm);
    exec_dev(handler.plan.0, ffi::NDFB_OP_C2R, norm_code(&handler.norm), input, output, axis, stream);
}
/// `nddct1..4` on device arrays (`kind` = 1..=4).
pub fn nddct_dev<T: FftNum + FloatConst, D: Dimension>(
    kind: u8, input: &DeviceArray<T, D>, output: &mut DeviceArray<T, D>, handler: &DctHandler<T>, axis: usize, stream: &Stream,
) {
    assert!((1..=4).contains(&kind), "DCT kind must be 1..=4");
    no_custom(&handler.norm);
    exec_dev(handler.plan.0, ffi::NDFB_OP_DCT1 + (kind as c_int - 1), norm_code(&handler.norm), input, output, axis, stream);
}

/// One step of a multi-axis chain: which transform, with which handler, along which axis.
pub enum Step<'a, T> {
    /// `ndfft`
    Fft(&'a FftHandler<T>, usize),
    /// `ndifft`
    Ifft(&'a FftHandler<T>, usize),
    /// `ndfft_r2c`
    R2c(&'a R2cFftHandler<T>, usize),
    /// `ndifft_r2c`
    C2r(&'a R2cFftHandler<T>, usize),
    /// `nddct1..4` (kind, handler, axis)
    Dct(u8, &'a DctHandler<T>, usize),
}

fn lower<T: FftNum>(steps: &[Step<T>]) -> Vec<ffi::NdfbStep> {
    steps
        .iter()
        .map(|s| match s {
            Step::Fft(h, ax) => ffi::NdfbStep { plan: h.plan.0, op: ffi::NDFB_OP_FFT, norm: ffi::NDFB_NORM_NONE, axis: *ax as c_int },
            Step::Ifft(h, ax) => {
                no_custom(&h.norm);
                ffi::NdfbStep { plan: h.plan.0, op: ffi::NDFB_OP_IFFT, norm: norm_code(&h.norm), axis: *ax as c_int }
            }
            Step::R2c(h, ax) => ffi::NdfbStep { plan: h.plan.0, op: ffi::NDFB_OP_R2C, norm: ffi::NDFB_NORM_NONE, axis: *ax as c_int },
            Step::C2r(h, ax) => {
                no_custom(&h.norm);
                ffi::NdfbStep { plan: h.plan.0, op: ffi::NDFB_OP_C2R, norm: norm_code(&h.norm), axis: *ax as c_int }
            }
            Step::Dct(k, h, ax) => {
                assert!((1..=4).contains(k), "DCT kind must be 1..=4");
                no_custom(&h.norm);
                ffi::NdfbStep { plan: h.plan.0, op: ffi::NDFB_OP_DCT1 + (*k as c_int - 1), norm: norm_code(&h.norm), axis: *ax as c_int }
            }
        })
        .collect()
}

/// Applies `steps` in order as ONE call (`ndfb_exec_chain`): same result as the separate `nd*` calls through `work`
/// arrays, but the intermediates stay on the GPU and the host arrays cross PCIe once each way.
pub fn ndchain<A, B, R, S, T, D>(input: &ArrayBase<R, D>, output: &mut ArrayBase<S, D>, steps: &[Step<T>])
where
    T: FftNum + FloatConst,
    R: Data<Elem = A>,
    S: Data<Elem = B> + DataMut,
    D: Dimension,
{
    let _ = dtype_of::<T>();
    let low = lower(steps);
    let rc = unsafe {
        ffi::ndfb_exec_chain(
            low.as_ptr(), low.len() as c_int, input.as_ptr() as *const c_void, output.as_mut_ptr() as *mut c_void, input.ndim() as c_int,
            input.shape().as_ptr(), input.strides().as_ptr(), output.shape().as_ptr(), output.strides().as_ptr(),
            ffi::NDFB_MEM_HOST, std::ptr::null_mut(),
        )
    };
    assert!(rc == 0, "{}", last_error());
}

/// `ndchain` on device arrays, asynchronous on `stream`.
pub fn ndchain_dev<A: Copy, B: Copy, T: FftNum + FloatConst, D: Dimension>(
    input: &DeviceArray<A, D>, output: &mut DeviceArray<B, D>, steps: &[Step<T>], stream: &Stream,
) {
    let low = lower(steps);
    let rc = unsafe {
        ffi::ndfb_exec_chain(
            low.as_ptr(), low.len() as c_int, input.ptr as *const c_void, output.ptr, input.dim.ndim() as c_int,
            input.dim.slice().as_ptr(), input.strides.as_ptr(), output.dim.slice().as_ptr(), output.strides.as_ptr(),
            ffi::NDFB_MEM_DEVICE, stream.raw(),
        )
    };
    assert!(rc == 0, "{}", last_error());
}

/// examples/fft2.rs:23-27 as one call: `ndfft` along axis 1, then along axis 0.
pub fn fft2<R, S, T>(input: &ArrayBase<R, ndarray::Ix2>, output: &mut ArrayBase<S, ndarray::Ix2>, handler_ax0: &FftHandler<T>, handler_ax1: &FftHandler<T>)
where
    T: FftNum + FloatConst,
    R: Data<Elem = Complex<T>>,
    S: Data<Elem = Complex<T>> + DataMut,
{
    ndchain(input, output, &[Step::Fft(handler_ax1, 1), Step::Fft(handler_ax0, 0)]);
}
/// examples/fft2.rs:55-59 as one call: `ndifft` along axis 0, then along axis 1.
pub fn ifft2<R, S, T>(input: &ArrayBase<R, ndarray::Ix2>, output: &mut ArrayBase<S, ndarray::Ix2>, handler_ax0: &FftHandler<T>, handler_ax1: &FftHandler<T>)
where
    T: FftNum + FloatConst,
    R: Data<Elem = Complex<T>>,
    S: Data<Elem = Complex<T>> + DataMut,
{
    ndchain(input, output, &[Step::Ifft(handler_ax0, 0), Step::Ifft(handler_ax1, 1)]);
}
/// examples/rfft2.rs:29-33 as one call: `ndfft_r2c` along axis 1, then `ndfft` along axis 0.
pub fn rfft2<R, S, T>(input: &ArrayBase<R, ndarray::Ix2>, output: &mut ArrayBase<S, ndarray::Ix2>, handler_ax0: &FftHandler<T>, handler_ax1: &R2cFftHandler<T>)
where
    T: FftNum + FloatConst,
    R: Data<Elem = T>,
    S: Data<Elem = Complex<T>> + DataMut,
{
    ndchain(input, output, &[Step::R2c(handler_ax1, 1), Step::Fft(handler_ax0, 0)]);
}
/// examples/rfft2.rs:49-53 as one call: `ndifft` along axis 0, then `ndifft_r2c` along axis 1.
pub fn irfft2<R, S, T>(input: &ArrayBase<R, ndarray::Ix2>, output: &mut ArrayBase<S, ndarray::Ix2>, handler_ax0: &FftHandler<T>, handler_ax1: &R2cFftHandler<T>)
where
    T: FftNum + FloatConst,
    R: Data<Elem = Complex<T>>,
    S: Data<Elem = T> + DataMut,
{
    ndchain(input, output, &[Step::Ifft(handler_ax0, 0), Step::C2r(handler_ax1, 1)]);
}
