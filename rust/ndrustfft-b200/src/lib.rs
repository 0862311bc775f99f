//! # ndrustfft-b200: ndrustfft's API on an NVIDIA B200
//!
//! Same public names as `ndrustfft` 0.5 (`Normalization`, `FftHandler`, `R2cFftHandler`, `DctHandler`,
//! `ndfft`, `ndifft`, `ndfft_r2c`, `ndifft_r2c`, `nddct1..4` and their `_par` twins) so that
//! `use ndrustfft_b200 as ndrustfft;` is the whole port.  Every call is one `ndfb_exec` on the C ABI of
//! `include/ndfft_b200.h`; the arithmetic runs in hand-written sm_100a kernels.  There is no CPU fallback:
//! without a usable GPU every transform panics with the library's error text.
//!
//! This crate is source-only in the build image (no Rust toolchain there); see INTEGRATION.md.
#![warn(missing_docs)]
mod device;
mod ffi;

pub use device::{
    fft2, ifft2, irfft2, ndchain, ndchain_dev, nddct_dev, ndfft_dev, ndfft_r2c_dev, ndifft_dev, ndifft_r2c_dev, rfft2, DeviceArray, Step, Stream,
};
use ndarray::{ArrayBase, Data, DataMut, Dimension};
use num_traits::FloatConst;
use std::any::TypeId;
use std::ffi::CStr;
use std::os::raw::{c_int, c_void};
use std::sync::Arc;

// The reference re-exports exactly these three names (src/lib.rs:83-85):
//     pub use rustfft::FftNum;  pub use rustfft::num_complex::Complex;  pub use rustfft::num_traits::Zero;
// With the `rustfft` feature the SAME trait object is re-exported, so code that is generic over `rustfft::FftNum`
// keeps compiling; without it (the dependency-light default) a trait of the same name and the same supertraits as
// rustfft 6's stands in.  The functions below bound `T: FftNum + FloatConst` exactly as the reference does (:111) and
// find the element type at run time, so no extra bound leaks into callers' generic code.
pub use num_complex::Complex;
pub use num_traits::Zero;
#[cfg(feature = "rustfft")]
pub use rustfft::FftNum;
/// Generic floating point number (`rustfft::FftNum`'s supertraits); implemented for `f32` and `f64`.
#[cfg(not(feature = "rustfft"))]
pub trait FftNum: Copy + num_traits::FromPrimitive + num_traits::Signed + Sync + Send + std::fmt::Debug + 'static {}
#[cfg(not(feature = "rustfft"))]
impl<T> FftNum for T where T: Copy + num_traits::FromPrimitive + num_traits::Signed + Sync + Send + std::fmt::Debug + 'static {}

/// `NDFB_F32` / `NDFB_F64` for `T`; the reference supports exactly these two (`FftNum + FloatConst`, src/lib.rs:111).
pub(crate) fn dtype_of<T: 'static>() -> c_int {
    if TypeId::of::<T>() == TypeId::of::<f32>() {
        ffi::NDFB_F32
    } else if TypeId::of::<T>() == TypeId::of::<f64>() {
        ffi::NDFB_F64
    } else {
        panic!("ndrustfft-b200 supports f32 and f64")
    }
}

/// Represents different types of normalization methods (reference src/lib.rs:89-98).
#[derive(Clone)]
pub enum Normalization<T> {
    /// No normalization applied, output equals `rustfft`, `realfft` or `rustdct`.
    None,
    /// Applies normalization similar to scipy's default behavior.
    Default,
    /// Applies a custom normalization function.  It is a host function: lanes are staged through the host.
    Custom(fn(&mut [T])),
}

pub(crate) struct PlanHandle(pub(crate) *mut ffi::NdfbPlan);
// The C plan is immutable after creation and internally synchronised (include/ndfft_b200.h).
unsafe impl Send for PlanHandle {}
unsafe impl Sync for PlanHandle {}
impl Drop for PlanHandle {
    fn drop(&mut self) {
        unsafe { ffi::ndfb_plan_destroy(self.0) }
    }
}

pub(crate) fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::ndfb_last_error()).to_string_lossy().into_owned() }
}

fn new_plan(kind: c_int, dtype: c_int, n: usize) -> Arc<PlanHandle> {
    let mut p: *mut ffi::NdfbPlan = std::ptr::null_mut();
    let device = std::env::var("NDFB_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
    let rc = unsafe { ffi::ndfb_plan_create(&mut p, kind, dtype, n, device) };
    assert!(rc == 0, "{}", last_error());
    Arc::new(PlanHandle(p))
}

#[allow(clippy::too_many_arguments)]
fn exec<A, B, R, S, D>(
    plan: &PlanHandle, op: c_int, norm: c_int, input: &ArrayBase<R, D>, output: &mut ArrayBase<S, D>, axis: usize,
) where
    R: Data<Elem = A>,
    S: Data<Elem = B> + DataMut,
    D: Dimension,
{
    let _ = output.shape()[axis]; // same index panic as the reference (src/lib.rs:116)
    let rc = unsafe {
        ffi::ndfb_exec(
            plan.0, op, norm,
            input.as_ptr() as *const c_void, output.as_mut_ptr() as *mut c_void,
            input.ndim() as c_int,
            input.shape().as_ptr(), input.strides().as_ptr(),     // ndarray strides are in elements, signed
            output.shape().as_ptr(), output.strides().as_ptr(),
            axis as c_int, ffi::NDFB_MEM_HOST, std::ptr::null_mut(),
        )
    };
    // assert_size text "Size mismatch in fft, got {} expected {}" comes back verbatim (src/lib.rs:340-347)
    assert!(rc == 0, "{}", last_error());
}

/// Apply a host normalisation function to every lane of `arr` along `axis` (the reference calls it per lane).
fn apply_lanes<T: Clone, S: DataMut<Elem = T>, D: Dimension>(f: fn(&mut [T]), arr: &mut ArrayBase<S, D>, axis: usize) {
    for mut lane in arr.lanes_mut(ndarray::Axis(axis)) {
        if let Some(s) = lane.as_slice_mut() {
            f(s);
        } else {
            let mut tmp = lane.to_vec();
            f(&mut tmp);
            lane.assign(&ndarray::ArrayView1::from(&tmp));
        }
    }
}

pub(crate) fn norm_code<T>(n: &Normalization<T>) -> c_int {
    match n {
        Normalization::Default => ffi::NDFB_NORM_DEFAULT,
        _ => ffi::NDFB_NORM_NONE,
    }
}

// ------------------------------------------------------------------------------------------------
/// *n*-dimensional complex-to-complex Fourier Transform handler (reference src/lib.rs:270-348).
#[derive(Clone)]
pub struct FftHandler<T> {
    n: usize,
    pub(crate) plan: Arc<PlanHandle>,
    pub(crate) norm: Normalization<Complex<T>>,
}

impl<T: FftNum> FftHandler<T> {
    /// Creates a new `FftHandler` for transforms of length `n` (reference src/lib.rs:294-304).
    #[must_use]
    pub fn new(n: usize) -> Self {
        Self { n, plan: new_plan(ffi::NDFB_C2C, dtype_of::<T>(), n), norm: Normalization::Default }
    }
    /// Modifies the normalization applied to the backward transform (reference src/lib.rs:308-311).
    #[must_use]
    pub fn normalization(mut self, norm: Normalization<Complex<T>>) -> Self {
        self.norm = norm;
        self
    }
    /// Transform length (not part of the reference's API).
    #[doc(hidden)]
    #[allow(clippy::len_without_is_empty)]
    pub fn len(&self) -> usize {
        self.n
    }
}

/// Complex-to-complex Fourier Transform (reference src/lib.rs:350-372).
pub fn ndfft<R, S, T, D>(input: &ArrayBase<R, D>, output: &mut ArrayBase<S, D>, handler: &FftHandler<T>, axis: usize)
where
    T: FftNum + FloatConst,
    R: Data<Elem = Complex<T>>,
    S: Data<Elem = Complex<T>> + DataMut,
    D: Dimension,
{
    exec(&handler.plan, ffi::NDFB_OP_FFT, ffi::NDFB_NORM_NONE, input, output, axis);
}

/// Complex-to-complex inverse Fourier Transform (reference src/lib.rs:374-397).
pub fn ndifft<R, S, T, D>(input: &ArrayBase<R, D>, output: &mut ArrayBase<S, D>, handler: &FftHandler<T>, axis: usize)
where
    T: FftNum + FloatConst,
    R: Data<Elem = Complex<T>>,
    S: Data<Elem = Complex<T>> + DataMut,
    D: Dimension,
{
    exec(&handler.plan, ffi::NDFB_OP_IFFT, norm_code(&handler.norm), input, output, axis);
    if let Normalization::Custom(f) = handler.norm {
        apply_lanes(f, output, axis); // after the transform, on the output lane (src/lib.rs:329)
    }
}

// ------------------------------------------------------------------------------------------------
/// *n*-dimensional real-to-complex Fourier Transform handler (reference src/lib.rs:452-541).
#[derive(Clone)]
pub struct R2cFftHandler<T> {
    n: usize,
    m: usize,
    pub(crate) plan: Arc<PlanHandle>,
    pub(crate) norm: Normalization<Complex<T>>,
}

impl<T: FftNum> R2cFftHandler<T> {
    /// Creates a new handler for real length `n`; the spectrum has `n / 2 + 1` entries (reference src/lib.rs:477-488).
    #[must_use]
    pub fn new(n: usize) -> Self {
        Self { n, m: n / 2 + 1, plan: new_plan(ffi::NDFB_R2C, dtype_of::<T>(), n), norm: Normalization::Default }
    }
    /// Modifies the normalization applied to the backward transform (reference src/lib.rs:492-495).
    #[must_use]
    pub fn normalization(mut self, norm: Normalization<Complex<T>>) -> Self {
        self.norm = norm;
        self
    }
    /// (real length, spectrum length) (not part of the reference's API).
    #[doc(hidden)]
    pub fn lens(&self) -> (usize, usize) {
        (self.n, self.m)
    }
}

/// Real-to-complex Fourier Transform (reference src/lib.rs:543-564).
pub fn ndfft_r2c<R, S, T, D>(input: &ArrayBase<R, D>, output: &mut ArrayBase<S, D>, handler: &R2cFftHandler<T>, axis: usize)
where
    T: FftNum + FloatConst,
    R: Data<Elem = T>,
    S: Data<Elem = Complex<T>> + DataMut,
    D: Dimension,
{
    exec(&handler.plan, ffi::NDFB_OP_R2C, ffi::NDFB_NORM_NONE, input, output, axis);
}

/// Complex-to-real inverse Fourier Transform (reference src/lib.rs:566-587).
pub fn ndifft_r2c<R, S, T, D>(input: &ArrayBase<R, D>, output: &mut ArrayBase<S, D>, handler: &R2cFftHandler<T>, axis: usize)
where
    T: FftNum + FloatConst,
    R: Data<Elem = Complex<T>>,
    S: Data<Elem = T> + DataMut,
    D: Dimension,
{
    if let Normalization::Custom(f) = handler.norm {
        // the closure sees the m-long spectrum copy BEFORE the transform (src/lib.rs:509-515)
        let mut staged = input.to_owned();
        apply_lanes(f, &mut staged, axis);
        exec(&handler.plan, ffi::NDFB_OP_C2R, ffi::NDFB_NORM_NONE, &staged, output, axis);
    } else {
        exec(&handler.plan, ffi::NDFB_OP_C2R, norm_code(&handler.norm), input, output, axis);
    }
}

// ------------------------------------------------------------------------------------------------
/// *n*-dimensional real-to-real Cosine Transform handler, DCT-I..IV (reference src/lib.rs:641-751).
#[derive(Clone)]
pub struct DctHandler<T> {
    n: usize,
    pub(crate) plan: Arc<PlanHandle>,
    pub(crate) norm: Normalization<T>,
}

impl<T: FftNum> DctHandler<T> {
    /// Creates a new `DctHandler` (reference src/lib.rs:665-679); the four schedules are built lazily on first use.
    #[must_use]
    pub fn new(n: usize) -> Self {
        Self { n, plan: new_plan(ffi::NDFB_DCT, dtype_of::<T>(), n), norm: Normalization::Default }
    }
    /// Modifies the normalization (reference src/lib.rs:683-686).
    #[must_use]
    pub fn normalization(mut self, norm: Normalization<T>) -> Self {
        self.norm = norm;
        self
    }
    /// Transform length (not part of the reference's API).
    #[doc(hidden)]
    #[allow(clippy::len_without_is_empty)]
    pub fn len(&self) -> usize {
        self.n
    }
}

fn dct<R, S, T, D>(op: c_int, input: &ArrayBase<R, D>, output: &mut ArrayBase<S, D>, handler: &DctHandler<T>, axis: usize)
where
    T: FftNum + FloatConst,
    R: Data<Elem = T>,
    S: Data<Elem = T> + DataMut,
    D: Dimension,
{
    if let Normalization::Custom(f) = handler.norm {
        let mut staged = input.to_owned(); // on the input copy, before the transform (src/lib.rs:691-696)
        apply_lanes(f, &mut staged, axis);
        exec(&handler.plan, op, ffi::NDFB_NORM_NONE, &staged, output, axis);
    } else {
        exec(&handler.plan, op, norm_code(&handler.norm), input, output, axis);
    }
}

macro_rules! dct_fn {
    ($(#[$m:meta])* $name:ident, $op:expr) => {
        $(#[$m])*
        pub fn $name<R, S, T, D>(input: &ArrayBase<R, D>, output: &mut ArrayBase<S, D>, handler: &DctHandler<T>, axis: usize)
        where
            T: FftNum + FloatConst,
            R: Data<Elem = T>,
            S: Data<Elem = T> + DataMut,
            D: Dimension,
        {
            dct($op, input, output, handler, axis)
        }
    };
}
dct_fn!(/// DCT-I (reference src/lib.rs:753-775).
    nddct1, ffi::NDFB_OP_DCT1);
dct_fn!(/// DCT-II (reference src/lib.rs:789-796).
    nddct2, ffi::NDFB_OP_DCT2);
dct_fn!(/// DCT-III (reference src/lib.rs:808-815).
    nddct3, ffi::NDFB_OP_DCT3);
dct_fn!(/// DCT-IV (reference src/lib.rs:827-834).
    nddct4, ffi::NDFB_OP_DCT4);

/// `_par` twins (reference src/lib.rs:399-421, 589-611, 777-844): the GPU already runs all lanes in parallel.
#[cfg(feature = "parallel")]
pub use self::{
    nddct1 as nddct1_par, nddct2 as nddct2_par, nddct3 as nddct3_par, nddct4 as nddct4_par, ndfft as ndfft_par,
    ndfft_r2c as ndfft_r2c_par, ndifft as ndifft_par, ndifft_r2c as ndifft_r2c_par,
};
