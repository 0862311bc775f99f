// ndrustfft_b200.hpp — header-only C++ mirror of ndrustfft's public API over the C ABI of ndfft_b200.h.
//
// Same names and argument order as the reference (src/lib.rs): Normalization {None, Default, Custom(fn)},
// FftHandler<T> / R2cFftHandler<T> / DctHandler<T> with new(n) + normalization(..) builder, and
// ndfft / ndifft / ndfft_r2c / ndifft_r2c / nddct1..4 (+ _par twins) taking (input, output, handler, axis).
// Arrays are described by ndview<E>: pointer + shape + signed element strides (what ndarray's ArrayBase carries).
// Errors: the reference panics; here std::runtime_error carries the same text ("Size mismatch in fft, got .. expected ..").
#pragma once
#include <complex>
#include <cstddef>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "ndfft_b200.h"

namespace ndrustfft_b200 {

template <typename T> struct dtype_of;
template <> struct dtype_of<float> { static constexpr int value = NDFB_F32; };
template <> struct dtype_of<double> { static constexpr int value = NDFB_F64; };

// A strided n-dimensional view (host memory by default; set device = true for device pointers).
template <typename E>
struct ndview {
    E* data = nullptr;
    std::vector<size_t> shape;
    std::vector<ptrdiff_t> strides;  // in elements, may be negative
    bool device = false;
    static ndview c_order(E* p, std::vector<size_t> shp, bool dev = false) {
        ndview v; v.data = p; v.shape = shp; v.strides.assign(shp.size(), 1); v.device = dev;
        ptrdiff_t s = 1;
        for (size_t d = shp.size(); d-- > 0;) { v.strides[d] = s; s *= (ptrdiff_t)shp[d]; }
        return v;
    }
    size_t ndim() const { return shape.size(); }
};

// Normalization<T> (src/lib.rs:89-98)
template <typename T>
struct Normalization {
    enum Kind { None, Default, Custom } kind = Default;
    std::function<void(T*, size_t)> func;  // fn(&mut [T])
    static Normalization none() { return {None, {}}; }
    static Normalization dflt() { return {Default, {}}; }
    static Normalization custom(std::function<void(T*, size_t)> f) { return {Custom, std::move(f)}; }
};

namespace detail {
struct PlanDeleter { void operator()(ndfb_plan* p) const { ndfb_plan_destroy(p); } };
inline std::shared_ptr<ndfb_plan> make_plan(int kind, int dtype, size_t n, int device) {
    ndfb_plan* p = nullptr;
    if (ndfb_plan_create(&p, kind, dtype, n, device) != 0) throw std::runtime_error(ndfb_last_error());
    return std::shared_ptr<ndfb_plan>(p, PlanDeleter());
}
template <typename A, typename B>
inline void exec(const ndfb_plan* plan, int op, int norm, const ndview<A>& in, ndview<B>& out, size_t axis, void* stream = nullptr) {
    if (in.ndim() != out.ndim()) throw std::runtime_error("input and output must have the same number of dimensions");
    if (in.device != out.device) throw std::runtime_error("input and output must both be host or both be device arrays");
    int rc = ndfb_exec(plan, op, norm, in.data, out.data, (int)in.ndim(), in.shape.data(), in.strides.data(),
                       out.shape.data(), out.strides.data(), (int)axis, in.device ? NDFB_MEM_DEVICE : NDFB_MEM_HOST, stream);
    if (rc != 0) throw std::runtime_error(ndfb_last_error());
}
// call f on every lane of a HOST view along `axis`
template <typename E, typename F>
inline void for_each_lane(ndview<E>& v, size_t axis, F f) {
    // Normalization::Custom is a HOST callback (src/lib.rs:95-97): it cannot walk device memory
    if (v.device) throw std::runtime_error("Normalization::Custom needs host arrays (the callback runs on the host); use None/Default or host views");
    const size_t n = v.shape[axis];
    size_t lanes = 1;
    for (size_t d = 0; d < v.ndim(); ++d) if (d != axis) lanes *= v.shape[d];
    std::vector<E> tmp(n);
    for (size_t g = 0; g < lanes; ++g) {
        size_t rem = g; ptrdiff_t off = 0;
        for (size_t d = v.ndim(); d-- > 0;) { if (d == axis) continue; off += (ptrdiff_t)(rem % v.shape[d]) * v.strides[d]; rem /= v.shape[d]; }
        for (size_t i = 0; i < n; ++i) tmp[i] = v.data[off + (ptrdiff_t)i * v.strides[axis]];
        f(tmp.data(), n);
        for (size_t i = 0; i < n; ++i) v.data[off + (ptrdiff_t)i * v.strides[axis]] = tmp[i];
    }
}
template <typename E>
inline std::vector<E> copy_dense(const ndview<E>& v, ndview<E>& dense) {
    if (v.device) throw std::runtime_error("Normalization::Custom needs host arrays (the callback runs on the host); use None/Default or host views");
    size_t total = 1;
    for (size_t s : v.shape) total *= s;
    std::vector<E> buf(total);
    dense = ndview<E>::c_order(buf.data(), v.shape);
    for (size_t g = 0; g < total; ++g) {
        size_t rem = g; ptrdiff_t off = 0;
        for (size_t d = v.ndim(); d-- > 0;) { off += (ptrdiff_t)(rem % v.shape[d]) * v.strides[d]; rem /= v.shape[d]; }
        buf[g] = v.data[off];
    }
    return buf;
}
}  // namespace detail

template <typename T>
class FftHandler {  // src/lib.rs:270-348
public:
    explicit FftHandler(size_t n, int device = 0) : n_(n), plan_(detail::make_plan(NDFB_C2C, dtype_of<T>::value, n, device)) {}
    static FftHandler make(size_t n) { return FftHandler(n); }   // FftHandler::new(n)
    FftHandler normalization(Normalization<std::complex<T>> norm) const { FftHandler h(*this); h.norm_ = std::move(norm); return h; }
    size_t n_; std::shared_ptr<ndfb_plan> plan_; Normalization<std::complex<T>> norm_;
};
template <typename T>
class R2cFftHandler {  // src/lib.rs:452-541
public:
    explicit R2cFftHandler(size_t n, int device = 0) : n_(n), m_(n / 2 + 1), plan_(detail::make_plan(NDFB_R2C, dtype_of<T>::value, n, device)) {}
    R2cFftHandler normalization(Normalization<std::complex<T>> norm) const { R2cFftHandler h(*this); h.norm_ = std::move(norm); return h; }
    size_t n_, m_; std::shared_ptr<ndfb_plan> plan_; Normalization<std::complex<T>> norm_;
};
template <typename T>
class DctHandler {  // src/lib.rs:641-751
public:
    explicit DctHandler(size_t n, int device = 0) : n_(n), plan_(detail::make_plan(NDFB_DCT, dtype_of<T>::value, n, device)) {}
    DctHandler normalization(Normalization<T> norm) const { DctHandler h(*this); h.norm_ = std::move(norm); return h; }
    size_t n_; std::shared_ptr<ndfb_plan> plan_; Normalization<T> norm_;
};

template <typename N> inline int norm_code(const N& n) { return n.kind == N::Default ? NDFB_NORM_DEFAULT : NDFB_NORM_NONE; }

template <typename T>
void ndfft(const ndview<std::complex<T>>& in, ndview<std::complex<T>>& out, const FftHandler<T>& h, size_t axis) {  // :350-372
    detail::exec(h.plan_.get(), NDFB_OP_FFT, NDFB_NORM_NONE, in, out, axis);
}
template <typename T>
void ndifft(const ndview<std::complex<T>>& in, ndview<std::complex<T>>& out, const FftHandler<T>& h, size_t axis) {  // :374-397
    detail::exec(h.plan_.get(), NDFB_OP_IFFT, norm_code(h.norm_), in, out, axis);
    if (h.norm_.kind == Normalization<std::complex<T>>::Custom) detail::for_each_lane(out, axis, h.norm_.func);   // :329
}
template <typename T>
void ndfft_r2c(const ndview<T>& in, ndview<std::complex<T>>& out, const R2cFftHandler<T>& h, size_t axis) {  // :543-564
    detail::exec(h.plan_.get(), NDFB_OP_R2C, NDFB_NORM_NONE, in, out, axis);
}
template <typename T>
void ndifft_r2c(const ndview<std::complex<T>>& in, ndview<T>& out, const R2cFftHandler<T>& h, size_t axis) {  // :566-587
    if (h.norm_.kind == Normalization<std::complex<T>>::Custom) {
        ndview<std::complex<T>> dense;
        auto buf = detail::copy_dense(in, dense);                   // the m-long spectrum copy (:509-515)
        detail::for_each_lane(dense, axis, h.norm_.func);
        detail::exec(h.plan_.get(), NDFB_OP_C2R, NDFB_NORM_NONE, dense, out, axis);
    } else {
        detail::exec(h.plan_.get(), NDFB_OP_C2R, norm_code(h.norm_), in, out, axis);
    }
}
template <typename T>
void nddct(int op, const ndview<T>& in, ndview<T>& out, const DctHandler<T>& h, size_t axis) {
    if (h.norm_.kind == Normalization<T>::Custom) {
        ndview<T> dense;
        auto buf = detail::copy_dense(in, dense);                   // the input copy (:691-696)
        detail::for_each_lane(dense, axis, h.norm_.func);
        detail::exec(h.plan_.get(), op, NDFB_NORM_NONE, dense, out, axis);
    } else {
        detail::exec(h.plan_.get(), op, norm_code(h.norm_), in, out, axis);
    }
}
template <typename T> void nddct1(const ndview<T>& i, ndview<T>& o, const DctHandler<T>& h, size_t a) { nddct(NDFB_OP_DCT1, i, o, h, a); }
template <typename T> void nddct2(const ndview<T>& i, ndview<T>& o, const DctHandler<T>& h, size_t a) { nddct(NDFB_OP_DCT2, i, o, h, a); }
template <typename T> void nddct3(const ndview<T>& i, ndview<T>& o, const DctHandler<T>& h, size_t a) { nddct(NDFB_OP_DCT3, i, o, h, a); }
template <typename T> void nddct4(const ndview<T>& i, ndview<T>& o, const DctHandler<T>& h, size_t a) { nddct(NDFB_OP_DCT4, i, o, h, a); }

// ---- multi-axis compositions in one call (ndfb_exec_chain): the fft2 / rfft2 patterns of examples/fft2.rs:23-27, 55-59
// and examples/rfft2.rs:29-33, 48-53, with the `work` array kept on the GPU.  Custom normalisation is a host callback:
// compose the single calls for handlers that carry one. ----
namespace detail {
template <typename A, typename B>
inline void chain(std::initializer_list<ndfb_step> steps, const ndview<A>& in, ndview<B>& out, void* stream = nullptr) {
    if (in.ndim() != out.ndim()) throw std::runtime_error("input and output must have the same number of dimensions");
    if (in.device != out.device) throw std::runtime_error("input and output must both be host or both be device arrays");
    int rc = ndfb_exec_chain(steps.begin(), (int)steps.size(), in.data, out.data, (int)in.ndim(), in.shape.data(), in.strides.data(),
                             out.shape.data(), out.strides.data(), in.device ? NDFB_MEM_DEVICE : NDFB_MEM_HOST, stream);
    if (rc != 0) throw std::runtime_error(ndfb_last_error());
}
template <typename H> inline void no_custom(const H& h) {
    if (h.norm_.kind == decltype(h.norm_)::Custom) throw std::runtime_error("chained transforms take None/Default normalisation; compose the single calls for Custom");
}
}  // namespace detail
template <typename T>
void fft2(const ndview<std::complex<T>>& in, ndview<std::complex<T>>& out, const FftHandler<T>& ax0, const FftHandler<T>& ax1) {
    detail::chain({{ax1.plan_.get(), NDFB_OP_FFT, NDFB_NORM_NONE, 1}, {ax0.plan_.get(), NDFB_OP_FFT, NDFB_NORM_NONE, 0}}, in, out);
}
template <typename T>
void ifft2(const ndview<std::complex<T>>& in, ndview<std::complex<T>>& out, const FftHandler<T>& ax0, const FftHandler<T>& ax1) {
    detail::no_custom(ax0); detail::no_custom(ax1);
    detail::chain({{ax0.plan_.get(), NDFB_OP_IFFT, norm_code(ax0.norm_), 0}, {ax1.plan_.get(), NDFB_OP_IFFT, norm_code(ax1.norm_), 1}}, in, out);
}
template <typename T>
void rfft2(const ndview<T>& in, ndview<std::complex<T>>& out, const FftHandler<T>& ax0, const R2cFftHandler<T>& ax1) {
    detail::chain({{ax1.plan_.get(), NDFB_OP_R2C, NDFB_NORM_NONE, 1}, {ax0.plan_.get(), NDFB_OP_FFT, NDFB_NORM_NONE, 0}}, in, out);
}
template <typename T>
void irfft2(const ndview<std::complex<T>>& in, ndview<T>& out, const FftHandler<T>& ax0, const R2cFftHandler<T>& ax1) {
    detail::no_custom(ax0); detail::no_custom(ax1);
    detail::chain({{ax0.plan_.get(), NDFB_OP_IFFT, norm_code(ax0.norm_), 0}, {ax1.plan_.get(), NDFB_OP_C2R, norm_code(ax1.norm_), 1}}, in, out);
}

// `_par` twins (src/lib.rs:399-421, 589-611, 777-844): on the GPU every call already runs all lanes in parallel.
template <typename... A> void ndfft_par(A&&... a) { ndfft(std::forward<A>(a)...); }
template <typename... A> void ndifft_par(A&&... a) { ndifft(std::forward<A>(a)...); }
template <typename... A> void ndfft_r2c_par(A&&... a) { ndfft_r2c(std::forward<A>(a)...); }
template <typename... A> void ndifft_r2c_par(A&&... a) { ndifft_r2c(std::forward<A>(a)...); }
template <typename... A> void nddct1_par(A&&... a) { nddct1(std::forward<A>(a)...); }
template <typename... A> void nddct2_par(A&&... a) { nddct2(std::forward<A>(a)...); }
template <typename... A> void nddct3_par(A&&... a) { nddct3(std::forward<A>(a)...); }
template <typename... A> void nddct4_par(A&&... a) { nddct4(std::forward<A>(a)...); }

}  // namespace ndrustfft_b200
