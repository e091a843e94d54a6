// ext/oneapi/experimental/sycl_complex.hpp -- host stand-in for the oneAPI complex extension.  TEST INFRASTRUCTURE ONLY.
// portBLAS's complex GEMM (BLAS_ENABLE_COMPLEX, include/blas_meta.h:32-38,206-222) uses
// sycl::ext::oneapi::experimental::complex<T> as its device-side complex type: a distinct class template (not std::complex)
// with real()/imag() and the four arithmetic operators written out in real arithmetic.  See sycl/sycl.hpp here for why
// this directory exists.
#pragma once
#include <complex>

namespace sycl {
namespace ext {
namespace oneapi {
namespace experimental {

template <typename T>
class complex {
 public:
  using value_type = T;
  constexpr complex(T re = T(), T im = T()) : re_(re), im_(im) {}
  template <typename U>
  constexpr complex(const complex<U>& o) : re_(static_cast<T>(o.real())), im_(static_cast<T>(o.imag())) {}
  constexpr complex(const std::complex<T>& o) : re_(o.real()), im_(o.imag()) {}
  constexpr operator std::complex<T>() const { return std::complex<T>(re_, im_); }
  constexpr T real() const { return re_; }
  constexpr T imag() const { return im_; }
  void real(T v) { re_ = v; }
  void imag(T v) { im_ = v; }
  complex& operator+=(const complex& o) { re_ += o.re_; im_ += o.im_; return *this; }
  complex& operator-=(const complex& o) { re_ -= o.re_; im_ -= o.im_; return *this; }
  complex& operator*=(const complex& o) { return *this = *this * o; }
  complex& operator/=(const complex& o) { return *this = *this / o; }
  friend constexpr complex operator+(const complex& a, const complex& b) { return complex(a.re_ + b.re_, a.im_ + b.im_); }
  friend constexpr complex operator-(const complex& a, const complex& b) { return complex(a.re_ - b.re_, a.im_ - b.im_); }
  friend constexpr complex operator-(const complex& a) { return complex(-a.re_, -a.im_); }
  // (a + bi)(c + di) = (ac - bd) + (ad + bc)i, the textbook form in the element's own precision
  friend constexpr complex operator*(const complex& a, const complex& b) {
    return complex(a.re_ * b.re_ - a.im_ * b.im_, a.re_ * b.im_ + a.im_ * b.re_);
  }
  friend constexpr complex operator/(const complex& a, const complex& b) {
    const T d = b.re_ * b.re_ + b.im_ * b.im_;
    return complex((a.re_ * b.re_ + a.im_ * b.im_) / d, (a.im_ * b.re_ - a.re_ * b.im_) / d);
  }
  friend constexpr bool operator==(const complex& a, const complex& b) { return a.re_ == b.re_ && a.im_ == b.im_; }
  friend constexpr bool operator!=(const complex& a, const complex& b) { return !(a == b); }

 private:
  T re_, im_;
};

template <typename T> constexpr T real(const complex<T>& z) { return z.real(); }
template <typename T> constexpr T imag(const complex<T>& z) { return z.imag(); }
template <typename T> constexpr complex<T> conj(const complex<T>& z) { return complex<T>(z.real(), -z.imag()); }

}  // namespace experimental
}  // namespace oneapi
}  // namespace ext
}  // namespace sycl
