// blas_meta.h -- host-side traits used by the GEMM interface (B200 build).
// Mirrors the subset of reference include/blas_meta.h:113-121,166-204,225-235 that the GEMM
// callers rely on: is_half, is_sycl_scalar, ValueType/RebindType hooks, concatenate_vectors /
// append_vector for event lists, unsupported_exception, PORTBLAS_INLINE.
#pragma once
#include <sycl/sycl.hpp>

#include <complex>
#include <stdexcept>
#include <type_traits>
#include <vector>

#ifndef PORTBLAS_INLINE
#define PORTBLAS_INLINE inline
#endif
#ifndef PORTBLAS_ALWAYS_INLINE
#define PORTBLAS_ALWAYS_INLINE inline
#endif

namespace blas {

template <typename T> struct is_half : std::is_same<std::remove_cv_t<T>, sycl::half> {};
template <typename T>
struct is_sycl_scalar : std::integral_constant<bool, std::is_floating_point_v<T> || is_half<T>::value> {};

// element type carried by a container (T* / const T* / BufferIterator<T>)
template <typename container_t> struct ValueType { using type = std::remove_cv_t<container_t>; };
template <typename T> struct ValueType<T*> { using type = std::remove_cv_t<T>; };
template <typename T> struct ValueType<const T*> { using type = std::remove_cv_t<T>; };

template <typename T, typename U> struct RebindType { using type = U; };
template <typename T, typename U> struct RebindType<T*, U> { using type = U*; };

// complex element types (reference include/blas_meta.h:205-225, BLAS_ENABLE_COMPLEX builds)
template <typename T> using complex_sycl = typename sycl::ext::oneapi::experimental::complex<T>;
template <class type>
struct is_complex_sycl : std::integral_constant<bool, std::is_same_v<type, complex_sycl<double>> ||
                                                          std::is_same_v<type, complex_sycl<float>>> {};
template <class type>
struct is_complex_std : std::integral_constant<bool, std::is_same_v<type, std::complex<double>> ||
                                                         std::is_same_v<type, std::complex<float>>> {};

struct unsupported_exception : public std::runtime_error {
  unsupported_exception(const char* msg = "Unsupported operation") : std::runtime_error(msg) {}
};

// event-list helpers (sb_handle_t::event_t is std::vector<sycl::event>)
template <typename T>
inline std::vector<T> concatenate_vectors(const std::vector<T>& a, const std::vector<T>& b) {
  std::vector<T> r(a);
  r.insert(r.end(), b.begin(), b.end());
  return r;
}
template <typename T, typename... Ts>
inline std::vector<T> concatenate_vectors(const std::vector<T>& a, const std::vector<T>& b, const Ts&... rest) {
  return concatenate_vectors(concatenate_vectors(a, b), rest...);
}
template <typename T> inline void append_vector(std::vector<T>& dst, const std::vector<T>& src) {
  dst.insert(dst.end(), src.begin(), src.end());
}

}  // namespace blas
