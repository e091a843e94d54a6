// container/sycl_iterator.h -- blas::BufferIterator<T> backed by a CUDA device allocation.
// Same surface as reference include/container/sycl_iterator.h:33-182,227-272 (value semantics,
// pointer arithmetic on an element offset, converting ctor to BufferIterator<T const>,
// make_sycl_iterator_buffer factories); the accessor machinery is replaced by a raw device
// pointer (get_device_ptr) because kernels are launched through the C-ABI, not SYCL handlers.
#pragma once
#include <sycl/sycl.hpp>

#include <cstddef>
#include <type_traits>
#include <vector>

#include "../blas_meta.h"

namespace blas {

template <typename element_t>
class BufferIterator {
 public:
  using scalar_t = element_t;
  using self_t = BufferIterator<scalar_t>;
  using buff_t = sycl::buffer<scalar_t, 1>;

  BufferIterator() = default;
  BufferIterator(const buff_t& buff, std::ptrdiff_t offset) : offset_(offset), buffer_(buff) {}
  explicit BufferIterator(const buff_t& buff) : offset_(0), buffer_(buff) {}
  // BufferIterator<T> -> BufferIterator<T const>
  template <typename other_t, typename = std::enable_if_t<std::is_const_v<scalar_t> &&
                                                         std::is_same_v<std::remove_const_t<scalar_t>, other_t>>>
  BufferIterator(const BufferIterator<other_t>& other)
      : offset_(other.get_offset()), buffer_(other.get_buffer().template reinterpret<scalar_t, 1>()) {}

  std::ptrdiff_t get_size() const { return static_cast<std::ptrdiff_t>(buffer_.size()) - offset_; }
  std::ptrdiff_t get_offset() const { return offset_; }
  void set_offset(std::ptrdiff_t offset) { offset_ = offset; }
  buff_t get_buffer() const { return buffer_; }
  // device address of the element this iterator designates
  scalar_t* get_device_ptr() const { return buffer_.device_ptr() + offset_; }

  self_t& operator+=(std::ptrdiff_t n) { offset_ += n; return *this; }
  self_t& operator-=(std::ptrdiff_t n) { offset_ -= n; return *this; }
  self_t operator+(std::ptrdiff_t n) const { return self_t(buffer_, offset_ + n); }
  self_t operator-(std::ptrdiff_t n) const { return self_t(buffer_, offset_ - n); }
  self_t& operator++() { ++offset_; return *this; }
  self_t operator++(int) { self_t t(*this); ++offset_; return t; }
  self_t& operator--() { --offset_; return *this; }
  // host code cannot dereference device memory (reference :175-177)
  scalar_t& operator*() = delete;
  scalar_t* operator->() = delete;

 private:
  std::ptrdiff_t offset_ = 0;
  buff_t buffer_;
};

template <typename T> struct ValueType<BufferIterator<T>> { using type = std::remove_cv_t<T>; };
template <typename T, typename U> struct RebindType<BufferIterator<T>, U> { using type = BufferIterator<U>; };

// host pointer: device copy initialised from the host data, written back when the last copy dies
template <typename scalar_t>
inline BufferIterator<scalar_t> make_sycl_iterator_buffer(scalar_t* data, size_t size) {
  return BufferIterator<scalar_t>(sycl::buffer<scalar_t, 1>(data, sycl::range<1>(size)));
}
template <typename scalar_t>
inline BufferIterator<scalar_t> make_sycl_iterator_buffer(std::vector<scalar_t>& data, size_t size) {
  return BufferIterator<scalar_t>(sycl::buffer<scalar_t, 1>(data.data(), sycl::range<1>(size)));
}
// uninitialised device storage
template <typename scalar_t, typename index_t,
          typename = std::enable_if_t<std::is_integral_v<index_t>>>
inline BufferIterator<scalar_t> make_sycl_iterator_buffer(index_t size) {
  return BufferIterator<scalar_t>(sycl::buffer<scalar_t, 1>(sycl::range<1>(static_cast<size_t>(size))));
}
template <typename scalar_t>
inline BufferIterator<scalar_t> make_sycl_iterator_buffer(sycl::buffer<scalar_t, 1> buff) {
  return BufferIterator<scalar_t>(buff);
}

// raw device pointer of any container kind
template <typename T> inline T* get_device_ptr(T* p) { return p; }
template <typename T> inline T* get_device_ptr(const BufferIterator<T>& it) { return it.get_device_ptr(); }

}  // namespace blas
