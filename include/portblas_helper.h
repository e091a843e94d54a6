// portblas_helper.h -- blas::helper:: allocation / copy helpers on CUDA device memory.
// Same names and argument meaning as reference include/portblas_helper.h:38-81,114-131,139-219.
#pragma once
#include <sycl/sycl.hpp>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "container/sycl_iterator.h"

#ifndef SB_ENABLE_USM
#define SB_ENABLE_USM 1
#endif

namespace blas {
namespace helper {

enum class AllocType : int { usm = 0, buffer = 1 };

template <typename value_t, AllocType mem_alloc> struct AllocHelper;
template <typename value_t> struct AllocHelper<value_t, AllocType::usm> { using type = value_t*; };
template <typename value_t> struct AllocHelper<value_t, AllocType::buffer> { using type = blas::BufferIterator<value_t>; };

template <AllocType alloc, typename value_t>
inline std::enable_if_t<alloc == AllocType::usm, value_t*> allocate(int size, sycl::queue q) {
  return sycl::malloc_device<value_t>(static_cast<size_t>(size), q);
}
template <AllocType alloc, typename value_t>
inline std::enable_if_t<alloc == AllocType::buffer, blas::BufferIterator<value_t>> allocate(int size, sycl::queue) {
  return make_sycl_iterator_buffer<value_t>(size);
}
template <AllocType alloc, typename container_t>
inline std::enable_if_t<alloc == AllocType::usm> deallocate(container_t mem, sycl::queue q) {
  if (mem != nullptr) sycl::free(reinterpret_cast<void*>(const_cast<std::remove_const_t<std::remove_pointer_t<container_t>>*>(mem)), q);
}
template <AllocType alloc, typename container_t>
inline std::enable_if_t<alloc == AllocType::buffer> deallocate(container_t, sycl::queue) {}

template <typename container_t,
          AllocType alloc = std::is_pointer<container_t>::value ? AllocType::usm : AllocType::buffer>
using add_const = std::conditional_t<alloc == AllocType::usm,
                                     std::add_pointer_t<std::add_const_t<std::remove_pointer_t<container_t>>>, container_t>;

inline bool has_local_memory(sycl::queue&) { return true; }
inline size_t get_work_group_size(sycl::queue&) { return 256; }
inline size_t get_num_compute_units(sycl::queue& q) { return static_cast<size_t>(pbx_get_num_compute_units(q.pbx())); }

// host -> device
template <typename element_t>
inline sycl::event copy_to_device(sycl::queue q, const element_t* src, BufferIterator<element_t> dst, size_t size,
                                  const std::vector<sycl::event>& deps = {}) {
  return q.memcpy(dst.get_device_ptr(), src, size * sizeof(element_t), deps);
}
template <typename element_t>
inline sycl::event copy_to_device(sycl::queue q, const element_t* src, element_t* dst, size_t size,
                                  const std::vector<sycl::event>& deps = {}) {
  return q.memcpy(dst, src, size * sizeof(element_t), deps);
}
// device -> host
template <typename element_t>
inline sycl::event copy_to_host(sycl::queue q, BufferIterator<element_t> src, element_t* dst, size_t size) {
  return q.memcpy(dst, src.get_device_ptr(), size * sizeof(element_t));
}
template <typename element_t>
inline sycl::event copy_to_host(sycl::queue q, element_t* src, element_t* dst, size_t size) {
  return q.memcpy(dst, src, size * sizeof(element_t));
}
template <typename element_t>
inline sycl::event copy_to_host(sycl::queue q, const element_t* src, element_t* dst, size_t size) {
  return q.memcpy(dst, src, size * sizeof(element_t));
}
template <typename element_t>
inline sycl::event fill(sycl::queue q, BufferIterator<element_t> buff, element_t value, size_t size,
                        const std::vector<sycl::event>& deps) {
  return q.fill(buff.get_device_ptr(), value, size, deps);
}
template <typename element_t>
inline sycl::event fill(sycl::queue q, element_t* buff, element_t value, size_t size,
                        const std::vector<sycl::event>& deps) {
  return q.fill(buff, value, size, deps);
}
template <typename sb_handle_t, typename containerT>
inline bool is_malloc_shared(sb_handle_t&, const containerT) { return false; }  // device USM only

}  // namespace helper
}  // namespace blas
