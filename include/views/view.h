// views/view.h -- MatrixView as a POD {ptr, rows, cols, ld}.
// The reference's views (include/views/view.h:139-279,343-392) are device-side accessors for the
// SYCL kernels; here only make_matrix_view<col_major> and the getters survive, as the host-side
// description of an operand that the launcher hands to the C-ABI
// (reference src/interface/gemm_launcher.hpp:54-56 builds (a_,M,K,lda), (b_,K,N,ldb), (C,M,N,ldc)).
#pragma once
#include "../container/sycl_iterator.h"

namespace blas {

struct col_major {};
struct row_major {};

template <typename element_t, typename index_t>
struct MatrixView {
  using value_t = element_t;
  element_t* data_;
  index_t sizeR_, sizeC_, sizeL_, inc_;
  element_t* get_data() const { return data_; }
  element_t* get_pointer() const { return data_; }
  index_t get_size() const { return sizeR_ * sizeC_; }
  index_t get_size_row() const { return sizeR_; }
  index_t get_size_col() const { return sizeC_; }
  index_t getSizeL() const { return sizeL_; }
};

template <typename layout = col_major, typename container_t, typename index_t>
inline auto make_matrix_view(container_t buff, index_t m, index_t n, index_t lda, index_t inc = 1) {
  static_assert(std::is_same_v<layout, col_major>, "only column-major is supported (reference README.md:282)");
  auto* p = get_device_ptr(buff);
  using elem_t = std::remove_pointer_t<decltype(p)>;
  return MatrixView<elem_t, index_t>{p, m, n, lda, inc};
}

}  // namespace blas
