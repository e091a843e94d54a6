// Compile-only check of the drop-in C++ surface: every entry point of the GEMM path is instantiated with the container
// kinds and element types the reference's explicit instantiations cover (src/interface/blas3/gemm.cpp.in:33-137:
// BufferIterator<T>, BufferIterator<T const> inputs, T*, const T* x float / double / half / half->float [/ complex])
// and must return sb_handle_t::event_t with the reference's argument order and defaults
// (include/interface/blas3_interface.h:86-147).  Nothing is executed; the file only has to compile and link.
#include "portblas.hpp"
#include <sycl/sycl.hpp>

#include <complex>
#include <type_traits>

namespace {

using blas::BufferIterator;
using blas::SB_Handle;
using event_t = SB_Handle::event_t;

template <typename in_a_t, typename in_b_t, typename out_t, typename scalar_t>
void gemm_family(SB_Handle& sb, in_a_t a, in_b_t b, out_t c, scalar_t alpha, scalar_t beta) {
  const int m = 8, n = 8, k = 8, ld = 8, batch = 2;
  static_assert(std::is_same_v<decltype(blas::_gemm(sb, 'n', 't', m, n, k, alpha, a, ld, b, ld, beta, c, ld)), event_t>);
  static_assert(std::is_same_v<decltype(blas::_gemm(sb, 'n', 't', m, n, k, alpha, a, ld, b, ld, beta, c, ld, event_t{})),
                               event_t>);
  static_assert(std::is_same_v<decltype(blas::_gemm_batched(sb, 'n', 'n', m, n, k, alpha, a, ld, b, ld, beta, c, ld, batch)),
                               event_t>);
  static_assert(std::is_same_v<decltype(blas::_gemm_batched(sb, 'n', 'n', m, n, k, alpha, a, ld, b, ld, beta, c, ld, batch,
                                                            blas::gemm_batch_type_t::interleaved, event_t{})),
                               event_t>);
  static_assert(std::is_same_v<decltype(blas::_gemm_strided_batched(sb, 't', 'n', m, n, k, alpha, a, ld, 64, b, ld, 64, beta,
                                                                    c, ld, 64, batch)),
                               event_t>);
  if (false) {   // instantiate the bodies as well
    blas::_gemm(sb, 'n', 't', m, n, k, alpha, a, ld, b, ld, beta, c, ld);
    blas::_gemm_batched(sb, 'n', 'n', m, n, k, alpha, a, ld, b, ld, beta, c, ld, batch);
    blas::_gemm_strided_batched(sb, 't', 'n', m, n, k, alpha, a, ld, 64, b, ld, 64, beta, c, ld, 64, batch, event_t{});
  }
}

template <typename T>
void real_type(SB_Handle& sb) {
  BufferIterator<T> bi;
  BufferIterator<T const> bci;
  T* p = nullptr;
  const T* cp = nullptr;
  const T one = T(1.0f);
  gemm_family(sb, bi, bi, bi, one, one);
  gemm_family(sb, bci, bci, bi, one, one);     // BLAS_ENABLE_CONST_INPUT
  gemm_family(sb, p, p, p, one, one);
  gemm_family(sb, cp, cp, p, one, one);
  gemm_family(sb, bi, p, p, one, one);         // mixed containers (samples / benchmarks do this)
}

template <typename T>
void symm_trsm(SB_Handle& sb) {
  BufferIterator<T> bi;
  T* p = nullptr;
  const T* cp = nullptr;
  const T one = T(1);
  static_assert(std::is_same_v<decltype(blas::_symm(sb, 'l', 'u', 8, 8, one, bi, 8, bi, 8, one, bi, 8)), event_t>);
  static_assert(std::is_same_v<decltype(blas::_trsm(sb, 'l', 'u', 'n', 'n', 8, 8, one, cp, 8, p, 8, event_t{})), event_t>);
  if (false) {
    blas::_symm(sb, 'r', 'l', 8, 8, one, cp, 8, cp, 8, one, p, 8);
    blas::_symm(sb, 'l', 'u', 8, 8, one, bi, 8, bi, 8, one, bi, 8, event_t{});
    blas::_trsm(sb, 'r', 'l', 't', 'u', 8, 8, one, bi, 8, bi, 8);
    blas::_trsm(sb, 'l', 'u', 'n', 'n', 8, 8, one, cp, 8, p, 8);
  }
}

}  // namespace

int main() {
  // types only: the handle is never constructed when the program runs without arguments
  static_assert(std::is_same_v<event_t, std::vector<sycl::event>>);
  static_assert(static_cast<int>(blas::gemm_batch_type_t::strided) == 0 &&
                static_cast<int>(blas::gemm_batch_type_t::interleaved) == 1);
  if (false) {
    sycl::queue q;
    SB_Handle sb(q);
    blas::Temp_Mem_Pool pool(q);
    SB_Handle sb2(&pool);                       // benchmark/portblas/main.cpp:68-72 (BLAS_MEMPOOL_BENCHMARK)
    real_type<float>(sb);
    real_type<double>(sb);
    real_type<sycl::half>(sb);
    {                                            // half in, float out (gemm.cpp.in:36-39)
      BufferIterator<sycl::half> hi;
      BufferIterator<float> fo;
      gemm_family(sb, hi, hi, fo, 1.0f, 1.0f);
      sycl::half* hp = nullptr;
      float* fp = nullptr;
      gemm_family(sb, hp, hp, fp, 1.0f, 1.0f);
    }
    {                                            // BLAS_ENABLE_COMPLEX
      std::complex<float>* cp = nullptr;
      std::complex<double>* zp = nullptr;
      gemm_family(sb, cp, cp, cp, std::complex<float>(1, 0), std::complex<float>(0, 1));
      gemm_family(sb, zp, zp, zp, std::complex<double>(1, 0), std::complex<double>(0, 1));
    }
    symm_trsm<float>(sb);
    symm_trsm<double>(sb);
    (void)sb.get_num_compute_units();
    (void)sb.get_work_group_size();
    (void)sb.has_local_memory();
    sb.wait();
  }
  return 0;
}
