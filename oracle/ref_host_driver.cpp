// ref_host_driver.cpp -- C entry points over the REFERENCE's own GEMM, compiled from its sources.  TEST INFRASTRUCTURE ONLY.
//
// Everything below `blas::_gemm` in this translation unit is portBLAS's unmodified code, included from where it lies
// under /root/reference (include/ and src/; nothing is copied into this repository):
//   include/interface/blas3_interface.h:86-123        blas::_gemm / _gemm_batched / _gemm_strided_batched
//   src/interface/gemm_interface.hpp:105-240           _gemm_backend: alpha == 0 -> _scal / _scal_matrix, validation, dispatch
//   src/interface/blas3/backend/{default,nvidia_gpu,intel_gpu}.hpp  tile heuristics (which one: -DNVIDIA_GPU / -DINTEL_GPU /
//                                                      nothing, one library each: see Makefile)
//   src/interface/gemm_launcher.hpp:39-64              views + make_gemm + SB_Handle::execute
//   src/interface/symm_interface.hpp:35-71, trsm_interface.hpp:150-387  the two callers of the path
//   src/sb_handle/portblas_handle.hpp:277-436          nd_range sizing, tall-skinny GemmPartial + Reduction
//   src/sb_handle/kernel_constructor.hpp:187-217       execute_tree (queue.submit / parallel_for)
//   src/operations/blas3/gemm_*.hpp                    the kernels themselves
// The only thing that is not the reference's is the SYCL runtime underneath: oracle/sycl_host/sycl/sycl.hpp, a host
// stand-in that runs every work-group of the nd_range on the CPU (fibers for barriers, OpenMP over groups,
// sycl::mad = fma).  So "reference outputs" here means: the reference's code on that executor.
//
// Used by tests/test_oracle_ref.py (pins oracle/gemm_oracle.c and the CUDA path against it) and by bench.py's
// --impl reference / cpu_baseline leg (kind "reference").  The product never links or loads it.
#include <sycl/sycl.hpp>

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>

#include "blas_meta.h"
#include "container/sycl_iterator.h"
#include "views/view.h"
#include "views/view.hpp"
#include "views/view_sycl.hpp"
#include "sb_handle/portblas_handle.h"
#include "sb_handle/kernel_constructor.hpp"
#include "sb_handle/portblas_handle.hpp"
#include "operations/blas_constants.hpp"
#include "operations/blas1_trees.hpp"
#include "operations/blas3_trees.hpp"
#include "operations/extension/reduction.hpp"
#include "interface/blas1_interface.hpp"
#include "interface/gemm_launcher.hpp"
#include "interface/gemm_interface.hpp"
#include "interface/symm_interface.hpp"
#include "interface/trsm_interface.hpp"
#include "interface/blas3_interface.h"

namespace {

thread_local std::string g_last_error;

blas::SB_Handle& handle() {
  static sycl::queue q;
  static blas::SB_Handle h(q);
  return h;
}

template <typename F>
int guarded(F&& f) {
  try {
    f();
    g_last_error.clear();
    return 0;
  } catch (const std::invalid_argument& e) {  // gemm_interface.hpp:144-165
    g_last_error = e.what();
    return 1;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return 2;
  }
}

}  // namespace

extern "C" {

const char* ref_last_error() { return g_last_error.c_str(); }
int ref_compute_units() { return (int)handle().get_num_compute_units(); }
// which backend header this library was compiled with (src/interface/blas3/backend/backend.hpp:25-33): "default" (the CPU /
// generic device), "nvidia_gpu", or "intel_gpu" (built with GEMM_TALL_SKINNY_SUPPORT: the only route to the reference's
// tall-skinny GemmPartial + Reduction path, intel_gpu.hpp:67-140)
const char* ref_backend() {
#if defined(INTEL_GPU)
  return "intel_gpu";
#elif defined(NVIDIA_GPU)
  return "nvidia_gpu";
#else
  return "default";
#endif
}
// 0: work-items of barrier-free kernels run as plain loop iterations (timing); 1: always fibers (default)
void ref_set_fibers(int on) { sycl::host_standin::kernels_use_barriers = (on != 0); }

// TI: element type of A and B, TO: element type of C, alpha and beta (gemm.cpp.in:33-137: (float,float), (double,double),
// (half,half), (half,float)).  Scalars cross the C boundary as TS (float for the half variants).
#define REF_DEFINE(SUFFIX, TI, TO, TS)                                                                              \
  int ref_gemm_##SUFFIX(char ta, char tb, int m, int n, int k, TS alpha, const TI* A, int lda, const TI* B,        \
                        int ldb, TS beta, TO* C, int ldc) {                                                         \
    return guarded([&] {                                                                                            \
      blas::_gemm(handle(), ta, tb, m, n, k, TO(alpha), A, lda, B, ldb, TO(beta), C, ldc, {});                      \
    });                                                                                                             \
  }                                                                                                                 \
  int ref_gemm_batched_##SUFFIX(char ta, char tb, int m, int n, int k, TS alpha, const TI* A, int lda,             \
                                const TI* B, int ldb, TS beta, TO* C, int ldc, int batch, int batch_type) {         \
    return guarded([&] {                                                                                            \
      blas::_gemm_batched(handle(), ta, tb, m, n, k, TO(alpha), A, lda, B, ldb, TO(beta), C, ldc, batch,            \
                          static_cast<blas::gemm_batch_type_t>(batch_type), {});                                    \
    });                                                                                                             \
  }                                                                                                                 \
  int ref_gemm_strided_batched_##SUFFIX(char ta, char tb, int m, int n, int k, TS alpha, const TI* A, int lda,     \
                                        int stride_a, const TI* B, int ldb, int stride_b, TS beta, TO* C, int ldc,  \
                                        int stride_c, int batch) {                                                  \
    return guarded([&] {                                                                                            \
      blas::_gemm_strided_batched(handle(), ta, tb, m, n, k, TO(alpha), A, lda, stride_a, B, ldb, stride_b,         \
                                  TO(beta), C, ldc, stride_c, batch, {});                                           \
    });                                                                                                             \
  }

// blas::_symm and blas::_trsm (float and double only, as in the reference: symm.cpp.in / trsm.cpp.in)
#define REF_DEFINE_EXT(SUFFIX, T)                                                                                   \
  int ref_symm_##SUFFIX(char side, char uplo, int m, int n, T alpha, const T* A, int lda, const T* B, int ldb,     \
                        T beta, T* C, int ldc) {                                                                    \
    return guarded([&] { blas::_symm(handle(), side, uplo, m, n, alpha, A, lda, B, ldb, beta, C, ldc, {}); });     \
  }                                                                                                                 \
  int ref_trsm_##SUFFIX(char side, char uplo, char trans, char diag, int m, int n, T alpha, const T* A, int lda,   \
                        T* B, int ldb) {                                                                            \
    return guarded([&] { blas::_trsm(handle(), side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb, {}); });      \
  }

REF_DEFINE(f32, float, float, float)
REF_DEFINE(f64, double, double, double)
REF_DEFINE_EXT(f32, float)
REF_DEFINE_EXT(f64, double)
#ifdef BLAS_ENABLE_HALF
REF_DEFINE(f16, sycl::half, sycl::half, float)
REF_DEFINE(f16f32, sycl::half, float, float)
#endif

#ifdef BLAS_ENABLE_COMPLEX
// complex<float> / complex<double> GEMM (gemm.cpp.in complex instantiations; blas3_gemm_test.cpp:143-259).  Buffers are
// interleaved (re, im) pairs, the layout of std::complex and of the reference's device type alike; alpha and beta arrive
// as two reals each.
#define REF_DEFINE_CPLX(SUFFIX, T)                                                                                  \
  int ref_gemm_##SUFFIX(char ta, char tb, int m, int n, int k, T alpha_re, T alpha_im, const void* A, int lda,     \
                        const void* B, int ldb, T beta_re, T beta_im, void* C, int ldc) {                           \
    using cplx = blas::complex_sycl<T>;                                                                            \
    return guarded([&] {                                                                                            \
      blas::_gemm(handle(), ta, tb, m, n, k, cplx(alpha_re, alpha_im), static_cast<const cplx*>(A), lda,            \
                  static_cast<const cplx*>(B), ldb, cplx(beta_re, beta_im), static_cast<cplx*>(C), ldc, {});        \
    });                                                                                                             \
  }                                                                                                                 \
  int ref_gemm_strided_batched_##SUFFIX(char ta, char tb, int m, int n, int k, T alpha_re, T alpha_im,             \
                                        const void* A, int lda, int stride_a, const void* B, int ldb, int stride_b, \
                                        T beta_re, T beta_im, void* C, int ldc, int stride_c, int batch) {          \
    using cplx = blas::complex_sycl<T>;                                                                            \
    return guarded([&] {                                                                                            \
      blas::_gemm_strided_batched(handle(), ta, tb, m, n, k, cplx(alpha_re, alpha_im), static_cast<const cplx*>(A), \
                                  lda, stride_a, static_cast<const cplx*>(B), ldb, stride_b, cplx(beta_re, beta_im), \
                                  static_cast<cplx*>(C), ldc, stride_c, batch, {});                                 \
    });                                                                                                             \
  }
REF_DEFINE_CPLX(c64, float)
REF_DEFINE_CPLX(c128, double)
#endif

}  // extern "C"
