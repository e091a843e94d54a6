// interface/gemm_launcher.h -- blas::Gemm_Launcher<...>::_select_gemm, THE SEAM of the GEMM path
// (reference include/interface/gemm_launcher.h:37-52, src/interface/gemm_launcher.hpp:39-64).
//
// In the reference this is where a backend's choice of kernel becomes a launch: the template arguments carry the
// tile, the memory type, the vectorisation, transposes, beta == 0 and the batch type; _select_gemm builds the views,
// makes the Gemm<> tree and hands it to SB_Handle::execute.  Here everything under the seam is libpbx_gemm.so, so the
// template keeps its full parameter list (callers such as the reference's test/unittest/joint_matrix/launch_gemm.hpp
// name all of it) and uses what still has a meaning:
//   TransA, TransB   -> the transpose characters of pbx_gemm
//   is_beta_zero     -> beta is not read and C is not read (the reference's compile-time specialisation,
//                       src/interface/gemm_interface.hpp:70-100)
//   BatchType        -> strided / interleaved
//   SymmA, SymmB     -> _symm's mirrored operand: not reachable through this seam here (pbx_symm is its own entry point)
//   UseJointMatrix   -> the reference's tensor-core kernels with reduced-precision fragments; this library's default
//                       fp32 path (3xTF32) is at least as accurate on the same inputs, so the flag selects nothing
//   WgSize, DoubleBuffer, ConflictA/B, ClSize, TileT, GemmMemoryType, GemmAlgorithm, GemmVectorization, VectorSize
//                    -> SYCL kernel shape; tile and pipeline are chosen inside the library (gemm_tcgen05.cu:make_plan)
#pragma once
#include "interface/blas3_interface.h"
#include "operations/blas3_trees.h"
#include "sb_handle/portblas_handle.h"

namespace blas {

template <typename container_0_t, typename container_1_t, typename container_2_t, int WgSize, bool DoubleBuffer,
          bool ConflictA, bool ConflictB, int ClSize, typename TileT, bool TransA, bool TransB, bool SymmA, bool SymmB,
          int GemmMemoryType, int GemmAlgorithm, int GemmVectorization, bool is_beta_zero, int VectorSize, int BatchType,
          bool UseJointMatrix = false>
struct Gemm_Launcher {
  template <typename sb_handle_t, typename element_t, typename index_t>
  static typename sb_handle_t::event_t _select_gemm(sb_handle_t& sb_handle, index_t _M, index_t _N, index_t _K,
                                                    element_t _alpha, container_0_t a_, index_t _lda, index_t _stridea,
                                                    container_1_t b_, index_t _ldb, index_t _strideb, element_t _beta,
                                                    container_2_t _C, index_t _ldc, index_t _stridec, index_t batch_size,
                                                    const typename sb_handle_t::event_t& _dependencies = {}) {
    static_assert(!SymmA && !SymmB, "symmetric operands go through blas::_symm (pbx_symm) in this build");
    return internal::_gemm_backend(sb_handle, TransA ? 't' : 'n', TransB ? 't' : 'n', _M, _N, _K, _alpha, a_, _lda,
                                   _stridea, b_, _ldb, _strideb, is_beta_zero ? element_t(0) : _beta, _C, _ldc, _stridec,
                                   batch_size, static_cast<gemm_batch_type_t>(BatchType), _dependencies);
  }
};

}  // namespace blas
