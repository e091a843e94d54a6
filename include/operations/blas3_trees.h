// operations/blas3_trees.h -- the public enum of the GEMM path.
// The reference's expression-tree types (Gemm<>, GemmPartial<>, make_gemm;
// reference include/operations/blas3_trees.h:140-278) describe SYCL kernels and have no
// counterpart here: tile shapes are chosen inside libpbx_gemm.so.  What callers see is kept.
#pragma once

namespace blas {

// reference include/operations/blas3_trees.h:59 (values are part of the C-ABI: pbx_gemm batch_type)
enum class gemm_batch_type_t : int { strided = 0, interleaved = 1 };

// kept for source compatibility with code that names them (reference :38,45,52)
enum class gemm_memory_t : int { local = 0, no_local = 1 };
enum class gemm_algorithm_t : int { naive = 0, standard = 1, tall_skinny = 2 };
enum class gemm_vectorization_t : int { none = 0, partial = 1, full = 2 };

// Tile<> (reference :113-138): the work-item / work-group / sub-group / joint_matrix shape of a SYCL GEMM kernel.  Kept
// as a type so that code naming a tile -- Gemm_Launcher's callers -- compiles; nothing reads it (interface/gemm_launcher.h).
template <int ItemRows = 8, int ItemCols = 8, int WgRows = 16, int WgCols = 16, int SgRows = 1, int SgCols = 1,
          int TlRows = 1, int TlCols = 1, int ItemBatchs = 1, int WgBatchs = 1, int jm_M = 1, int jm_N = 1, int jm_K = 1,
          typename inp_jmT = float, typename out_jmT = float>
struct Tile {
  static constexpr int item_rows = ItemRows, item_cols = ItemCols, item_batchs = ItemBatchs;
  static constexpr int wg_rows = WgRows, wg_cols = WgCols, wg_batchs = WgBatchs;
  static constexpr int sg_rows = SgRows, sg_cols = SgCols, tl_rows = TlRows, tl_cols = TlCols;
  static constexpr int joint_matrix_M = jm_M, joint_matrix_N = jm_N, joint_matrix_K = jm_K;
  using jmInpType = inp_jmT;
  using jmOutType = out_jmT;
};

}  // namespace blas
