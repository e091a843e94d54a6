// operations/blas3_trees.h -- the public enum of the GEMM path.
// The reference's expression-tree types (Tile<>, Gemm<>, GemmPartial<>, make_gemm;
// reference include/operations/blas3_trees.h:113-278) describe SYCL kernels and have no
// counterpart here: tile shapes are chosen inside libpbx_gemm.so.  What callers see is kept.
#pragma once

namespace blas {

// reference include/operations/blas3_trees.h:59 (values are part of the C-ABI: pbx_gemm batch_type)
enum class gemm_batch_type_t : int { strided = 0, interleaved = 1 };

// kept for source compatibility with code that names them (reference :38,45,52)
enum class gemm_memory_t : int { local = 0, no_local = 1 };
enum class gemm_algorithm_t : int { naive = 0, standard = 1, tall_skinny = 2 };
enum class gemm_vectorization_t : int { none = 0, partial = 1, full = 2 };

}  // namespace blas
