// portblas.h -- umbrella header of the B200 GEMM path (reference include/portblas.h:25-66).
// Only the GEMM-path headers exist in this build (GEMM, and _symm / _trsm / complex GEMM built on it);
// BLAS-1/2 and the extension operators of the reference are out of scope.
#pragma once
#include "blas_meta.h"
#include "container/sycl_iterator.h"
#include "interface/blas3_interface.h"
#include "operations/blas3_trees.h"
#include "portblas_helper.h"
#include "sb_handle/portblas_handle.h"
#include "views/view.h"
