"""portblas_b200 -- B200-native implementation of portBLAS's GEMM path.

The product is ``libpbx_gemm.so`` (C-ABI in ``include/pbx_gemm.h``; CUDA sources under
``portblas_b200/csrc``).  This package holds the ctypes binding (``_lib``), the Python mirror of the
reference interface (``blas``) and the multi-GPU partitioning helpers (``sharding``).
"""
from .blas import (SB_Handle, gemm_batch_type_t, _gemm, _gemm_batched, _gemm_strided_batched, gemm_host,  # noqa: F401
                   _symm, _trsm, PbxError, SB_Handle_Group)
