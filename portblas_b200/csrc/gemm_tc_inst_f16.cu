// gemm_tc_inst_f16.cu -- one instantiation unit of the tcgen05 GEMM kernel (gemm_tcgen05_kernel.cuh):
// TIn = __half, TOut = __half, fp32 split mode 0; 4 tile configurations x 4 operand-major combinations.
#include "gemm_tcgen05_kernel.cuh"

PBX_TC_INST_DEFINE(pbx_tc_inst_f16, __half, __half, 0)
