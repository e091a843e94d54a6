// gemm_tc_inst_bf16.cu -- one instantiation unit of the tcgen05 GEMM kernel (gemm_tcgen05_kernel.cuh):
// TIn = __nv_bfloat16, TOut = __nv_bfloat16, fp32 split mode 0; 4 tile configurations x 4 operand-major combinations.
#include "gemm_tcgen05_kernel.cuh"

PBX_TC_INST_DEFINE(pbx_tc_inst_bf16, __nv_bfloat16, __nv_bfloat16, 0)
