// gemm_tc_inst_f32_pre3.cu -- one instantiation unit of the tcgen05 GEMM kernel (gemm_tcgen05_kernel.cuh):
// TIn = float, TOut = float, fp32 split mode 3; 4 tile configurations x 4 operand-major combinations.
#include "gemm_tcgen05_kernel.cuh"

PBX_TC_INST_DEFINE(pbx_tc_inst_f32_pre3, float, float, 3)
