// gemm_tc_inst_f32_pre4.cu -- one instantiation unit of the tcgen05 GEMM kernel (gemm_tcgen05_kernel.cuh):
// TIn = float, TOut = float, fp32 split mode 4 (tf32 + 2 x bf16, bf16 tiles made in the kernel); 4 tile configurations x
// 4 operand-major combinations.
#include "gemm_tcgen05_kernel.cuh"

PBX_TC_INST_DEFINE(pbx_tc_inst_f32_pre4, float, float, 4)
