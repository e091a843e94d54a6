"""CPU oracle for the GEMM path -- test infrastructure only (see gemm_oracle.c)."""
