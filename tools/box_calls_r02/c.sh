#!/bin/bash
# Round-2 third box call: A/B of the launch features (burst + sustained) with cuBLAS beside them, ncu of the new
# fp32 kernel and of cfg4, the rest of the GPU suite.
O=gpurun_out/r02c; mkdir -p $O
timeout 300 python tools/ab_variants.py --workload bf16gemm_batched --variants default,static,nopdl,static_nopdl,cg2_256,static_cg2_256,cublas > $O/ab_cfg4.jsonl 2> $O/ab_cfg4.err; cat $O/ab_cfg4.jsonl
timeout 300 python tools/ab_variants.py --workload sgemm8192 --variants default,static,cg2_128,split16_off,cublas --burst-steps 5 --rounds 5 > $O/ab_sgemm8192.jsonl 2> $O/ab_sgemm8192.err; cat $O/ab_sgemm8192.jsonl
timeout 300 python tools/ab_variants.py --workload bf16gemm8192 --variants default,static,cublas --burst-steps 10 --rounds 5 > $O/ab_bf16gemm8192.jsonl 2> $O/ab_bf16.err; cat $O/ab_bf16gemm8192.jsonl
timeout 300 python tools/ab_variants.py --workload sgemm1024 --variants default,static,nopdl,cublas --burst-steps 200 --rounds 5 > $O/ab_sgemm1024.jsonl 2> $O/ab_1024.err; cat $O/ab_sgemm1024.jsonl
timeout 300 python tools/ab_variants.py --workload sgemm_splitk --variants default,static,cublas --burst-steps 5 --rounds 5 > $O/ab_splitk.jsonl 2> $O/ab_splitk.err; cat $O/ab_splitk.jsonl
bash tools/gpu_ncu.sh r02c sgemm8192 bf16gemm_batched
python tools/ncu_summary.py $O/prof_sgemm8192_raw.csv $O/prof_bf16gemm_batched_raw.csv
timeout 1500 python -m pytest tests -m gpu -q -k "not reference_unit_tests and not joint_matrix and not beta_zero" > $O/pytest_gpu.txt 2>&1; echo "gpu rc=$?"; tail -15 $O/pytest_gpu.txt
