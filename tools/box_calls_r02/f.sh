#!/bin/bash
# Round-2 sixth box call: decoupled raw / derived rings of the in-kernel split modes, cost-model planner, fp32 sweep.
set -o pipefail
O=gpurun_out/r02f; mkdir -p $O
python -m portblas_b200.build > /dev/null || { echo "BUILD BROKEN"; exit 9; }
timeout 300 python -m pytest tests/test_split16_gpu.py -q -x > $O/pytest_split.txt 2>&1; echo "split rc=$?"; tail -5 $O/pytest_split.txt
if ! grep -q " passed" $O/pytest_split.txt || grep -q "failed\|error" $O/pytest_split.txt; then tail -60 $O/pytest_split.txt; exit 1; fi
timeout 300 python tools/ab_variants.py --workload sgemm_splitk --variants default,split16_off,cublas --burst-steps 5 --rounds 3 > $O/ab_splitk.jsonl 2> $O/ab_splitk.err; cat $O/ab_splitk.jsonl
timeout 300 python tools/ab_variants.py --workload sgemm1024 --variants default,split16_off,cublas --burst-steps 200 --rounds 3 > $O/ab_sgemm1024.jsonl 2> $O/ab_1024.err; cat $O/ab_sgemm1024.jsonl
timeout 900 python tools/csv_sweep.py --dtype f32 --graph > $O/sweep_f32.jsonl 2> $O/sweep_f32.err; echo "sweep rc=$?"; tail -1 $O/sweep_f32.jsonl
PBX_PLAN_MODEL=0 timeout 900 python tools/csv_sweep.py --dtype f32 --graph > $O/sweep_f32_oldplan.jsonl 2> $O/sweep_f32_oldplan.err; echo "sweep old rc=$?"; tail -1 $O/sweep_f32_oldplan.jsonl
timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_beta_zero_nan_gpu.py tests/test_nonfinite_gpu.py tests/test_ref_parity_gpu.py tests/test_full_size_gpu.py -q > $O/pytest_gemm.txt 2>&1; echo "gemm rc=$?"; tail -12 $O/pytest_gemm.txt
