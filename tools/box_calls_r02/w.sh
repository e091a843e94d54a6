#!/bin/bash
# Round-2: skinny-M (operand swap) rows with one round of K slices; graph-timed like the sweep.
set -o pipefail
O=gpurun_out/r02w; mkdir -p $O
python -m portblas_b200.build > /dev/null || { echo "BUILD BROKEN"; exit 9; }
timeout 600 python tools/csv_sweep.py --dtype f32 --graph --api gemm > $O/sweep_f32_gemm.jsonl 2> $O/sweep.err; echo "sweep rc=$?"; tail -1 $O/sweep_f32_gemm.jsonl
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -k "split_k or skinny or tall" > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.txt
