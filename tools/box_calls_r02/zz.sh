#!/bin/bash
# Round-2: the reference executables the time-limited last-tree suite run did not reach (gemm_batched test, joint_matrix, benches).
set -o pipefail
O=gpurun_out/r02zz; mkdir -p $O
timeout 170 python -m pytest tests/test_zz_reference_unittests_gpu.py -m gpu -q -n 4 -k "joint_matrix or benchmark_harness or gemm_batched_test" > $O/pytest_ref.txt 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_ref.txt
