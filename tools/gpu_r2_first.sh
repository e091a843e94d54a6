#!/bin/bash
# Round-2 first box call: the reference's verifying benchmark binaries (logs kept), the beta==0 NaN tests, the split16 experiment.
O=gpurun_out/r02a; mkdir -p $O
timeout 900 python -m pytest tests/test_zz_reference_unittests_gpu.py -q -k "benchmark_harness" -rxXs > $O/pytest_refbench.txt 2>&1
tail -15 $O/pytest_refbench.txt
timeout 600 python -m pytest tests/test_beta_zero_nan_gpu.py -q -x > $O/pytest_nan.txt 2>&1
tail -30 $O/pytest_nan.txt
bash tools/gpu_split16.sh
