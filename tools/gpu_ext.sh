#!/bin/bash
# GPU pass for the routines built on the GEMM path: parity tests, the C++ caller, device-timed throughput.
# usage: gpurun --timeout 1200 -- 'bash tools/gpu_ext.sh <tag>'
TAG=${1:-r01_ext}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 900 python -m pytest tests/test_blas3_ext_gpu.py tests/test_host_api.py -m gpu -q --tb=short -s > $OUT/pytest_ext.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_ext.log )
tail -60 $OUT/pytest_ext.log
timeout 300 python tools/ext_bench.py --steps 5 --warmup 2 > $OUT/ext_bench.jsonl 2> $OUT/ext_bench.err
cat $OUT/ext_bench.jsonl; tail -5 $OUT/ext_bench.err
