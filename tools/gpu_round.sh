#!/bin/bash
# One GPU-box pass: parity tests, bench lines for each BASELINE config, ncu launch list.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log )
tail -5 $OUT/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 > $OUT/bench_dgemm8192.json 2> $OUT/bench_dgemm8192.err
cat $OUT/bench_dgemm8192.json
for wl in sgemm8192 sgemm16384 bf16gemm8192 bf16gemm_batched hgemm_batched sgemm_splitk sgemm1024; do
  timeout 300 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err
  cat $OUT/bench_$wl.json
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_dgemm8192.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch.log 2>&1
tail -3 $OUT/ncu_launch.log
