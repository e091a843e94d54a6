#!/bin/bash
# ncu --set full capture of the dominant kernel of each workload (one GPU, one launch each).
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_ncu.sh <tag> wl1 wl2 ...'
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for wl in "$@"; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_ -s 3 -c 1 -f -o $OUT/prof_$wl \
    python bench.py --workload $wl --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/ncu_$wl.log 2>&1
  tail -2 $OUT/ncu_$wl.log
done
