#!/bin/bash
# Full GPU pass: parity tests, bench line per workload, cuBLAS bar, ncu launch list, ncu --set full captures.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round_full.sh <tag> "<ncu workloads>"'
TAG=${1:-r01}
NCU_WL=${2:-dgemm8192}
bash tools/gpu_round.sh $TAG
OUT=gpurun_out/$TAG
timeout 240 python tools/cublas_compare.py --iters 10 > $OUT/cublas.jsonl 2> $OUT/cublas.err
cat $OUT/cublas.jsonl
bash tools/gpu_ncu.sh $TAG $NCU_WL
python tools/ncu_summary.py $OUT/prof_*_raw.csv > $OUT/ncu_summary.txt 2>&1
head -60 $OUT/ncu_summary.txt
