#!/bin/bash
# Full GPU-box pass: every parity test, a bench line per BASELINE workload (e2e + cpu baseline on the default one),
# the reference arm, the routines built on the GEMM path, the gbench-style harness, an ncu launch list.
# usage: gpurun --timeout 2400 -- 'bash tools/gpu_round_full.sh <tag>'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
( timeout 1500 python -m pytest tests -m gpu -q --tb=short > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log )
tail -6 $OUT/pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 400 python bench.py --steps 10 --warmup 3 > $OUT/bench_dgemm8192.json 2> $OUT/bench_dgemm8192.err
for wl in sgemm8192 sgemm16384 bf16gemm8192 bf16gemm_batched hgemm_batched sgemm_splitk sgemm1024; do
  timeout 300 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f.split("bench_")[1][:-5], d["value"], d["unit"], "ms", d["ms_per_step"], "frac", r.get("frac"), "sust", r.get("frac_of_sustained"),
              "e2e", (d.get("e2e") or {}).get("value"), "launches", d.get("gpu_launches"), (d.get("clocks") or {}).get("sm_mhz"), (d.get("clocks") or {}).get("reasons"))
    except Exception as e:
        print(f, "FAILED", e)
PY
timeout 300 python tools/ext_bench.py --steps 5 --warmup 2 > $OUT/ext_bench.jsonl 2> $OUT/ext_bench.err; cat $OUT/ext_bench.jsonl
for op in symm trsm; do
  timeout 200 python tools/portblas_bench.py --op $op --types float,double --benchmark_format=json --benchmark_out $OUT/gbench_$op.json > /dev/null 2> $OUT/gbench_$op.err
done
timeout 300 python tools/portblas_bench.py --op gemm_batched --from-fixture --types float --benchmark_out $OUT/gbench_gemm_batched.json > $OUT/gbench_gemm_batched.txt 2> $OUT/gbench_gemm_batched.err
tail -4 $OUT/gbench_gemm_batched.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_dgemm8192.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch.log 2>&1
tail -2 $OUT/ncu_launch.log
