#!/bin/bash
# quick GPU pass: selected tests + selected bench workloads
# usage: gpurun -- 'bash tools/gpu_quick.sh <tag> "<pytest -k expr or empty for all>" wl1 wl2 ...'
TAG=$1; KEXPR=$2; shift; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ -n "$KEXPR" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
else
  timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
fi
tail -15 $OUT/pytest_gpu.log
for wl in "$@"; do
  timeout 300 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$wl.json"))
    print("$wl", d["value"], d["unit"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"] if d.get("e2e") else None, d["clocks"])
except Exception as e:
    print("$wl FAILED", e); print(open("$OUT/bench_$wl.err").read()[-1500:])
PY
done
