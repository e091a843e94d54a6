#!/bin/bash
# Round-2: where do the skinny-M split-K rows lose against round 1?
set -o pipefail
O=gpurun_out/r02v; mkdir -p $O
python -m portblas_b200.build > /dev/null || { echo "BUILD BROKEN"; exit 9; }
for shp in f32,64,147,13225 f32,64,64,12544 f32,64,363,103968; do
  timeout 200 python tools/ab_variants.py --shape $shp --variants default,r01like,static,nopdl,split16_off,noswap --burst-steps 100 --rounds 3 --sustained-s 0.2 > $O/ab_$shp.jsonl 2> $O/ab_$shp.err
  python - <<PY
import json
for l in open("$O/ab_$shp.jsonl"):
    if l.startswith("{"):
        d = json.loads(l); print(d["workload"], d["variant"], "burst", d["burst_ms"], "sustained", d["sustained_ms"])
PY
done
