#!/bin/bash
# Round-2: PDL on the split-K reduce kernel -- help or harm; pre-pass (mode 3) vs in-kernel (mode 4) at the large sizes.
set -o pipefail
O=gpurun_out/r02m; mkdir -p $O
python -m portblas_b200.build > /dev/null || { echo "BUILD BROKEN"; exit 9; }
for shp in f32,64,147,13225 f32,64,64,12544 f32,128,256,25088 f32,1024,1024,1024 f32,512,512,1048576 f32,384,5408,3456 bf16,256,256,25088; do
  timeout 200 python tools/ab_variants.py --shape $shp --variants default,nopdl_reduce,nopdl --burst-steps 50 --rounds 3 --sustained-s 0.2 > $O/ab_$shp.jsonl 2> $O/ab_$shp.err
  python - <<PY
import json
for l in open("$O/ab_$shp.jsonl"):
    if l.startswith("{"):
        d = json.loads(l); print(d["workload"], d["variant"], "burst", d["burst_ms"], "sustained", d["sustained_ms"])
PY
done
timeout 300 python tools/ab_variants.py --workload sgemm8192 --variants default,presplit_off --burst-steps 5 --rounds 4 > $O/ab_pre_8192.jsonl 2> $O/ab_pre.err; cat $O/ab_pre_8192.jsonl | cut -c1-300
timeout 300 python tools/ab_variants.py --workload sgemm16384 --variants default,presplit_off --burst-steps 3 --rounds 3 --sustained-s 1.0 > $O/ab_pre_16384.jsonl 2> $O/ab_pre2.err; cat $O/ab_pre_16384.jsonl | cut -c1-300
