#!/bin/bash
# Round-2: 4 epilogue + 8 splitter warps in the in-kernel split modes -- parity and timing.
set -o pipefail
O=gpurun_out/r02o; mkdir -p $O
python -m portblas_b200.build > /dev/null || { echo "BUILD BROKEN"; exit 9; }
timeout 400 python -m pytest tests/test_split16_gpu.py tests/test_beta_zero_nan_gpu.py tests/test_nonfinite_gpu.py -q -x > $O/pytest_a.txt 2>&1; echo "a rc=$?"; tail -3 $O/pytest_a.txt
if ! grep -q " passed" $O/pytest_a.txt || grep -q "failed\|error" $O/pytest_a.txt; then tail -60 $O/pytest_a.txt; exit 1; fi
for shp in f32,512,512,1048576 f32,1024,1024,1024 f32,384,5408,3456 f32,512,4608,6272 f32,64,147,13225 f32,384,384,64,896 f32,230,49,230,1000; do
  timeout 200 python tools/ab_variants.py --shape $shp --variants default,split16_off --burst-steps 30 --rounds 3 --sustained-s 0.3 > $O/ab_$shp.jsonl 2> $O/ab_$shp.err
  python - <<PY
import json
for l in open("$O/ab_$shp.jsonl"):
    if l.startswith("{"):
        d = json.loads(l); print(d["workload"], d["variant"], "burst", d["burst_ms"], d["burst_tflops"], "sustained", d["sustained_ms"], d["sustained_tflops"])
PY
done
timeout 300 python tools/ab_variants.py --workload sgemm8192 --variants default,presplit_off --burst-steps 5 --rounds 3 > $O/ab_pre_8192.jsonl 2> $O/ab_pre.err; cat $O/ab_pre_8192.jsonl | cut -c1-260
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_ref_parity_gpu.py -q > $O/pytest_b.txt 2>&1; echo "b rc=$?"; tail -4 $O/pytest_b.txt
