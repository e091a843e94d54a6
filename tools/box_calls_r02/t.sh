#!/bin/bash
# Round-2: raster group height (L2 reuse / DRAM traffic) under the power cap; new host-path tests.
set -o pipefail
O=gpurun_out/r02t; mkdir -p $O
python -m portblas_b200.build > /dev/null || { echo "BUILD BROKEN"; exit 9; }
timeout 300 python -m pytest tests/test_host_path_gpu.py -q > $O/pytest_host.txt 2>&1; echo "host rc=$?"; tail -3 $O/pytest_host.txt
timeout 300 python tools/ab_variants.py --workload bf16gemm8192 --variants default,gm4,gm6,gm12,gm16,gm32 --burst-steps 10 --rounds 3 --sustained-s 1.0 > $O/ab_gm_bf16.jsonl 2> $O/ab_gm_bf16.err
timeout 400 python tools/ab_variants.py --workload sgemm16384 --variants default,gm4,gm16,gm32 --burst-steps 3 --rounds 2 --sustained-s 1.0 > $O/ab_gm_sgemm.jsonl 2> $O/ab_gm_sgemm.err
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/ab_*.jsonl")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); print(f"{d['workload']:14s} {d['variant']:8s} burst {d['burst_ms']:8.4f} {d['burst_tflops']:7.1f} | sustained {d['sustained_ms']:8.4f} {d['sustained_tflops']:7.1f}")
PY
