#!/bin/bash
# Round-2: do longer mbarrier suspensions (fewer polling instructions) buy clocks under the power cap?
set -o pipefail
O=gpurun_out/r02p; mkdir -p $O
python -m portblas_b200.build > /dev/null || { echo "BUILD BROKEN"; exit 9; }
timeout 300 python tools/ab_variants.py --workload bf16gemm_batched --variants default,hint1us,hint20us,hint1ms,cublas --rounds 5 --sustained-s 1.0 > $O/ab_cfg4.jsonl 2> $O/ab_cfg4.err
timeout 300 python tools/ab_variants.py --workload bf16gemm8192 --variants default,hint1us,hint20us,hint1ms,cublas --burst-steps 10 --rounds 4 --sustained-s 1.0 > $O/ab_bf16_8192.jsonl 2> $O/ab_bf16.err
timeout 300 python tools/ab_variants.py --workload sgemm8192 --variants default,hint20us,hint1ms --burst-steps 5 --rounds 3 --sustained-s 1.0 > $O/ab_sgemm8192.jsonl 2> $O/ab_s.err
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/ab_*.jsonl")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); print(f"{d['workload']:18s} {d['variant']:10s} burst {d['burst_ms']:8.4f} {d['burst_tflops']:7.1f} | sustained {d['sustained_ms']:8.4f} {d['sustained_tflops']:7.1f}")
PY
