#!/bin/bash
# Round-2 eighth box call: relaxed splitter hand-over -- parity, planner probe, cfg5 / 1024 A/B.
set -o pipefail
O=gpurun_out/r02h; mkdir -p $O
python -m portblas_b200.build > /dev/null || { echo "BUILD BROKEN"; exit 9; }
timeout 200 python -m pytest tests/test_split16_gpu.py -q -x > $O/pytest_split.txt 2>&1; echo "split rc=$?"; tail -3 $O/pytest_split.txt
if ! grep -q " passed" $O/pytest_split.txt || grep -q "failed\|error" $O/pytest_split.txt; then tail -60 $O/pytest_split.txt; exit 1; fi
timeout 600 python tools/plan_probe.py --dtype f32 > $O/plan_probe_f32.jsonl 2> $O/plan_probe_f32.err; echo "probe rc=$?"
timeout 300 python tools/ab_variants.py --workload sgemm_splitk --variants default,split16_off --burst-steps 5 --rounds 3 > $O/ab_splitk.jsonl 2> $O/ab_splitk.err; cat $O/ab_splitk.jsonl
