#!/bin/bash
# Round-2 seventh box call: planner calibration, ncu of cfg5 and of the default bench command (launch list).
set -o pipefail
O=gpurun_out/r02g; mkdir -p $O
python -m portblas_b200.build > /dev/null || { echo "BUILD BROKEN"; exit 9; }
timeout 200 python -m pytest tests/test_split16_gpu.py -q -x > $O/pytest_split.txt 2>&1; echo "split rc=$?"; tail -3 $O/pytest_split.txt
if ! grep -q " passed" $O/pytest_split.txt || grep -q "failed\|error" $O/pytest_split.txt; then tail -60 $O/pytest_split.txt; exit 1; fi
timeout 600 python tools/plan_probe.py --dtype f32 > $O/plan_probe_f32.jsonl 2> $O/plan_probe_f32.err; echo "probe rc=$?"
timeout 300 python tools/plan_probe.py --dtype bf16 > $O/plan_probe_bf16.jsonl 2> $O/plan_probe_bf16.err; echo "probe bf16 rc=$?"
timeout 300 python tools/ab_variants.py --workload sgemm_splitk --variants default,split16_off --burst-steps 5 --rounds 3 > $O/ab_splitk.jsonl 2> $O/ab_splitk.err; cat $O/ab_splitk.jsonl
bash tools/gpu_ncu.sh r02g sgemm_splitk
python tools/ncu_summary.py $O/prof_sgemm_splitk_raw.csv
