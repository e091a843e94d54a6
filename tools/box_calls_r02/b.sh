#!/bin/bash
# Round-2 second box call: smoke of the new scheduler under a timeout, parity suites, reference bench harness, bench A/B.
O=gpurun_out/r02b; mkdir -p $O
timeout 300 python -m pytest tests/test_beta_zero_nan_gpu.py -q -x > $O/pytest_nan.txt 2>&1; echo "nan rc=$?"; tail -3 $O/pytest_nan.txt
if ! grep -q "passed" $O/pytest_nan.txt || grep -q "failed" $O/pytest_nan.txt; then tail -40 $O/pytest_nan.txt; fi
timeout 1200 python -m pytest tests -m gpu -q -x -k "not reference_unit_tests and not joint_matrix" > $O/pytest_gpu.txt 2>&1; echo "gpu rc=$?"; tail -15 $O/pytest_gpu.txt
for mode in "1 1" "0 0" "1 0" "0 1"; do set -- $mode
  PBX_DYNAMIC_SCHED=$1 PBX_PDL=$2 timeout 300 python bench.py --workload bf16gemm_batched --no-sub --no-e2e --no-cpu-baseline --steps 2000 > $O/bench_cfg4_dyn$1_pdl$2.json 2>&1
  python - <<PY
import json
for l in open("$O/bench_cfg4_dyn$1_pdl$2.json"):
    if l.startswith("{"):
        d=json.loads(l); print("cfg4 dyn=$1 pdl=$2", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"])
PY
done
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -c 300 $O/bench_default.err
python - <<PY
import json
for l in open("$O/bench_default.json"):
    if l.startswith("{"):
        d=json.loads(l)
        print("MAIN", d["value"], d["ms_per_step"], d["roofline"], d["clocks"], d["e2e"], d["cpu_baseline"])
        for s in d.get("sub", []):
            print("SUB", s["workload"], s["value"], s["ms_per_step"], s["steps"], s["roofline"]["frac"], s["roofline"]["bound"], s["clocks"]["sm_mhz"], s["clocks"]["reasons"], s["clocks"].get("samples"))
PY
