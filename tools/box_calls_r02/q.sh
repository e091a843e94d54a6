#!/bin/bash
# Round-2: full GPU suite + smoke + default bench on the final code of the round.
set -o pipefail
O=gpurun_out/r02q; mkdir -p $O
python -m portblas_b200.build > /dev/null || { echo "BUILD BROKEN"; exit 9; }
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?"; tail -12 $O/smoke.txt
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_gpu.txt
mkdir -p $O/ref_unittests; cp gpurun_out/ref_unittests/*.log $O/ref_unittests/ 2>/dev/null
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -c 300 $O/bench_n1.err
python - <<PY
import json
for l in open("$O/bench_n1.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("MAIN", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["e2e"]["value"], d["cpu_baseline"]["value"])
        for s in d.get("sub", []):
            print("SUB", s["workload"], s["value"], s["ms_per_step"], s["steps"], s["roofline"]["frac"], s["clocks"]["sm_mhz"], s["clocks"]["reasons"])
PY
