#!/bin/bash
# Round-2 8-GPU record on the final code: bench.py at N = 8, 4, 2 exactly as the driver launches it, C++ sample on 8 devices.
set -o pipefail
O=gpurun_out/r02n8b; mkdir -p $O
python -m portblas_b200.build > /dev/null || { echo "BUILD BROKEN"; exit 9; }
for n in 8 4 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2956$n bench.py --gpus $n --steps 20 --warmup 3 > $O/bench_n$n.json 2> $O/bench_n$n.err; echo "bench n$n rc=$?"
done
python - <<PY
import json
for f in ("bench_n8.json", "bench_n4.json", "bench_n2.json"):
    for l in open("$O/" + f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, "MAIN", d["value"], d["ms_per_step"], "compute", d["compute_only"]["value"], d["compute_only"]["ms_per_step"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "e2e", d["e2e"]["value"], d["e2e"]["host_gb_per_s_this_rank"], d["gather"]["all_ranks_hold_identical_c"])
            for s in d.get("sub", []):
                print("   SUB", s["workload"], s["value"], s["ms_per_step"], s["steps"], s["roofline"]["frac"], s["clocks"]["sm_mhz"], s["clocks"]["reasons"])
PY
timeout 300 ./build/gemm_multi_b200 16384 8 0 > $O/sample_multi_16384_8.txt 2>&1; cat $O/sample_multi_16384_8.txt
