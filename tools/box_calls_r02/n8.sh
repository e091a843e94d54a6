#!/bin/bash
# Round-2 8-GPU box call: bench.py at N=8 and N=4 exactly as the driver launches it (ours + reference arm), C++ sample.
set -o pipefail
O=gpurun_out/r02n8; mkdir -p $O
python -m portblas_b200.build > /dev/null || { echo "BUILD BROKEN"; exit 9; }
nvidia-smi --query-gpu=index,name,power.limit --format=csv > $O/gpus.txt; nvidia-smi topo -m >> $O/gpus.txt 2>&1; numactl -H >> $O/gpus.txt 2>&1
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > $O/bench_n$n.json 2> $O/bench_n$n.err; echo "bench n$n rc=$?"; tail -c 600 $O/bench_n$n.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > $O/bench_ref_n8.json 2> $O/bench_ref_n8.err; echo "ref n8 rc=$?"
python - <<PY
import json
for f in ("bench_n8.json", "bench_n4.json", "bench_ref_n8.json"):
    for l in open("$O/" + f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, "MAIN", d["value"], d["ms_per_step"], d.get("compute_only"), d.get("gather"), d.get("roofline", {}).get("frac"), d.get("clocks"), d.get("e2e"), d.get("cpu_baseline"))
            for s in d.get("sub", []):
                print("SUB", s["workload"], s["value"], s["ms_per_step"], s["steps"], s["roofline"]["frac"], s["clocks"]["sm_mhz"], s["clocks"]["reasons"])
PY
timeout 300 ./build/gemm_multi_b200 16384 8 0 > $O/sample_multi_16384_8.txt 2>&1; cat $O/sample_multi_16384_8.txt
