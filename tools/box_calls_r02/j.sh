#!/bin/bash
# Round-2 (2 GPUs): paced pusher on/off, multicast correctness across peers.
set -o pipefail
O=gpurun_out/r02j; mkdir -p $O
python -m portblas_b200.build > /dev/null || { echo "BUILD BROKEN"; exit 9; }
timeout 400 python -m pytest tests/test_multicast_gpu.py -q -x > $O/pytest_multicast.txt 2>&1; echo "multicast rc=$?"; tail -4 $O/pytest_multicast.txt
for pace in 1 0; do
  PBX_MULTICAST_PACE=$pace timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$pace bench.py --gpus 2 --steps 20 --warmup 3 --no-sub --no-e2e > $O/bench_n2_pace$pace.json 2> $O/bench_n2_pace$pace.err; echo "bench pace=$pace rc=$?"
done
python - <<PY
import json
for f in ("bench_n2_pace1.json", "bench_n2_pace0.json"):
    for l in open("$O/" + f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, "fused", d["value"], d["ms_per_step"], "compute", d["compute_only"]["value"], d["compute_only"]["ms_per_step"], d["clocks"]["sm_mhz"], d["compute_only"]["clocks"]["sm_mhz"])
PY
