#!/bin/bash
# Round-2 (8 GPUs): compute + gather at N=8 with the paced pusher, the unpaced pusher and the round-1 epilogue stores.
set -o pipefail
O=gpurun_out/r02k; mkdir -p $O
python -m portblas_b200.build > /dev/null || { echo "BUILD BROKEN"; exit 9; }
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 8 --steps 20 --warmup 3 --no-sub --no-e2e > $O/bench_n8_$name.json 2> $O/bench_n8_$name.err
  echo "$name rc=$?"
}
run paced PBX_MULTICAST_PACE=1
run unpaced PBX_MULTICAST_PACE=0
run epilogue PBX_MULTICAST_PUSH=0
python - <<PY
import json
for f in ("paced", "unpaced", "epilogue"):
    for l in open("$O/bench_n8_" + f + ".json"):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, "fused", d["value"], d["ms_per_step"], "compute", d["compute_only"]["value"], d["compute_only"]["ms_per_step"], d["clocks"]["sm_mhz"], d["compute_only"]["clocks"]["sm_mhz"], d["gather"]["all_ranks_hold_identical_c"])
PY
