#!/bin/bash
# Round-2 fifth box call (2 GPUs): the pusher warp over real NVLink, the device group on real peers, bench.py at N=2.
O=gpurun_out/r02e; mkdir -p $O
set -o pipefail
python -m portblas_b200.build > /dev/null || { echo "BUILD BROKEN"; exit 9; }
nvidia-smi --query-gpu=index,name,power.limit --format=csv > $O/gpus.txt; nvidia-smi topo -m >> $O/gpus.txt 2>&1
timeout 300 python tools/ab_variants.py --workload sgemm_splitk --variants default,split16_off --burst-steps 5 --rounds 3 > $O/ab_splitk.jsonl 2> $O/ab_splitk.err; cat $O/ab_splitk.jsonl
timeout 300 python tools/ab_variants.py --workload sgemm1024 --variants default,split16_off --burst-steps 200 --rounds 3 > $O/ab_sgemm1024.jsonl 2> $O/ab_1024.err; cat $O/ab_sgemm1024.jsonl
timeout 600 python -m pytest tests/test_multicast_gpu.py tests/test_multi_group_gpu.py tests/test_split16_gpu.py -q -x > $O/pytest_multi.txt 2>&1; echo "multi rc=$?"; tail -25 $O/pytest_multi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"; tail -c 1500 $O/bench_n2.err
PBX_MULTICAST_PUSH=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 3 --no-sub --no-e2e > $O/bench_n2_nopush.json 2> $O/bench_n2_nopush.err; echo "bench n2 nopush rc=$?"
python - <<PY
import json
for f in ("bench_n2.json", "bench_n2_nopush.json"):
    for l in open("$O/" + f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, "MAIN", d["value"], d["ms_per_step"], d.get("compute_only"), d.get("gather"), d["roofline"]["frac"], d["clocks"], d["e2e"])
            for s in d.get("sub", []):
                print("SUB", s["workload"], s["value"], s["ms_per_step"], s["steps"], s["roofline"]["frac"], s["clocks"]["sm_mhz"], s["clocks"]["reasons"])
PY
timeout 300 ./build/gemm_multi_b200 8192 2 0 > $O/sample_multi.txt 2>&1; cat $O/sample_multi.txt
