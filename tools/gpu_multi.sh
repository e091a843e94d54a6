#!/bin/bash
# multi-GPU pass: bench.py under torchrun for the sharded configurations (+ the three gather forms on cfg3)
# usage: gpurun --gpus N -- 'bash tools/gpu_multi.sh <tag> N'
TAG=$1; N=$2
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.txt
run() { # workload extra-args
  wl=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 --workload $wl --no-cpu-baseline "$@" > $OUT/bench_${wl}_n$N.json 2> $OUT/bench_${wl}_n$N.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/bench_${wl}_n$N.json") if l.startswith("{")][-1])
    print("$wl n=$N", d["value"], d["unit"], d["scaling"], "ms", d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"),
          {k: v for k, v in d.items() if "gather" in k}, d["clocks"])
except Exception as e:
    print("$wl FAILED", e); print(open("$OUT/bench_${wl}_n$N.err").read()[-2000:])
PY
}
run dgemm8192
run sgemm16384 --gather --no-e2e
run bf16gemm_batched --no-e2e
