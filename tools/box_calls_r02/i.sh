#!/bin/bash
# Round-2: cost of the pusher on one GPU (local copies), interleaved-via-strided tests + sweep rows, full GPU suite.
set -o pipefail
O=gpurun_out/r02i; mkdir -p $O
python -m portblas_b200.build > /dev/null || { echo "BUILD BROKEN"; exit 9; }
timeout 200 python tools/push_probe.py --copies 2 > $O/push_probe.txt 2>&1; cat $O/push_probe.txt
timeout 200 python tools/push_probe.py --copies 8 >> $O/push_probe.txt 2>&1; tail -4 $O/push_probe.txt
timeout 200 python tools/push_probe.py --copies 2 --m 8192 --n 8192 --k 8192 >> $O/push_probe.txt 2>&1; tail -4 $O/push_probe.txt
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -x -k "interleaved or batched" > $O/pytest_ilv.txt 2>&1; echo "ilv rc=$?"; tail -5 $O/pytest_ilv.txt
timeout 600 python tools/csv_sweep.py --dtype f32 --graph --api gemm_batched,gemm_batched_strided > $O/sweep_f32_batched.jsonl 2> $O/sweep_b.err; echo "sweep rc=$?"
python - <<PY
import json
for l in open("$O/sweep_f32_batched.jsonl"):
    if l.startswith("{"):
        r = json.loads(l)
        if "ms" in r and 2.0 * r["m"] * r["n"] * r["k"] * r["batch"] > 5e8:
            print(r["api"][5:], r["m"], r["n"], r["k"], r["batch"], r["batch_type"][:5], r["kernel"], "ms", r["ms"], "eager", r["eager_ms"], "frac", r["frac_of_roof"], "rp", r["repack"], "ok", r["ok"])
PY
