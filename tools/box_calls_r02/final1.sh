#!/bin/bash
# Round-2 one-GPU record: full GPU suite, bench.py (both arms) as the driver runs it, ncu launch list + full capture of the
# dominant kernel of the default command, fp32 / bf16 sweeps of the reference's config_csv shapes.
set -o pipefail
O=gpurun_out/r02final; mkdir -p $O
python -m portblas_b200.build > /dev/null || { echo "BUILD BROKEN"; exit 9; }
nvidia-smi --query-gpu=index,name,power.limit,clocks.max.sm --format=csv > $O/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -12 $O/pytest_gpu.txt
mkdir -p $O/ref_unittests; cp gpurun_out/ref_unittests/*.log $O/ref_unittests/ 2>/dev/null
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -c 300 $O/bench_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref_n1.json 2> $O/bench_ref_n1.err; echo "ref rc=$?"
python - <<PY
import json
for f in ("bench_n1.json", "bench_ref_n1.json"):
    for l in open("$O/" + f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, "MAIN", d["value"], d["ms_per_step"], d.get("roofline"), d.get("clocks"), d.get("e2e"), d.get("cpu_baseline"))
            for s in d.get("sub", []):
                print("SUB", s["workload"], s["value"], s["ms_per_step"], s["steps"], s["roofline"]["frac"], s["roofline"]["bound"], s["clocks"]["sm_mhz"], s["clocks"]["reasons"], s["clocks"].get("samples"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default_main.csv python bench.py --steps 3 --warmup 3 --no-sub --no-e2e --no-cpu-baseline > $O/ncu_launches.log 2>&1; echo "launch list rc=$?"
python tools/launch_summary.py $O/launches_default_main.csv > $O/launch_summary_default_main.txt 2>&1; cat $O/launch_summary_default_main.txt | tail -12
bash tools/gpu_ncu.sh r02final sgemm16384 bf16gemm_batched
python tools/ncu_summary.py $O/prof_sgemm16384_raw.csv $O/prof_bf16gemm_batched_raw.csv > $O/ncu_summary.txt; cat $O/ncu_summary.txt
timeout 900 python tools/csv_sweep.py --dtype f32 --graph > $O/sweep_f32.jsonl 2> $O/sweep_f32.err; echo "sweep f32 rc=$?"; tail -1 $O/sweep_f32.jsonl
timeout 900 python tools/csv_sweep.py --dtype bf16 --graph > $O/sweep_bf16.jsonl 2> $O/sweep_bf16.err; echo "sweep bf16 rc=$?"; tail -1 $O/sweep_bf16.jsonl
