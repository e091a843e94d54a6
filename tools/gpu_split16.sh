#!/bin/bash
# Round-2 first step for the experimental tf32 + 2 x bf16 fp32 path (never run on a GPU yet): parity under a timeout
# (a wrong mbarrier byte count would hang), then SGEMM 8192 / 16384 with and without it.
O=gpurun_out/split16; mkdir -p $O
PBX_RUN_EXPERIMENTAL=1 timeout 240 python -m pytest tests/test_split16_experimental_gpu.py -x -q > $O/pytest.txt 2>&1
echo "pytest rc=$?" >> $O/pytest.txt; tail -5 $O/pytest.txt
if grep -q "passed" $O/pytest.txt && ! grep -q "failed" $O/pytest.txt; then
  for w in sgemm8192 sgemm16384; do
    timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_${w}_base.json 2>&1
    PBX_F32_SPLIT16=1 timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_${w}_split16.json 2>&1
    tail -c 600 $O/bench_${w}_base.json; echo; tail -c 600 $O/bench_${w}_split16.json; echo
  done
fi
