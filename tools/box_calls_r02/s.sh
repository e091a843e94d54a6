#!/bin/bash
# Round-2: 128-byte-row TMA-store epilogue for 16-bit outputs -- parity and cfg4 / bf16 8192 timing next to cuBLAS.
set -o pipefail
O=gpurun_out/r02s; mkdir -p $O
python -m portblas_b200.build > /dev/null || { echo "BUILD BROKEN"; exit 9; }
timeout 600 python -m pytest tests/test_beta_zero_nan_gpu.py tests/test_multicast_gpu.py -q -x > $O/pytest_a.txt 2>&1; echo "a rc=$?"; tail -3 $O/pytest_a.txt
if ! grep -q " passed" $O/pytest_a.txt || grep -q "failed\|error" $O/pytest_a.txt; then tail -60 $O/pytest_a.txt; exit 1; fi
timeout 300 python tools/ab_variants.py --workload bf16gemm_batched --variants default,static,cublas --rounds 7 --sustained-s 1.0 > $O/ab_cfg4.jsonl 2> $O/ab_cfg4.err
timeout 300 python tools/ab_variants.py --workload hgemm_batched --variants default,cublas --rounds 5 --sustained-s 1.0 > $O/ab_cfg4h.jsonl 2> $O/ab_cfg4h.err
timeout 300 python tools/ab_variants.py --workload bf16gemm8192 --variants default,cublas --burst-steps 10 --rounds 4 --sustained-s 1.0 > $O/ab_bf16_8192.jsonl 2> $O/ab_bf16.err
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/ab_*.jsonl")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); print(f"{d['workload']:18s} {d['variant']:10s} burst {d['burst_ms']:8.4f} {d['burst_tflops']:7.1f} | sustained {d['sustained_ms']:8.4f} {d['sustained_tflops']:7.1f}")
PY
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_full_size_gpu.py tests/test_ref_parity_gpu.py tests/test_blas3_ext_gpu.py -q -k "not f64" > $O/pytest_b.txt 2>&1; echo "b rc=$?"; tail -4 $O/pytest_b.txt
