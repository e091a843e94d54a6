"""On-box comparison bar: the vendor library (cuBLAS via torch.matmul / torch.bmm) on the same shapes
as bench.py's workloads, timed the same way (CUDA events, warm-up, mean of N launches).  This is the
"cuBLAS benchmark" the reference ships beside its own (benchmark/cublas/blas3/gemm*.cpp:34-49,132-146:
cublas{S,D,H}gemm[StridedBatched] timed with cudaEvent).  Measurement aid only: nothing in the product
path calls cuBLAS.

    python tools/cublas_compare.py [--iters 10] > gpurun_out/cublas.jsonl
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--workloads", default=",".join(WORKLOADS))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    tdt = {"f64": torch.float64, "f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}
    for name in args.workloads.split(","):
        w = WORKLOADS[name]
        m, n, k, batch = w["m"], w["n"], w["k"], w["batch"]
        dt = tdt[w["dt"]]
        modes = [("default", False)] + ([("tf32", True)] if w["dt"] == "f32" else [])
        # row-major (k x m)^T view == column-major m x k: C^T = B^T A^T, the same GEMM cuBLAS sees
        bt = (torch.rand(batch, n, k, device=dev, dtype=torch.float32) * 7 - 2).to(dt)
        at = (torch.rand(batch, k, m, device=dev, dtype=torch.float32) * 7 - 2).to(dt)
        ct = torch.empty(batch, n, m, device=dev, dtype=dt)
        for mode, tf32 in modes:
            torch.backends.cuda.matmul.allow_tf32 = tf32
            for _ in range(3):
                torch.bmm(bt, at, out=ct) if batch > 1 else torch.matmul(bt[0], at[0], out=ct[0])
            torch.cuda.synchronize()
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.iters + 1)]
            evs[0].record()
            for i in range(args.iters):
                torch.bmm(bt, at, out=ct) if batch > 1 else torch.matmul(bt[0], at[0], out=ct[0])
                evs[i + 1].record()
            torch.cuda.synchronize()
            ts = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.iters)]
            fl = 2.0 * m * n * k * batch
            mean = sum(ts) / len(ts)
            print(json.dumps(dict(workload=name, lib="cuBLAS (torch %s)" % torch.__version__, mode=mode,
                                  mean_ms=round(mean, 4), best_ms=round(min(ts), 4),
                                  tflops_mean=round(fl / mean / 1e9, 2), tflops_best=round(fl / min(ts) / 1e9, 2))),
                  flush=True)
        del at, bt, ct
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
