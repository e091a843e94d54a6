#!/usr/bin/env python
"""Device-timed throughput of the routines built on the GEMM path (_symm, _trsm, complex _gemm) with cuBLAS
(through torch) timed the same way beside them.  One JSON line per case.

    python tools/ext_bench.py [--steps 10] [--warmup 3] [--size 8192]

flops: symm 2*K*M*N, trsm K*M*N (K = M left / N right), complex gemm 8*M*N*K real flops.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from portblas_b200 import SB_Handle, blas  # noqa: E402


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=8192)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    h = SB_Handle(0)
    n = args.size
    out = []
    for dt, name in ((torch.float32, "f32"), (torch.float64, "f64")):
        a = torch.rand(n * n, device=dev, dtype=dt) * 7 - 2
        b = torch.rand(n * n, device=dev, dtype=dt) * 7 - 2
        c = torch.zeros(n * n, device=dev, dtype=dt)
        # ---- symm
        l0 = h.launch_count
        ms = timed(lambda: blas._symm(h, "l", "u", n, n, 1.0, a, n, b, n, 0.0, c, n), args.steps, args.warmup)
        launches = (h.launch_count - l0) // (args.steps + args.warmup)
        a2 = a.view(n, n)
        sym = torch.triu(a2.T) + torch.triu(a2.T, 1).T   # col-major upper triangle mirrored (row-major view is the transpose)
        ms_ref = timed(lambda: torch.matmul(sym, b.view(n, n)), args.steps, args.warmup)
        out.append(dict(op="symm", dtype=name, m=n, n=n, ms=round(ms, 4), tflops=round(2.0 * n ** 3 / ms / 1e9, 2),
                        launches_per_call=launches, kernel=h.last_kernel,
                        cublas_gemm_same_shape_tflops=round(2.0 * n ** 3 / ms_ref / 1e9, 2)))
        print(json.dumps(out[-1]), flush=True)
        # ---- trsm: well-conditioned lower-triangular A (column-major)
        t = torch.tril(torch.rand(n, n, device=dev, dtype=dt) * 2 - 1) / n + torch.eye(n, device=dev, dtype=dt) * 4
        a_cm = t.T.contiguous().view(-1)           # column-major storage of t
        b0 = b.clone()
        l0 = h.launch_count

        def run_trsm():
            b.copy_(b0)
            blas._trsm(h, "l", "l", "n", "n", n, n, 1.0, a_cm, n, b, n)

        ms_all = timed(run_trsm, args.steps, args.warmup)
        ms_copy = timed(lambda: b.copy_(b0), args.steps, args.warmup)
        launches = (h.launch_count - l0) // (args.steps + args.warmup)
        ms = ms_all - ms_copy
        bt = b0.view(n, n).T.contiguous()           # row-major view of the col-major B
        ms_ref = timed(lambda: torch.linalg.solve_triangular(t, bt, upper=False), args.steps, args.warmup)
        out.append(dict(op="trsm", dtype=name, m=n, n=n, ms=round(ms, 4), tflops=round(1.0 * n ** 3 / ms / 1e9, 2),
                        launches_per_call=launches, cublas_trsm_tflops=round(1.0 * n ** 3 / ms_ref / 1e9, 2)))
        print(json.dumps(out[-1]), flush=True)
        del a, b, c, t, a_cm, b0, bt, sym
    for dt, name in ((torch.complex64, "c64"), (torch.complex128, "c128")):
        nn = n // 2
        a = torch.randn(nn * nn, device=dev, dtype=dt)
        b = torch.randn(nn * nn, device=dev, dtype=dt)
        c = torch.zeros(nn * nn, device=dev, dtype=dt)
        l0 = h.launch_count
        ms = timed(lambda: blas._gemm(h, "n", "n", nn, nn, nn, 1.0 + 0.5j, a, nn, b, nn, 0j, c, nn), args.steps,
                   args.warmup)
        launches = (h.launch_count - l0) // (args.steps + args.warmup)
        ms_ref = timed(lambda: torch.matmul(a.view(nn, nn), b.view(nn, nn)), args.steps, args.warmup)
        out.append(dict(op="cgemm", dtype=name, m=nn, n=nn, k=nn, ms=round(ms, 4),
                        real_tflops=round(8.0 * nn ** 3 / ms / 1e9, 2), launches_per_call=launches,
                        cublas_tflops=round(8.0 * nn ** 3 / ms_ref / 1e9, 2)))
        print(json.dumps(out[-1]), flush=True)
    h.close()


if __name__ == "__main__":
    main()
