"""A/B timing of launch-feature variants of the tcgen05 kernel against each other and against cuBLAS (torch.bmm /
torch.matmul) in ONE process, interleaved, in two regimes:

  burst     : `--burst-steps` back-to-back launches after a `--rest` pause (clocks at maximum), repeated `--rounds` times;
              the median round is reported
  sustained : back-to-back launches for `--sustained-s` seconds (the board settles at its power cap)

Variants are handles created under different environment switches (every PBX_* switch is read into the handle by
pbx_create).  Measurement aid only.

    python tools/ab_variants.py --workload bf16gemm_batched > gpurun_out/ab_cfg4.jsonl
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from bench import WORKLOADS, algorithmic  # noqa: E402
from portblas_b200 import SB_Handle, blas  # noqa: E402

VARIANTS = {
    "default": ({}, {}),
    "static": ({"PBX_DYNAMIC_SCHED": "0"}, {}),
    "nopdl": ({"PBX_PDL": "0"}, {}),
    "nopdl_reduce": ({"PBX_PDL_REDUCE": "0"}, {}),
    "static_nopdl": ({"PBX_DYNAMIC_SCHED": "0", "PBX_PDL": "0"}, {}),
    "cg2_256": ({}, {"PBX_TC_CONFIG": "2,256"}),
    "cg2_128": ({}, {"PBX_TC_CONFIG": "2,128"}),
    "cg1_128": ({}, {"PBX_TC_CONFIG": "1,128"}),
    "static_cg2_256": ({"PBX_DYNAMIC_SCHED": "0"}, {"PBX_TC_CONFIG": "2,256"}),
    "split16_off": ({}, {"PBX_F32_SPLIT16": "0"}),
    "chunk32": ({}, {"PBX_TF32_CHUNK_KB": "32"}),      # timing probes only: longer tensor-core accumulation chains
    "chunk64": ({}, {"PBX_TF32_CHUNK_KB": "64"}),      # exceed the fp32 error budget
    "chunk_inf": ({}, {"PBX_TF32_CHUNK_KB": "1000000"}),
    "presplit_off": ({}, {"PBX_TF32_PRESPLIT": "0"}),
    "r01like": ({"PBX_DYNAMIC_SCHED": "0", "PBX_PDL": "0"}, {"PBX_F32_SPLIT16": "0"}),
    "noswap": ({}, {"PBX_TC_SWAP": "0"}),
    "gm4": ({}, {"PBX_GROUP_M": "4"}),
    "gm6": ({}, {"PBX_GROUP_M": "6"}),
    "gm12": ({}, {"PBX_GROUP_M": "12"}),
    "gm16": ({}, {"PBX_GROUP_M": "16"}),
    "gm32": ({}, {"PBX_GROUP_M": "32"}),
    "hint1us": ({}, {"PBX_WAIT_HINT_NS": "1000"}),
    "hint20us": ({}, {"PBX_WAIT_HINT_NS": "20000"}),
    "hint1ms": ({}, {"PBX_WAIT_HINT_NS": "1000000"}),
}


class Env:
    def __init__(self, kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="bf16gemm_batched", choices=sorted(WORKLOADS))
    ap.add_argument("--variants", default="default,static,nopdl,static_nopdl,cublas")
    ap.add_argument("--burst-steps", type=int, default=20)
    ap.add_argument("--rounds", type=int, default=9)
    ap.add_argument("--rest", type=float, default=0.4)
    ap.add_argument("--sustained-s", type=float, default=0.6)
    ap.add_argument("--shape", default="", help="dtype,m,n,k[,batch]: an ad-hoc NN problem instead of a bench.py workload")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.shape:
        f = args.shape.split(",")
        w = dict(dt=f[0], m=int(f[1]), n=int(f[2]), k=int(f[3]), batch=int(f[4]) if len(f) > 4 else 1, ta="n", tb="n", alpha=1.0,
                 beta=0.0)
        args.workload = args.shape
    dev = torch.device("cuda", 0)
    tdt = {"f64": torch.float64, "f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}[w["dt"]]
    m, n, k, batch = w["m"], w["n"], w["k"], w["batch"]
    a = (torch.rand(m * k * batch, device=dev) * 7 - 2).to(tdt)
    b = (torch.rand(k * n * batch, device=dev) * 7 - 2).to(tdt)
    c = torch.zeros(m * n * batch, device=dev, dtype=tdt)
    flops, byts = algorithmic(w)
    runners = {}
    for name in args.variants.split(","):
        if name == "cublas":
            at, bt, ct = a.view(batch, k, m), b.view(batch, n, k), c.view(batch, n, m)   # row-major views of the column-major operands

            def run(at=at, bt=bt, ct=ct):
                torch.bmm(bt, at, out=ct) if batch > 1 else torch.matmul(bt[0], at[0], out=ct[0])
            runners[name] = run
            continue
        create_env, call_env = VARIANTS[name]
        with Env({**create_env, **call_env}):   # every PBX_* switch is read into the handle when it is created
            h = SB_Handle(0)

        def run(h=h):
            if batch == 1:
                blas._gemm(h, w["ta"], w["tb"], m, n, k, w["alpha"], a, m, b, k, w["beta"], c, m)
            else:
                blas._gemm_strided_batched(h, w["ta"], w["tb"], m, n, k, w["alpha"], a, m, m * k, b, k, k * n, w["beta"], c, m,
                                           m * n, batch)
        runners[name] = run
    for run in runners.values():
        for _ in range(3):
            run()
    torch.cuda.synchronize()

    def timed(run, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            run()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    burst = {name: [] for name in runners}
    for _ in range(args.rounds):
        for name, run in runners.items():
            time.sleep(args.rest)
            burst[name].append(timed(run, args.burst_steps))
    sustained = {}
    for name, run in runners.items():
        time.sleep(1.0)
        est = timed(run, 5)
        steps = max(20, int(args.sustained_s * 1e3 / est))
        sustained[name] = (timed(run, steps), steps)
    for name in runners:
        bm = statistics.median(burst[name])
        sm, steps = sustained[name]
        print(json.dumps(dict(workload=args.workload, variant=name, burst_ms=round(bm, 4), burst_best_ms=round(min(burst[name]), 4),
                              burst_tflops=round(flops / bm / 1e9, 1), burst_gbs=round(byts / bm / 1e6, 1),
                              sustained_ms=round(sm, 4), sustained_steps=steps, sustained_tflops=round(flops / sm / 1e9, 1),
                              sustained_gbs=round(byts / sm / 1e6, 1))), flush=True)


if __name__ == "__main__":
    main()
