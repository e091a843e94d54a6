"""Calibration data for the tile planner (gemm_tcgen05.cu make_plan): device time of a few mid-size shapes under every
forced tile configuration x K-slice count, next to what the planner picks on its own.  One JSON line per measurement.

    python tools/plan_probe.py [--dtype f32] > gpurun_out/plan_probe.jsonl
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from portblas_b200 import SB_Handle, blas  # noqa: E402

SHAPES = [(1024, 1024, 1024), (384, 5408, 3456), (512, 4608, 6272), (2048, 1024, 1568), (256, 6272, 2304), (512, 512, 1 << 20),
          (256, 256, 25088), (128, 12544, 1152), (2240, 1024, 4096), (4096, 4096, 4096), (512, 2304, 784), (1120, 1024, 1024)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="f32")
    args = ap.parse_args()
    dt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[args.dtype]
    dev = torch.device("cuda", 0)
    h = SB_Handle(0)
    for m, n, k in SHAPES:
        a = (torch.rand(m * k, device=dev) * 7 - 2).to(dt)
        b = (torch.rand(k * n, device=dev) * 7 - 2).to(dt)
        c = torch.zeros(m * n, device=dev, dtype=dt)
        kb = (k + (31 if dt == torch.float32 else 63)) // (32 if dt == torch.float32 else 64)

        def run():
            blas._gemm(h, "n", "n", m, n, k, 1.0, a, m, b, k, 0.0, c, m)

        def timed():
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            reps = 20 if 2.0 * m * n * k < 5e10 else 5
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e9
            for _ in range(3):
                e0.record()
                for _ in range(reps):
                    run()
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / reps)
            return best

        os.environ.pop("PBX_TC_CONFIG", None)
        h.reload_env()
        h.set_split_k(0)
        t = timed()
        print(json.dumps(dict(m=m, n=n, k=k, dtype=args.dtype, cfg="auto", slices=h.last_split_k, presplit=h.last_presplit,
                              us=round(t * 1e3, 2), tflops=round(2.0 * m * n * k / t / 1e9, 1))), flush=True)
        for cfg in ("2,256", "2,128", "1,128"):
            cg, bn = (int(x) for x in cfg.split(","))
            if (cg == 2 and m <= 128) or (bn == 256 and n <= 128):
                continue
            tiles = -(-m // (128 * cg)) * -(-n // bn)
            units = 148 // cg
            cand = sorted({s for s in (1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64, max(1, units // tiles), max(1, 2 * units // tiles))
                           if s <= max(1, kb // 4) and tiles * s <= 4 * units})
            for s in cand:
                os.environ["PBX_TC_CONFIG"] = cfg
                h.reload_env()
                h.set_split_k(s if s > 1 else 1)
                t = timed()
                print(json.dumps(dict(m=m, n=n, k=k, dtype=args.dtype, cfg=cfg, slices=h.last_split_k, presplit=h.last_presplit,
                                      us=round(t * 1e3, 2), tflops=round(2.0 * m * n * k / t / 1e9, 1))), flush=True)
        os.environ.pop("PBX_TC_CONFIG", None)
        h.reload_env()
        h.set_split_k(0)
        del a, b, c
    h.close()


if __name__ == "__main__":
    main()
