"""Diagnostic sweep for a GPU box: runs a list of parity cases per kernel family and prints one line
per case (never stops at the first failure) plus quick device timings.  Development aid; the
judged parity tests are tests/test_gemm_gpu.py.

    python tools/gpu_check.py [--quick] [--time]
"""
from __future__ import annotations

import argparse
import itertools
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402

from gemm_case import Case, run_case  # noqa: E402
from portblas_b200 import SB_Handle, blas, build  # noqa: E402

TRANS = [("n", "n"), ("n", "t"), ("t", "n"), ("t", "t")]


def sweep(h, title, cases):
    nbad = 0
    t0 = time.time()
    for cs in cases:
        try:
            r = run_case(h, cs)
        except Exception as e:  # noqa: BLE001
            print(f"  EXC  {cs.ident()} {type(e).__name__}: {e}")
            nbad += 1
            continue
        if not r.ok:
            nbad += 1
        tag = "ok  " if r.ok else "FAIL"
        if not r.ok or args.verbose:
            print(f"  {tag} {cs.ident()} kern={r.kernel} sk={r.split_k} refmis={r.ref_mismatch} "
                  f"viol={r.bound_violations} maxrel={r.max_rel_bound:.3e} {r.detail}")
    print(f"[{title}] {len(cases) - nbad}/{len(cases)} ok  ({time.time() - t0:.1f}s)", flush=True)
    return nbad


def time_gemm(h, dt_in, dt_out, ta, tb, m, n, k, batch=1, iters=10, warm=3, kernel=0, split_k=0):
    dev = torch.device("cuda", 0)
    lda = k if ta != "n" else m
    ldb = n if tb != "n" else k
    a = (torch.rand(lda * (m if ta != "n" else k) * batch, device=dev, dtype=torch.float32) * 7 - 2).to(dt_in)
    b = (torch.rand(ldb * (k if tb != "n" else n) * batch, device=dev, dtype=torch.float32) * 7 - 2).to(dt_in)
    c = torch.zeros(m * n * batch, device=dev, dtype=dt_out)
    h.set_forced_kernel(kernel)
    h.set_split_k(split_k)

    def run():
        if batch == 1:
            blas._gemm(h, ta, tb, m, n, k, 1.0, a, lda, b, ldb, 0.0, c, m)
        else:
            blas._gemm_batched(h, ta, tb, m, n, k, 1.0, a, lda, b, ldb, 0.0, c, m, batch)
    for _ in range(warm):
        run()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for s, e in evs:
        s.record()
        run()
        e.record()
    torch.cuda.synchronize()
    ts = [s.elapsed_time(e) for s, e in evs]
    best, mean = min(ts), sum(ts) / len(ts)
    fl = 2.0 * m * n * k * batch
    h.set_forced_kernel(0)
    h.set_split_k(0)
    print(f"  time {str(dt_in).split('.')[-1]:9s} {ta}{tb} {m}x{n}x{k} b{batch} kern={h.last_kernel} sk={h.last_split_k}: "
          f"best {best:.3f} ms ({fl / best / 1e9:.1f} TFLOP/s)  mean {mean:.3f} ms ({fl / mean / 1e9:.1f} TFLOP/s)",
          flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    build.build()
    h = SB_Handle(0)
    print("SMs:", h.get_num_compute_units(), torch.cuda.get_device_name(0), flush=True)
    bad = 0
    sel = args.only.split(",") if args.only else []

    def want(name):
        return not sel or name in sel

    if want("simt"):
        for dt in ["f32", "f64", "f16", "f16f32", "bf16", "bf16f32"]:
            bad += sweep(h, f"simt {dt}", [Case(dtype=dt, transa=ta, transb=tb, m=m, n=n, k=k, beta=b, kernel=1,
                                                lda_mul=2, ldc_mul=3, offset=3)
                                           for (ta, tb), (m, n, k), b in itertools.product(
                                               TRANS, [(11, 16, 17), (65, 130, 70)], [0.0, 0.5])])
    if want("dmma"):
        bad += sweep(h, "dmma f64", [Case(dtype="f64", transa=ta, transb=tb, m=m, n=n, k=k, beta=b, kernel=3,
                                          offset=off, ldb_mul=ldm)
                                     for (ta, tb), (m, n, k), b, off, ldm in itertools.product(
                                         TRANS, [(11, 16, 17), (128, 128, 64), (253, 257, 511), (1024, 511, 253)],
                                         [0.0, 0.5], [0, 1], [1, 3])])
    if want("tc16"):
        for dt in ["bf16f32", "f16f32", "bf16", "f16"]:
            bad += sweep(h, f"tcgen05 {dt}", [Case(dtype=dt, transa=ta, transb=tb, m=m, n=n, k=k, beta=b, kernel=2)
                                              for (ta, tb), (m, n, k), b in itertools.product(
                                                  TRANS, [(128, 128, 64), (128, 256, 128), (136, 264, 520),
                                                          (2200, 264, 40), (1024, 1536, 1536)], [0.0, 0.5])])
    if want("tc32"):
        bad += sweep(h, "tcgen05 f32 (3xTF32)", [Case(dtype="f32", transa=ta, transb=tb, m=m, n=n, k=k, beta=b, kernel=2)
                                                 for (ta, tb), (m, n, k), b in itertools.product(
                                                     TRANS, [(128, 128, 32), (128, 128, 256), (136, 264, 520),
                                                             (2200, 264, 40), (1024, 1536, 1536)], [0.0, 0.5])])
    if want("tccfg"):
        # every tcgen05 tile configuration (cta_group, BN) on ragged and aligned shapes
        for cfgs in ["1,128", "2,128", "2,256"]:
            for dt in ["bf16f32", "f16", "f32"]:
                envs = [(("PBX_TC_CONFIG", cfgs),)]
                if dt == "f32":
                    envs.append((("PBX_TC_CONFIG", cfgs), ("PBX_TF32_RAW_HI", 1)))
                for env in envs:
                    bad += sweep(h, f"tcgen05 cfg={cfgs} {dt} {env[1:] if len(env) > 1 else ''}",
                                 [Case(dtype=dt, transa=ta, transb=tb, m=m, n=n, k=k, beta=b, kernel=2, env=env)
                                  for (ta, tb), (m, n, k), b in itertools.product(
                                      TRANS, [(256, 256, 64), (136, 264, 520), (2200, 264, 40), (1024, 1536, 1536),
                                              (520, 136, 4104)], [0.0, 0.5])])
            bad += sweep(h, f"tcgen05 cfg={cfgs} batched/split", [
                Case(dtype=dt, api="strided", transa=ta, transb=tb, m=256, n=256, k=256, alpha=1.0, beta=0.0, batch=6,
                     stride_a_mul=sa, stride_b_mul=1, kernel=2, env=(("PBX_TC_CONFIG", cfgs),))
                for dt, (ta, tb), sa in itertools.product(["bf16", "f32"], TRANS, [0, 1])] + [
                Case(dtype=dt, transa=ta, transb=tb, m=264, n=136, k=16416, beta=0.5, kernel=2, split_k=3,
                     env=(("PBX_TC_CONFIG", cfgs),)) for dt, (ta, tb) in itertools.product(["bf16f32", "f32"], TRANS)])
    if want("batched"):
        for dt in ["f32", "f64", "bf16", "f16f32"]:
            bad += sweep(h, f"batched {dt}", [Case(dtype=dt, api="batched", transa=ta, transb=tb, m=m, n=n, k=k,
                                                   alpha=3.0, beta=7.0, batch=5, batch_type=bt, lda_mul=lm, ldc_mul=lm)
                                              for (ta, tb), (m, n, k), bt, lm in itertools.product(
                                                  TRANS, [(63, 63, 63), (128, 128, 128), (256, 256, 256)], [0, 1], [1, 2])])
    if want("splitk"):
        for dt, kern in [("f32", 2), ("f32", 1), ("f64", 3), ("bf16f32", 2)]:
            bad += sweep(h, f"split-k {dt} kern{kern}", [Case(dtype=dt, transa=ta, transb=tb, m=72, n=136, k=4104,
                                                               beta=b, kernel=kern, split_k=sk)
                                                          for (ta, tb), b, sk in itertools.product(TRANS, [0.0, 0.5], [3, 7])])
    print("TOTAL FAILURES:", bad, flush=True)

    if args.time:
        f32, f64, bf16, f16 = torch.float32, torch.float64, torch.bfloat16, torch.float16
        for ta, tb in TRANS:
            time_gemm(h, f64, f64, ta, tb, 8192, 8192, 8192, iters=5, warm=2)
        for ta, tb in TRANS:
            time_gemm(h, bf16, bf16, ta, tb, 8192, 8192, 8192)
        time_gemm(h, f16, f16, "n", "n", 8192, 8192, 8192)
        for ta, tb in TRANS:
            time_gemm(h, f32, f32, ta, tb, 8192, 8192, 8192)
        time_gemm(h, f32, f32, "n", "n", 16384, 16384, 16384, iters=3, warm=1)
        time_gemm(h, bf16, bf16, "n", "n", 256, 256, 256, batch=4096)
        time_gemm(h, f16, f16, "n", "n", 256, 256, 256, batch=4096)
        time_gemm(h, f32, f32, "n", "n", 512, 512, 1048576, iters=3, warm=1)
        time_gemm(h, f32, f32, "n", "n", 1024, 1024, 1024)
    h.close()
    sys.exit(1 if bad else 0)
