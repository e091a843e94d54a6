"""Replays the reference benchmark's GEMM shape sweeps (tests/golden/config_csv_shapes.json, generated from
benchmark/config_csv/blas3/gemm*/ by tests/golden/make_config_shapes.py) through the public API on one B200:
parity on sampled C entries against an fp64 recomputation, device time, TFLOP/s, GB/s and the roof that binds
(min(tensor peak, AI x HBM)).  One JSON line per row.

    python tools/csv_sweep.py [--dtype f32] [--api gemm,gemm_batched,gemm_batched_strided] [--max-rows N] > sweep.jsonl
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from portblas_b200 import SB_Handle, blas, gemm_batch_type_t  # noqa: E402

TDT = {"f32": torch.float32, "f64": torch.float64, "f16": torch.float16, "bf16": torch.bfloat16}
TOL = {"f32": 1e-5, "f64": 1e-12, "f16": 2e-3, "bf16": 1.6e-2}
PEAK_TF = {"f32": 265.0, "f64": 40.0, "f16": 1590.0, "bf16": 1590.0}
HBM_GBS = 6650.0


def rand(count, dt, dev, gen):
    out = torch.empty(count, device=dev, dtype=dt)
    chunk = 1 << 26
    for s in range(0, count, chunk):
        e = min(count, s + chunk)
        out[s:e] = (torch.rand(e - s, device=dev, dtype=torch.float32, generator=gen) * 7.0 - 2.0).to(dt)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--api", default="gemm,gemm_batched,gemm_batched_strided")
    ap.add_argument("--max-rows", type=int, default=0)
    ap.add_argument("--max-gib", type=float, default=60.0)
    ap.add_argument("--graph", action="store_true",
                    help="time CUDA-graph replays of the call (device time without the Python/ctypes launch cost, "
                         "which is ~20 us per call and hides every small shape)")
    args = ap.parse_args()
    rows = json.loads((ROOT / "tests" / "golden" / "config_csv_shapes.json").read_text())["rows"]
    rows = [r for r in rows if r["api"] in args.api.split(",")]
    if args.max_rows:
        rows = rows[:args.max_rows]
    dev = torch.device("cuda", 0)
    h = SB_Handle(0)
    side = torch.cuda.Stream(device=dev)
    dt = TDT[args.dtype]
    es = torch.empty(0, dtype=dt).element_size()
    gen = torch.Generator(device=dev)
    gen.manual_seed(12345)
    nbad = 0
    for r in rows:
        ta, tb, m, n, k = r["ta"], r["tb"], r["m"], r["n"], r["k"]
        batch = r.get("batch", 1)
        ilv = r.get("batch_type") == "interleaved"
        a_rows, a_cols = (k, m) if ta != "n" else (m, k)
        b_rows, b_cols = (n, k) if tb != "n" else (k, n)
        lda, ldb, ldc = a_rows, b_rows, m
        sa = lda * a_cols * r.get("stride_a_mul", 1)
        sb = ldb * b_cols * r.get("stride_b_mul", 1)
        sc = ldc * n * r.get("stride_c_mul", 1)
        na = max(sa * (batch - 1), 0) + lda * a_cols
        nb = max(sb * (batch - 1), 0) + ldb * b_cols
        nc = max(sc * (batch - 1), 0) + ldc * n
        gib = (na + nb + nc) * es / 2**30
        base = dict(api=r["api"], dtype=args.dtype, ta=ta, tb=tb, m=m, n=n, k=k, batch=batch,
                    batch_type=r.get("batch_type", "strided"), alpha=r["alpha"], beta=r["beta"])
        if gib > args.max_gib:
            print(json.dumps(dict(base, skipped=f"{gib:.1f} GiB")), flush=True)
            continue
        a = rand(na, dt, dev, gen)
        b = rand(nb, dt, dev, gen)
        if args.dtype == "f16" and k * 2.25 > 2.0e4:
            # U(-2,5) products average 2.25: a K of tens of thousands overflows fp16 storage of C (65504) -- a property
            # of the data type, not of the kernel -- so deep-K rows run on inputs scaled into range
            sc_in = (2.0e4 / (k * 2.25)) ** 0.5
            a.mul_(sc_in)
            b.mul_(sc_in)
        c0 = rand(nc, dt, dev, gen) if r["beta"] != 0 else torch.zeros(nc, device=dev, dtype=dt)
        c = c0.clone()
        if ilv:  # strided (batch, cols, ld) -> interleaved (cols, ld, batch): element (r,c,b) at (c*ld + r)*batch + b
            a_i = a.view(batch, a_cols, lda).permute(1, 2, 0).contiguous().view(-1)
            b_i = b.view(batch, b_cols, ldb).permute(1, 2, 0).contiguous().view(-1)
            c_i = c.view(batch, n, ldc).permute(1, 2, 0).contiguous().view(-1)

        def run():
            if r["api"] == "gemm":
                blas._gemm(h, ta, tb, m, n, k, r["alpha"], a, lda, b, ldb, r["beta"], c, ldc)
            elif r["api"] == "gemm_batched":
                if ilv:
                    blas._gemm_batched(h, ta, tb, m, n, k, r["alpha"], a_i, lda, b_i, ldb, r["beta"], c_i, ldc, batch,
                                       gemm_batch_type_t.interleaved)
                else:
                    blas._gemm_batched(h, ta, tb, m, n, k, r["alpha"], a, lda, b, ldb, r["beta"], c, ldc, batch)
            else:
                blas._gemm_strided_batched(h, ta, tb, m, n, k, r["alpha"], a, lda, sa, b, ldb, sb, r["beta"], c, ldc, sc,
                                           batch)
        run()
        torch.cuda.synchronize()
        kernel, split_k = h.last_kernel, h.last_split_k
        # ---- parity on a sample: 48 rows x 48 cols of up to 3 batch entries, fp64 recomputation ----
        got = c_i.view(n, ldc, batch).permute(2, 0, 1).contiguous().view(-1) if ilv else c
        g2 = torch.Generator(device="cpu"); g2.manual_seed(m * 31 + n * 17 + k)
        ri = torch.randint(0, m, (min(m, 48),), generator=g2).to(dev)
        ci = torch.randint(0, n, (min(n, 48),), generator=g2).to(dev)
        worst = 0.0
        for bi in sorted(set([0, batch // 2, batch - 1])):
            A = a[bi * sa: bi * sa + lda * a_cols].view(a_cols, lda).t().to(torch.float64)   # stored rows x cols
            B = b[bi * sb: bi * sb + ldb * b_cols].view(b_cols, ldb).t().to(torch.float64)
            opA = (A.t() if ta != "n" else A)[ri, :]
            opB = (B.t() if tb != "n" else B)[:, ci]
            C0 = c0[bi * sc: bi * sc + ldc * n].view(n, ldc).t().to(torch.float64)[ri][:, ci]
            G = got[bi * sc: bi * sc + ldc * n].view(n, ldc).t().to(torch.float64)[ri][:, ci]
            want = r["alpha"] * (opA @ opB) + r["beta"] * C0
            bound = abs(r["alpha"]) * (opA.abs() @ opB.abs()) + abs(r["beta"]) * C0.abs()
            worst = max(worst, float(((G - want).abs() / bound.clamp_min(1e-300)).max()))
        ok = worst <= TOL[args.dtype]
        nbad += (not ok)
        # ---- timing ----
        flops = 2.0 * m * n * k * batch
        iters = 3 if flops > 2e12 else (10 if flops > 5e10 else 30)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        if args.graph:
            reps = 1 if flops > 5e10 else 8
            h.set_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                for _ in range(reps):
                    run()
            h.set_stream(torch.cuda.current_stream(dev))
            graph.replay()
            torch.cuda.synchronize()
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
            evs[0].record()
            for i in range(iters):
                graph.replay()
                evs[i + 1].record()
            torch.cuda.synchronize()
            ts = sorted(evs[i].elapsed_time(evs[i + 1]) / reps for i in range(iters))
            del graph
            # the same call launched eagerly, back to back (what a caller of blas::_gemm in a loop sees per call: the
            # larger of the device time and the host-side cost of the call -- tensor-map encodes, selector, ctypes)
            e_it = iters * (8 if flops <= 5e10 else 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(e_it):
                run()
            e1.record()
            torch.cuda.synchronize()
            eager_ms = e0.elapsed_time(e1) / e_it
        else:
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
            evs[0].record()
            for i in range(iters):
                run()
                evs[i + 1].record()
            torch.cuda.synchronize()
            ts = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(iters))
        ms = ts[len(ts) // 2]
        byts = (m * k + k * n + m * n * (2 if r["beta"] != 0 else 1)) * es * batch
        ai = flops / byts
        roof_tf = min(PEAK_TF[args.dtype], ai * HBM_GBS / 1e3)
        tf = flops / ms / 1e9
        print(json.dumps(dict(base, kernel=kernel, split_k=split_k, ms=round(ms, 4), tflops=round(tf, 2),
                              gbs=round(byts / ms / 1e6, 1), ai=round(ai, 1),
                              bound="hbm" if ai * HBM_GBS / 1e3 < PEAK_TF[args.dtype] else "tensor",
                              frac_of_roof=round(tf / roof_tf, 3), timing="graph" if args.graph else "launch",
                              eager_ms=round(eager_ms, 4) if args.graph else None, presplit=h.last_presplit, repack=h.last_repack, max_rel_err=float(f"{worst:.3e}"), ok=ok)), flush=True)
        del a, b, c, c0
        if ilv:
            del a_i, b_i, c_i
        torch.cuda.empty_cache()
    print(json.dumps(dict(summary=True, rows=len(rows), parity_failures=nbad)), flush=True)
    h.close()


if __name__ == "__main__":
    main()
