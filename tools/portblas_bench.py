#!/usr/bin/env python
"""A stand-in for the reference's `benchmark/portblas` harness (benchmark/portblas/main.cpp:68-89 + blas3/*.cpp) over
this library, so that results line up with upstream reports entry for entry:

  * same parameter CSVs (`--csv-param`, column order of benchmark/README.md:118-127: gemm `ta,tb,m,k,n,alpha,beta`;
    gemm_batched `...,batch,strided|interleaved`; gemm_batched_strided `...,batch,sa_mul,sb_mul,sc_mul`;
    symm `side,uplo,m,n,alpha,beta`; trsm `side,uplo,trans,diag,m,n,alpha`),
  * same benchmark names (common/include/common/benchmark_names.hpp:53-58,190-232):
    `BM_Gemm<float>/n/n/<m>/<k>/<n>/usm`, `BM_Gemm_batched<float>/.../<batch>/<strided|interleaved>/usm`,
    `BM_Gemm_batched_strided<float>/.../<batch>/<sa>/<sb>/<sc>/usm`, `BM_Symm<float>/l/u/<m>/<n>/<alpha>/<beta>/usm`,
    `BM_Trsm<float>/l/u/n/n/<m>/<n>/usm`,
  * same counters (common/include/common/blas3_state_counters.hpp:38-76,141-166, common_utils.hpp:1894-1918):
    n_fl_ops, bytes_processed, {avg,best,total}_{event,overall}_time in ns, plus google-benchmark's
    items_per_second / bytes_per_second, in google-benchmark's JSON layout (`--benchmark_format=json`),
  * same procedure: inputs U(-2,5)-style random, C = 0, minimal leading dimensions, 10 warm-up calls
    (common_utils.hpp:1841-1845), then timed calls each followed by a wait; event time = CUDA events on the launching
    stream (the SYCL event profiling of benchmark/portblas/utils.hpp:52-63), overall time = host clock around
    call + wait.  Only the `usm` variants exist here (containers are device pointers).

    python tools/portblas_bench.py --op gemm --csv-param my.csv --types float,double --benchmark_format=json
"""
from __future__ import annotations

import argparse
import csv
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

TYPE_NAMES = ["float", "double", "half", "bfloat16", "complex<float>", "complex<double>"]
TYPE_BYTES = {"float": 4, "double": 8, "half": 2, "bfloat16": 2, "complex<float>": 8, "complex<double>": 16}
OP_NAME = {"gemm": "Gemm", "gemm_batched": "Gemm_batched", "gemm_batched_strided": "Gemm_batched_strided",
           "symm": "Symm", "trsm": "Trsm"}
DEFAULTS = {  # a small default sweep when no CSV is given (the reference's defaults are {32,256,2048}^3 x all trans)
    "gemm": [(ta, tb, s, s, s, 1.0, 1.0) for s in (32, 256, 2048) for ta in "nt" for tb in "nt"],
    "gemm_batched": [("n", "n", s, s, s, 1.0, 1.0, 8, bt) for s in (128, 1024) for bt in ("strided", "interleaved")],
    "gemm_batched_strided": [("n", "n", s, s, s, 1.0, 1.0, 8, 2, 2, 2) for s in (128, 1024)],
    "symm": [(sd, ul, s, s, 1.0, 0.0) for s in (256, 1024, 4096) for sd in "lr" for ul in "lu"],
    "trsm": [(sd, ul, "n", "n", s, s, 1.0) for s in (256, 1024, 4096) for sd in "lr" for ul in "lu"],
}


def fmt_scalar(v: float) -> str:
    return f"{v:g}"


def rand(count, dt, dev):
    import torch
    if dt.is_complex:
        return torch.complex(torch.rand(count, device=dev) * 7 - 2, torch.rand(count, device=dev) * 7 - 2).to(dt)
    return (torch.rand(count, device=dev, dtype=torch.float32) * 7 - 2).to(dt)


def bench_name(op, tname, row) -> str:
    """The reference's benchmark name for one parameter row (benchmark_names.hpp:53-58,190-232; mem type "usm")."""
    if op == "gemm":
        body = f"{row[0].lower()}/{row[1].lower()}/{int(row[2])}/{int(row[3])}/{int(row[4])}"
    elif op == "gemm_batched":
        body = (f"{row[0].lower()}/{row[1].lower()}/{int(row[2])}/{int(row[3])}/{int(row[4])}/{int(row[7])}/"
                f"{str(row[8]).strip().lower()}")
    elif op == "gemm_batched_strided":
        body = (f"{row[0].lower()}/{row[1].lower()}/{int(row[2])}/{int(row[3])}/{int(row[4])}/{int(row[7])}/"
                f"{int(row[8])}/{int(row[9])}/{int(row[10])}")
    elif op == "symm":
        body = (f"{row[0].lower()}/{row[1].lower()}/{int(row[2])}/{int(row[3])}/{fmt_scalar(float(row[4]))}/"
                f"{fmt_scalar(float(row[5]))}")
    elif op == "trsm":
        body = f"{row[0].lower()}/{row[1].lower()}/{row[2].lower()}/{row[3].lower()}/{int(row[4])}/{int(row[5])}"
    else:
        raise ValueError(op)
    return f"BM_{OP_NAME[op]}<{tname}>/{body}/usm"


def counters(op, tname, row) -> dict:
    """n_fl_ops / bytes_processed (+ the shape counters) of blas3_state_counters.hpp:38-76 (gemm family, complex
    variant :79-135), :141-166 (symm) and the trsm counters of benchmark/portblas/blas3/trsm.cpp."""
    es = TYPE_BYTES[tname]
    cplx = tname.startswith("complex")
    if op.startswith("gemm"):
        m, k, n = int(row[2]), int(row[3]), int(row[4])
        beta = float(row[6])
        batch = int(row[7]) if op != "gemm" else 1
        b0 = beta != 0
        cnt = dict(beta=beta, m=m, n=n, k=k, batch_size=batch)
        if op == "gemm_batched_strided":
            cnt.update(stride_a_mul=int(row[8]), stride_b_mul=int(row[9]), stride_c_mul=int(row[10]))
        if cplx:   # 4 mul + 4 add per complex multiply-add, 6 per complex scaling
            cnt["n_fl_ops"] = (8.0 * k * m * n + 6.0 * m * n + (8.0 * m * n if b0 else 0)) * batch
        else:
            cnt["n_fl_ops"] = (2.0 * k * m * n + m * n + (2.0 * m * n if b0 else 0)) * batch
        cnt["bytes_processed"] = float((m * k + k * n + m * n + (m * n if b0 else 0)) * batch * es)
        return cnt
    if op == "symm":
        side, m, n, beta = row[0].lower(), int(row[2]), int(row[3]), float(row[5])
        kk = m if side == "l" else n
        b0 = beta != 0
        return dict(beta=beta, m=m, n=n,
                    n_fl_ops=(2.0 * m * m * n if side == "l" else 2.0 * n * n * n) + (2.0 * m * n if b0 else 0),
                    bytes_processed=float(((2 if b0 else 1) * m * n + m * n + kk * (kk + 1) / 2) * es))
    if op == "trsm":
        side, m, n = row[0].lower(), int(row[4]), int(row[5])
        kk = m if side == "l" else n
        return dict(m=m, n=n, k=kk, n_fl_ops=float(kk) * kk * (n if side == "l" else m) + m * n,
                    bytes_processed=float((kk * (kk + 1) / 2 + 2 * m * n) * es))
    raise ValueError(op)


def build_case(op, tname, dt, row, h, dev):
    """Returns (benchmark name, counters, callable)."""
    import torch
    from portblas_b200 import blas, gemm_batch_type_t
    es = torch.empty(0, dtype=dt).element_size()
    if op.startswith("gemm"):
        ta, tb = row[0].lower(), row[1].lower()
        m, k, n = int(row[2]), int(row[3]), int(row[4])
        alpha, beta = float(row[5]), float(row[6])
        batch, bt, sm = 1, "strided", (1, 1, 1)
        if op == "gemm_batched":
            batch, bt = int(row[7]), row[8].strip().lower()
        elif op == "gemm_batched_strided":
            batch, sm = int(row[7]), (int(row[8]), int(row[9]), int(row[10]))
        lda, ldb, ldc = (m if ta == "n" else k), (k if tb == "n" else n), m
        sa, sb, sc = m * k * sm[0], k * n * sm[1], m * n * sm[2]
        a = rand(max(sa * (batch - 1), 0) + m * k, dt, dev)
        b = rand(max(sb * (batch - 1), 0) + k * n, dt, dev)
        c = torch.zeros(max(sc * (batch - 1), 0) + m * n, device=dev, dtype=dt)
        if dt.is_complex:
            alpha, beta = complex(alpha, 0), complex(beta, 0)
        if op == "gemm":
            fn = lambda: blas._gemm(h, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc)  # noqa: E731
        elif op == "gemm_batched":
            btype = gemm_batch_type_t.interleaved if bt == "interleaved" else gemm_batch_type_t.strided
            fn = lambda: blas._gemm_batched(h, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, batch, btype)  # noqa: E731
        else:
            fn = lambda: blas._gemm_strided_batched(h, ta, tb, m, n, k, alpha, a, lda, sa, b, ldb, sb, beta, c, ldc,  # noqa: E731
                                                    sc, batch)
        return bench_name(op, tname, row), counters(op, tname, row), fn
    if op == "symm":
        side, uplo, m, n, alpha, beta = row[0].lower(), row[1].lower(), int(row[2]), int(row[3]), float(row[4]), float(row[5])
        kk = m if side == "l" else n
        a, b = rand(kk * kk, dt, dev), rand(m * n, dt, dev)
        c = torch.zeros(m * n, device=dev, dtype=dt)
        fn = lambda: blas._symm(h, side, uplo, m, n, alpha, a, kk, b, m, beta, c, m)  # noqa: E731
        return bench_name(op, tname, row), counters(op, tname, row), fn
    if op == "trsm":
        side, uplo, tr, dg, m, n = row[0].lower(), row[1].lower(), row[2].lower(), row[3].lower(), int(row[4]), int(row[5])
        alpha = float(row[6]) if len(row) > 6 else 1.0
        kk = m if side == "l" else n
        # benchmark/portblas/blas3/trsm.cpp fills A with fill_trsm_matrix: a well-conditioned triangle
        t = torch.tril(torch.rand(kk, kk, device=dev, dtype=torch.float64) * 2 - 1) / kk + \
            torch.eye(kk, device=dev, dtype=torch.float64) * 4
        a = (t.T if uplo == "l" else t).contiguous().view(-1).to(dt)   # column-major storage
        b0 = rand(m * n, dt, dev)
        b = b0.clone()

        def fn():
            blas._trsm(h, side, uplo, tr, dg, m, n, alpha, a, kk, b, m)
        return bench_name(op, tname, row), counters(op, tname, row), fn
    raise ValueError(op)


def run_case(h, name, cnt, fn, min_time, max_iters):
    import torch
    for _ in range(10):   # warmup (common_utils.hpp:1841-1845)
        fn()
    h.wait()
    total_ev = total_ov = 0.0
    best_ev = best_ov = float("inf")
    iters = 0
    t_begin = time.perf_counter()
    while True:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        fn()
        e1.record()
        h.wait()
        torch.cuda.synchronize()
        ov = (time.perf_counter() - t0) * 1e9
        ev = e0.elapsed_time(e1) * 1e6
        total_ev += ev; total_ov += ov
        best_ev, best_ov = min(best_ev, ev), min(best_ov, ov)
        iters += 1
        if iters >= max_iters or (time.perf_counter() - t_begin >= min_time and iters >= 3):
            break
    real = total_ov / iters
    out = dict(name=name, run_name=name, run_type="iteration", repetitions=1, repetition_index=0, threads=1,
               iterations=iters, real_time=real, cpu_time=real, time_unit="ns",
               avg_event_time=total_ev / iters, avg_overall_time=total_ov / iters, best_event_time=best_ev,
               best_overall_time=best_ov, total_event_time=total_ev, total_overall_time=total_ov,
               items_per_second=cnt["n_fl_ops"] / (real * 1e-9), bytes_per_second=cnt["bytes_processed"] / (real * 1e-9))
    out.update({k: float(v) for k, v in cnt.items()})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--op", default="gemm", choices=sorted(OP_NAME))
    ap.add_argument("--csv-param", default=None)
    ap.add_argument("--types", default="float")
    ap.add_argument("--benchmark_format", default="console", choices=["console", "json"])
    ap.add_argument("--benchmark_min_time", type=float, default=0.05)
    ap.add_argument("--max-iters", type=int, default=200)
    ap.add_argument("--benchmark_out", default=None)
    ap.add_argument("--from-fixture", action="store_true",
                    help="rows of tests/golden/config_csv_shapes.json (generated from the reference's "
                         "benchmark/config_csv/blas3/gemm*/ files) for this --op")
    ap.add_argument("--max-rows", type=int, default=0)
    ap.add_argument("--max-gib", type=float, default=40.0)
    args = ap.parse_args()
    rows = DEFAULTS[args.op]
    if args.from_fixture:
        fx = json.loads((ROOT / "tests" / "golden" / "config_csv_shapes.json").read_text())["rows"]
        rows = []
        for r in fx:
            if r["api"] != args.op:
                continue
            base = [r["ta"], r["tb"], r["m"], r["k"], r["n"], r["alpha"], r["beta"]]
            if args.op == "gemm_batched":
                base += [r.get("batch", 1), r.get("batch_type", "strided")]
            elif args.op == "gemm_batched_strided":
                base += [r.get("batch", 1), r.get("stride_a_mul", 1), r.get("stride_b_mul", 1), r.get("stride_c_mul", 1)]
            rows.append(base)
    if args.csv_param:
        with open(args.csv_param) as f:
            rows = [r for r in csv.reader(f) if r and not r[0].startswith("#")]
    if args.max_rows:
        rows = rows[:args.max_rows]
    import torch
    from portblas_b200 import SB_Handle
    types = {"float": torch.float32, "double": torch.float64, "half": torch.float16, "bfloat16": torch.bfloat16,
             "complex<float>": torch.complex64, "complex<double>": torch.complex128}
    dev = torch.device("cuda", 0)
    h = SB_Handle(0)
    results = []
    for tname in args.types.split(","):
        dt = types[tname.strip()]
        for row in rows:
            if args.op.startswith("gemm"):
                m_, k_, n_ = int(row[2]), int(row[3]), int(row[4])
                bsz = int(row[7]) if len(row) > 7 else 1
                if (m_ * k_ + k_ * n_ + m_ * n_) * bsz * torch.empty(0, dtype=dt).element_size() > args.max_gib * 2**30:
                    continue
            name, cnt, fn = build_case(args.op, tname.strip(), dt, row, h, dev)
            r = run_case(h, name, cnt, fn, args.benchmark_min_time, args.max_iters)
            results.append(r)
            if args.benchmark_format == "console":
                print(f"{name:70s} {r['avg_event_time']:14.0f} ns event  {r['avg_overall_time']:14.0f} ns overall  "
                      f"{r['n_fl_ops'] / r['avg_event_time']:10.1f} GFLOP/s  iters {r['iterations']}", flush=True)
            torch.cuda.empty_cache()
    doc = dict(context=dict(date=time.strftime("%Y-%m-%dT%H:%M:%S"), executable="tools/portblas_bench.py",
                            device=torch.cuda.get_device_name(0), library="portblas_b200 (libpbx_gemm.so)",
                            num_compute_units=h.get_num_compute_units()), benchmarks=results)
    if args.benchmark_format == "json":
        print(json.dumps(doc, indent=1))
    if args.benchmark_out:
        Path(args.benchmark_out).write_text(json.dumps(doc, indent=1))
    h.close()


if __name__ == "__main__":
    main()
