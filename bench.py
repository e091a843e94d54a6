#!/usr/bin/env python
"""bench.py -- headline benchmark of the GEMM path (BASELINE.json: "GEMM TFLOP/s and % of
per-dtype tensor peak").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A "step" is one pass of the hot path (one blas::_gemm / _gemm_strided_batched call) over one
synthetic batch of inputs.  Default workload at N=1 is BASELINE configs[1], DGEMM 8192^3 NN on one
B200.  For N>1 the problem is sharded with no data-path collective (portblas_b200/sharding.py):
every rank owns one M-block (or batch range) of the SAME per-GPU size, i.e. weak scaling of a
(8192*N) x 8192 x 8192 GEMM with B replicated.  Timing: CUDA events on the launching stream,
barrier + synchronize on both sides, max over ranks.  One JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# name -> dict(dtype key, in/out torch dtype names, m, n, k, batch, transa, transb, alpha, beta)
WORKLOADS = {
    # BASELINE configs[1]
    "dgemm8192": dict(dt="f64", m=8192, n=8192, k=8192, batch=1, ta="n", tb="n", alpha=1.0, beta=0.0,
                      desc="DGEMM 8192x8192x8192 NN alpha=1 beta=0 (BASELINE configs[1]); N>1: one such M-block per GPU"),
    # BASELINE configs[2] (per-GPU M-block of the 16384^3 problem when sharded 1/2/4/8 ways is 16384/N rows)
    "sgemm16384": dict(dt="f32", m=16384, n=16384, k=16384, batch=1, ta="n", tb="n", alpha=1.0, beta=0.0, strong=True,
                       desc="SGEMM 16384^3 NN fp32 via 3xTF32 (BASELINE configs[2]); N>1: M-block shards of the SAME problem"),
    "sgemm8192": dict(dt="f32", m=8192, n=8192, k=8192, batch=1, ta="n", tb="n", alpha=1.0, beta=0.0,
                      desc="SGEMM 8192^3 NN fp32 via 3xTF32"),
    # BASELINE configs[3]
    "hgemm_batched": dict(dt="f16", m=256, n=256, k=256, batch=4096, ta="n", tb="n", alpha=1.0, beta=0.0, strong=True,
                          desc="strided-batched HGEMM, batch 4096 of 256^3 (BASELINE configs[3]); N>1: batch shards of the SAME problem"),
    "bf16gemm_batched": dict(dt="bf16", m=256, n=256, k=256, batch=4096, ta="n", tb="n", alpha=1.0, beta=0.0, strong=True,
                             desc="strided-batched BF16 GEMM, batch 4096 of 256^3 (BASELINE configs[3]); N>1: batch shards of the SAME problem"),
    # BASELINE configs[4]
    "sgemm_splitk": dict(dt="f32", m=512, n=512, k=1048576, batch=1, ta="n", tb="n", alpha=1.0, beta=0.0,
                         desc="tall-skinny SGEMM M=N=512 K=1048576, split-K (BASELINE configs[4])"),
    # BASELINE configs[0] shape (the reference's CPU-runnable case)
    "sgemm1024": dict(dt="f32", m=1024, n=1024, k=1024, batch=1, ta="n", tb="n", alpha=1.0, beta=0.0,
                      desc="SGEMM 1024^3 NN (BASELINE configs[0] shape)"),
    "bf16gemm8192": dict(dt="bf16", m=8192, n=8192, k=8192, batch=1, ta="n", tb="n", alpha=1.0, beta=0.0,
                         desc="BF16 GEMM 8192^3 NN"),
}
ES_IN = {"f64": 8, "f32": 4, "f16": 2, "bf16": 2}
NOMINAL_FP64_TFLOPS = 40.0      # B200 datasheet FP64 (tensor == vector); MEASURED_PEAKS.json has no fp64 entry
MEASURED_FP64_PIPE_TFLOPS = 36.98  # register-resident DMMA.8x8x4 loop on this pool's B200 (profiles/r01/dmma_rate.txt)
NOMINAL_TF32_TFLOPS = 1100.0    # dense; the 3xTF32 roof is a third of the tf32 rate


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def algorithmic(w):
    flops = 2.0 * w["m"] * w["n"] * w["k"] * w["batch"]
    es = ES_IN[w["dt"]]
    byts = (w["m"] * w["k"] + w["k"] * w["n"] + w["m"] * w["n"] * (2 if w["beta"] != 0 else 1)) * es * w["batch"]
    return flops, byts


def roofline_for(w, avg_ms, traffic):
    """Roof that binds the dominant kernel: tensor pipe for deep contractions, HBM when the
    arithmetic intensity is below the ridge."""
    pk = _peaks()
    flops, byts = algorithmic(w)
    ai = flops / byts
    if w["dt"] == "f64":
        tensor_peak = NOMINAL_FP64_TFLOPS
        src = ("nominal fp64 40 TF (no fp64 entry in MEASURED_PEAKS.json); measured DMMA pipe ceiling "
               f"{MEASURED_FP64_PIPE_TFLOPS} TF (tools/micro/dmma_rate.cu, profiles/r01/dmma_rate.txt)")
    elif w["dt"] == "f32" and os.environ.get("SB_ENABLE_JOINT_MATRIX", "")[:1] == "1":
        tensor_peak = pk["bf16"] / 2.0
        src = f"single-tf32 roof (SB_ENABLE_JOINT_MATRIX=1: reduced-precision fragments) = bf16 {pk['source']} / 2"
    elif w["dt"] == "f32":
        tensor_peak = pk["bf16"] / 2.0 / 3.0
        src = f"3xTF32 roof = bf16 {pk['source']} / 2 (tf32 rate) / 3 (three MMAs per product)"
    else:
        tensor_peak, src = pk["bf16"], f"bf16 {pk['source']}"
    hbm_tf = ai * pk["hbm"] / 1e3
    if hbm_tf < tensor_peak:
        ach = byts / (avg_ms * 1e-3) / 1e9
        return dict(bound="hbm", achieved=round(ach, 1), peak=pk["hbm"], unit="GB/s", frac=round(ach / pk["hbm"], 4),
                    traffic=traffic, peak_source=f"HBM copy {pk['source']}", ai_flop_per_byte=round(ai, 1))
    ach = flops / (avg_ms * 1e-3) / 1e12
    out = dict(bound="tensor", achieved=round(ach, 2), peak=round(tensor_peak, 1), unit="TFLOP/s",
               frac=round(ach / tensor_peak, 4), traffic=traffic, peak_source=src, ai_flop_per_byte=round(ai, 1))
    if w["dt"] != "f64" and pk.get("bf16_sustained"):
        # a kernel that runs for tens of milliseconds back to back sits at the board's power cap: the same ratio
        # against the SUSTAINED bf16 figure of MEASURED_PEAKS.json (burst stays the headline `peak` / `frac`)
        div = 1.0 if w["dt"] != "f32" else (2.0 if os.environ.get("SB_ENABLE_JOINT_MATRIX", "")[:1] == "1" else 6.0)
        sus = pk["bf16_sustained"] / div
        out["peak_sustained"] = round(sus, 1)
        out["frac_of_sustained"] = round(ach / sus, 4)
    return out


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), power_w_max=max(pw), samples=len(sm),
                    reasons=sorted(reasons))


def _cpu_gemm():
    """The CPU GEMM both CPU legs time: the REFERENCE's own code when oracle/_ref was built (portBLAS's blas::_gemm,
    default backend, compiled from its sources over a host stand-in for the SYCL runtime: oracle/ref_host_driver.cpp;
    work-items of its barrier-free kernels run as plain loop iterations, work-groups over OpenMP threads), else the C
    port of the kernel that backend picks (oracle/gemm_oracle.c -- bit-identical results, tests/test_oracle_ref.py).
    Returns (fn(ms, n, k, a, b, c), kind, cores, label)."""
    from oracle import oracle, ref_host
    if ref_host.usable("default"):   # probed in a child process: a prebuilt library that cannot run here costs a fallback
        ref_host.set_fibers(False)

        def fn(ms, n, k, a, b, c):
            ref_host.gemm("n", "n", ms, n, k, 1.0, a, ms, b, k, 0.0, c, ms)
        return fn, "reference", ref_host.compute_units("default"), (
            "portBLAS's own blas::_gemm (default backend, compiled from the reference's sources over a host "
            "stand-in for the SYCL runtime, oracle/_ref), OpenMP over work-groups")

    def fn(ms, n, k, a, b, c):
        oracle.gemm_default_cpu(False, False, ms, n, k, 1.0, a, ms, b, k, 0.0, c, ms)
    return fn, "port", oracle.num_threads(), "the restated DEFAULT-backend kernel (oracle/gemm_oracle.c, OpenMP)"


def cpu_sample(w, target_s=12.0):
    """Bounded CPU sample of the same workload on the first ``ms`` rows of C with full N and K (see _cpu_gemm)."""
    import numpy as np
    from oracle import oracle
    npdt = np.float64 if w["dt"] == "f64" else np.float32
    n, k = min(w["n"], 8192), min(w["k"], 8192)
    rng = np.random.default_rng(12345)
    gemm, kind, cores, label = _cpu_gemm()

    def run(ms):
        a = oracle.random_uniform(rng, ms * k, npdt)
        b = oracle.random_uniform(rng, k * n, npdt)
        c = np.zeros(ms * n, dtype=npdt)
        t0 = time.perf_counter()
        gemm(ms, n, k, a, b, c)
        dt = time.perf_counter() - t0
        t1 = time.perf_counter()
        cc = a.reshape(k, ms).T @ b.reshape(n, k).T  # OpenBLAS (the reference tests' CBLAS oracle)
        dt_blas = time.perf_counter() - t1
        del cc
        return dt, dt_blas

    probe_ms = 64
    dt, _ = run(probe_ms)
    rate = 2.0 * probe_ms * n * k / dt
    ms = int(min(w["m"], max(64, (target_s * rate / (2.0 * n * k)) // 16 * 16)))
    dt, dt_blas = run(ms)
    fl = 2.0 * ms * n * k
    return dict(value=round(fl / dt / 1e12, 5), unit="TFLOP/s", cores=cores, kind=kind,
                sample=f"first {ms} rows of C x N={n} x K={k} ({fl / 1e9:.1f} GFLOP, {dt:.1f} s) with {label}",
                cblas_tflops=round(fl / dt_blas / 1e12, 5))


def run_reference(args, w, rank, world):
    """--impl reference: the reference's own CPU implementation of the path with all host threads, each step a bounded
    sample of the workload (see _cpu_gemm for what runs)."""
    if rank != 0:
        return
    import numpy as np
    from oracle import oracle
    npdt = np.float64 if w["dt"] == "f64" else np.float32
    n, k = min(w["n"], 4096), min(w["k"], 4096)
    ms = 256
    rng = np.random.default_rng(12345)
    gemm, kind, cores, label = _cpu_gemm()
    a = oracle.random_uniform(rng, ms * k, npdt)
    b = oracle.random_uniform(rng, k * n, npdt)
    c = np.zeros(ms * n, dtype=npdt)
    for _ in range(args.warmup):
        gemm(ms, n, k, a, b, c)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        gemm(ms, n, k, a, b, c)
    dt = (time.perf_counter() - t0) / args.steps
    fl = 2.0 * ms * n * k
    val = fl / dt / 1e12
    sample = f"{ms} rows of C x N={n} x K={k} per step ({fl / 1e9:.1f} GFLOP)"
    line = dict(impl="reference", metric="gemm_tflops", value=round(val, 5), unit="TFLOP/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=round(dt * 1e3, 3), higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype=w["dt"], data="synthetic U(-2,5) seed 12345",
                config=dict(workload=w["desc"], sample=sample),
                cpu_baseline=dict(value=round(val, 5), unit="TFLOP/s", cores=cores, kind=kind, sample=sample),
                e2e=dict(value=round(val, 5), unit="TFLOP/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0,
                note=f"reference arm = {label}; portBLAS needs a SYCL compiler (absent in this image), so its kernels run "
                     "on a host executor rather than on a SYCL CPU device")
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="dgemm8192", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--gather", action="store_true", help="also time the NCCL gather of C (N>1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GEMM path has no CPU fallback")
    from portblas_b200 import SB_Handle, blas, build, sharding
    build.build()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    h = SB_Handle(local_rank)

    tdt = {"f64": torch.float64, "f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}[w["dt"]]
    m, n, k, batch = w["m"], w["n"], w["k"], w["batch"]
    lda, ldb, ldc = m, k, m
    gen = torch.Generator(device=dev)
    gen.manual_seed(12345 + rank)

    def rand(count):
        chunk = 1 << 26
        out = torch.empty(count, device=dev, dtype=tdt)
        for s in range(0, count, chunk):
            e = min(count, s + chunk)
            out[s:e] = (torch.rand(e - s, device=dev, dtype=torch.float32, generator=gen) * 7.0 - 2.0).to(tdt)
        return out

    a = rand(lda * k * batch)
    b = rand(ldb * n * batch)
    c = torch.zeros(ldc * n * batch, device=dev, dtype=tdt)

    # N > 1.  weak: every rank runs the whole per-GPU shape on its own data.  strong (BASELINE configs[2],[3]):
    # the SAME problem is cut into M-blocks / batch ranges (portblas_b200/sharding.py): a rank's shard is a
    # pointer offset into the full operands with the ORIGINAL leading dimensions -- no copy, no collective.
    strong = bool(w.get("strong")) and world > 1
    m_loc, batch_loc, a_off, b_off, c_off = m, batch, 0, 0, 0
    if strong and batch == 1:
        sh = sharding.shard_mblock(w["ta"], m, lda, world, rank, align=256)
        m_loc, a_off, c_off = sh.rows, sh.a_offset, sh.c_offset
    elif strong:
        bs = sharding.shard_batch(batch, m * k, k * n, m * n, world, rank)
        batch_loc, a_off, b_off, c_off = bs.batches, bs.a_offset, bs.b_offset, bs.c_offset
    a_v, b_v, c_v = a[a_off:], b[b_off:], c[c_off:]

    def step():
        if batch == 1:
            blas._gemm(h, w["ta"], w["tb"], m_loc, n, k, w["alpha"], a_v, lda, b_v, ldb, w["beta"], c_v, ldc)
        else:
            blas._gemm_strided_batched(h, w["ta"], w["tb"], m, n, k, w["alpha"], a_v, lda, m * k, b_v, ldb, k * n,
                                       w["beta"], c_v, ldc, m * n, batch_loc)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = h.launch_count
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    evs[0].record()
    for i in range(args.steps):
        step()
        evs[i + 1].record()
    barrier()
    launches = h.launch_count - launches0
    clocks = sampler.stop() if sampler else None
    total_ms = evs[0].elapsed_time(evs[-1])
    per_launch = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    flops, byts = algorithmic(w)
    job_flops = flops if strong else flops * world   # strong: the one problem; weak: one problem per GPU
    value = job_flops / (ms_per_step * 1e-3) / 1e12
    kernel_used, split_used = h.last_kernel, h.last_split_k

    # ---- optional C gather (the only collective of the path) ----
    gather_ms = None
    if args.gather and world > 1:
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        warm = torch.zeros(1 << 20, device=dev, dtype=tdt)      # NCCL channel set-up happens on the first collective
        dist.all_gather_into_tensor(torch.empty(world << 20, device=dev, dtype=tdt), warm)
        torch.cuda.synchronize()
        dist.barrier()
        g0.record()
        if batch == 1 and strong:
            # compact (rows x n) copy of this rank's row block, then all-gather + interleave
            c_loc = c.view(n, ldc)[:, c_off:c_off + m_loc].contiguous().view(-1)
            full = sharding.gather_c_mblocks(c_loc, m, n, world, align=256)
        elif batch == 1:
            full = sharding.gather_c_mblocks(c, m * world, n, world, align=m)
        elif strong:
            full = sharding.gather_c_batches(c_v[:batch_loc * m * n], m * n, batch, world)
        else:
            full = sharding.gather_c_batches(c, m * n, batch * world, world)
        g1.record()
        torch.cuda.synchronize()
        tg = torch.tensor([g0.elapsed_time(g1)], device=dev, dtype=torch.float64)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        gather_ms = float(tg.item())
        del full
        # the same gather overlapped with the compute: column panels, panel j on NVLink while panel j+1 computes
        overlap_ms = None
        if batch == 1 and strong and m_loc * world == m:
            def gemm_panel(n0, nb, c_panel):
                b_off_p = n0 * ldb if w["tb"] == "n" else n0
                blas._gemm(h, w["ta"], w["tb"], m_loc, nb, k, w["alpha"], a_v, lda, b[b_off_p:], ldb, 0.0, c_panel, m_loc)
            side = torch.cuda.Stream(device=dev)
            for it in range(2):   # first pass warms the allocator and NCCL for the panel sizes
                torch.cuda.synchronize()
                dist.barrier()
                g0.record()
                full = sharding.gemm_mblock_gather_overlapped(gemm_panel, m, n, m_loc, world, tdt, dev, panels=8,
                                                              side_stream=side)
                g1.record()
                torch.cuda.synchronize()
                del full
            to = torch.tensor([g0.elapsed_time(g1)], device=dev, dtype=torch.float64)
            dist.all_reduce(to, op=dist.ReduceOp.MAX)
            overlap_ms = float(to.item())
            # fused form: the epilogue of this rank's M-block stores every tile into ALL ranks' C (CUDA IPC peer
            # pointers), no staging buffer and no collective
            c_full = torch.zeros(m * n, device=dev, dtype=tdt)
            ptrs = sharding.share_full_c(h, c_full)
            for it in range(2):
                torch.cuda.synchronize()
                dist.barrier()
                g0.record()
                sharding.gemm_mblock_fused_gather(h, w["ta"], w["tb"], m, n, k, w["alpha"], a_v, lda, b, ldb, 0.0, ptrs, m,
                                                  tdt, world, rank, align=256)
                g1.record()
                torch.cuda.synchronize()
            dist.barrier()
            tf_ = torch.tensor([g0.elapsed_time(g1)], device=dev, dtype=torch.float64)
            dist.all_reduce(tf_, op=dist.ReduceOp.MAX)
            fused_ms = float(tf_.item())
            # every rank must now hold the SAME full C (bitwise: compare checksums across ranks), and this rank's rows
            # must equal its own plain GEMM
            cs = torch.tensor([float(c_full.double().sum()), float(c_full.double().abs().sum())], device=dev,
                              dtype=torch.float64)
            all_cs = [torch.empty_like(cs) for _ in range(world)]
            dist.all_gather(all_cs, cs)
            same = all(bool(torch.equal(all_cs[0], x)) for x in all_cs)
            mine = c_full.view(n, m)[:, c_off:c_off + m_loc]
            local_ok = bool(torch.equal(mine, c.view(n, ldc)[:, c_off:c_off + m_loc]))
            fused_ok = same and local_ok
            del c_full

    # ---- end-to-end: HOST (pinned) buffers through the public host-buffer call, H2D + D2H in the timed region ----
    e2e = None
    if not args.no_e2e:
        es = a.element_size()
        a_h = torch.empty(a.numel(), dtype=tdt, pin_memory=True); a_h.copy_(a)
        b_h = torch.empty(b.numel(), dtype=tdt, pin_memory=True); b_h.copy_(b)
        c_h = torch.zeros(c.numel(), dtype=tdt, pin_memory=True)
        torch.cuda.synchronize()

        def e2e_step():
            blas.gemm_host(h, w["ta"], w["tb"], m_loc, n, k, w["alpha"], a_h[a_off:], lda, b_h[b_off:], ldb, w["beta"],
                           c_h[c_off:], ldc, stridea=m * k if batch > 1 else 0, strideb=k * n if batch > 1 else 0,
                           stridec=m * n if batch > 1 else 0, batch_size=batch_loc)
        e2e_step()
        e2e_steps = max(1, min(args.steps, 5))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()  # synchronous: returns after the D2H copy of C completed
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        te = torch.tensor([el], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        el = float(te.item()) / e2e_steps
        c_el = m_loc * n * batch_loc
        h2d = (m_loc * k * batch_loc + k * n * batch_loc) * es + (c_el * es if w["beta"] != 0 else 0)
        e2e = dict(value=round(job_flops / el / 1e12, 3), unit="TFLOP/s", h2d_bytes_per_step=int(h2d),
                   d2h_bytes_per_step=int(c_el * es), ms_per_step=round(el * 1e3, 3), steps=e2e_steps,
                   api="pbx_gemm_host: pinned host buffers, H2D panels | GEMM | D2H panels pipelined on 3 streams "
                       "(= copy_to_device + _gemm + copy_to_host + wait of samples/gemm.cpp)")
        del a_h, b_h, c_h

    if rank == 0:
        traffic = None
        tp = ROOT / "profiles" / "roofline_traffic.json"
        if tp.exists():
            traffic = json.loads(tp.read_text()).get(args.workload)
        avg_launch_ms = sum(per_launch) / len(per_launch)
        roof = roofline_for(dict(w, m=m_loc, batch=batch_loc), avg_launch_ms, traffic)  # rank 0's launch
        roof["kernel"] = kernel_used
        roof["avg_launch_ms"] = round(avg_launch_ms, 4)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_sample(w)
        line = dict(metric="gemm_tflops", value=round(value, 3), unit="TFLOP/s", n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=round(ms_per_step, 4), higher_is_better=True,
                    scaling="strong" if strong else "weak",
                    vs_baseline=None, dtype=w["dt"], data="synthetic U(-2,5), seed 12345+rank, generated on device",
                    config=dict(workload=w["desc"], per_gpu_shape=[m_loc, n, k, batch_loc],
                                parallelism=f"mblock{world}" if batch == 1 else f"batchshard{world}",
                                l2="inputs+output %.0f MiB per GPU > 126 MB L2, no flush needed" % (byts / 2**20)
                                if byts > 200e6 else "working set fits L2: back-to-back launches reuse L2 (noted)",
                                kernel=kernel_used, split_k=split_used),
                    roofline=roof, cpu_baseline=cpu, e2e=e2e, gpu_launches=int(launches), clocks=clocks)
        if gather_ms is not None:
            line["gather_c_ms"] = round(gather_ms, 3)
            if overlap_ms is not None:
                line["gemm_plus_overlapped_gather_ms"] = round(overlap_ms, 3)
                line["gemm_fused_gather_ms"] = round(fused_ms, 3)
                line["gemm_fused_gather_ok"] = fused_ok
        print(json.dumps(line), flush=True)
    h.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
