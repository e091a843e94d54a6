#!/usr/bin/env python
"""bench.py -- headline benchmark of the GEMM path (BASELINE.json: "GEMM TFLOP/s and % of per-dtype tensor peak
(fp32/fp64/bf16), 1-8 B200").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference] [--no-sub]

A "step" is one pass of the hot path (one blas::_gemm / _gemm_strided_batched call) over one synthetic batch of inputs.

Main line (what `value` is): BASELINE configs[2], SGEMM 16384^3 NN fp32 -- the configuration the north star shards.
  N = 1: one pbx_gemm call on one B200.
  N > 1: STRONG scaling of the SAME problem, M-block shards (portblas_b200/sharding.py): rank g owns rows
         [g*M/N, (g+1)*M/N) of A and C, B is replicated.  `value` times compute + gather: the GEMM's epilogue stores
         every finished tile into the same rows of EVERY rank's full C over NVLink (pbx_gemm_multicast, CUDA IPC peer
         pointers), so when the step ends each rank holds the whole 16384 x 16384 product.  `compute_only` in the same
         line is the same step without the peer stores (each rank keeps only its row block).
`sub` in the same JSON line: the other BASELINE configurations, each timed the same way with its own step count
(scaled so that the timed region lasts >= 0.6 s), roofline and clock record:
  N = 1: cfg2 DGEMM 8192^3 (NN alpha=1 beta=0, and TT alpha=1.5 beta=0.5), cfg4 strided-batched bf16 and f16
         4096 x 256^3, cfg5 tall-skinny SGEMM 512x512x2^20 (split-K), cfg1's shape SGEMM 1024^3, BF16 8192^3.
  N > 1: cfg4 bf16 and f16, batch-sharded (strong; compute only -- every rank keeps its batch range of C).
Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks; inputs larger than
L2 or rotated through enough buffer sets to exceed it (stated per line in `config.l2`); clocks and throttle reasons
sampled through NVML every ~5 ms during each timed region.  One JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# name -> dict(dtype key, m, n, k, batch, transa, transb, alpha, beta[, strong])
WORKLOADS = {
    # BASELINE configs[1]
    "dgemm8192": dict(dt="f64", m=8192, n=8192, k=8192, batch=1, ta="n", tb="n", alpha=1.0, beta=0.0,
                      desc="DGEMM 8192x8192x8192 NN alpha=1 beta=0 (BASELINE configs[1])"),
    "dgemm8192_tt": dict(dt="f64", m=8192, n=8192, k=8192, batch=1, ta="t", tb="t", alpha=1.5, beta=0.5,
                         desc="DGEMM 8192x8192x8192 TT alpha=1.5 beta=0.5 (BASELINE configs[1])"),
    # BASELINE configs[2] (per-GPU M-block of the 16384^3 problem when sharded 1/2/4/8 ways is 16384/N rows)
    "sgemm16384": dict(dt="f32", m=16384, n=16384, k=16384, batch=1, ta="n", tb="n", alpha=1.0, beta=0.0, strong=True,
                       desc="SGEMM 16384^3 NN fp32 (BASELINE configs[2]); N>1: M-block shards of the SAME problem"),
    "sgemm8192": dict(dt="f32", m=8192, n=8192, k=8192, batch=1, ta="n", tb="n", alpha=1.0, beta=0.0,
                      desc="SGEMM 8192^3 NN fp32"),
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
DEFAULT_WORKLOAD = "sgemm16384"
SUBS_ONE_GPU = ["dgemm8192", "dgemm8192_tt", "bf16gemm_batched", "hgemm_batched", "sgemm_splitk", "sgemm1024", "bf16gemm8192"]
SUBS_MULTI_GPU = ["bf16gemm_batched", "hgemm_batched"]
ES_IN = {"f64": 8, "f32": 4, "f16": 2, "bf16": 2}
NOMINAL_FP64_TFLOPS = 40.0      # B200 datasheet FP64 (tensor == vector); MEASURED_PEAKS.json has no fp64 entry
MEASURED_FP64_PIPE_TFLOPS = 36.98  # register-resident DMMA.8x8x4 loop on this pool's B200 (profiles/r01/dmma_rate.txt)
L2_BYTES = 126e6
MIN_REGION_S = 0.6              # sub-workloads: timed region at least this long


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


def roofline_for(w, avg_ms, traffic, presplit=0):
    """Roof that binds the dominant kernel: tensor pipe for deep contractions, HBM when the arithmetic intensity is
    below the ridge.  fp32 runs as split products on the tf32 / bf16 tensor pipes, so its roof is
    the bf16 figure divided by the tf32-MMA-equivalents one k-step costs: 3xTF32 (presplit 0/1) = 3 tf32 MMAs = 6 bf16
    times; tf32 + 2 x bf16 (presplit 3 / 4) = 1 tf32 + 2 bf16 = 4 bf16 times; single tf32 (presplit 2) = 2."""
    pk = _peaks()
    flops, byts = algorithmic(w)
    ai = flops / byts
    div = 1.0
    if w["dt"] == "f64":
        tensor_peak = NOMINAL_FP64_TFLOPS
        src = ("nominal fp64 40 TF (no fp64 entry in MEASURED_PEAKS.json); measured DMMA pipe ceiling "
               f"{MEASURED_FP64_PIPE_TFLOPS} TF (tools/micro/dmma_rate.cu, profiles/r01/dmma_rate.txt)")
    elif w["dt"] == "f32":
        div = {2: 2.0, 3: 4.0, 4: 4.0}.get(presplit, 6.0)
        how = {2: "single tf32 product (SB_ENABLE_JOINT_MATRIX=1: reduced-precision fragments): bf16 / 2",
               3: "tf32 + 2 x bf16 split (1 tf32 MMA + 2 bf16 MMAs per k-step; bf16 copies made by a pre-pass): bf16 / 4",
               4: "tf32 + 2 x bf16 split (1 tf32 MMA + 2 bf16 MMAs per k-step; bf16 tiles made in the kernel): bf16 / 4"}.get(
                   presplit, "3xTF32 (three tf32 MMAs per k-step): bf16 / 6")
        tensor_peak = pk["bf16"] / div
        src = f"fp32 roof for {how}; bf16 {pk['source']}"
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
        sus = pk["bf16_sustained"] / div
        out["peak_sustained"] = round(sus, 1)
        out["frac_of_sustained"] = round(ach / sus, 4)
    if w["dt"] == "f32" and presplit in (3, 4):
        # the round-1 denominator, for comparison across rounds: three tf32 MMAs per product
        out["frac_of_3xtf32_roof"] = round(ach / (pk["bf16"] / 6.0), 4)
    return out


class ClockSampler:
    """SM clock, power and throttle reasons during a timed region.  NVML polled from a thread every ~5 ms (so that even a
    50 ms region gets >= 5 samples); nvidia-smi -lms 100 (the B200_PROFILING.md recipe) when NVML is not importable."""
    SMI_Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, cuda_index, period_s=0.005):
        self.samples, self.reasons, self.proc, self.thread = [], set(), None, None
        self.stop_flag = threading.Event()
        self.mx = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(cuda_index).uuid)
            try:
                self.dev = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.dev = pynvml.nvmlDeviceGetHandleByIndex(cuda_index)
            self.nv = pynvml
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM))
            self.period = period_s
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            self.how = "nvml"
        except Exception:
            self.thread = None
            try:
                self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.SMI_Q}", "--format=csv,noheader,nounits",
                                              "-lms", "100", "-i", str(cuda_index)], stdout=subprocess.PIPE,
                                             stderr=subprocess.DEVNULL, text=True)
                self.how = "nvidia-smi"
            except Exception:
                self.proc = None

    def _poll(self):
        nv = self.nv
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap),
                 ("hw_power_brake_slowdown", nv.nvmlClocksEventReasonHwPowerBrakeSlowdown))
        while not self.stop_flag.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(self.dev) / 1000.0
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
                self.samples.append((sm, pw))
                for name, bit in names:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(self.period)

    def stop(self):
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            if not self.samples:
                return dict(sm_mhz=None, sm_max_mhz=self.mx, reasons=["no samples"], how="nvml")
            sm = sorted(s[0] for s in self.samples)
            return dict(sm_mhz=sm[len(sm) // 2], sm_min_mhz=sm[0], sm_max_mhz=self.mx,
                        power_w_max=round(max(s[1] for s in self.samples), 1), samples=len(sm),
                        reasons=sorted(self.reasons), how="nvml, 5 ms period")
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
                    reasons=sorted(reasons), how="nvidia-smi -lms 100")


def host_threads() -> int:
    """Threads the CPU legs may use: the cores this process may run on -- NOT an inherited OMP_NUM_THREADS (torchrun exports
    OMP_NUM_THREADS=1 to its workers, which would turn the reference arm into a single-core run)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def _claim_host_threads() -> int:
    n = host_threads()
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(n)   # must happen before the OpenMP runtime / OpenBLAS are loaded
    return n


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
    """Bounded CPU sample of the same workload on the first ``ms`` rows of C with full N and K capped at 8192
    (see _cpu_gemm)."""
    _claim_host_threads()
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
    sample of the workload (see _cpu_gemm for what runs).  Rank 0 only; the thread count comes from the process's CPU
    affinity, not from the OMP_NUM_THREADS=1 that torchrun hands its workers."""
    if rank != 0:
        return
    threads = _claim_host_threads()
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
                scaling="strong" if w.get("strong") and args.gpus > 1 else "weak", vs_baseline=None, dtype=w["dt"],
                data="synthetic U(-2,5) seed 12345",
                config=dict(workload=w["desc"], sample=sample, host_threads=threads),
                cpu_baseline=dict(value=round(val, 5), unit="TFLOP/s", cores=cores, kind=kind, sample=sample),
                e2e=dict(value=round(val, 5), unit="TFLOP/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0,
                note=f"reference arm = {label}; portBLAS needs a SYCL compiler (absent in this image), so its kernels run "
                     "on a host executor rather than on a SYCL CPU device")
    print(json.dumps(line), flush=True)


class Bench:
    """One process = one rank = one GPU.  Holds the handle and the torch plumbing shared by the main line and `sub`."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the GEMM path has no CPU fallback")
        from portblas_b200 import SB_Handle, blas, build, sharding
        build.build()
        self.blas, self.sharding = blas, sharding
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.h = SB_Handle(self.local_rank)
        self.gen = torch.Generator(device=self.dev)
        self.gen.manual_seed(12345 + self.rank)
        tp = ROOT / "profiles" / "roofline_traffic.json"
        self.traffic = json.loads(tp.read_text()) if tp.exists() else {}

    def rand(self, count, tdt):
        torch = self.torch
        chunk = 1 << 26
        out = torch.empty(count, device=self.dev, dtype=tdt)
        for s in range(0, count, chunk):
            e = min(count, s + chunk)
            out[s:e] = (torch.rand(e - s, device=self.dev, dtype=torch.float32, generator=self.gen) * 7.0 - 2.0).to(tdt)
        return out

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def time_steps(self, step, steps, warmup, min_region_s=0.0):
        """W warm-up steps, then `steps` timed ones (scaled up when min_region_s asks for a longer region) bracketed by
        barrier + synchronize; CUDA events on the launching stream; max over ranks.  Returns
        (steps, ms_per_step over ranks, this rank's avg launch ms, launches, clocks)."""
        torch = self.torch
        for i in range(warmup):
            step(i)
        self.barrier()
        if min_region_s > 0:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(3):
                step(i)
            e1.record()
            torch.cuda.synchronize()
            est_ms = self.max_over_ranks(max(e0.elapsed_time(e1) / 3.0, 1e-3))
            steps = max(steps, int(math.ceil(min_region_s * 1e3 / est_ms)))
        sampler = ClockSampler(self.local_rank) if self.rank == 0 else None
        launches0 = self.h.launch_count
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        ev0.record()
        for i in range(steps):
            step(i)
        ev1.record()
        self.barrier()
        launches = self.h.launch_count - launches0
        clocks = sampler.stop() if sampler else None
        mine = ev0.elapsed_time(ev1)
        total_ms = self.max_over_ranks(mine)
        return steps, total_ms / steps, mine / steps, int(launches), clocks

    def run_workload(self, name, steps, warmup, min_region_s=0.0, fused_gather=False):
        """Times one workload.  Returns the result dict (value, ms_per_step, roofline, clocks, config ...)."""
        torch, blas, sharding, h = self.torch, self.blas, self.sharding, self.h
        w = WORKLOADS[name]
        world, rank = self.world, self.rank
        tdt = {"f64": torch.float64, "f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}[w["dt"]]
        m, n, k, batch = w["m"], w["n"], w["k"], w["batch"]
        ta, tb = w["ta"] != "n", w["tb"] != "n"
        lda, ldb, ldc = (k if ta else m), (n if tb else k), m
        strong = bool(w.get("strong")) and world > 1
        flops, byts = algorithmic(w)

        # N > 1, strong: the SAME problem is cut into M-blocks / batch ranges: a rank's shard is a pointer offset into
        # the full operands with the ORIGINAL leading dimensions -- no copy.  Not strong: every rank runs the whole shape.
        m_loc, batch_loc, a_off, b_off, c_off = m, batch, 0, 0, 0
        if strong and batch == 1:
            sh = sharding.shard_mblock(w["ta"], m, lda, world, rank, align=256)
            m_loc, a_off, c_off = sh.rows, sh.a_offset, sh.c_offset
        elif strong:
            bs = sharding.shard_batch(batch, m * k, k * n, m * n, world, rank)
            batch_loc, a_off, b_off, c_off = bs.batches, bs.a_offset, bs.b_offset, bs.c_offset
        local_w = dict(w, m=m_loc, batch=batch_loc)
        _, local_bytes = algorithmic(local_w)
        # working set below ~2 x L2: rotate through enough operand sets that a step never finds its inputs in L2
        sets = 1 if local_bytes > 2 * L2_BYTES else int(math.ceil(2.5 * L2_BYTES / local_bytes))
        ops = []
        for _ in range(sets):
            a = self.rand(lda * (m if ta else k) * batch, tdt)
            b = self.rand(ldb * (k if tb else n) * batch, tdt)
            c = (self.rand(ldc * n * batch, tdt) if w["beta"] != 0 else torch.zeros(ldc * n * batch, device=self.dev, dtype=tdt))
            ops.append((a, b, c))

        ptrs = None
        if fused_gather and strong and batch == 1:
            ptrs = sharding.share_full_c(h, ops[0][2])   # every rank's full C, mapped through CUDA IPC

        def step(i, fused=False):
            a, b, c = ops[i % sets]
            if batch == 1 and fused:
                sharding.gemm_mblock_fused_gather(h, w["ta"], w["tb"], m, n, k, w["alpha"], a[a_off:], lda, b, ldb, w["beta"],
                                                  ptrs, ldc, tdt, world, rank, align=256)
            elif batch == 1:
                blas._gemm(h, w["ta"], w["tb"], m_loc, n, k, w["alpha"], a[a_off:], lda, b[b_off:], ldb, w["beta"], c[c_off:], ldc)
            else:
                blas._gemm_strided_batched(h, w["ta"], w["tb"], m, n, k, w["alpha"], a[a_off:], lda, m * k, b[b_off:], ldb, k * n,
                                           w["beta"], c[c_off:], ldc, m * n, batch_loc)

        out = {}
        job_flops = flops if strong else flops * world   # strong: the one problem; otherwise one problem per GPU
        if ptrs is not None:
            # the headline form first (compute + gather), the compute-only form after it: on a power-capped board the
            # first of two back-to-back regions runs at higher clocks, and the headline must not be the favoured one
            steps_f, ms_f, mine_f, launches_f, clocks_f = self.time_steps(lambda i: step(i, True), steps, warmup, min_region_s)
            kernel_used, split_used, presplit = h.last_kernel, h.last_split_k, h.last_presplit
            # every rank must now hold the SAME full C (checksums compared across ranks)
            c_full = ops[0][2]
            cs = torch.tensor([float(c_full.double().sum()), float(c_full.double().abs().sum())], device=self.dev,
                              dtype=torch.float64)
            all_cs = [torch.empty_like(cs) for _ in range(world)]
            self.dist.all_gather(all_cs, cs)
            out["gather"] = dict(how="fused into the GEMM: every finished tile is stored into the local C and pushed to all "
                                     "ranks' C over NVLink by TMA while the next tile computes (pbx_gemm_multicast); no "
                                     "collective, no staging buffer",
                                 all_ranks_hold_identical_c=all(bool(torch.equal(all_cs[0], x)) for x in all_cs),
                                 nvlink_egress_gb_per_rank_per_step=round(m_loc * n * ES_IN[w["dt"]] * (world - 1) / 1e9, 3))
            steps_c, ms_c, mine_c, launches_c, clocks_c = self.time_steps(lambda i: step(i), steps, warmup, min_region_s)
            out["compute_only"] = dict(value=round(job_flops / (ms_c * 1e-3) / 1e12, 3), ms_per_step=round(ms_c, 4),
                                       steps=steps_c, clocks=clocks_c, note="timed after the headline region")
            res_steps, res_ms, res_mine, res_launches, res_clocks = steps_f, ms_f, mine_f, launches_f, clocks_f
        else:
            res_steps, res_ms, res_mine, res_launches, res_clocks = self.time_steps(lambda i: step(i), steps, warmup,
                                                                                    min_region_s)
            kernel_used, split_used, presplit = h.last_kernel, h.last_split_k, h.last_presplit
        value = job_flops / (res_ms * 1e-3) / 1e12
        roof = roofline_for(local_w, res_mine, self.traffic.get(name), presplit)   # rank 0's launch
        roof["kernel"] = kernel_used
        roof["avg_launch_ms"] = round(res_mine, 4)
        if sets > 1:
            l2 = (f"per-GPU working set {local_bytes / 2**20:.0f} MiB: steps rotate through {sets} operand sets "
                  f"({sets * local_bytes / 2**20:.0f} MiB > 126 MB L2), so no step finds its inputs in L2")
        else:
            l2 = f"inputs+output {local_bytes / 2**20:.0f} MiB per GPU > 126 MB L2, no flush needed"
        out.update(workload=name, value=round(value, 3), unit="TFLOP/s", ms_per_step=round(res_ms, 4), steps=res_steps,
                   warmup=warmup, dtype=w["dt"], scaling="strong" if strong else "weak",
                   config=dict(workload=w["desc"], per_gpu_shape=[m_loc, n, k, batch_loc],
                               parallelism=(f"mblock{world}" if batch == 1 else f"batchshard{world}") if strong else f"replica{world}",
                               l2=l2, kernel=kernel_used, split_k=split_used, f32_split_mode=presplit if w["dt"] == "f32" else None),
                   roofline=roof, gpu_launches=res_launches, clocks=res_clocks)
        self._last = (w, ops, (m_loc, batch_loc, a_off, b_off, c_off, lda, ldb, ldc), job_flops, tdt)
        return out

    def e2e(self, steps):
        """The last workload through pbx_gemm_host: pinned HOST buffers in, host result out; H2D + D2H inside the timed
        region, wall clock (the call is synchronous), max over ranks."""
        torch, blas, h = self.torch, self.blas, self.h
        w, ops, (m_loc, batch_loc, a_off, b_off, c_off, lda, ldb, ldc), job_flops, tdt = self._last
        a, b, c = ops[0]
        m, n, k, batch = w["m"], w["n"], w["k"], w["batch"]
        es = a.element_size()
        a_h = torch.empty(a.numel(), dtype=tdt, pin_memory=True); a_h.copy_(a)
        b_h = torch.empty(b.numel(), dtype=tdt, pin_memory=True); b_h.copy_(b)
        c_h = torch.zeros(c.numel(), dtype=tdt, pin_memory=True)
        torch.cuda.synchronize()

        exchange = self.world > 1 and batch == 1 and bool(w.get("strong")) and w["tb"] == "n" and n % (256 * self.world) == 0

        def e2e_step():
            if exchange:
                # M-block shards with HOST operands: own rows of A + ONE panel of B over PCIe, the panels exchanged over
                # NVLink (NCCL all-gather: the path's one real exchange step), own rows of C back
                self.sharding.gemm_mblock_host(h, w["ta"], w["tb"], m, n, k, w["alpha"], a_h, lda, b_h, ldb, w["beta"], c_h, ldc,
                                               self.world, self.rank, align=256)
            else:
                blas.gemm_host(h, w["ta"], w["tb"], m_loc, n, k, w["alpha"], a_h[a_off:], lda, b_h[b_off:], ldb, w["beta"],
                               c_h[c_off:], ldc, stridea=m * k if batch > 1 else 0, strideb=k * n if batch > 1 else 0,
                               stridec=m * n if batch > 1 else 0, batch_size=batch_loc)
        e2e_step()
        e2e_steps = max(1, min(steps, 5))
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()  # synchronous: returns after the D2H copy of C completed
        torch.cuda.synchronize()
        mine = (time.perf_counter() - t0) / e2e_steps
        el = self.max_over_ranks(mine)
        c_el = m_loc * n * batch_loc
        b_el = k * n * batch_loc // (self.world if exchange else 1)
        h2d = (m_loc * k * batch_loc + b_el) * es + (c_el * es if w["beta"] != 0 else 0)
        d2h = c_el * es
        del a_h, b_h, c_h
        api = ("pbx_gemm_host: pinned host buffers, H2D panels | GEMM | D2H panels pipelined on 3 streams "
               "(= copy_to_device + _gemm + copy_to_host + wait of samples/gemm.cpp)")
        if exchange:
            api = ("sharding.gemm_mblock_host: every rank uploads its M-block of A and 1/N of B from pinned host memory, the B "
                   "panels are exchanged over NVLink (NCCL all-gather), the rank computes its M-block and downloads it")
        return dict(value=round(job_flops / el / 1e12, 3), unit="TFLOP/s", h2d_bytes_per_step=int(h2d),
                    d2h_bytes_per_step=int(d2h), bytes_are="per rank", ms_per_step=round(el * 1e3, 3), steps=e2e_steps,
                    host_gb_per_s_this_rank=round((h2d + d2h) / mine / 1e9, 1), api=api)

    def close(self):
        self.h.close()
        if self.world > 1:
            self.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="main line only (skip the other BASELINE configurations)")
    ap.add_argument("--no-gather", action="store_true", help="N>1: time the compute only (no fused gather of C)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return

    bn = Bench(args)
    main_res = bn.run_workload(args.workload, args.steps, args.warmup, 0.0, fused_gather=not args.no_gather)
    e2e = None if args.no_e2e else bn.e2e(args.steps)
    subs = []
    if not args.no_sub and args.workload == DEFAULT_WORKLOAD:
        for name in (SUBS_ONE_GPU if world == 1 else SUBS_MULTI_GPU):
            bn._last = None
            bn.torch.cuda.empty_cache()
            r = bn.run_workload(name, args.steps, args.warmup, MIN_REGION_S)
            subs.append(r)
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_sample(w)
        line = dict(metric="gemm_tflops", value=main_res["value"], unit="TFLOP/s", n_gpus=world, steps=main_res["steps"],
                    warmup=args.warmup, ms_per_step=main_res["ms_per_step"], higher_is_better=True,
                    scaling=main_res["scaling"], vs_baseline=None, dtype=main_res["dtype"],
                    data="synthetic U(-2,5), seed 12345+rank, generated on device", config=main_res["config"],
                    roofline=main_res["roofline"], cpu_baseline=cpu, e2e=e2e,
                    gpu_launches=main_res["gpu_launches"] + sum(s["gpu_launches"] for s in subs), clocks=main_res["clocks"])
        if world > 1 and "gather" in main_res:
            line["value_is"] = "compute + gather: every rank ends the step holding the full C (fused into the GEMM epilogue)"
            line["gather"] = main_res["gather"]
            line["compute_only"] = main_res["compute_only"]
        elif world > 1:
            line["value_is"] = "compute only: every rank keeps its shard of C"
        line["gpu_launches_main"] = main_res["gpu_launches"]
        if subs:
            line["sub"] = subs
        print(json.dumps(line), flush=True)
    bn.close()


if __name__ == "__main__":
    main()
