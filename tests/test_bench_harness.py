"""CPU checks of tools/portblas_bench.py, the stand-in for the reference's benchmark/portblas harness: benchmark
names (common/include/common/benchmark_names.hpp:53-58,190-232) and counters
(common/include/common/blas3_state_counters.hpp:38-76,141-166) for rows of the reference's CSV parameter files."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))
import portblas_bench as pb  # noqa: E402


def test_benchmark_names_follow_the_reference():
    # benchmark/config_csv/blas3/gemm/*.csv rows are `ta,tb,m,k,n,alpha,beta` (note m,k,n)
    assert pb.bench_name("gemm", "float", ["n", "N", "1024", "56", "2", "1", "0"]) == "BM_Gemm<float>/n/n/1024/56/2/usm"
    assert pb.bench_name("gemm", "complex<double>", ["t", "n", 8, 9, 10, 1, 1]) == "BM_Gemm<complex<double>>/t/n/8/9/10/usm"
    assert pb.bench_name("gemm_batched", "float", ["n", "n", 64, 64, 5000, 1, 0, 32, "interleaved"]) == \
        "BM_Gemm_batched<float>/n/n/64/64/5000/32/interleaved/usm"
    assert pb.bench_name("gemm_batched_strided", "half", ["n", "t", 384, 64, 384, 1, 0, 896, 1, 2, 3]) == \
        "BM_Gemm_batched_strided<half>/n/t/384/64/384/896/1/2/3/usm"
    assert pb.bench_name("symm", "double", ["l", "u", 256, 512, 1, 0]) == "BM_Symm<double>/l/u/256/512/1/0/usm"
    assert pb.bench_name("symm", "float", ["R", "L", 16, 8, 1.5, 0.5]) == "BM_Symm<float>/r/l/16/8/1.5/0.5/usm"
    assert pb.bench_name("trsm", "float", ["l", "u", "n", "n", 345, 560, 1]) == "BM_Trsm<float>/l/u/n/n/345/560/usm"


def test_counters_follow_the_reference():
    c = pb.counters("gemm", "float", ["n", "n", 4, 3, 2, 1, 1])          # m=4 k=3 n=2, beta != 0
    assert c["n_fl_ops"] == 2 * 3 * 4 * 2 + 4 * 2 + 2 * 4 * 2 and c["bytes_processed"] == (12 + 6 + 8 + 8) * 4
    c = pb.counters("gemm", "double", ["n", "n", 4, 3, 2, 1, 0])         # beta == 0: C is not read
    assert c["n_fl_ops"] == 48 + 8 and c["bytes_processed"] == (12 + 6 + 8) * 8
    c = pb.counters("gemm_batched_strided", "half", ["n", "n", 4, 3, 2, 1, 0, 5, 2, 2, 2])
    assert c["batch_size"] == 5 and c["stride_c_mul"] == 2 and c["n_fl_ops"] == 56 * 5 and c["bytes_processed"] == 26 * 5 * 2
    c = pb.counters("symm", "float", ["r", "u", 4, 3, 1, 0])
    assert c["n_fl_ops"] == 2 * 3 * 3 * 3 and c["bytes_processed"] == (12 + 12 + 6) * 4
    c = pb.counters("trsm", "double", ["l", "l", "n", "n", 4, 3, 1])
    assert c["k"] == 4 and c["n_fl_ops"] == 16 * 3 + 12


def test_fixture_rows_map_to_names():
    import json
    rows = json.loads((ROOT / "tests" / "golden" / "config_csv_shapes.json").read_text())["rows"]
    r = next(x for x in rows if x["api"] == "gemm")
    name = pb.bench_name("gemm", "float", [r["ta"], r["tb"], r["m"], r["k"], r["n"], r["alpha"], r["beta"]])
    assert name == f"BM_Gemm<float>/{r['ta']}/{r['tb']}/{r['m']}/{r['k']}/{r['n']}/usm"


def _run_bench(extra_args, env_extra=None):
    import os
    import subprocess
    env = dict(os.environ)
    for key in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(key, None)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *extra_args], cwd=str(ROOT), env=env, capture_output=True,
                          text=True, timeout=600)


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` needs no GPU: it times the reference's own CPU GEMM (oracle/_ref, else the port) on a
    bounded sample and prints ONE JSON line in the driver's format."""
    import json
    r = _run_bench(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "gemm_tflops" and d["unit"] == "TFLOP/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert "16384" in d["config"]["workload"] and d["config"]["host_threads"] >= 1
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == dict(value=d["value"], unit="TFLOP/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun (N > 1) rank 0 alone runs the reference arm; the other ranks exit 0 without work or output."""
    r = _run_bench(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                   {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_bench_algorithmic_work_and_roofline_choice():
    """bench.py's per-launch algorithmic work is the reference's counter pair (blas3_state_counters.hpp:62-75: 2MNK flops,
    (MK + KN + MN(1 or 2)) elements) and the roof is the one the arithmetic intensity selects."""
    sys.path.insert(0, str(ROOT))
    import bench
    w = bench.WORKLOADS["dgemm8192"]
    fl, by = bench.algorithmic(w)
    assert fl == 2.0 * 8192 ** 3 and by == 3 * 8192 * 8192 * 8
    fl_tt, by_tt = bench.algorithmic(bench.WORKLOADS["dgemm8192_tt"])          # beta != 0: C is read as well
    assert fl_tt == fl and by_tt == 4 * 8192 * 8192 * 8
    w4 = bench.WORKLOADS["bf16gemm_batched"]                                    # cfg4: 85 flop/B -> HBM bound
    fl4, by4 = bench.algorithmic(w4)
    assert fl4 == 2.0 * 256 ** 3 * 4096 and by4 == 3 * 256 * 256 * 2 * 4096
    r4 = bench.roofline_for(w4, 0.25, None)
    assert r4["bound"] == "hbm" and r4["unit"] == "GB/s" and abs(r4["achieved"] - by4 / 0.25e-3 / 1e9) < 0.1
    assert abs(r4["frac"] - r4["achieved"] / r4["peak"]) < 1e-3
    pk = bench._peaks()
    main = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
    assert (main["m"], main["n"], main["k"], main["dt"]) == (16384, 16384, 16384, "f32")
    r = bench.roofline_for(main, 30.0, 49.4e9, presplit=3)                      # tf32 + 2 x bf16: bf16 / 4
    assert r["bound"] == "tensor" and abs(r["peak"] - pk["bf16"] / 4) < 0.06 and r["traffic"] == 49.4e9
    assert abs(r["achieved"] - 2.0 * 16384 ** 3 / 30e-3 / 1e12) < 0.01
    assert abs(bench.roofline_for(main, 30.0, None, presplit=0)["peak"] - pk["bf16"] / 6) < 0.06   # 3xTF32
    assert abs(bench.roofline_for(main, 30.0, None, presplit=2)["peak"] - pk["bf16"] / 2) < 0.06   # single tf32
    assert bench.roofline_for(w, 31.0, None)["peak"] == bench.NOMINAL_FP64_TFLOPS
