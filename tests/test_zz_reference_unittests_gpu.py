"""The reference's OWN unit tests, compiled unchanged (portblas_b200/build_host.py: build/ref_unittest_*), run on the B200
against libpbx_gemm.so: test/unittest/blas3/blas3_gemm_test.cpp (float, double, half, half->float, complex),
blas3_gemm_batched_test.cpp (strided + interleaved), blas3_gemm_tall_skinny_test.cpp, blas3_symm_test.cpp and
blas3_trsm_test.cpp, each comparing with CBLAS through the reference's own verifier and tolerance.

Round 2: every suite ran to its end on the box (logs: profiles/r02/ref_unittests/), so all of them are plain tests --
no xfail.  The reference's benchmark executables (build/ref_bench_*) run here too, with the verification the reference
builds them with by default (BLAS_VERIFY_BENCHMARK: each benchmark first checks its result against CBLAS).
The file sorts last on purpose.
"""
from __future__ import annotations

import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu

# binary -> (--gtest_filter, minimum number of tests that must have run).  The filters keep the file to a few minutes:
# USM containers only where the buffer variants repeat the same grid, and for the 17 078 batched cases the float,
# half->float and complex<float> instantiations.
SUITES = {
    "blas3_gemm_test": ("*alloc_usm*", 1800),
    "blas3_gemm_tall_skinny_test": ("*", 300),
    "blas3_symm_test": ("*alloc_usm*", 450),
    "blas3_trsm_test": ("*alloc_usm*", 550),
    "blas3_gemm_batched_test": ("*FloatFloat.test/alloc_usm*:*HalfFloat.test/alloc_usm*:*CplxFloat.test/alloc_usm*", 4000),
}


@pytest.mark.parametrize("name", list(SUITES))
def test_reference_unit_tests_pass_on_the_gpu(handle, name):
    exe = ROOT / "build" / f"ref_unittest_{name}"
    if not exe.exists():
        pytest.skip("reference unit tests were not prebuilt (needs /root/reference at build time)")
    flt, at_least = SUITES[name]
    r = subprocess.run([str(exe), f"--gtest_filter={flt}"], capture_output=True, text=True, timeout=420)
    tail = "\n".join(r.stdout.splitlines()[-15:])
    out_dir = ROOT / "gpurun_out" / "ref_unittests"
    out_dir.mkdir(parents=True, exist_ok=True)
    (out_dir / f"{name}.log").write_text(r.stdout[-200000:] + "\n--- stderr ---\n" + r.stderr[-20000:])
    ran = [ln for ln in r.stdout.splitlines() if ln.startswith("[==========]")]
    assert ran, tail + r.stderr[-2000:]
    n_ran = int(ran[-1].split()[1])
    assert "[  FAILED  ]" not in r.stdout, tail
    assert r.returncode == 0 and n_ran >= at_least, (r.returncode, n_ran, tail)


JOINT_MATRIX = ["half_half_16_16_16", "half_half_32_8_16", "half_half_8_32_16", "half_float_16_16_16", "half_float_32_8_16",
                "half_float_8_32_16", "bfloat16_float_16_16_16", "bfloat16_float_32_8_16", "bfloat16_float_8_32_16",
                "tf32_float_16_16_8"]


@pytest.mark.parametrize("name", JOINT_MATRIX)
def test_reference_joint_matrix_tests_pass_through_the_seam(handle, name):
    """build/ref_unittest_joint_matrix_<name>: test/unittest/joint_matrix/<name>.cpp unchanged; it calls
    blas::Gemm_Launcher<...>::_select_gemm (include/interface/gemm_launcher.h) with float storage whose low 13 / 16
    mantissa bits are zero and compares with CBLAS.  USM containers (the buffer variants repeat the grid)."""
    exe = ROOT / "build" / f"ref_unittest_joint_matrix_{name}"
    if not exe.exists():
        pytest.skip("reference joint_matrix tests were not prebuilt (needs /root/reference at build time)")
    r = subprocess.run([str(exe), "--gtest_filter=*usm*"], capture_output=True, text=True, timeout=180)
    tail = "\n".join(r.stdout.splitlines()[-15:])
    out_dir = ROOT / "gpurun_out" / "ref_unittests"
    out_dir.mkdir(parents=True, exist_ok=True)
    (out_dir / f"joint_matrix_{name}.log").write_text(r.stdout[-200000:] + "\n--- stderr ---\n" + r.stderr[-20000:])
    ran = [ln for ln in r.stdout.splitlines() if ln.startswith("[==========]")]
    assert ran, tail + r.stderr[-2000:]
    assert "[  FAILED  ]" not in r.stdout, tail
    assert r.returncode == 0 and int(ran[-1].split()[1]) >= 700, (r.returncode, ran[-1], tail)


# reference benchmark executable -> rows of its --csv-param file (benchmark/README.md: the column order per operator)
BENCH_CSV = {
    "gemm": "n,n,1024,1024,1024,1.5,0.5\nt,n,512,333,257,1,0\nn,t,63,1025,129,1,1\n",
    # batch type is spelled out (common_utils.hpp:240-247); the interleaved rows are the ones the reference ships in
    # benchmark/config_csv/blas3/gemm_batched/gemm_batched_interleaved.csv plus a ragged one
    "gemm_batched": ("n,n,64,64,64,1,0,32,strided\nt,n,33,65,17,1.5,0.5,7,strided\nn,n,32,32,32,1,1,64,interleaved\n"
                     "n,n,65,3,49,1,0,32,interleaved\nt,t,230,230,49,1,0,32,interleaved\nn,t,17,9,5,1.5,0.5,7,interleaved\n"),
    "gemm_batched_strided": "n,n,128,128,128,1,0,16,2,2,2\nt,t,33,65,17,1.5,0.5,5,1,1,1\n",
    "symm": "l,u,512,256,1,0\nr,l,127,255,1.5,0.5\n",
    "trsm": "l,u,n,n,512,256,1\nr,l,t,u,127,255,2\n",
}


@pytest.mark.parametrize("name", list(BENCH_CSV))
def test_reference_benchmark_harness_runs_and_verifies(handle, name, tmp_path):
    """build/ref_bench_<name>: the reference's benchmark/portblas/blas3/<name>.cpp + main.cpp, unchanged, with
    BLAS_VERIFY_BENCHMARK (each benchmark first checks its result against CBLAS and reports an error otherwise)."""
    import json
    exe = ROOT / "build" / f"ref_bench_{name}"
    if not exe.exists():
        pytest.skip("reference benchmarks were not prebuilt (needs /root/reference at build time)")
    csv = tmp_path / "params.csv"
    csv.write_text(BENCH_CSV[name])
    out = tmp_path / "report.json"
    r = subprocess.run([str(exe), "--csv-param", str(csv), "--benchmark_min_time=0.05", f"--benchmark_out={out}"],
                       capture_output=True, text=True, timeout=420)
    out_dir = ROOT / "gpurun_out" / "ref_unittests"
    out_dir.mkdir(parents=True, exist_ok=True)
    (out_dir / f"bench_{name}.log").write_text(r.stdout[-200000:] + "\n--- stderr ---\n" + r.stderr[-20000:])
    assert "ERROR OCCURRED" not in r.stdout, r.stdout[-3000:]
    assert r.returncode == 0, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    rep = json.loads(out.read_text())["benchmarks"]
    assert len(rep) >= len(BENCH_CSV[name].splitlines()) * 2
    for b in rep:
        assert "error_occurred" not in b and b["real_time"] > 0 and b["n_fl_ops"] > 0, b
