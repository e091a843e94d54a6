"""The reference's OWN unit tests, compiled unchanged (portblas_b200/build_host.py: build/ref_unittest_*), run on the B200
against libpbx_gemm.so: test/unittest/blas3/blas3_gemm_test.cpp (float, double, half, half->float, complex),
blas3_gemm_batched_test.cpp (strided + interleaved), blas3_gemm_tall_skinny_test.cpp, blas3_symm_test.cpp and
blas3_trsm_test.cpp, each comparing with CBLAS through the reference's own verifier and tolerance.

The binaries were first built after this round's GPU budget was spent, so this file has not run on a GPU yet: the tests
are non-strict xfail until a box run confirms them (a pass shows as XPASS).  It sorts last on purpose.
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


@pytest.mark.xfail(strict=False, reason="first GPU run of the reference's unit-test binaries happens after this round")
@pytest.mark.parametrize("name", list(SUITES))
def test_reference_unit_tests_pass_on_the_gpu(handle, name):
    exe = ROOT / "build" / f"ref_unittest_{name}"
    if not exe.exists():
        pytest.skip("reference unit tests were not prebuilt (needs /root/reference at build time)")
    flt, at_least = SUITES[name]
    r = subprocess.run([str(exe), f"--gtest_filter={flt}"], capture_output=True, text=True, timeout=900)
    tail = "\n".join(r.stdout.splitlines()[-15:])
    out_dir = ROOT / "gpurun_out" / "ref_unittests"
    out_dir.mkdir(parents=True, exist_ok=True)
    (out_dir / f"{name}.log").write_text(r.stdout[-200000:] + "\n--- stderr ---\n" + r.stderr[-20000:])
    ran = [ln for ln in r.stdout.splitlines() if ln.startswith("[==========]")]
    assert ran, tail + r.stderr[-2000:]
    n_ran = int(ran[-1].split()[1])
    assert "[  FAILED  ]" not in r.stdout, tail
    assert r.returncode == 0 and n_ran >= at_least, (r.returncode, n_ran, tail)
