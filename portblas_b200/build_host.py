"""Build the C++ host-API callers against include/ + libpbx_gemm.so (plain g++, no nvcc):

  build/gemm_b200        samples/gemm_b200.cpp (this repo's own caller / self-check)
  build/gemm_multi_b200  samples/gemm_multi_b200.cpp (one SGEMM over all B200s of the box through blas::multi)
  build/ref_sample_gemm  the REFERENCE's samples/gemm.cpp compiled UNCHANGED from
                         /root/reference (only when that tree is present: it proves the drop-in
                         claim "existing callers relink unchanged"; the source is not copied)
  build/ref_unittest_*   the REFERENCE's own unit tests -- test/unittest/main.cpp + test/unittest/blas3/
                         {blas3_gemm,blas3_gemm_batched,blas3_gemm_tall_skinny,blas3_symm,blas3_trsm}_test.cpp --
                         compiled UNCHANGED from /root/reference with the compile definitions its CMake uses
                         (test/unittest/CMakeLists.txt:109-133), against include/, a minimal GoogleTest stand-in
                         (tests/cpp/shim/gtest/gtest.h), the reference's vendored cblas.h / clara.hpp and the
                         OpenBLAS of this image as the CBLAS the tests compare with

  build/ref_unittest_joint_matrix_*  the REFERENCE's joint_matrix unit tests (test/unittest/joint_matrix/*.cpp), which call the
                         path's seam blas::Gemm_Launcher<...>::_select_gemm directly (include/interface/gemm_launcher.h)
  build/ref_bench_*      the REFERENCE's own benchmark harness -- benchmark/portblas/main.cpp + benchmark/portblas/blas3/
                         {gemm,gemm_batched,gemm_batched_strided,symm,trsm}.cpp -- compiled UNCHANGED from
                         /root/reference with the definitions of benchmark/portblas/CMakeLists.txt:99-133
                         (BLAS_VERIFY_BENCHMARK on, the reference's default: every benchmark first checks its result
                         against CBLAS), against include/, a minimal Google Benchmark stand-in
                         (tests/cpp/shim/benchmark/benchmark.h: console + JSON reports, --benchmark_filter/min_time/out)
                         and tests/cpp/shim/bench_info_stub.cc for the two labels the reference generates at configure time

    python -m portblas_b200.build_host
"""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
OUT = ROOT / "build"
CXX = "/usr/bin/g++"
REF = Path("/root/reference")


def _compile(src: Path, exe: Path, extra_inc=()) -> None:
    cmd = [CXX, "-std=c++17", "-O2", "-I", str(ROOT / "include"), "-I", "/usr/local/cuda/include"]
    for inc in extra_inc:
        cmd += ["-I", str(inc)]
    cmd += [str(src), "-o", str(exe), "-L", str(HERE), "-lpbx_gemm", f"-Wl,-rpath,{HERE}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"host build failed for {src}:\n{r.stdout}\n{r.stderr}")


SCIPY_LIBS = Path(sys.prefix) / "lib" / f"python{sys.version_info.major}.{sys.version_info.minor}" / "site-packages" / "scipy.libs"

# reference unit-test sources -> extra compile definitions (test/unittest/CMakeLists.txt:109-133: half only for the
# HALF_DATA_OPS list, complex only for executables whose name contains "gemm")
REF_UNITTESTS = {
    "blas3_gemm_test": ["-DBLAS_ENABLE_HALF=1", "-DBLAS_ENABLE_COMPLEX=1"],
    "blas3_gemm_batched_test": ["-DBLAS_ENABLE_HALF=1", "-DBLAS_ENABLE_COMPLEX=1"],
    "blas3_gemm_tall_skinny_test": ["-DBLAS_ENABLE_COMPLEX=1"],
    "blas3_symm_test": [],
    "blas3_trsm_test": [],
}


def build_reference_unittests() -> list:
    """The reference's own blas3 unit tests, unchanged, against this repository's headers and library."""
    openblas = sorted(SCIPY_LIBS.glob("libscipy_openblas*.so"))
    if not (REF / "test" / "unittest" / "main.cpp").exists() or not openblas:
        return []
    shim = ROOT / "tests" / "cpp" / "shim"
    built = []
    # up to date = newer than every source and header; libpbx_gemm.so is linked dynamically (rpath), so a rebuilt library
    # with the same include/pbx_gemm.h needs no relink
    for name, defs in REF_UNITTESTS.items():
        exe = OUT / f"ref_unittest_{name}"
        srcs = [REF / "test" / "unittest" / "main.cpp", REF / "test" / "unittest" / "blas3" / f"{name}.cpp"]
        newest = max(p.stat().st_mtime for p in [*srcs, *ROOT.glob("include/**/*.h*"), *shim.rglob("*.h")])
        if exe.exists() and exe.stat().st_mtime >= newest:
            built.append(exe)
            continue
        cmd = [CXX, "-std=c++17", "-O1", "-I", str(ROOT / "include"), "-I", "/usr/local/cuda/include", "-I", str(shim),
               "-I", str(REF / "test"), "-I", str(REF / "common" / "include"), "-I", str(REF / "external" / "cblas" / "include"),
               "-I", str(REF / "external" / "clara" / "include"), "-include", str(shim / "cblas_scipy_rename.h"),
               "-DBLAS_INDEX_T=int", "-DBLAS_DATA_TYPE_DOUBLE", "-DSB_ENABLE_USM", *defs, *[str(s) for s in srcs],
               "-o", str(exe), "-L", str(HERE), "-lpbx_gemm", f"-Wl,-rpath,{HERE}", "-L", str(SCIPY_LIBS),
               f"-l:{openblas[0].name}", f"-Wl,-rpath,{SCIPY_LIBS}"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"reference unit test {name} does not build against include/:\n{r.stderr[-4000:]}")
        built.append(exe)
    return built


# reference benchmark sources -> extra compile definitions (benchmark/portblas/CMakeLists.txt:99-133: complex for
# CPLX_OPS, half for HALF_DATA_OPS)
REF_BENCHMARKS = {
    "gemm": ["-DBLAS_ENABLE_HALF=1", "-DBLAS_ENABLE_COMPLEX=1"],
    "gemm_batched": ["-DBLAS_ENABLE_HALF=1", "-DBLAS_ENABLE_COMPLEX=1"],
    "gemm_batched_strided": ["-DBLAS_ENABLE_HALF=1", "-DBLAS_ENABLE_COMPLEX=1"],
    # symm.cpp registers a lambda that captures loop locals by reference (upstream bug): run at registration, see the shim
    "symm": ["-DBENCHMARK_SHIM_RUN_AT_REGISTRATION"],
    "trsm": [],
}


# the reference's joint_matrix unit tests (test/unittest/joint_matrix/CMakeLists.txt:41-52): they reach the GEMM path through
# its seam, blas::Gemm_Launcher<...>::_select_gemm (launch_gemm.hpp), not through blas::_gemm
REF_JOINT_MATRIX_TESTS = ["half_half_16_16_16", "half_half_32_8_16", "half_half_8_32_16", "half_float_16_16_16",
                          "half_float_32_8_16", "half_float_8_32_16", "bfloat16_float_16_16_16", "bfloat16_float_32_8_16",
                          "bfloat16_float_8_32_16", "tf32_float_16_16_8"]


def build_reference_joint_matrix_tests() -> list:
    """test/unittest/main.cpp + test/unittest/joint_matrix/<name>.cpp, unchanged, against include/ (whose
    interface/gemm_launcher.h keeps the seam's full template signature) -> build/ref_unittest_joint_matrix_<name>."""
    from concurrent.futures import ThreadPoolExecutor
    openblas = sorted(SCIPY_LIBS.glob("libscipy_openblas*.so"))
    jdir = REF / "test" / "unittest" / "joint_matrix"
    if not (jdir / "launch_gemm.hpp").exists() or not openblas:
        return []
    shim = ROOT / "tests" / "cpp" / "shim"

    def one(name: str) -> Path:
        exe = OUT / f"ref_unittest_joint_matrix_{name}"
        srcs = [REF / "test" / "unittest" / "main.cpp", jdir / f"{name}.cpp"]
        newest = max(p.stat().st_mtime for p in [*srcs, *ROOT.glob("include/**/*.h*"), *shim.rglob("*.h")])
        if exe.exists() and exe.stat().st_mtime >= newest:
            return exe
        cmd = [CXX, "-std=c++17", "-O1", "-I", str(ROOT / "include"), "-I", "/usr/local/cuda/include", "-I", str(shim),
               "-I", str(REF / "test"), "-I", str(jdir), "-I", str(REF / "common" / "include"),
               "-I", str(REF / "external" / "cblas" / "include"), "-I", str(REF / "external" / "clara" / "include"),
               "-include", str(shim / "cblas_scipy_rename.h"), "-DBLAS_INDEX_T=int", "-DBLAS_DATA_TYPE_DOUBLE",
               "-DSB_ENABLE_USM", *[str(s) for s in srcs], "-o", str(exe), "-L", str(HERE), "-lpbx_gemm",
               f"-Wl,-rpath,{HERE}", "-L", str(SCIPY_LIBS), f"-l:{openblas[0].name}", f"-Wl,-rpath,{SCIPY_LIBS}"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"reference joint_matrix test {name} does not build against include/:\n{r.stderr[-4000:]}")
        return exe

    with ThreadPoolExecutor(max_workers=5) as ex:
        return list(ex.map(one, REF_JOINT_MATRIX_TESTS))


def build_reference_benchmarks() -> list:
    """The reference's own benchmark executables (bench_gemm ...), unchanged, against this repository's headers and library."""
    openblas = sorted(SCIPY_LIBS.glob("libscipy_openblas*.so"))
    bdir = REF / "benchmark" / "portblas"
    if not (bdir / "main.cpp").exists() or not openblas:
        return []
    shim = ROOT / "tests" / "cpp" / "shim"
    built = []
    for name, defs in REF_BENCHMARKS.items():
        exe = OUT / f"ref_bench_{name}"
        srcs = [bdir / "main.cpp", bdir / "blas3" / f"{name}.cpp", shim / "bench_info_stub.cc"]
        newest = max(p.stat().st_mtime for p in [*srcs, *ROOT.glob("include/**/*.h*"), *shim.rglob("*.h")])
        if exe.exists() and exe.stat().st_mtime >= newest:
            built.append(exe)
            continue
        cmd = [CXX, "-std=c++17", "-O1", "-I", str(ROOT / "include"), "-I", "/usr/local/cuda/include", "-I", str(shim),
               "-I", str(bdir), "-I", str(REF / "common" / "include"), "-I", str(REF / "external" / "cblas" / "include"),
               "-I", str(REF / "external" / "clara" / "include"), "-include", str(shim / "cblas_scipy_rename.h"),
               "-DBLAS_INDEX_T=int", "-DBLAS_DATA_TYPE_DOUBLE", "-DSB_ENABLE_USM", "-DBLAS_VERIFY_BENCHMARK", *defs,
               *[str(s) for s in srcs], "-o", str(exe), "-L", str(HERE), "-lpbx_gemm", f"-Wl,-rpath,{HERE}",
               "-L", str(SCIPY_LIBS), f"-l:{openblas[0].name}", f"-Wl,-rpath,{SCIPY_LIBS}"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"reference benchmark {name} does not build against include/:\n{r.stderr[-4000:]}")
        built.append(exe)
    return built


def build() -> list:
    from . import build as libbuild
    libbuild.build()
    OUT.mkdir(exist_ok=True)
    built = []
    exe = OUT / "gemm_b200"
    src = ROOT / "samples" / "gemm_b200.cpp"
    if not exe.exists() or exe.stat().st_mtime < max(p.stat().st_mtime for p in [src, *ROOT.glob("include/**/*.h*")]):
        _compile(src, exe)
    built.append(exe)
    exe = OUT / "gemm_multi_b200"
    src = ROOT / "samples" / "gemm_multi_b200.cpp"
    if not exe.exists() or exe.stat().st_mtime < max(p.stat().st_mtime for p in [src, *ROOT.glob("include/**/*.h*")]):
        _compile(src, exe)
    built.append(exe)
    ref_src = REF / "samples" / "gemm.cpp"
    if ref_src.exists():
        exe = OUT / "ref_sample_gemm"
        _compile(ref_src, exe, extra_inc=[REF / "samples"])
        built.append(exe)
    built += build_reference_unittests()
    built += build_reference_joint_matrix_tests()
    built += build_reference_benchmarks()
    return built


if __name__ == "__main__":
    for p in build():
        print(p)
