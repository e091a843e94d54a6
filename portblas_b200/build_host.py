"""Build the C++ host-API callers against include/ + libpbx_gemm.so (plain g++, no nvcc):

  build/gemm_b200        samples/gemm_b200.cpp (this repo's own caller / self-check)
  build/ref_sample_gemm  the REFERENCE's samples/gemm.cpp compiled UNCHANGED from
                         /root/reference (only when that tree is present: it proves the drop-in
                         claim "existing callers relink unchanged"; the source is not copied)

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
    ref_src = REF / "samples" / "gemm.cpp"
    if ref_src.exists():
        exe = OUT / "ref_sample_gemm"
        _compile(ref_src, exe, extra_inc=[REF / "samples"])
        built.append(exe)
    return built


if __name__ == "__main__":
    for p in build():
        print(p)
