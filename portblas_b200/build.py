"""Build libpbx_gemm.so (the C-ABI GEMM library) in-tree with nvcc for sm_100a.

    python -m portblas_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so lands next to this file so that it
travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "libpbx_gemm.so"
OBJ_DIR = HERE / "csrc" / "_obj"
SOURCES = ["gemm_tc_inst_f32_pre0.cu", "gemm_tc_inst_f32_pre1.cu", "gemm_tc_inst_f32_pre2.cu", "gemm_tc_inst_f32_pre3.cu", "gemm_tc_inst_f32_pre4.cu",
           "gemm_tc_inst_f16.cu", "gemm_tc_inst_f16f32.cu", "gemm_tc_inst_bf16.cu", "gemm_tc_inst_bf16f32.cu",
           "gemm_simt.cu", "gemm_dmma.cu", "blas3_ext.cu", "pbx_api.cu", "pbx_host.cu", "pbx_multi.cu", "gemm_tcgen05.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr", "--extended-lambda",
]


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "pbx_gemm.h"]
    stamp = OBJ_DIR / "stamp.txt"
    want = _digest(deps)
    if not force and OUT.exists() and stamp.exists() and stamp.read_text() == want:
        return OUT
    OBJ_DIR.mkdir(parents=True, exist_ok=True)

    def compile_one(src: str) -> str:
        obj = OBJ_DIR / (src + ".o")
        cmd = [NVCC, *FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return str(obj)

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [NVCC, "-shared", "-o", str(OUT), *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(want)
    return OUT


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
