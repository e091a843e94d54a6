"""ctypes / numpy front for the CPU oracle (oracle/gemm_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import
this module.  Nothing under portblas_b200/ does.

Two oracles live here:
  * the C restatement of portBLAS's kernels (gemm_ref / gemm_local ordering, front-end rules,
    the DEFAULT-backend CPU kernel) -- see the header of gemm_oracle.c for file:line citations;
  * ``cblas_gemm``: the oracle the reference's own unit tests use, ``reference_blas::gemm`` =
    cblas_{s,d}gemm ColMajor (common/include/common/system_reference_blas.hpp:402-431; half is
    up-cast to float, run through sgemm and down-cast :410-430).  The CBLAS implementation here is
    numpy's bundled OpenBLAS 0.3.30 (the reference README asks for OpenBLAS >= 0.3.0).
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_LIB_PATH = HERE / "_build" / "liboracle.so"
_lib = None

STATUS_TEXT = {
    0: "ok",
    1: "invalid _TransA",
    2: "invalid _TransB",
    3: "invalid _stridec",
    4: "invalid _stridea",
    5: "invalid _strideb",
}


def build(force: bool = False) -> Path:
    src = HERE / "gemm_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        r = subprocess.run(["make", "-C", str(HERE), "-B" if force else "-s"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(_LIB_PATH))
        i64, i32 = ctypes.c_int64, ctypes.c_int
        for sfx, ct in (("f32", ctypes.c_float), ("f64", ctypes.c_double)):
            p = ctypes.POINTER(ct)
            f = getattr(_lib, f"oracle_gemm_frontend_{sfx}")
            f.restype = i32
            f.argtypes = [i32, ctypes.c_char, ctypes.c_char, i64, i64, i64, ct, p, i64, i64, p, i64, i64,
                          ct, p, i64, i64, i64, i32]
            g = getattr(_lib, f"oracle_gemm_default_cpu_{sfx}")
            g.restype = None
            g.argtypes = [i32, i32, i64, i64, i64, ct, p, i64, p, i64, ct, p, i64]
        _lib.oracle_compare_f64.restype = i64
        _lib.oracle_compare_f64.argtypes = [ctypes.POINTER(ctypes.c_double)] * 2 + [i64, i32, i32]
        _lib.oracle_num_threads.restype = i32
    return _lib


MODE_REF, MODE_LOCAL, MODE_TRUTH = 0, 1, 2


def _ptr(a: np.ndarray, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def gemm(transa: str, transb: str, m: int, n: int, k: int, alpha, A: np.ndarray, lda: int,
         B: np.ndarray, ldb: int, beta, C: np.ndarray, ldc: int, *, stridea: int = 0, strideb: int = 0,
         stridec: int = 0, batch: int = 1, interleaved: bool = False, mode: int = MODE_REF) -> int:
    """In-place on flat column-major buffer ``C``; returns the front-end status code.

    A, B, C are 1-D float32 or float64 arrays (the whole allocation, like the reference's
    device buffers).  For 16-bit storage types pass the values widened to float32 and round the
    result yourself (see ``round_to``)."""
    assert A.dtype == B.dtype == C.dtype and A.dtype in (np.float32, np.float64)
    assert A.flags.c_contiguous and B.flags.c_contiguous and C.flags.c_contiguous
    L = lib()
    if A.dtype == np.float32:
        f, ct = L.oracle_gemm_frontend_f32, ctypes.c_float
    else:
        f, ct = L.oracle_gemm_frontend_f64, ctypes.c_double
    return f(mode, transa.encode()[:1], transb.encode()[:1], m, n, k, ct(alpha), _ptr(A, ct), lda, stridea,
             _ptr(B, ct), ldb, strideb, ct(beta), _ptr(C, ct), ldc, stridec, batch, int(interleaved))


def gemm_default_cpu(ta: bool, tb: bool, m: int, n: int, k: int, alpha, A, lda, B, ldb, beta, C, ldc) -> None:
    """The DEFAULT-backend (CPU) kernel restated; OpenMP over work-groups (timed CPU baseline)."""
    L = lib()
    if A.dtype == np.float32:
        f, ct = L.oracle_gemm_default_cpu_f32, ctypes.c_float
    else:
        f, ct = L.oracle_gemm_default_cpu_f64, ctypes.c_double
    f(int(ta), int(tb), m, n, k, ct(alpha), _ptr(A, ct), lda, _ptr(B, ct), ldb, ct(beta), _ptr(C, ct), ldc)


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def compare(a: np.ndarray, b: np.ndarray, kind: str = "float", margin_mul: int = 1) -> int:
    """Number of elements failing the reference's ``almost_equal`` (float_comparison.hpp:163-188)."""
    a64 = np.ascontiguousarray(a, dtype=np.float64).ravel()
    b64 = np.ascontiguousarray(b, dtype=np.float64).ravel()
    assert a64.size == b64.size
    k = {"float": 0, "double": 1, "half": 2}[kind]
    d = ctypes.c_double
    return int(lib().oracle_compare_f64(_ptr(a64, d), _ptr(b64, d), a64.size, k, margin_mul))


# ---- the reference tests' own oracle: CBLAS ------------------------------------------------
def _view(buf: np.ndarray, rows: int, cols: int, ld: int, off: int = 0) -> np.ndarray:
    """rows x cols column-major window at element offset ``off`` with leading dimension ``ld``."""
    return np.lib.stride_tricks.as_strided(buf[off:], shape=(rows, cols),
                                           strides=(buf.itemsize, ld * buf.itemsize), writeable=True)


def cblas_gemm(transa: str, transb: str, m: int, n: int, k: int, alpha, A, lda, B, ldb, beta, C, ldc,
               *, stridea=0, strideb=0, stridec=0, batch=1) -> None:
    """reference_blas::gemm per batch entry (blas3_gemm_common.hpp:164-169), via numpy/OpenBLAS."""
    ta, tb = transa.lower() != "n", transb.lower() != "n"
    for b in range(batch):
        a = _view(A, k if ta else m, m if ta else k, lda, b * stridea)
        bb = _view(B, n if tb else k, k if tb else n, ldb, b * strideb)
        c = _view(C, m, n, ldc, b * stridec)
        opa = a.T if ta else a
        opb = bb.T if tb else bb
        prod = np.matmul(opa, opb)  # sgemm / dgemm
        c[...] = (alpha * prod + beta * c).astype(C.dtype) if beta != 0 else (alpha * prod).astype(C.dtype)


# ---- helpers shared by the tests ---------------------------------------------------------------
def round_to(x: np.ndarray, storage: str) -> np.ndarray:
    """Round float32 values to a 16-bit storage type and widen back (exactly representable)."""
    x = np.asarray(x, dtype=np.float32)
    if storage in ("f16", "half"):
        return x.astype(np.float16).astype(np.float32)
    if storage == "bf16":
        u = x.view(np.uint32).astype(np.uint64)
        rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16  # round-to-nearest-even
        return rounded.astype(np.uint32).view(np.float32)
    return x


def to_bf16_bits(x: np.ndarray) -> np.ndarray:
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return (((u + 0x7FFF + ((u >> 16) & 1)) >> 16) & 0xFFFF).astype(np.uint16)


def from_bf16_bits(b: np.ndarray) -> np.ndarray:
    return (np.asarray(b, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


def random_uniform(rng: np.random.Generator, size: int, dtype) -> np.ndarray:
    """U(-2, 5): the reference's fill_random (test/blas_test.hpp:138-140)."""
    return rng.uniform(-2.0, 5.0, size=size).astype(dtype)


def interleave(buf: np.ndarray, rows: int, cols: int, ld: int, batch: int, stride: int) -> np.ndarray:
    """Strided -> interleaved re-layout, as the reference's tests do on the host
    (test/unittest/blas3/blas3_gemm_common.hpp:55-69): element (r,c,b) -> (c*ld + r)*batch + b."""
    out = np.zeros(ld * cols * batch, dtype=buf.dtype)
    o = out.reshape(cols, ld, batch)
    for b in range(batch):
        o[:, :rows, b] = _view(buf, rows, cols, ld, b * stride).T
    return out


def deinterleave(ibuf: np.ndarray, rows: int, cols: int, ld: int, batch: int) -> np.ndarray:
    """Interleaved -> [batch][cols][ld] strided copy (padding rows zero)."""
    o = ibuf.reshape(cols, ld, batch)
    return np.ascontiguousarray(np.transpose(o, (2, 0, 1))).reshape(-1)
