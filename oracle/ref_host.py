"""ctypes front for oracle/_ref/libportblas_ref_<backend>.so -- the REFERENCE's own GEMM on the host.  TEST INFRASTRUCTURE ONLY.

What is inside the libraries: portBLAS's unmodified `blas::_gemm / _gemm_batched / _gemm_strided_batched` and the two
callers of that path, `blas::_symm` and `blas::_trsm` (include/interface/blas3_interface.h:86-146) with everything under
them -- `_gemm_backend`
(src/interface/gemm_interface.hpp:105-240), the backend heuristics (src/interface/blas3/backend/default.hpp or
nvidia_gpu.hpp or intel_gpu.hpp with GEMM_TALL_SKINNY_SUPPORT -- one library per backend header), `Gemm_Launcher` (src/interface/gemm_launcher.hpp:39-64), `SB_Handle::execute`
(src/sb_handle/portblas_handle.hpp:277-436), `execute_tree` (src/sb_handle/kernel_constructor.hpp:187-217) and the
`Gemm<>` kernels (src/operations/blas3/gemm_*.hpp) -- compiled from /root/reference by oracle/ref_host_driver.cpp
(`make -C oracle ref`) over a host stand-in for the SYCL runtime (oracle/sycl_host/sycl/sycl.hpp).  /root/reference only
exists where the libraries are BUILT; the built .so files travel to the GPU box.

Only tests/, __graft_entry__ (build + smoke check) and bench.py's reference / cpu_baseline leg import this module.
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"
REFERENCE = Path("/root/reference")
BACKENDS = ("default", "nvidia_gpu", "intel_gpu")
_libs: dict = {}

# suffix -> (numpy dtype of A/B, numpy dtype of C, ctypes type of alpha/beta)
TYPES = {
    "f32": (np.float32, np.float32, ctypes.c_float),
    "f64": (np.float64, np.float64, ctypes.c_double),
    "f16": (np.float16, np.float16, ctypes.c_float),
    "f16f32": (np.float16, np.float32, ctypes.c_float),
}


def lib_path(backend: str) -> Path:
    return REF_DIR / f"libportblas_ref_{backend}.so"


def build(force: bool = False) -> list:
    """Builds both libraries when the reference tree is present (idempotent); returns the libraries that exist."""
    if (REFERENCE / "src" / "interface" / "gemm_interface.hpp").exists():
        srcs = [HERE / "ref_host_driver.cpp", HERE / "Makefile", *(HERE / "sycl_host").rglob("*.hpp")]
        newest = max(p.stat().st_mtime for p in srcs)
        stale = force or any(not lib_path(b).exists() or lib_path(b).stat().st_mtime < newest for b in BACKENDS)
        if stale:
            r = subprocess.run(["make", "-C", str(HERE), "-j3", "-B", "ref"], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("reference host build failed:\n" + r.stdout[-3000:] + r.stderr[-3000:])
    return [lib_path(b) for b in BACKENDS if lib_path(b).exists()]


REF_TESTS = ("blas3_gemm_test", "blas3_gemm_batched_test", "blas3_gemm_tall_skinny_test", "blas3_symm_test",
             "blas3_trsm_test")


def unittest_path(name: str) -> Path:
    return REF_DIR / f"ref_unittest_host_{name}"


def build_unittests() -> list:
    """The reference's own unit tests linked with the reference's own (header-only) library on the host stand-in
    (`make -C oracle ref_tests`); built when the reference tree is present, returns the binaries that exist."""
    if (REFERENCE / "test" / "unittest" / "main.cpp").exists():
        newest = max(p.stat().st_mtime for p in [HERE / "Makefile", *(HERE / "sycl_host").rglob("*.hpp")])
        if any(not unittest_path(t).exists() or unittest_path(t).stat().st_mtime < newest for t in REF_TESTS):
            r = subprocess.run(["make", "-C", str(HERE), "-j5", "-B", "ref_tests"], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("reference unit tests (host) build failed:\n" + r.stdout[-3000:] + r.stderr[-3000:])
    return [unittest_path(t) for t in REF_TESTS if unittest_path(t).exists()]


def build_bench() -> list:
    """The reference's own bench_gemm and samples/gemm.cpp on its own library on the stand-in (`make -C oracle ref_bench`)."""
    outs = [REF_DIR / "ref_bench_host_gemm", REF_DIR / "ref_sample_gemm_host"]
    if (REFERENCE / "benchmark" / "portblas" / "main.cpp").exists():
        newest = max(p.stat().st_mtime for p in [HERE / "Makefile", *(HERE / "sycl_host").rglob("*.hpp")])
        if any(not o.exists() or o.stat().st_mtime < newest for o in outs):
            r = subprocess.run(["make", "-C", str(HERE), "-j2", "-B", "ref_bench"], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("reference benchmark (host) build failed:\n" + r.stdout[-3000:] + r.stderr[-3000:])
    return [o for o in outs if o.exists()]


def available(backend: str = "default") -> bool:
    return lib_path(backend).exists()


_usable: dict = {}


def usable(backend: str = "default") -> bool:
    """available() AND the library loads and multiplies correctly on THIS machine -- probed once in a child process, so
    that a prebuilt library that cannot run here (other CPU, missing runtime) costs a skip / a fallback, not the caller."""
    if backend not in _usable:
        ok = available(backend)
        if ok:
            import sys
            code = ("import numpy as np; from oracle import ref_host as r; "
                    "a = np.arange(1, 13, dtype=np.float32); b = np.arange(1, 13, dtype=np.float32)[::-1].copy(); "
                    "c = np.zeros(9, dtype=np.float32); "
                    f"r.gemm('n', 'n', 3, 3, 4, 1.0, a, 3, b, 4, 0.0, c, 3, backend={backend!r}); "
                    "want = (a.reshape(4, 3).T @ b.reshape(3, 4).T).T.ravel(); "
                    "raise SystemExit(0 if np.array_equal(c, want) else 3)")
            try:
                ok = subprocess.run([sys.executable, "-c", code], cwd=str(HERE.parent), capture_output=True,
                                    timeout=120).returncode == 0
            except Exception:
                ok = False
        _usable[backend] = ok
    return _usable[backend]


def lib(backend: str = "default") -> ctypes.CDLL:
    if backend not in _libs:
        if not available(backend):
            raise FileNotFoundError(f"{lib_path(backend)} is not built (make -C oracle ref; needs /root/reference)")
        L = ctypes.CDLL(str(lib_path(backend)))
        L.ref_last_error.restype = ctypes.c_char_p
        L.ref_backend.restype = ctypes.c_char_p
        L.ref_compute_units.restype = ctypes.c_int
        L.ref_set_fibers.argtypes = [ctypes.c_int]
        c, i, vp = ctypes.c_char, ctypes.c_int, ctypes.c_void_p
        for sfx, (_, _, ct) in TYPES.items():
            getattr(L, f"ref_gemm_{sfx}").argtypes = [c, c, i, i, i, ct, vp, i, vp, i, ct, vp, i]
            getattr(L, f"ref_gemm_batched_{sfx}").argtypes = [c, c, i, i, i, ct, vp, i, vp, i, ct, vp, i, i, i]
            getattr(L, f"ref_gemm_strided_batched_{sfx}").argtypes = [c, c, i, i, i, ct, vp, i, i, vp, i, i, ct, vp, i, i, i]
        for sfx, ct in (("f32", ctypes.c_float), ("f64", ctypes.c_double)):
            getattr(L, f"ref_symm_{sfx}").argtypes = [c, c, i, i, ct, vp, i, vp, i, ct, vp, i]
            getattr(L, f"ref_trsm_{sfx}").argtypes = [c, c, c, c, i, i, ct, vp, i, vp, i]
        for sfx, ct in (("c64", ctypes.c_float), ("c128", ctypes.c_double)):
            getattr(L, f"ref_gemm_{sfx}").argtypes = [c, c, i, i, i, ct, ct, vp, i, vp, i, ct, ct, vp, i]
            getattr(L, f"ref_gemm_strided_batched_{sfx}").argtypes = [c, c, i, i, i, ct, ct, vp, i, i, vp, i, i, ct, ct, vp,
                                                                      i, i, i]
        assert L.ref_backend().decode() == backend
        _libs[backend] = L
    return _libs[backend]


class ReferenceError_(Exception):
    """std::invalid_argument (rc 1) or another exception (rc 2) thrown by the reference; the text is its what()."""


def _check(L, rc):
    if rc:
        raise ReferenceError_(L.ref_last_error().decode())


_CPLX = {np.dtype(np.complex64): "c64", np.dtype(np.complex128): "c128"}


def _suffix(A: np.ndarray, C: np.ndarray) -> str:
    for sfx, (ti, to, _) in TYPES.items():
        if A.dtype == ti and C.dtype == to:
            return sfx
    raise TypeError(f"no reference instantiation for ({A.dtype}, {C.dtype})")


def _b(ch: str) -> bytes:
    return ch.encode()[:1]


def gemm(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, *, backend="default") -> None:
    """blas::_gemm on flat column-major numpy buffers, C in place (complex64 / complex128 buffers take the reference's
    complex instantiations, BLAS_ENABLE_COMPLEX)."""
    L = lib(backend)
    if C.dtype in _CPLX:
        al, be = complex(alpha), complex(beta)
        return _check(L, getattr(L, f"ref_gemm_{_CPLX[C.dtype]}")(_b(ta), _b(tb), m, n, k, al.real, al.imag, A.ctypes.data,
                                                                 lda, B.ctypes.data, ldb, be.real, be.imag,
                                                                 C.ctypes.data, ldc))
    sfx = _suffix(A, C)
    _check(L, getattr(L, f"ref_gemm_{sfx}")(_b(ta), _b(tb), m, n, k, alpha, A.ctypes.data, lda, B.ctypes.data, ldb,
                                            beta, C.ctypes.data, ldc))


def gemm_batched(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, batch, interleaved=False, *,
                 backend="default") -> None:
    """blas::_gemm_batched (default strides = matrix footprints; batch_type strided 0 / interleaved 1)."""
    L = lib(backend)
    sfx = _suffix(A, C)
    _check(L, getattr(L, f"ref_gemm_batched_{sfx}")(_b(ta), _b(tb), m, n, k, alpha, A.ctypes.data, lda, B.ctypes.data,
                                                    ldb, beta, C.ctypes.data, ldc, batch, 1 if interleaved else 0))


def gemm_strided_batched(ta, tb, m, n, k, alpha, A, lda, stride_a, B, ldb, stride_b, beta, C, ldc, stride_c, batch, *,
                         backend="default") -> None:
    """blas::_gemm_strided_batched."""
    L = lib(backend)
    if C.dtype in _CPLX:
        al, be = complex(alpha), complex(beta)
        return _check(L, getattr(L, f"ref_gemm_strided_batched_{_CPLX[C.dtype]}")(
            _b(ta), _b(tb), m, n, k, al.real, al.imag, A.ctypes.data, lda, stride_a, B.ctypes.data, ldb, stride_b, be.real,
            be.imag, C.ctypes.data, ldc, stride_c, batch))
    sfx = _suffix(A, C)
    _check(L, getattr(L, f"ref_gemm_strided_batched_{sfx}")(_b(ta), _b(tb), m, n, k, alpha, A.ctypes.data, lda, stride_a,
                                                            B.ctypes.data, ldb, stride_b, beta, C.ctypes.data, ldc,
                                                            stride_c, batch))


def symm(side, uplo, m, n, alpha, A, lda, B, ldb, beta, C, ldc, *, backend="default") -> None:
    """blas::_symm (src/interface/symm_interface.hpp:35-71), C in place."""
    L = lib(backend)
    sfx = _suffix(A, C)
    _check(L, getattr(L, f"ref_symm_{sfx}")(_b(side), _b(uplo), m, n, alpha, A.ctypes.data, lda, B.ctypes.data, ldb, beta,
                                            C.ctypes.data, ldc))


def trsm(side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb, *, backend="default") -> None:
    """blas::_trsm (src/interface/trsm_interface.hpp:150-387), B in place."""
    L = lib(backend)
    sfx = _suffix(A, B)
    _check(L, getattr(L, f"ref_trsm_{sfx}")(_b(side), _b(uplo), _b(trans), _b(diag), m, n, alpha, A.ctypes.data, lda,
                                            B.ctypes.data, ldb))


def set_fibers(on: bool, backend: str = "default") -> None:
    """False: work-items of barrier-free kernels run as plain loop iterations (every non-symm GEMM of the default backend
    is barrier-free: default.hpp:52-147); a barrier reached in that mode aborts.  Used only for timing."""
    lib(backend).ref_set_fibers(1 if on else 0)


def compute_units(backend: str = "default") -> int:
    return lib(backend).ref_compute_units()
