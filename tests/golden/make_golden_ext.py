"""Generate tests/golden/ext_golden.npz -- small fixed input/output vectors for _symm, _trsm and complex _gemm.

Same rationale as make_golden.py: the reference ships no golden vectors and cannot be run here, so the fixtures come
from the oracle the reference's own tests use (CBLAS symm / trsm / cgemm -- OpenBLAS through scipy) on seeded inputs over
shapes of the reference's grids (blas3_symm_test.cpp:155-209, blas3_trsm_test.cpp:130-163 with fill_trsm_matrix and NaN
in the unused triangle, blas3_gemm_test.cpp:143-259).  They pin the numpy restatements (oracle/blas3_ext.py) on CPU and
the CUDA path on GPU to the same committed numbers.

    python tests/golden/make_golden_ext.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import blas3_ext as ox  # noqa: E402
from oracle import oracle  # noqa: E402

SYMM = [  # dtype, side, uplo, m, n, alpha, beta, lda_mul, ldb_mul, ldc_mul
    ("f32", "l", "l", 11, 16, 1.5, 0.5, 1, 1, 1), ("f32", "r", "u", 32, 11, 1.5, 0.5, 1, 1, 1),
    ("f64", "l", "u", 63, 16, 1.0, 1.0, 2, 1, 2), ("f64", "r", "l", 16, 63, 1.0, 1.0, 1, 2, 2),
    ("f32", "l", "u", 72, 65, 1.0, 1.0, 1, 1, 1),
]
TRSM = [  # dtype, side, uplo, trans, diag, m, n, alpha, unused
    ("f32", "l", "l", "n", "n", 7, 7, 2.0, 0.0), ("f32", "r", "u", "t", "u", 7, 7, 2.0, float("nan")),
    ("f64", "l", "u", "t", "n", 72, 24, 2.0, float("nan")), ("f64", "r", "l", "n", "u", 24, 72, 2.0, float("nan")),
    ("f32", "l", "u", "n", "n", 136, 24, -0.5, float("nan")),
]
CGEMM = [  # dtype, ta, tb, m, n, k, alpha, beta, ld muls (a, b, c)
    ("c64", "n", "n", 11, 11, 16, 1.5 + 1j, 1.5 + 3j, (1, 1, 1)), ("c64", "t", "n", 33, 11, 17, 1.5 + 3j, 0j, (2, 2, 3)),
    ("c128", "n", "t", 11, 33, 17, 1.5 + 1j, 0j, (1, 1, 1)), ("c128", "t", "t", 33, 31, 40, 1 + 1.5j, 1.5 + 1j, (1, 1, 1)),
]


def main():
    out = {}
    rng = np.random.default_rng(12345)
    for i, (dt, side, uplo, m, n, al, be, la, lb, lc) in enumerate(SYMM):
        npdt = np.float64 if dt == "f64" else np.float32
        k = m if side == "l" else n
        lda, ldb, ldc = k * la, m * lb, m * lc
        A, B, C = (oracle.random_uniform(rng, s, npdt) for s in (k * lda, n * ldb, n * ldc))
        res = C.copy()
        ox.cblas_symm(side, uplo, m, n, al, A, lda, B, ldb, be, res, ldc)
        out[f"symm{i}_meta"] = np.array([dt, side, uplo, m, n, al, be, la, lb, lc]).astype(str)
        out[f"symm{i}_A"], out[f"symm{i}_B"], out[f"symm{i}_C"], out[f"symm{i}_out"] = A, B, C, res
    for i, (dt, side, uplo, tr, dg, m, n, al, unused) in enumerate(TRSM):
        npdt = np.float64 if dt == "f64" else np.float32
        k = m if side == "l" else n
        lda, ldb = 2 * k, 2 * m
        A = ox.fill_trsm_matrix(rng, k, lda, uplo, dg, float(rng.uniform(1, 10)), unused, npdt)
        B = oracle.random_uniform(rng, n * ldb, npdt)
        res = B.copy()
        ox.cblas_trsm(side, uplo, tr, dg, m, n, al, A, lda, res, ldb)
        out[f"trsm{i}_meta"] = np.array([dt, side, uplo, tr, dg, m, n, al]).astype(str)
        out[f"trsm{i}_A"], out[f"trsm{i}_B"], out[f"trsm{i}_out"] = A, B, res
    for i, (dt, ta, tb, m, n, k, al, be, (la, lb, lc)) in enumerate(CGEMM):
        npdt, rdt = (np.complex64, np.float32) if dt == "c64" else (np.complex128, np.float64)
        lda, ldb, ldc = (k if ta != "n" else m) * la, (n if tb != "n" else k) * lb, m * lc

        def rand(cnt):
            return (oracle.random_uniform(rng, cnt, rdt) + 1j * oracle.random_uniform(rng, cnt, rdt)).astype(npdt)
        A, B, C = rand(lda * (m if ta != "n" else k)), rand(ldb * (k if tb != "n" else n)), rand(ldc * n)
        res = C.copy()
        ox.cblas_cgemm(ta, tb, m, n, k, al, A, lda, B, ldb, be, res, ldc)
        out[f"cgemm{i}_meta"] = np.array([dt, ta, tb, m, n, k, la, lb, lc]).astype(str)
        out[f"cgemm{i}_scal"] = np.array([al, be], dtype=np.complex128)
        out[f"cgemm{i}_A"], out[f"cgemm{i}_B"], out[f"cgemm{i}_C"], out[f"cgemm{i}_out"] = A, B, C, res
    np.savez_compressed(Path(__file__).with_name("ext_golden.npz"), **out)
    print("wrote", len(SYMM), "symm,", len(TRSM), "trsm,", len(CGEMM), "cgemm cases")


if __name__ == "__main__":
    main()
