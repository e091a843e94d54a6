"""Generate tests/golden/gemm_golden.npz -- small fixed input/output vectors for the GEMM path.

The reference ships no golden vectors (its tests draw random inputs and compare with a system
CBLAS at run time, test/unittest/blas3/blas3_gemm_common.hpp:164-169,223-225) and cannot be run
in this image (SYCL).  These fixtures are therefore produced by the oracle the reference's tests
use -- CBLAS, here numpy's OpenBLAS 0.3.30 -- on seeded U(-2,5) inputs over shapes taken from the
reference's test grids.  They pin the C restatement (oracle/gemm_oracle.c) on CPU and the CUDA path
on GPU to the same committed numbers.

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle  # noqa: E402

CASES = [
    # (dtype, ta, tb, m, n, k, alpha, beta, lda_mul, ldb_mul, ldc_mul, batch)
    ("f32", "n", "n", 7, 5, 9, 1.5, 0.5, 1, 1, 1, 1),        # samples/gemm.cpp shape
    ("f32", "t", "n", 11, 16, 17, 1.5, 1.5, 1, 1, 1, 1),     # SmallBetaNonZeroLDMatch
    ("f32", "n", "t", 32, 11, 17, 1.5, 0.0, 2, 3, 4, 1),     # SmallBetaZeroLDMultiplied
    ("f32", "t", "t", 31, 33, 35, 1.0, 1.0, 2, 2, 2, 1),     # OffsetNonZero-like odd shapes
    ("f64", "n", "n", 16, 16, 17, 1.5, 1.5, 1, 1, 1, 1),
    ("f64", "t", "n", 7, 9, 257, 1.5, 0.5, 2, 3, 4, 1),      # TallSkinny m,n (7x9) with reduced k
    ("f64", "n", "t", 15, 17, 32, 3.0, 7.0, 2, 3, 4, 3),     # AllStridedBatched scalars / ld muls, reduced
    ("f32", "n", "n", 15, 32, 15, 3.0, 7.0, 1, 1, 1, 3),     # BatchGemm BetaNonZeroLDMatch scalars, reduced
    ("f16", "n", "n", 32, 32, 16, 1.5, 1.5, 1, 1, 1, 1),
    ("bf16", "t", "n", 64, 32, 64, 1.0, 0.0, 1, 1, 1, 1),
]


def main():
    out = {}
    rng = np.random.default_rng(12345)
    for i, (dt, ta, tb, m, n, k, al, be, la, lb, lc, batch) in enumerate(CASES):
        npdt = np.float64 if dt == "f64" else np.float32
        lda = (k if ta == "t" else m) * la
        ldb = (n if tb == "t" else k) * lb
        ldc = m * lc
        sa, sb, sc = m * k * la, k * n * lb, m * n * lc
        A = oracle.random_uniform(rng, sa * batch, npdt)
        B = oracle.random_uniform(rng, sb * batch, npdt)
        C = oracle.random_uniform(rng, sc * batch, npdt)
        if dt in ("f16", "bf16"):
            A, B, C = (oracle.round_to(x, dt) for x in (A, B, C))
        out_c = C.copy()
        oracle.cblas_gemm(ta, tb, m, n, k, al, A, lda, B, ldb, be, out_c, ldc, stridea=sa, strideb=sb, stridec=sc,
                          batch=batch)
        if dt in ("f16", "bf16"):
            out_c = oracle.round_to(out_c, dt)
        out[f"case{i}_meta"] = np.array([dt, ta, tb, m, n, k, al, be, la, lb, lc, batch], dtype=object).astype(str)
        out[f"case{i}_A"], out[f"case{i}_B"], out[f"case{i}_C"], out[f"case{i}_out"] = A, B, C, out_c
    np.savez_compressed(Path(__file__).with_name("gemm_golden.npz"), **out)
    print("wrote", len(CASES), "cases")


if __name__ == "__main__":
    main()
