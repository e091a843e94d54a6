"""Generate tests/golden/ref_host_golden.npz -- outputs of the REFERENCE ITSELF, run here.

oracle/_ref/libportblas_ref_<backend>.so is portBLAS's own GEMM path (blas::_gemm down to the Gemm<> kernels) compiled
unchanged from /root/reference over a host stand-in for the SYCL runtime (oracle/ref_host_driver.cpp,
oracle/sycl_host/sycl/sycl.hpp; `make -C oracle ref`).  This script feeds it seeded U(-2,5) inputs over shapes from the
reference's test grids and stores inputs + outputs; the CPU test (tests/test_oracle_ref.py::test_ref_host_golden) checks
the C restatement against them bit for bit, the GPU test (tests/test_ref_parity_gpu.py) checks the CUDA path against them
within the north-star tolerances.  Needs /root/reference (to build oracle/_ref) -- the fixtures do not.

    python tests/golden/make_ref_host_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle, ref_host  # noqa: E402

CASES = [
    # (backend, dtype, ta, tb, m, n, k, alpha, beta, lda_mul, ldb_mul, ldc_mul, batch)
    ("default", "f32", "n", "n", 7, 5, 9, 1.5, 0.5, 1, 1, 1, 1),          # samples/gemm.cpp shape
    ("default", "f32", "t", "n", 11, 16, 17, 1.5, 1.5, 1, 1, 1, 1),       # SmallBetaNonZeroLDMatch, Tile<2,2,2,2>
    ("default", "f32", "n", "t", 32, 11, 17, 1.5, 0.0, 2, 3, 4, 1),       # SmallBetaZeroLDMultiplied
    ("default", "f32", "t", "t", 131, 65, 35, 1.0, 1.0, 2, 2, 2, 1) ,     # Tile<4,4,8,8> full-vec, ragged
    # (Tile<4,4,4,4>, M*N >= 524288 -- configs[0]'s kernel -- needs 2 MiB per matrix: checked live, not stored:
    #  tests/test_oracle_ref.py::test_default_cpu_port_is_bit_exact_with_reference)
    ("default", "f64", "n", "n", 16, 16, 17, 1.5, 1.5, 1, 1, 1, 1),
    ("default", "f64", "t", "n", 7, 9, 257, 1.5, 0.5, 2, 3, 4, 1),
    ("default", "f64", "n", "t", 15, 17, 32, 3.0, 7.0, 2, 3, 4, 3),       # AllStridedBatched scalars / ld muls
    ("nvidia_gpu", "f32", "n", "n", 63, 63, 63, 1.5, 1.5, 1, 1, 1, 1),    # local-memory kernel, M,N <= 256
    ("nvidia_gpu", "f32", "t", "n", 260, 24, 65, 1.5, 0.0, 1, 2, 1, 1),   # local-memory kernel, M,N <= 1024
    ("nvidia_gpu", "f64", "n", "t", 70, 60, 64, 2.0, 3.0, 3, 1, 2, 1),
    ("nvidia_gpu", "f32", "n", "n", 15, 32, 15, 3.0, 7.0, 1, 1, 1, 3),    # BatchGemm BetaNonZeroLDMatch, batch_size > 1 tile
    ("nvidia_gpu", "f64", "t", "t", 33, 31, 40, 1.0, 1.0, 2, 3, 4, 2),
]


def main():
    ref_host.build()
    out = {}
    rng = np.random.default_rng(54321)
    for i, (backend, dt, ta, tb, m, n, k, al, be, la, lb, lc, batch) in enumerate(CASES):
        npdt = np.float64 if dt == "f64" else np.float32
        lda = (k if ta == "t" else m) * la
        ldb = (n if tb == "t" else k) * lb
        ldc = m * lc
        sa, sb, sc = m * k * la, k * n * lb, m * n * lc
        A = oracle.random_uniform(rng, sa * batch, npdt)
        B = oracle.random_uniform(rng, sb * batch, npdt)
        C = oracle.random_uniform(rng, sc * batch, npdt)
        out_c = C.copy()
        if batch > 1:
            ref_host.gemm_strided_batched(ta, tb, m, n, k, al, A, lda, sa, B, ldb, sb, be, out_c, ldc, sc, batch,
                                          backend=backend)
        else:
            ref_host.gemm(ta, tb, m, n, k, al, A, lda, B, ldb, be, out_c, ldc, backend=backend)
        out[f"case{i}_meta"] = np.array([backend, dt, ta, tb, m, n, k, al, be, la, lb, lc, batch], dtype=object).astype(str)
        out[f"case{i}_A"], out[f"case{i}_B"], out[f"case{i}_C"], out[f"case{i}_out"] = A, B, C, out_c
    np.savez_compressed(Path(__file__).with_name("ref_host_golden.npz"), **out)
    print("wrote", len(CASES), "cases")


if __name__ == "__main__":
    main()
