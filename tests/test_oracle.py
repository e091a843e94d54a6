"""CPU tests of the oracle (oracle/gemm_oracle.c): the restatement of portBLAS's GEMM semantics
is pinned to the oracle the reference's own tests use (CBLAS; here numpy's OpenBLAS) on the
reference's parameter grids, with the reference's comparison predicate, and to the committed
golden fixtures.  No GPU involved."""
import itertools
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle

TRANS = [("n", "n"), ("n", "t"), ("t", "n"), ("t", "t")]
GOLDEN = Path(__file__).parent / "golden" / "gemm_golden.npz"


def _case(rng, npdt, ta, tb, m, n, k, lam=1, lbm=1, lcm=1, batch=1, off=0):
    lda = (k if ta == "t" else m) * lam
    ldb = (n if tb == "t" else k) * lbm
    ldc = m * lcm
    sa, sb, sc = m * k * lam, k * n * lbm, m * n * lcm
    A = oracle.random_uniform(rng, sa * batch + off, npdt)
    B = oracle.random_uniform(rng, sb * batch + off, npdt)
    C = oracle.random_uniform(rng, sc * batch + off, npdt)
    return A, B, C, lda, ldb, ldc, sa, sb, sc


@pytest.mark.parametrize("npdt,kind", [(np.float32, "float"), (np.float64, "double")])
def test_restatement_matches_cblas_on_reference_small_grids(npdt, kind):
    """Gemm/Small* and OffsetNonZero grids (blas3_gemm_test.cpp:30-122)."""
    rng = np.random.default_rng(12345)
    n_cases = 0
    for (ta, tb), m, n, k, (al, be), lds, off in itertools.product(
            TRANS, [11, 16, 32, 63], [11, 16, 63], [16, 17, 63], [(1.5, 1.5), (1.5, 0.0), (1.0, 1.0)],
            [(1, 1, 1), (2, 3, 4)], [0, 10]):
        A, B, C, lda, ldb, ldc, *_ = _case(rng, npdt, ta, tb, m, n, k, *lds, off=off)
        want = C.copy()
        oracle.cblas_gemm(ta, tb, m, n, k, al, A[off:], lda, B[off:], ldb, be, want[off:], ldc)
        for mode in (oracle.MODE_REF, oracle.MODE_LOCAL, oracle.MODE_TRUTH):
            got = C.copy()
            assert oracle.gemm(ta, tb, m, n, k, al, A[off:], lda, B[off:], ldb, be, got[off:], ldc, mode=mode) == 0
            assert oracle.compare(got, want, kind) == 0  # whole buffer: padding and offset prefix untouched
        n_cases += 1
    assert n_cases > 1000


def test_restatement_matches_cblas_large_and_batched():
    """LargeBetaNonZeroLDMatch (reduced), BatchGemm / BatchStridedGemm grids
    (blas3_gemm_test.cpp:125-141, blas3_gemm_batched_test.cpp:30-147)."""
    rng = np.random.default_rng(7)
    for (ta, tb), (m, n, k) in itertools.product(TRANS, [(253, 257, 253), (511, 257, 511)]):
        A, B, C, lda, ldb, ldc, *_ = _case(rng, np.float32, ta, tb, m, n, k)
        want, got = C.copy(), C.copy()
        oracle.cblas_gemm(ta, tb, m, n, k, 1.0, A, lda, B, ldb, 1.0, want, ldc)
        oracle.gemm(ta, tb, m, n, k, 1.0, A, lda, B, ldb, 1.0, got, ldc, mode=oracle.MODE_LOCAL)
        assert oracle.compare(got, want, "float") == 0
    for (ta, tb), sam, sbm, scm in itertools.product(TRANS, [0, 1, 2], [0, 1, 2], [1, 3]):
        m, n, k, batch = 63, 63, 128, 5
        A, B, C, lda, ldb, ldc, sa, sb, sc = _case(rng, np.float64, ta, tb, m, n, k, 2, 3, 4, batch=3 * batch)
        want, got = C.copy(), C.copy()
        oracle.cblas_gemm(ta, tb, m, n, k, 3.0, A, lda, B, ldb, 7.0, want, ldc, stridea=sa * sam, strideb=sb * sbm,
                          stridec=sc * scm, batch=batch)
        assert oracle.gemm(ta, tb, m, n, k, 3.0, A, lda, B, ldb, 7.0, got, ldc, stridea=sa * sam, strideb=sb * sbm,
                           stridec=sc * scm, batch=batch) == 0
        assert oracle.compare(got, want, "double") == 0


def test_interleaved_layout_equals_strided():
    """gemm_interleaved.hpp:265-271 + the host re-layout of blas3_gemm_common.hpp:55-69."""
    rng = np.random.default_rng(3)
    for (ta, tb), (m, n, k) in itertools.product(TRANS, [(63, 40, 17), (5, 129, 33)]):
        batch = 5
        A, B, C, lda, ldb, ldc, sa, sb, sc = _case(rng, np.float32, ta, tb, m, n, k, 2, 3, 4, batch=batch)
        want = C.copy()
        oracle.gemm(ta, tb, m, n, k, 3.0, A, lda, B, ldb, 7.0, want, ldc, stridea=sa, strideb=sb, stridec=sc,
                    batch=batch)
        a_rows, a_cols = (k, m) if ta == "t" else (m, k)
        b_rows, b_cols = (n, k) if tb == "t" else (k, n)
        Ai = oracle.interleave(A, a_rows, a_cols, lda, batch, sa)
        Bi = oracle.interleave(B, b_rows, b_cols, ldb, batch, sb)
        Ci = oracle.interleave(C, m, n, ldc, batch, sc)
        assert oracle.gemm(ta, tb, m, n, k, 3.0, Ai, lda, Bi, ldb, 7.0, Ci, ldc, batch=batch, interleaved=True) == 0
        back = oracle.deinterleave(Ci, m, n, ldc, batch)
        w = want.reshape(batch, n, ldc)[:, :, :m]
        assert np.array_equal(back.reshape(batch, n, ldc)[:, :, :m], w)


def test_front_end_rules():
    """gemm_interface.hpp:105-185: alpha==0 first, trans / stride validation, 'c' == 't'."""
    rng = np.random.default_rng(5)
    A, B, C, lda, ldb, ldc, sa, sb, sc = _case(rng, np.float32, "n", "n", 16, 16, 17, 2, 2, 2, batch=3)
    # invalid arguments -> the reference's exception texts
    assert oracle.STATUS_TEXT[oracle.gemm("x", "n", 16, 16, 17, 1.0, A, lda, B, ldb, 0.0, C.copy(), ldc)] == "invalid _TransA"
    assert oracle.STATUS_TEXT[oracle.gemm("n", "y", 16, 16, 17, 1.0, A, lda, B, ldb, 0.0, C.copy(), ldc)] == "invalid _TransB"
    assert oracle.STATUS_TEXT[oracle.gemm("n", "n", 16, 16, 17, 1.0, A, lda, B, ldb, 0.0, C.copy(), ldc, stridea=sa,
                                          strideb=sb, stridec=ldc * 16 - 1, batch=3)] == "invalid _stridec"
    assert oracle.STATUS_TEXT[oracle.gemm("n", "n", 16, 16, 17, 1.0, A, lda, B, ldb, 0.0, C.copy(), ldc, stridea=-1,
                                          strideb=sb, stridec=sc, batch=3)] == "invalid _stridea"
    assert oracle.STATUS_TEXT[oracle.gemm("n", "n", 16, 16, 17, 1.0, A, lda, B, ldb, 0.0, C.copy(), ldc, stridea=sa,
                                          strideb=-2, stridec=sc, batch=3)] == "invalid _strideb"
    # alpha == 0 is tested before validation: invalid trans is NOT rejected (appendix A.1)
    c = C.copy()
    assert oracle.gemm("x", "y", 16, 16, 17, 0.0, A, lda, B, ldb, 2.0, c, ldc) == 0
    win = np.zeros_like(C, dtype=bool)
    win.reshape(-1)[: ldc * 16].reshape(16, ldc)[:, :16] = True
    assert np.array_equal(c[win], 2.0 * C[win]) and np.array_equal(c[~win], C[~win])
    # beta == 1 with alpha == 0 is a no-op; 'c' behaves as 't'
    c = C.copy()
    oracle.gemm("n", "n", 16, 16, 17, 0.0, A, lda, B, ldb, 1.0, c, ldc)
    assert np.array_equal(c, C)
    c1, c2 = C.copy(), C.copy()
    oracle.gemm("c", "C", 16, 16, 17, 1.5, A, 17 * 2, B, 16 * 2, 0.5, c1, ldc)
    oracle.gemm("t", "t", 16, 16, 17, 1.5, A, 17 * 2, B, 16 * 2, 0.5, c2, ldc)
    assert np.array_equal(c1, c2)
    # beta == 0 never reads C: NaNs in C do not propagate (gemm_ref.hpp:245-251)
    cn = np.full_like(C, np.nan)
    oracle.gemm("n", "n", 16, 16, 17, 1.0, A, lda, B, ldb, 0.0, cn, ldc)
    assert not np.isnan(cn[win]).any()


def test_comparison_predicate_matches_reference_margins():
    """float_comparison.hpp:101-188."""
    one = np.array([1.0])
    assert oracle.compare(one, one * (1 + 0.009), "float") == 0       # |d|/(|a|+|b|) = 0.0045 < 0.005
    assert oracle.compare(one * 10, one * 10 * (1 + 0.011), "float") == 1
    assert oracle.compare(np.array([0.0]), np.array([0.0009]), "float") == 0   # abs margin 1e-3
    assert oracle.compare(np.array([0.0]), np.array([0.0011]), "float") == 1
    assert oracle.compare(one, one * (1 + 1e-9), "double") == 1 and oracle.compare(one, one * (1 + 1e-11), "double") == 0
    assert oracle.compare(np.array([100.0]), np.array([100.9]), "half") == 0   # abs margin 1.0
    assert oracle.compare(np.array([np.nan]), np.array([np.nan]), "float") == 0
    assert oracle.compare(np.array([np.inf]), np.array([np.inf]), "float") == 0
    assert oracle.compare(one, one * 1.012, "float", 3) == 0                   # joint_matrix x3 multiplier


def test_default_cpu_kernel_restatement():
    """The timed CPU baseline kernel (DEFAULT backend, Tile<4,4,4,4>) computes the same GEMM."""
    rng = np.random.default_rng(11)
    for npdt, kind in ((np.float32, "float"), (np.float64, "double")):
        for ta, tb in TRANS:
            m, n, k = 70, 37, 129
            A, B, C, lda, ldb, ldc, *_ = _case(rng, npdt, ta, tb, m, n, k, 2, 2, 2)
            want, got = C.copy(), C.copy()
            oracle.cblas_gemm(ta, tb, m, n, k, 1.5, A, lda, B, ldb, 0.5, want, ldc)
            oracle.gemm_default_cpu(ta == "t", tb == "t", m, n, k, 1.5, A, lda, B, ldb, 0.5, got, ldc)
            assert oracle.compare(got, want, kind) == 0


def test_golden_fixtures():
    """The committed vectors (tests/golden/make_golden.py) pin the restatement; storage-type
    cases (f16 / bf16) are compared after rounding to the storage type."""
    g = np.load(GOLDEN, allow_pickle=False)
    n = len([k for k in g.files if k.endswith("_meta")])
    assert n >= 10
    for i in range(n):
        dt, ta, tb, m, nn, k, al, be, la, lb, lc, batch = g[f"case{i}_meta"]
        m, nn, k, la, lb, lc, batch = (int(x) for x in (m, nn, k, la, lb, lc, batch))
        al, be = float(al), float(be)
        A, B, C, want = g[f"case{i}_A"], g[f"case{i}_B"], g[f"case{i}_C"], g[f"case{i}_out"]
        lda = (k if ta == "t" else m) * la
        ldb = (nn if tb == "t" else k) * lb
        ldc = m * lc
        got = C.copy()
        assert oracle.gemm(ta, tb, m, nn, k, al, A, lda, B, ldb, be, got, ldc, stridea=m * k * la, strideb=k * nn * lb,
                           stridec=m * nn * lc, batch=batch, mode=oracle.MODE_LOCAL) == 0
        if dt in ("f16", "bf16"):
            got = oracle.round_to(got, dt)
        kind = {"f32": "float", "f64": "double", "f16": "half", "bf16": "half"}[dt]
        assert oracle.compare(got, want, kind) == 0, f"golden case {i}"
        if dt == "f64":
            assert np.max(np.abs(got - want) / (np.abs(want) + 1)) < 1e-13


def test_bf16_rounding_helpers():
    x = np.array([1.0, 1.00390625, 3.1415927, -2.7182817, 65504.0, 1e-8], dtype=np.float32)
    r = oracle.round_to(x, "bf16")
    assert np.array_equal(oracle.from_bf16_bits(oracle.to_bf16_bits(x)), r)
    import torch
    assert np.array_equal(torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy(), r)
