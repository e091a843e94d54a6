"""CPU tests of the oracle for _symm / _trsm / complex _gemm (oracle/blas3_ext.py) and of the numpy model of
pbx_trsm's host logic.  The restatements are pinned to the oracle the reference's own tests use -- CBLAS symm /
trsm / gemm (OpenBLAS through scipy) -- on the reference's parameter grids
(test/unittest/blas3/blas3_symm_test.cpp:155-209, blas3_trsm_test.cpp:130-163, blas3_gemm_test.cpp:143-259)
with the reference's tolerance predicate."""
from __future__ import annotations

import itertools

import numpy as np
import pytest

from oracle import blas3_ext as ox
from oracle import oracle
from trsm_model import trsm_model


def _kind(dt):
    return "double" if dt in (np.float64, np.complex128) else "float"


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_symm_restatement_matches_cblas(dt):
    rng = np.random.default_rng(12345)
    grid = itertools.product([11, 16, 32, 63], [11, 16, 32, 63], "lr", "lu", [(1.5, 0.5), (1.0, 1.0), (1.5, 0.0)],
                             [1, 2])
    for m, n, side, uplo, (al, be), ldm in grid:
        k = m if side == "l" else n
        lda, ldb, ldc = k * ldm, m * ldm, m * ldm
        A, B, C = (oracle.random_uniform(rng, s, dt) for s in (k * lda, n * ldb, n * ldc))
        c1, c2 = C.copy(), C.copy()
        assert ox.symm(side, uplo, m, n, al, A, lda, B, ldb, be, c1, ldc) == 0
        ox.cblas_symm(side, uplo, m, n, al, A, lda, B, ldb, be, c2, ldc)
        assert oracle.compare(c1, c2, _kind(dt)) == 0
        pad = np.ones(n * ldc, bool)
        ox.view(pad, m, n, ldc)[...] = False
        assert np.array_equal(c1[pad], C[pad])


def test_symm_validation_order_and_alpha_zero():
    z = np.zeros(16)
    assert ox.symm("x", "q", 4, 4, 1.0, z, 4, z, 4, 0.0, z.copy(), 4) == 10   # uplo is checked first
    assert ox.symm("x", "u", 4, 4, 1.0, z, 4, z, 4, 0.0, z.copy(), 4) == 11
    c = np.full(16, np.nan)
    assert ox.symm("l", "u", 4, 4, 0.0, z, 4, z, 4, 0.0, c, 4) == 0 and np.all(c == 0)   # beta == 0 stores zeros


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_trsm_restatement_and_model_match_cblas(dt):
    rng = np.random.default_rng(12345)
    nb = 128 if dt == np.float32 else 64
    for m, n, tr, side, dg, uplo, unused in itertools.product([7, 130], [7, 257], "nt", "lr", "un", "lu", [0.0, np.nan]):
        k = m if side == "l" else n
        lda, ldb = 2 * k, 2 * m
        A = ox.fill_trsm_matrix(rng, k, lda, uplo, dg, float(rng.uniform(1, 10)), unused, dt)
        B = oracle.random_uniform(rng, n * ldb, dt)
        b_ref, b_alg, b_mod = B.copy(), B.copy(), B.copy()
        ox.cblas_trsm(side, uplo, tr, dg, m, n, 2.0, A, lda, b_ref, ldb)
        assert ox.trsm_ref_algorithm(side, uplo, tr, dg, m, n, 2.0, A, lda, b_alg, ldb) == 0
        trsm_model(side, uplo, tr, dg, m, n, 2.0, A, lda, b_mod, ldb, nb)
        assert oracle.compare(b_alg, b_ref, _kind(dt)) == 0, (m, n, tr, side, dg, uplo)
        assert oracle.compare(b_mod, b_ref, _kind(dt)) == 0, (m, n, tr, side, dg, uplo)
        truth = ox.trsm_truth(side, uplo, tr, dg, m, n, 2.0, A, lda, B, ldb)
        assert oracle.compare(ox.view(b_ref, m, n, ldb).copy(), truth.astype(dt), _kind(dt)) == 0
        assert np.isfinite(b_mod).all()


def test_trsm_validation_order():
    args = dict(m=4, n=4, lda=4, ldb=4)
    assert ox.trsm_status("l", "l", "n", "n", 0, 4, 4, 4) == 12
    assert ox.trsm_status("x", "x", "x", "x", **args) == 13
    assert ox.trsm_status("l", "x", "x", "x", **args) == 14
    assert ox.trsm_status("l", "u", "c", "x", **args) == 15   # 'c' is rejected (trsm_interface.hpp:125)
    assert ox.trsm_status("l", "u", "t", "x", **args) == 16
    assert ox.trsm_status("R", "U", "T", "U", **args) == 0


@pytest.mark.parametrize("dt", [np.complex64, np.complex128])
def test_cgemm_restatement_matches_cblas(dt):
    rng = np.random.default_rng(12345)
    rdt = np.float32 if dt == np.complex64 else np.float64

    def rand(cnt):
        return (oracle.random_uniform(rng, cnt, rdt) + 1j * oracle.random_uniform(rng, cnt, rdt)).astype(dt)

    for m, n, k, ta, tb, (al, be) in itertools.product([11, 33], [11, 33], [16, 17], "ntc", "ntc",
                                                       [(1.5 + 1j, 1.5 + 3j), (1.5 + 3j, 0j)]):
        lda, ldb, ldc = (k if ta != "n" else m) * 2, (n if tb != "n" else k) * 2, m * 3
        A, B, C = rand(lda * (m if ta != "n" else k)), rand(ldb * (k if tb != "n" else n)), rand(ldc * n)
        c1, c2 = C.copy(), C.copy()
        assert ox.cgemm(ta, tb, m, n, k, al, A, lda, B, ldb, be, c1, ldc, conj=True) == 0
        ox.cblas_cgemm(ta, tb, m, n, k, al, A, lda, B, ldb, be, c2, ldc)
        assert oracle.compare(c1.real.copy(), c2.real.copy(), _kind(dt)) == 0
        assert oracle.compare(c1.imag.copy(), c2.imag.copy(), _kind(dt)) == 0
        if "c" in (ta, tb):   # the reference's quirk: 'c' behaves as 't'
            c3, c4 = C.copy(), C.copy()
            ox.cgemm(ta, tb, m, n, k, al, A, lda, B, ldb, be, c3, ldc, conj=False)
            ox.cgemm(ta.replace("c", "t"), tb.replace("c", "t"), m, n, k, al, A, lda, B, ldb, be, c4, ldc)
            assert np.array_equal(c3, c4)


def test_cgemm_front_end_rules():
    z = np.zeros(16, np.complex64)
    c = np.full(16, np.nan + 0j, np.complex64)
    # alpha == 0 comes before validation, beta == 0 stores zeros without reading C
    assert ox.cgemm("x", "y", 4, 4, 4, 0j, z, 4, z, 4, 0j, c, 4) == 0 and np.all(c == 0)
    assert ox.cgemm("x", "n", 4, 4, 4, 1 + 0j, z, 4, z, 4, 0j, c, 4) == 1
    assert ox.cgemm("n", "y", 4, 4, 4, 1 + 0j, z, 4, z, 4, 0j, c, 4) == 2
    assert ox.cgemm("n", "n", 4, 4, 4, 1 + 0j, z, 4, z, 4, 0j, c, 4, stridec=3, batch=2) == 3


def _golden():
    from pathlib import Path
    return np.load(Path(__file__).parent / "golden" / "ext_golden.npz", allow_pickle=False)


def test_restatements_reproduce_committed_golden_vectors():
    """tests/golden/ext_golden.npz (make_golden_ext.py: CBLAS on seeded inputs) against the numpy restatements."""
    g = _golden()
    i = 0
    while f"symm{i}_meta" in g.files:
        dt, side, uplo, m, n, al, be, la, lb, lc = g[f"symm{i}_meta"]
        m, n, la, lb, lc = (int(x) for x in (m, n, la, lb, lc))
        k = m if side == "l" else n
        c = g[f"symm{i}_C"].copy()
        assert ox.symm(str(side), str(uplo), m, n, float(al), g[f"symm{i}_A"], k * la, g[f"symm{i}_B"], m * lb, float(be),
                       c, m * lc) == 0
        assert oracle.compare(c, g[f"symm{i}_out"], "double" if dt == "f64" else "float") == 0, f"symm case {i}"
        i += 1
    assert i >= 5
    i = 0
    while f"trsm{i}_meta" in g.files:
        dt, side, uplo, tr, dg, m, n, al = g[f"trsm{i}_meta"]
        m, n = int(m), int(n)
        k = m if side == "l" else n
        b = g[f"trsm{i}_B"].copy()
        assert ox.trsm_ref_algorithm(str(side), str(uplo), str(tr), str(dg), m, n, float(al), g[f"trsm{i}_A"], 2 * k, b,
                                     2 * m) == 0
        assert oracle.compare(b, g[f"trsm{i}_out"], "double" if dt == "f64" else "float") == 0, f"trsm case {i}"
        b2 = g[f"trsm{i}_B"].copy()
        trsm_model(str(side), str(uplo), str(tr), str(dg), m, n, float(al), g[f"trsm{i}_A"], 2 * k, b2, 2 * m,
                   64 if dt == "f64" else 128)
        assert oracle.compare(b2, g[f"trsm{i}_out"], "double" if dt == "f64" else "float") == 0, f"trsm model case {i}"
        i += 1
    assert i >= 5
    i = 0
    while f"cgemm{i}_meta" in g.files:
        dt, ta, tb, m, n, k, la, lb, lc = g[f"cgemm{i}_meta"]
        m, n, k, la, lb, lc = (int(x) for x in (m, n, k, la, lb, lc))
        al, be = (complex(x) for x in g[f"cgemm{i}_scal"])
        lda, ldb, ldc = (k if ta != "n" else m) * la, (n if tb != "n" else k) * lb, m * lc
        c = g[f"cgemm{i}_C"].copy()
        assert ox.cgemm(str(ta), str(tb), m, n, k, al, g[f"cgemm{i}_A"], lda, g[f"cgemm{i}_B"], ldb, be, c, ldc) == 0
        kind = "double" if dt == "c128" else "float"
        want = g[f"cgemm{i}_out"]
        assert oracle.compare(c.real.copy(), want.real.copy(), kind) == 0, f"cgemm case {i}"
        assert oracle.compare(c.imag.copy(), want.imag.copy(), kind) == 0, f"cgemm case {i}"
        i += 1
    assert i >= 4
