"""GPU parity tests of the routines built on the GEMM path -- _symm, _trsm, complex _gemm (SURVEY.md section 8
rows f1-f3) -- through the C-ABI, against the CPU oracle (oracle/blas3_ext.py) and CBLAS.

Grids restate the reference's suites: test/unittest/blas3/blas3_symm_test.cpp:155-209,
blas3_trsm_test.cpp:130-163 (m, n in {7, 513, 1027}, the unused triangle 0 or NaN) and the complex GEMM suites
blas3_gemm_test.cpp:143-259.  Bars: the reference's almost_equal margins against CBLAS and, against the
long-double truth, 1e-5 (fp32 / complex64) and 1e-12 (fp64 / complex128) relative to |alpha||A||B| + |beta||C|
(for TRSM: of the residual |op(A)X - alpha B| relative to |op(A)||X| + |alpha||B|).
"""
from __future__ import annotations

import itertools

import pytest

from ext_case import CgemmCase, SymmCase, TrsmCase, run_cgemm, run_symm, run_trsm

pytestmark = pytest.mark.gpu


def _run_all(handle, runner, cases):
    failures = []
    worst = 0.0
    for cs in cases:
        r = runner(handle, cs)
        worst = max(worst, r.max_rel)
        if not r.ok:
            failures.append(f"{cs.ident()} kernel={r.kernel} max_rel={r.max_rel:.3e} {r.detail}")
    assert not failures, f"{len(failures)}/{len(cases)} cases failed:\n" + "\n".join(failures[:20])
    return worst


# ---- symm (blas3_symm_test.cpp:155-209) -----------------------------------------------------------
@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_symm_small_and_alpha_zero(handle, dt):
    cases = [SymmCase(dtype=dt, side=s, uplo=u, m=m, n=n, alpha=1.5, beta=0.5)
             for m, n, s, u in itertools.product([11, 16, 32], [11, 16, 32], "lr", "lu")]
    cases += [SymmCase(dtype=dt, side=s, uplo=u, m=16, n=16, alpha=0.0, beta=b, lda_mul=la, ldb_mul=lb, ldc_mul=lc)
              for s, u, b, la, lb, lc in itertools.product("lr", "lu", [0.0, 1.0], [1, 2], [1, 2], [1, 2])]
    _run_all(handle, run_symm, cases)


@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_symm_ld_multipliers_and_large(handle, dt):
    cases = [SymmCase(dtype=dt, side=s, uplo=u, m=m, n=n, alpha=1.0, beta=1.0, lda_mul=la, ldb_mul=lb, ldc_mul=lc)
             for m, n, s, u, la, lb, lc in itertools.product([16, 63], [16, 63], "lr", "lu", [1, 2], [1, 2], [1, 2])]
    cases += [SymmCase(dtype=dt, side=s, uplo=u, m=m, n=n, alpha=1.0, beta=1.0)
              for m, n, s, u in itertools.product([253, 511], [257, 511], "lr", "lu")]
    _run_all(handle, run_symm, cases)


@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_symm_never_reads_the_other_triangle_and_beta_zero(handle, dt):
    cases = [SymmCase(dtype=dt, side=s, uplo=u, m=m, n=n, alpha=1.5, beta=be, nan_unused=True, ldc_mul=2)
             for m, n, s, u, be in itertools.product([33, 300], [17, 260], "lr", "LU", [0.0, 0.5])]
    _run_all(handle, run_symm, cases)


def test_symm_invalid_arguments(handle):
    # uplo is validated before side (symm_interface.hpp:51-72)
    cases = [SymmCase(side="x", uplo="q"), SymmCase(side="x", uplo="u"), SymmCase(side="l", uplo="q")]
    _run_all(handle, run_symm, cases)


def test_symm_large_runs_on_tensor_cores(handle):
    r = run_symm(handle, SymmCase(dtype="f32", side="l", uplo="u", m=1024, n=768, alpha=1.0, beta=0.0))
    assert r.ok, r.detail
    assert r.kernel == "tcgen05"
    r = run_symm(handle, SymmCase(dtype="f64", side="r", uplo="l", m=640, n=1024, alpha=-1.0, beta=2.0))
    assert r.ok and r.kernel == "dmma", r.detail


# ---- trsm (blas3_trsm_test.cpp:130-163) --------------------------------------------------------------
def _trsm_grid(dt, sizes, unused_values):
    return [TrsmCase(dtype=dt, side=s, uplo=u, trans=t, diag=d, m=m, n=n, alpha=2.0, unused=un)
            for m, n, t, s, d, u, un in itertools.product(sizes, sizes, "nt", "lr", "un", "lu", unused_values)]


@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_trsm_small(handle, dt):
    _run_all(handle, run_trsm, _trsm_grid(dt, [7], [0.0, float("nan")]) +
             [TrsmCase(dtype=dt, side=s, uplo=u, trans=t, diag=d, m=m, n=n, alpha=al, lda_mul=1, ldb_mul=1)
              for (m, n), t, s, d, u, al in itertools.product([(64, 65), (129, 16), (200, 130)], "nT", "lR", "Un", "Lu",
                                                              [1.0, -0.5])])


@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_trsm_reference_grid(handle, dt):
    cases = [c for c in _trsm_grid(dt, [7, 513, 1027], [float("nan")]) if not (c.m == 7 and c.n == 7)]
    worst = _run_all(handle, run_trsm, cases)
    print(f"trsm {dt}: worst residual / bound = {worst:.3e}")


def test_trsm_alpha_zero_and_invalid_arguments(handle):
    cases = [TrsmCase(side="l", alpha=0.0, m=40, n=9), TrsmCase(side="r", alpha=0.0, m=9, n=140, dtype="f64"),
             TrsmCase(m=0), TrsmCase(n=0), TrsmCase(side="x"), TrsmCase(uplo="x"), TrsmCase(trans="c"),
             TrsmCase(diag="x"), TrsmCase(side="x", uplo="x", trans="x", diag="x")]
    _run_all(handle, run_trsm, cases)


# ---- complex gemm (blas3_gemm_test.cpp:143-259) ------------------------------------------------------
NT = ["n", "t"]


@pytest.mark.parametrize("dt", ["c64", "c128"])
def test_cgemm_small_suites(handle, dt):
    cases = [CgemmCase(dtype=dt, transa=ta, transb=tb, m=m, n=n, k=k, alpha=1.5 + 1j, beta=1.5 + 3j)
             for m, n, k, ta, tb in itertools.product([11, 33], [11, 33], [16, 17], NT, NT)]
    cases += [CgemmCase(dtype=dt, transa=ta, transb=tb, m=m, n=n, k=17, alpha=1.5 + 1j, beta=0j)
              for m, n, ta, tb in itertools.product([11, 32], [11, 32], NT, NT)]
    cases += [CgemmCase(dtype=dt, transa=ta, transb=tb, m=m, n=n, k=17, alpha=1.5 + 3j, beta=0j, lda_mul=2, ldb_mul=2,
                        ldc_mul=3) for m, n, ta, tb in itertools.product([11, 33], [11, 33], NT, NT)]
    _run_all(handle, run_cgemm, cases)


@pytest.mark.parametrize("dt", ["c64", "c128"])
def test_cgemm_alpha_zero_offsets_and_large(handle, dt):
    cases = [CgemmCase(dtype=dt, m=16, n=16, k=17, alpha=0j, beta=be, offset=off, lda_mul=la, ldb_mul=lb, ldc_mul=lc)
             for off, be, la, lb, lc in itertools.product([0, 10], [0j, 1 + 0j], [1, 2], [1, 2], [1, 2])]
    cases += [CgemmCase(dtype=dt, m=m, n=n, k=k, alpha=1 + 1j, beta=1 + 1j, offset=off, lda_mul=la, ldb_mul=lb,
                        ldc_mul=lc)
              for off, m, n, k, la, lb, lc in itertools.product([1, 10], [16, 63], [16, 63], [17, 63], [1, 2], [1, 2],
                                                                [1, 2])]
    cases += [CgemmCase(dtype=dt, transa=ta, transb=tb, m=m, n=n, k=k, alpha=1 + 1.5j, beta=1.5 + 1j)
              for m, n, k, ta, tb in itertools.product([63, 253], [63, 253], [63, 253], NT, NT)]
    _run_all(handle, run_cgemm, cases)


@pytest.mark.parametrize("dt", ["c64", "c128"])
def test_cgemm_conj_quirk_batches_and_errors(handle, dt):
    # 'c' == 't' by default (the reference's behaviour); BLAS conjugation behind the handle option
    cases = [CgemmCase(dtype=dt, transa=ta, transb=tb, m=33, n=20, k=40, conj=cj)
             for ta, tb, cj in itertools.product("ntc", "ntC", [False, True])]
    # strided batches incl. broadcast-free padded strides; NaN in C with beta == 0 is covered by the oracle's zeros
    cases += [CgemmCase(dtype=dt, transa=ta, transb=tb, m=40, n=24, k=56, batch=5, stride_mul=sm, beta=be)
              for ta, tb, sm, be in itertools.product(NT, NT, [1, 2], [0j, 0.5 - 1j])]
    cases += [CgemmCase(dtype=dt, m=300, n=200, k=520, alpha=-1 + 0.25j, beta=0j),
              CgemmCase(dtype=dt, transa="x"), CgemmCase(dtype=dt, transb="x"),
              CgemmCase(dtype=dt, m=0), CgemmCase(dtype=dt, k=0)]
    _run_all(handle, run_cgemm, cases)


@pytest.mark.parametrize("dt_name", ["f16", "bf16"])
def test_symm_16bit_storage(handle, dt_name):
    """16-bit storage: the mirror pass is type-agnostic (2-byte elements) and the product runs on tcgen05 kind::f16;
    checked on the device against the fp64 product of the mirrored matrix (bars of the 16-bit GEMM tests)."""
    import torch
    from portblas_b200 import blas
    dt, tol = (torch.float16, 2e-3) if dt_name == "f16" else (torch.bfloat16, 1.6e-2)
    gen = torch.Generator(device="cuda").manual_seed(11)
    for side, uplo, m, n in [("l", "u", 264, 200), ("r", "l", 136, 392), ("l", "l", 63, 17)]:
        kk = m if side == "l" else n
        lda, ldb, ldc = kk + 8, m, m + 8
        a = (torch.rand(lda * kk, device="cuda", generator=gen) * 7 - 2).to(dt)
        b = (torch.rand(ldb * n, device="cuda", generator=gen) * 7 - 2).to(dt)
        c0 = (torch.rand(ldc * n, device="cuda", generator=gen) * 7 - 2).to(dt)
        c = c0.clone()
        blas._symm(handle, side, uplo, m, n, 1.5, a, lda, b, ldb, 0.5, c, ldc)
        handle.wait()
        f64 = torch.float64
        A = a.view(kk, lda).T[:kk].to(f64)
        tri = torch.tril(A) if uplo == "l" else torch.triu(A)
        S = tri + tri.T - torch.diag(torch.diag(A))
        B, C0, C = b.view(n, ldb).T[:m].to(f64), c0.view(n, ldc).T.to(f64), c.view(n, ldc).T.to(f64)
        want = 1.5 * (S @ B if side == "l" else B @ S) + 0.5 * C0[:m]
        bound = 1.5 * (S.abs() @ B.abs() if side == "l" else B.abs() @ S.abs()) + 0.5 * C0[:m].abs()
        assert float(((C[:m] - want).abs() / bound).max()) <= tol
        assert torch.equal(C[m:], C0[m:])      # ld padding untouched


def test_ext_golden_fixtures_gpu(handle):
    """The committed golden vectors (tests/golden/ext_golden.npz, CBLAS outputs on seeded inputs) through the CUDA
    path: _symm, _trsm and complex _gemm must reproduce them within the reference's almost_equal margins."""
    from pathlib import Path

    import numpy as np
    import torch

    from oracle import oracle
    from portblas_b200 import blas

    g = np.load(Path(__file__).parent / "golden" / "ext_golden.npz", allow_pickle=False)
    dev = torch.device("cuda", handle.device)
    i = 0
    while f"symm{i}_meta" in g.files:
        dt, side, uplo, m, n, al, be, la, lb, lc = g[f"symm{i}_meta"]
        m, n, la, lb, lc = (int(x) for x in (m, n, la, lb, lc))
        k = m if side == "l" else n
        a, b, c = (torch.from_numpy(g[f"symm{i}_{x}"]).to(dev) for x in "ABC")
        blas._symm(handle, str(side), str(uplo), m, n, float(al), a, k * la, b, m * lb, float(be), c, m * lc)
        handle.wait()
        assert oracle.compare(c.cpu().numpy(), g[f"symm{i}_out"], "double" if dt == "f64" else "float") == 0, f"symm {i}"
        i += 1
    i = 0
    while f"trsm{i}_meta" in g.files:
        dt, side, uplo, tr, dg, m, n, al = g[f"trsm{i}_meta"]
        m, n = int(m), int(n)
        k = m if side == "l" else n
        a, b = torch.from_numpy(g[f"trsm{i}_A"]).to(dev), torch.from_numpy(g[f"trsm{i}_B"]).to(dev)
        blas._trsm(handle, str(side), str(uplo), str(tr), str(dg), m, n, float(al), a, 2 * k, b, 2 * m)
        handle.wait()
        assert oracle.compare(b.cpu().numpy(), g[f"trsm{i}_out"], "double" if dt == "f64" else "float") == 0, f"trsm {i}"
        i += 1
    i = 0
    while f"cgemm{i}_meta" in g.files:
        dt, ta, tb, m, n, k, la, lb, lc = g[f"cgemm{i}_meta"]
        m, n, k, la, lb, lc = (int(x) for x in (m, n, k, la, lb, lc))
        al, be = (complex(x) for x in g[f"cgemm{i}_scal"])
        lda, ldb, ldc = (k if ta != "n" else m) * la, (n if tb != "n" else k) * lb, m * lc
        a, b, c = (torch.from_numpy(g[f"cgemm{i}_{x}"]).to(dev) for x in "ABC")
        blas._gemm(handle, str(ta), str(tb), m, n, k, al, a, lda, b, ldb, be, c, ldc)
        handle.wait()
        got, want = c.cpu().numpy(), g[f"cgemm{i}_out"]
        kind = "double" if dt == "c128" else "float"
        assert oracle.compare(got.real.copy(), want.real.copy(), kind) == 0, f"cgemm {i}"
        assert oracle.compare(got.imag.copy(), want.imag.copy(), kind) == 0, f"cgemm {i}"
        i += 1
