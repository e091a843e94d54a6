"""Case runners for the routines built on the GEMM path (_symm, _trsm, complex _gemm), shared by the GPU
parity tests and __graft_entry__.smoke().

They follow the reference's own verifiers: ``verify_symm`` (test/unittest/blas3/blas3_symm_test.cpp:34-106),
``run_test`` of the TRSM suite (blas3_trsm_test.cpp:31-106, matrices from ``fill_trsm_matrix``
test/blas_test.hpp:186-214) and ``verify_gemm`` for complex (blas3_gemm_common.hpp:232-400): leading dimensions
are ``rows * ld_mul``, inputs U(-2,5), expected values from the CPU oracle (oracle/blas3_ext.py) and from CBLAS,
whole output buffer compared (padding must stay bit-identical), the reference's ``almost_equal`` as predicate
plus an error bound against a long-double truth.
"""
from __future__ import annotations

import dataclasses

import numpy as np
import torch

from oracle import blas3_ext as ox
from oracle import oracle
from portblas_b200 import blas

REAL = {"f32": (torch.float32, np.float32, "float", 1e-5), "f64": (torch.float64, np.float64, "double", 1e-12)}
CPLX = {"c64": (torch.complex64, np.complex64, np.float32, "float", 1e-5),
        "c128": (torch.complex128, np.complex128, np.float64, "double", 1e-12)}
# TRSM residual bar: |op(A) X - alpha B| <= tol * (|op(A)||X| + |alpha||B|) elementwise.  The scheme multiplies
# by explicitly inverted 128 (fp32) / 64 (fp64) wide diagonal blocks, as the reference does with 16-wide ones, so
# the bound carries the condition number of those blocks (the test matrices are row diagonally dominant; the numpy
# model of the scheme, tests/trsm_model.py, sits at 4e-7 / 6e-16 on the reference's grid).
TRSM_TOL = {"f32": 1e-5, "f64": 1e-12}


@dataclasses.dataclass
class ExtResult:
    ok: bool
    detail: str
    max_rel: float = 0.0
    kernel: str = ""


def _status_result(want: int, got_text: str) -> ExtResult:
    ok = ox.STATUS_TEXT.get(want, "?") == got_text
    return ExtResult(ok, f"status oracle='{ox.STATUS_TEXT.get(want)}' got='{got_text}'")


# ------------------------------------------------------------------------------------------------ symm
@dataclasses.dataclass
class SymmCase:
    dtype: str = "f32"
    side: str = "l"
    uplo: str = "l"
    m: int = 16
    n: int = 16
    alpha: float = 1.5
    beta: float = 0.5
    lda_mul: int = 1
    ldb_mul: int = 1
    ldc_mul: int = 1
    nan_unused: bool = False   # poison the triangle that must never be read
    seed: int = 12345

    def ident(self) -> str:
        return (f"symm-{self.dtype}-{self.side}{self.uplo}-{self.m}x{self.n}-a{self.alpha}b{self.beta}"
                f"-ld{self.lda_mul}{self.ldb_mul}{self.ldc_mul}" + ("-nan" if self.nan_unused else ""))


def run_symm(handle: blas.SB_Handle, cs: SymmCase) -> ExtResult:
    tt, npdt, kind, tol = REAL[cs.dtype]
    m, n = cs.m, cs.n
    k = m if cs.side.lower() == "l" else n
    lda, ldb, ldc = max(k, 1) * cs.lda_mul, max(m, 1) * cs.ldb_mul, max(m, 1) * cs.ldc_mul
    rng = np.random.default_rng(cs.seed)
    a_h = oracle.random_uniform(rng, k * lda, npdt)
    b_h = oracle.random_uniform(rng, n * ldb, npdt)
    c_h = oracle.random_uniform(rng, n * ldc, npdt)
    st_want = ox.symm_status(cs.side, cs.uplo)
    if cs.nan_unused and st_want == 0:
        av = ox.view(a_h, k, k, lda)
        mask = np.triu(np.ones((k, k), bool), 1) if cs.uplo.lower() == "l" else np.tril(np.ones((k, k), bool), -1)
        av[mask] = np.nan
    dev = torch.device("cuda", handle.device)
    a_d, b_d, c_d = (torch.from_numpy(x).to(dev) for x in (a_h, b_h, c_h))
    text = ""
    try:
        blas._symm(handle, cs.side, cs.uplo, m, n, cs.alpha, a_d, lda, b_d, ldb, cs.beta, c_d, ldc)
        handle.wait()
    except ValueError as e:
        text = str(e)
    if st_want or text:
        return _status_result(st_want, text)
    got = c_d.cpu().numpy()
    a_clean = a_h.copy()
    if cs.nan_unused:
        full = ox.symm_full(cs.uplo, k, np.nan_to_num(a_h, nan=0.0), lda)
        a_clean = np.zeros_like(a_h)
        ox.view(a_clean, k, k, lda)[...] = full
    truth = c_h.copy()
    ox.symm(cs.side, cs.uplo, m, n, cs.alpha, a_clean, lda, b_h, ldb, cs.beta, truth, ldc)
    bound = np.abs(c_h).copy()
    ox.symm(cs.side, cs.uplo, m, n, abs(cs.alpha), np.abs(a_clean), lda, np.abs(b_h), ldb, abs(cs.beta), bound, ldc)
    exp = c_h.copy()
    ox.cblas_symm(cs.side, cs.uplo, m, n, cs.alpha, a_clean, lda, b_h, ldb, cs.beta, exp, ldc)  # the tests' oracle
    mism = oracle.compare(got, exp, kind)
    err = np.abs(got.astype(np.float64) - truth.astype(np.float64))
    untouched = truth == c_h
    window = np.zeros(n * ldc, bool)
    ox.view(window, m, n, ldc)[...] = True
    viol = np.where(window, err > tol * bound.astype(np.float64), got != c_h)
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(window & (bound > 0), err / bound.astype(np.float64), 0.0)
    nv = int(viol.sum())
    del untouched
    return ExtResult(mism == 0 and nv == 0, f"ref_mismatch={mism} bound_viol={nv}", float(rel.max(initial=0.0)),
                     handle.last_kernel)


# ------------------------------------------------------------------------------------------------ trsm
@dataclasses.dataclass
class TrsmCase:
    dtype: str = "f32"
    side: str = "l"
    uplo: str = "l"
    trans: str = "n"
    diag: str = "n"
    m: int = 7
    n: int = 7
    alpha: float = 2.0
    lda_mul: int = 2
    ldb_mul: int = 2
    unused: float = float("nan")   # value placed in the triangle TRSM must not read (0 or NaN in the reference)
    seed: int = 12345

    def ident(self) -> str:
        return (f"trsm-{self.dtype}-{self.side}{self.uplo}{self.trans}{self.diag}-{self.m}x{self.n}-a{self.alpha}"
                f"-ld{self.lda_mul}{self.ldb_mul}-u{self.unused}")


def run_trsm(handle: blas.SB_Handle, cs: TrsmCase) -> ExtResult:
    tt, npdt, kind, _ = REAL[cs.dtype]
    tol = TRSM_TOL[cs.dtype]
    m, n = cs.m, cs.n
    k = m if cs.side.lower() == "l" else n
    lda, ldb = k * cs.lda_mul, m * cs.ldb_mul
    st_want = ox.trsm_status(cs.side, cs.uplo, cs.trans, cs.diag, m, n, lda, ldb)
    rng = np.random.default_rng(cs.seed)
    diag_value = float(rng.uniform(1.0, 10.0))
    if st_want == 0:
        a_h = ox.fill_trsm_matrix(rng, k, lda, cs.uplo, cs.diag, diag_value, cs.unused, npdt)
    else:
        a_h = np.ones(max(k * lda, 1), npdt)
    b_h = oracle.random_uniform(rng, max(n * ldb, 1), npdt)
    dev = torch.device("cuda", handle.device)
    a_d, b_d = torch.from_numpy(a_h).to(dev), torch.from_numpy(b_h).to(dev)
    text = ""
    try:
        blas._trsm(handle, cs.side, cs.uplo, cs.trans, cs.diag, m, n, cs.alpha, a_d, lda, b_d, ldb)
        handle.wait()
    except ValueError as e:
        text = str(e)
    if st_want or text:
        return _status_result(st_want, text)
    got = b_d.cpu().numpy()
    exp = b_h.copy()
    ox.cblas_trsm(cs.side, cs.uplo, cs.trans, cs.diag, m, n, cs.alpha, a_h, lda, exp, ldb)   # the tests' oracle
    mism = oracle.compare(got, exp, kind)
    # padding rows of B must be untouched
    window = np.zeros(n * ldb, bool)
    ox.view(window, m, n, ldb)[...] = True
    pad_bad = int((got[~window] != b_h[~window]).sum())
    # residual against the exact operator: long double for small systems, float64 (BLAS) for large ones, whose
    # own rounding (~sqrt(K) * 1.1e-16 of the denominator) stays far below both bars
    wide = np.longdouble if k <= 160 else np.float64
    a = ox.view(a_h, k, k, lda).astype(wide)
    t = np.triu(a) if cs.uplo.lower() == "u" else np.tril(a)
    if cs.diag.lower() == "u":
        np.fill_diagonal(t, 1)
    if cs.trans.lower() == "t":
        t = t.T
    x = ox.view(got, m, n, ldb).astype(wide)
    rhs = wide(cs.alpha) * ox.view(b_h, m, n, ldb).astype(wide)
    if cs.side.lower() == "l":
        res, den = t @ x - rhs, np.abs(t) @ np.abs(x) + np.abs(rhs)
    else:
        res, den = x @ t - rhs, np.abs(x) @ np.abs(t) + np.abs(rhs)
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(den > 0, np.abs(res) / den, 0.0)
    finite = bool(np.isfinite(x.astype(np.float64)).all())
    nv = int((rel > tol).sum())
    ok = mism == 0 and nv == 0 and pad_bad == 0 and finite
    return ExtResult(ok, f"ref_mismatch={mism} resid_viol={nv} pad_bad={pad_bad} finite={finite}",
                     float(rel.max(initial=0.0)), handle.last_kernel)


# ------------------------------------------------------------------------------------------------ complex gemm
@dataclasses.dataclass
class CgemmCase:
    dtype: str = "c64"
    transa: str = "n"
    transb: str = "n"
    m: int = 11
    n: int = 11
    k: int = 16
    alpha: complex = 1.5 + 1.0j
    beta: complex = 1.5 + 3.0j
    lda_mul: int = 1
    ldb_mul: int = 1
    ldc_mul: int = 1
    offset: int = 0
    batch: int = 1
    stride_mul: int = 1
    conj: bool = False     # handle option: BLAS conjugate transposes instead of the reference's 'c' == 't'
    seed: int = 12345

    def ident(self) -> str:
        return (f"cgemm-{self.dtype}-{self.transa}{self.transb}-{self.m}x{self.n}x{self.k}-a{self.alpha}b{self.beta}"
                f"-ld{self.lda_mul}{self.ldb_mul}{self.ldc_mul}-off{self.offset}-bs{self.batch}s{self.stride_mul}"
                + ("-conj" if self.conj else ""))


def run_cgemm(handle: blas.SB_Handle, cs: CgemmCase) -> ExtResult:
    tt, npdt, rdt, kind, tol = CPLX[cs.dtype]
    ta, tb = cs.transa.lower() != "n", cs.transb.lower() != "n"
    m, n, k, batch, off = cs.m, cs.n, cs.k, cs.batch, cs.offset
    lda = max((k if ta else m), 1) * cs.lda_mul
    ldb = max((n if tb else k), 1) * cs.ldb_mul
    ldc = max(m, 1) * cs.ldc_mul
    size_a, size_b, size_c = (m if ta else k) * lda, (k if tb else n) * ldb, n * ldc
    sa, sb, sc = size_a * cs.stride_mul, size_b * cs.stride_mul, size_c * cs.stride_mul
    rng = np.random.default_rng(cs.seed)

    def rand(count):
        return (oracle.random_uniform(rng, count, rdt) + 1j * oracle.random_uniform(rng, count, rdt)).astype(npdt)

    a_h = rand(max((batch - 1) * sa + size_a, 1) + off)
    b_h = rand(max((batch - 1) * sb + size_b, 1) + off)
    c_h = rand(max((batch - 1) * sc + size_c, 1) + off)
    osa, osb, osc = (sa, sb, sc) if batch > 1 else (0, 0, 0)
    truth = c_h.copy()
    st_want = ox.cgemm(cs.transa, cs.transb, m, n, k, cs.alpha, a_h[off:], lda, b_h[off:], ldb, cs.beta, truth[off:], ldc,
                       stridea=osa, strideb=osb, stridec=osc, batch=batch, conj=cs.conj)
    dev = torch.device("cuda", handle.device)
    a_d, b_d, c_d = (torch.from_numpy(x).to(dev) for x in (a_h, b_h, c_h))
    text = ""
    handle.set_conj_transpose(cs.conj)
    try:
        if batch == 1:
            blas._gemm(handle, cs.transa, cs.transb, m, n, k, cs.alpha, a_d[off:], lda, b_d[off:], ldb, cs.beta,
                       c_d[off:], ldc)
        else:
            blas._gemm_strided_batched(handle, cs.transa, cs.transb, m, n, k, cs.alpha, a_d[off:], lda, sa, b_d[off:],
                                       ldb, sb, cs.beta, c_d[off:], ldc, sc, batch)
        handle.wait()
    except ValueError as e:
        text = str(e)
    finally:
        handle.set_conj_transpose(False)
    if st_want or text:
        return _status_result(st_want, text)
    got = c_d.cpu().numpy()
    # bound: |alpha| |A||B| + |beta||C| with |z| taken as |re| + |im|
    l1 = lambda z: (np.abs(z.real) + np.abs(z.imag)).astype(np.complex128)   # noqa: E731
    bound = l1(c_h)
    ox.cgemm(cs.transa, cs.transb, m, n, k, abs(cs.alpha.real) + abs(cs.alpha.imag), l1(a_h)[off:], lda, l1(b_h)[off:],
             ldb, abs(complex(cs.beta).real) + abs(complex(cs.beta).imag), bound[off:], ldc, stridea=osa, strideb=osb,
             stridec=osc, batch=batch, conj=False)
    bound = bound.real
    untouched = truth == c_h
    err = np.abs(got.astype(np.complex128) - truth.astype(np.complex128))
    viol = np.where(untouched & (got == c_h), False, err > tol * bound)
    # elements outside every window must be bit-identical
    window = np.zeros(c_h.size, bool)
    for b in range(batch):
        ox.view(window[off + b * osc:], m, n, ldc)[...] = True
    viol = np.where(window, viol, got != c_h)
    # the reference's predicate, applied to the real and imaginary parts (float_comparison.hpp:163-188)
    mism = oracle.compare(got.real.copy(), truth.real.astype(rdt), kind) + \
        oracle.compare(got.imag.copy(), truth.imag.astype(rdt), kind)
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(window & (bound > 0), err / bound, 0.0)
    nv = int(viol.sum())
    return ExtResult(mism == 0 and nv == 0, f"ref_mismatch={mism} bound_viol={nv}", float(rel.max(initial=0.0)),
                     handle.last_kernel)
