"""CPU oracle for the routines the reference builds on its GEMM path: _symm, _trsm, complex _gemm.
TEST INFRASTRUCTURE ONLY -- imported by tests/ and __graft_entry__.smoke(), never by portblas_b200/.

Restated from the reference (file:line are relative to /root/reference):
  * ``symm``      src/interface/symm_interface.hpp:35-75 (uplo checked before side; side 'l' -> GEMM M,N,K=M with a
                  symmetric A, side 'r' -> GEMM M,N,K=N with a symmetric B) and the mirroring loader
                  src/operations/blas3/gemm_local.hpp:813-873 (an element outside the stored triangle is read
                  from its mirror image).  The product itself goes through the existing GEMM oracle.
  * ``trsm_ref_algorithm``  src/interface/trsm_interface.hpp:105-387: invert the 16x16 diagonal blocks
                  (DiagonalBlocksInverter, src/operations/blas3/trsm.hpp:60-160), then the block substitution made
                  of _gemm calls, X copied back to B at the end.
  * ``cgemm``     complex GEMM as the naive reference kernel computes it (src/operations/blas3/gemm_ref.hpp:204-260)
                  with the front end of src/interface/gemm_interface.hpp:105-185 -- including its quirk that
                  'c' is NOT a conjugate transpose (``_TrA = _TransA != 'n'``, :150-151).
The oracles the reference's own tests use are CBLAS symm / trsm / gemm (reference_blas::symm / trsm / cgemm,
common/include/common/system_reference_blas.hpp); here scipy.linalg.blas = OpenBLAS 0.3.x.

Parity unpinned against reference OUTPUTS (no SYCL compiler here, the reference ships no golden vectors); pinned
against CBLAS on the reference's own parameter grids and tolerance predicate in tests/test_blas3_ext_oracle.py.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg.blas as sblas

STATUS_TEXT = {
    0: "ok",
    1: "invalid _TransA", 2: "invalid _TransB", 3: "invalid _stridec", 4: "invalid _stridea", 5: "invalid _strideb",
    10: "invalid _uplo", 11: "invalid _side",
    12: "invalid matrix size argument", 13: "invalid Side argument", 14: "invalid Triangle argument",
    15: "invalid Transpose argument", 16: "invalid Diagonal argument",
}


def view(buf: np.ndarray, rows: int, cols: int, ld: int) -> np.ndarray:
    """Column-major rows x cols window of a flat buffer (a writable view)."""
    return np.lib.stride_tricks.as_strided(buf, shape=(rows, cols), strides=(buf.itemsize, ld * buf.itemsize),
                                           writeable=True)


# ------------------------------------------------------------------------------------------------ symm
def symm_full(uplo: str, k: int, A: np.ndarray, lda: int) -> np.ndarray:
    """The k x k matrix the mirroring loader presents to the GEMM (gemm_local.hpp:813-873)."""
    a = view(A, k, k, lda)
    tri = np.tril(a) if uplo.lower() == "l" else np.triu(a)
    return tri + tri.T - np.diag(np.diag(a))


def symm_status(side: str, uplo: str) -> int:
    if uplo.lower() not in ("u", "l"):
        return 10
    if side.lower() not in ("l", "r"):
        return 11
    return 0


def symm(side, uplo, m, n, alpha, A, lda, B, ldb, beta, C, ldc, acc=np.longdouble) -> int:
    """C <- alpha*sym(A)*B + beta*C (side l) or alpha*B*sym(A) + beta*C (side r), accumulated in ``acc``."""
    st = symm_status(side, uplo)
    if st:
        return st
    k = m if side.lower() == "l" else n
    c = view(C, m, n, ldc)
    if alpha == 0:  # the GEMM front end's shortcut: C <- beta*C (beta == 0 stores zeros)
        c[...] = (acc(beta) * c.astype(acc)).astype(C.dtype) if beta != 0 else 0
        return 0
    full = symm_full(uplo, k, A, lda).astype(acc)
    b = view(B, m, n, ldb).astype(acc)
    prod = full @ b if side.lower() == "l" else b @ full
    out = acc(alpha) * prod
    if beta != 0:
        out = out + acc(beta) * c.astype(acc)
    c[...] = out.astype(C.dtype)
    return 0


def cblas_symm(side, uplo, m, n, alpha, A, lda, B, ldb, beta, C, ldc) -> None:
    f = sblas.dsymm if C.dtype == np.float64 else sblas.ssymm
    k = m if side.lower() == "l" else n
    a = np.asfortranarray(view(A, k, k, lda))
    b = np.asfortranarray(view(B, m, n, ldb))
    c = np.asfortranarray(view(C, m, n, ldc))
    r = f(alpha, a, b, beta=beta, c=c, side=0 if side.lower() == "l" else 1, lower=1 if uplo.lower() == "l" else 0)
    view(C, m, n, ldc)[...] = r


# ------------------------------------------------------------------------------------------------ trsm
def trsm_status(side, uplo, trans, diag, m, n, lda, ldb) -> int:
    if m == 0 or n == 0 or lda == 0 or ldb == 0:
        return 12
    if side.lower() not in ("l", "r"):
        return 13
    if uplo.lower() not in ("u", "l"):
        return 14
    if trans.lower() not in ("n", "t"):
        return 15
    if diag.lower() not in ("u", "n"):
        return 16
    return 0


def _invert_diag_block(blk: np.ndarray, upper: bool, unit: bool) -> np.ndarray:
    """DiagonalBlocksInverter::eval (trsm.hpp:60-160) for one 16x16 block already masked to its triangle
    (identity on out-of-range rows is NOT added by the reference: out-of-range entries are zero)."""
    bs = blk.shape[0]
    loc = blk.copy()
    if not unit:
        for i in range(bs):
            loc[i, i] = blk.dtype.type(1) / loc[i, i] if loc[i, i] != 0 else loc[i, i]
    if upper:
        for j in range(1, bs):
            col = loc[:j, :j] @ loc[:j, j]           # sum_k local(i,k) * local(k,j), k < j
            loc[:j, j] = col * (-1 if unit else -loc[j, j])
    else:
        for j in range(bs - 2, -1, -1):
            col = loc[j + 1:, j + 1:] @ loc[j + 1:, j]
            loc[j + 1:, j] = col * (-1 if unit else -loc[j, j])
    return loc


def trsm_ref_algorithm(side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb) -> int:
    """The reference's algorithm in B's own precision: 16-wide inverted diagonal blocks + GEMM substitution."""
    st = trsm_status(side, uplo, trans, diag, m, n, lda, ldb)
    if st:
        return st
    dt = B.dtype.type
    left, upper = side.lower() == "l", uplo.lower() == "u"
    tr, unit = trans.lower() == "t", diag.lower() == "u"
    K = m if left else n
    bs = 16
    a = view(A, K, K, lda)
    b = view(B, m, n, ldb)
    x = np.zeros((m, n), dtype=B.dtype)
    nblk = (K + bs - 1) // bs
    inv = []
    for blk in range(nblk):
        i0 = blk * bs
        cur = min(bs, K - i0)
        t = np.zeros((bs, bs), dtype=B.dtype)
        sub = a[i0:i0 + cur, i0:i0 + cur]
        t[:cur, :cur] = np.triu(sub) if upper else np.tril(sub)
        if unit:
            for i in range(cur):
                t[i, i] = 1
        # out-of-range diagonal entries stay zero in the reference; give them 1 so 1/x is defined (unused rows)
        for i in range(cur, bs):
            t[i, i] = 1
        inv.append(_invert_diag_block(t, upper, unit))
    op = (lambda mat: mat.T) if tr else (lambda mat: mat)
    op_lower = (not upper) != tr
    one = dt(1)
    if left:
        order = range(nblk) if op_lower else range(nblk - 1, -1, -1)
        first = True
        for blk in order:
            i0 = blk * bs
            cur = min(bs, K - i0)
            al = dt(alpha) if first else one
            x[i0:i0 + cur] = al * (op(inv[blk][:cur, :cur]) @ b[i0:i0 + cur])
            opa = op(a)
            if op_lower and i0 + bs < m:
                b[i0 + bs:] = -(opa[i0 + bs:, i0:i0 + cur] @ x[i0:i0 + cur]) + al * b[i0 + bs:]
            elif not op_lower and i0 > 0:
                b[:i0] = -(opa[:i0, i0:i0 + cur] @ x[i0:i0 + cur]) + al * b[:i0]
            first = False
    else:
        order = range(nblk) if not op_lower else range(nblk - 1, -1, -1)
        first = True
        for blk in order:
            i0 = blk * bs
            cur = min(bs, K - i0)
            al = dt(alpha) if first else one
            x[:, i0:i0 + cur] = al * (b[:, i0:i0 + cur] @ op(inv[blk][:cur, :cur]))
            opa = op(a)
            if not op_lower and i0 + bs < n:
                b[:, i0 + bs:] = -(x[:, i0:i0 + cur] @ opa[i0:i0 + cur, i0 + bs:]) + al * b[:, i0 + bs:]
            elif op_lower and i0 > 0:
                b[:, :i0] = -(x[:, i0:i0 + cur] @ opa[i0:i0 + cur, :i0]) + al * b[:, :i0]
            first = False
    b[...] = x
    return 0


def trsm_truth(side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb) -> np.ndarray:
    """Long-double substitution; returns the m x n solution (B is not modified)."""
    left, upper = side.lower() == "l", uplo.lower() == "u"
    tr, unit = trans.lower() == "t", diag.lower() == "u"
    K = m if left else n
    a = view(A, K, K, lda).astype(np.longdouble)
    t = np.triu(a) if upper else np.tril(a)
    if unit:
        np.fill_diagonal(t, 1)
    if tr:
        t = t.T
    rhs = np.longdouble(alpha) * view(B, m, n, ldb).astype(np.longdouble)
    if not left:  # X op(A) = R  <=>  op(A)^T X^T = R^T
        t, rhs = t.T, rhs.T
    lower = bool(np.allclose(np.triu(t, 1), 0))
    x = np.zeros_like(rhs)
    rng = range(K) if lower else range(K - 1, -1, -1)
    for i in rng:
        if lower:
            s = rhs[i] - t[i, :i] @ x[:i]
        else:
            s = rhs[i] - t[i, i + 1:] @ x[i + 1:]
        x[i] = s / t[i, i]
    return x if left else x.T


def cblas_trsm(side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb) -> None:
    f = sblas.dtrsm if B.dtype == np.float64 else sblas.strsm
    K = m if side.lower() == "l" else n
    a = np.asfortranarray(view(A, K, K, lda))
    # the unused triangle may hold NaN; BLAS never reads it
    b = np.asfortranarray(view(B, m, n, ldb))
    r = f(alpha, a, b, side=0 if side.lower() == "l" else 1, lower=1 if uplo.lower() == "l" else 0,
          trans_a=1 if trans.lower() == "t" else 0, diag=1 if diag.lower() == "u" else 0)
    view(B, m, n, ldb)[...] = r


def fill_trsm_matrix(rng: np.random.Generator, k: int, lda: int, uplo: str, diag: str, diag_value: float,
                     unused: float, dtype) -> np.ndarray:
    """test/blas_test.hpp:186-214: a well-conditioned triangular matrix, the other triangle set to ``unused``."""
    A = np.zeros(k * lda, dtype=dtype)
    a = view(A, k, k, lda)
    full = np.full((k, k), unused, dtype=np.float64)     # full[i, j], i = row of the LOWER form
    s = np.full(k, 1.0 if diag.lower() == "u" else abs(diag_value))
    for j in range(k):                                   # column by column, all rows i > j at once
        full[j, j] = diag_value
        rows = slice(j + 1, k)
        limit = s[rows] / np.sqrt(float(k) - float(j))
        v = np.where(s[rows] >= 1.0, rng.uniform(-1.0, 1.0, size=k - j - 1) * limit, 0.0)
        s[rows] -= np.abs(v)
        full[rows, j] = v
    a[...] = (full if uplo.lower() == "l" else full.T).astype(dtype)
    return A


# ------------------------------------------------------------------------------------------------ complex gemm
def cgemm_status(transa, transb, n, ldc, stridea, strideb, stridec, batch) -> int:
    if transa.lower() not in "ntc":
        return 1
    if transb.lower() not in "ntc":
        return 2
    if batch > 1:
        if stridec < ldc * n or stridec < 0:
            return 3
        if stridea < 0:
            return 4
        if strideb < 0:
            return 5
    return 0


def cgemm(transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, stridea=0, strideb=0, stridec=0, batch=1,
          conj=False, acc=np.clongdouble) -> int:
    """Complex C <- alpha*op(A)*op(B) + beta*C on flat complex buffers, per batch entry.  ``conj`` False restates
    the reference ('c' == 't'); True is the BLAS meaning."""
    if alpha == 0:
        if m == 0 or n == 0 or batch == 0 or beta == 1:
            return 0
        for b in range(batch):
            c = view(C[b * stridec:], m, n, ldc)
            c[...] = (acc(beta) * c.astype(acc)).astype(C.dtype) if beta != 0 else 0
        return 0
    st = cgemm_status(transa, transb, n, ldc, stridea, strideb, stridec, batch)
    if st:
        return st
    if m == 0 or n == 0 or batch == 0:
        return 0

    def op(buf, t, rows, cols, ld):
        t = t.lower()
        if t == "n":
            return view(buf, rows, cols, ld).astype(acc)
        s = view(buf, cols, rows, ld).astype(acc).T
        return np.conj(s) if (t == "c" and conj) else s

    for b in range(batch):
        c = view(C[b * stridec:], m, n, ldc)
        out = np.zeros((m, n), dtype=acc)
        if k > 0:
            out = acc(alpha) * (op(A[b * stridea:], transa, m, k, lda) @ op(B[b * strideb:], transb, k, n, ldb))
        if beta != 0:
            out = out + acc(beta) * c.astype(acc)
        c[...] = out.astype(C.dtype)
    return 0


def cblas_cgemm(transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc) -> None:
    f = sblas.zgemm if C.dtype == np.complex128 else sblas.cgemm
    code = {"n": 0, "t": 1, "c": 2}
    ta, tb = transa.lower(), transb.lower()
    a = np.asfortranarray(view(A, k if ta != "n" else m, m if ta != "n" else k, lda))
    b = np.asfortranarray(view(B, n if tb != "n" else k, k if tb != "n" else n, ldb))
    c = np.asfortranarray(view(C, m, n, ldc))
    r = f(alpha, a, b, beta=beta, c=c, trans_a=code[ta], trans_b=code[tb])
    view(C, m, n, ldc)[...] = r
