"""numpy model of pbx_trsm's host logic (portblas_b200/csrc/blas3_ext.cu: trsm_impl): the same diagonal-block
inverses, block ranges, operand offsets, recursion order and alpha propagation, with numpy matmuls standing in
for pbx_gemm.  Lets the CPU test-suite check the scheme (and calibrate its residual bound) without a GPU."""
from __future__ import annotations

import numpy as np

from oracle import blas3_ext as ox


def _trtri_lower(L: np.ndarray) -> np.ndarray:
    """Row-oriented forward substitution of trtri_diag_kernel, in L's precision."""
    nb = L.shape[0]
    X = np.zeros_like(L)
    for i in range(nb):
        s = -(L[i, :i] @ X[:i, :])
        s[i] += 1
        X[i, :] = s / L[i, i]
        X[i, i + 1:] = 0
    return X


def trsm_model(side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb, nb) -> None:
    dt = B.dtype.type
    left, lower = side.lower() == "l", uplo.lower() == "l"
    tr, unit = trans.lower() == "t", diag.lower() == "u"
    K = m if left else n
    nblk = (K + nb - 1) // nb
    a, b = ox.view(A, K, K, lda), ox.view(B, m, n, ldb)
    inv = []
    for blk in range(nblk):
        i0 = blk * nb
        cur = min(nb, K - i0)
        L = np.eye(nb, dtype=B.dtype)
        sub = a[i0:i0 + cur, i0:i0 + cur]
        L[:cur, :cur] = np.tril(sub) if lower else np.triu(sub).T
        if unit:
            np.fill_diagonal(L, 1)
        Xi = _trtri_lower(L)
        inv.append(Xi if lower else Xi.T)
    X = np.zeros((m, n), dtype=B.dtype)
    op = (lambda z: z.T) if tr else (lambda z: z)
    op_lower = lower != tr
    forward = op_lower if left else not op_lower

    def a_sub(p0, plen, q0, qlen):   # op(A)[P, Q] without touching anything outside the stored block
        return a[q0:q0 + qlen, p0:p0 + plen].T if tr else a[p0:p0 + plen, q0:q0 + qlen]

    def rec(b0, b1, scaled):
        if b1 - b0 == 1:
            i0 = b0 * nb
            bs = min(nb, K - i0)
            al = dt(1) if scaled else dt(alpha)
            iv = op(inv[b0][:bs, :bs])
            if left:
                X[i0:i0 + bs] = al * (iv @ b[i0:i0 + bs])
            else:
                X[:, i0:i0 + bs] = al * (b[:, i0:i0 + bs] @ iv)
            return
        mid = b0 + (b1 - b0 + 1) // 2
        (p0, p1), (q0, q1) = ((b0, mid), (mid, b1)) if forward else ((mid, b1), (b0, mid))
        rec(p0, p1, scaled)
        pi, plen = p0 * nb, min(p1 * nb, K) - p0 * nb
        qi, qlen = q0 * nb, min(q1 * nb, K) - q0 * nb
        be = dt(1) if scaled else dt(alpha)
        if left:
            b[qi:qi + qlen] = -(a_sub(qi, qlen, pi, plen) @ X[pi:pi + plen]) + be * b[qi:qi + qlen]
        else:
            b[:, qi:qi + qlen] = -(X[:, pi:pi + plen] @ a_sub(pi, plen, qi, qlen)) + be * b[:, qi:qi + qlen]
        rec(q0, q1, True)

    rec(0, nblk, False)
    b[...] = X
