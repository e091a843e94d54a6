"""Parity at BASELINE.json's FULL sizes, where the CPU oracle would take hours: size-independent properties of
C = alpha*op(A)*op(B) + beta*C0, evaluated on the device in fp64 (torch is plumbing here, not the product):

  * checksum of checksums (Huang-Abraham): the column sums of C must equal alpha*(e^T op(A))*op(B) + beta*e^T C0 and
    the row sums alpha*op(A)*(op(B) e) + beta*C0 e -- two O(N^2) fp64 products that involve EVERY element of C;
  * 64 x 64 sampled entries recomputed in fp64.

Both against the per-dtype bound of the small-size parity tests (fp64 1e-12, fp32 1e-5, f16 2e-3, bf16 1.6e-2)
relative to |alpha||op(A)||op(B)| + |beta||C0| (summed the same way for the checksums).  Configurations:
BASELINE configs[1] DGEMM 8192^3 x {NN,NT,TN,TT} x {(1,0),(1.5,0.5)}, configs[2] SGEMM 16384^3 (3xTF32), configs[3]
strided-batched f16 / bf16 4096 x 256^3, configs[4] SGEMM 512 x 512 x 2^20 (split-K).
"""
from __future__ import annotations

import itertools

import pytest
import torch

from portblas_b200 import blas

pytestmark = pytest.mark.gpu

TOL = {torch.float64: 1e-12, torch.float32: 1e-5, torch.float16: 2e-3, torch.bfloat16: 1.6e-2}


def _rand(count, dt, gen):
    out = torch.empty(count, device="cuda", dtype=dt)
    chunk = 1 << 26
    for s in range(0, count, chunk):
        e = min(count, s + chunk)
        out[s:e] = (torch.rand(e - s, device="cuda", dtype=torch.float32, generator=gen) * 7.0 - 2.0).to(dt)
    return out


def _op(buf, trans, rows, cols, batch):
    """Logical op(X) (batch, rows, cols) view of a column-major flat buffer with minimal ld."""
    if trans == "n":
        return buf.view(batch, cols, rows).transpose(1, 2)
    return buf.view(batch, rows, cols)


def _check(handle, dt, ta, tb, m, n, k, alpha, beta, batch=1, seed=1):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    a = _rand(m * k * batch, dt, gen)
    b = _rand(k * n * batch, dt, gen)
    c0 = _rand(m * n * batch, dt, gen) if beta != 0 else torch.zeros(m * n * batch, device="cuda", dtype=dt)
    c = c0.clone()
    lda, ldb = (m if ta == "n" else k), (k if tb == "n" else n)
    if batch == 1:
        blas._gemm(handle, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, m)
    else:
        blas._gemm_strided_batched(handle, ta, tb, m, n, k, alpha, a, lda, m * k, b, ldb, k * n, beta, c, m, m * n, batch)
    handle.wait()
    tol = TOL[dt]
    A, B = _op(a, ta, m, k, batch), _op(b, tb, k, n, batch)            # (batch, m, k), (batch, k, n) views
    C, C0 = _op(c, "n", m, n, batch), _op(c0, "n", m, n, batch)
    f64 = torch.float64
    # ---- column sums: e^T C = alpha (e^T A) B + beta e^T C0
    u = A.sum(dim=1, dtype=f64).unsqueeze(1)                            # (batch, 1, k)
    ua = A.abs().sum(dim=1, dtype=f64).unsqueeze(1)
    want = alpha * _chunked_bmm(u, B) + beta * C0.sum(dim=1, dtype=f64).unsqueeze(1)
    bound = abs(alpha) * _chunked_bmm(ua, B, absolute=True) + abs(beta) * C0.abs().sum(dim=1, dtype=f64).unsqueeze(1)
    got = C.sum(dim=1, dtype=f64).unsqueeze(1)
    rel_col = float(((got - want).abs() / bound).max())
    # ---- row sums: C e = alpha A (B e) + beta C0 e
    v = B.sum(dim=2, dtype=f64).unsqueeze(2)                            # (batch, k, 1)
    va = B.abs().sum(dim=2, dtype=f64).unsqueeze(2)
    want = alpha * _chunked_bmm_left(A, v) + beta * C0.sum(dim=2, dtype=f64).unsqueeze(2)
    bound = abs(alpha) * _chunked_bmm_left(A, va, absolute=True) + abs(beta) * C0.abs().sum(dim=2, dtype=f64).unsqueeze(2)
    got = C.sum(dim=2, dtype=f64).unsqueeze(2)
    rel_row = float(((got - want).abs() / bound).max())
    # ---- sampled entries
    g2 = torch.Generator(device="cpu").manual_seed(m * 31 + n * 17 + k)
    ri = torch.randint(0, m, (min(m, 64),), generator=g2).cuda()
    ci = torch.randint(0, n, (min(n, 64),), generator=g2).cuda()
    rel_s = 0.0
    for bi in sorted({0, batch // 2, batch - 1}):
        As, Bs = A[bi][ri, :].to(f64), B[bi][:, ci].to(f64)
        want = alpha * (As @ Bs) + beta * C0[bi][ri][:, ci].to(f64)
        bound = abs(alpha) * (As.abs() @ Bs.abs()) + abs(beta) * C0[bi][ri][:, ci].to(f64).abs()
        rel_s = max(rel_s, float(((C[bi][ri][:, ci].to(f64) - want).abs() / bound).max()))
    assert rel_col <= tol and rel_row <= tol and rel_s <= tol, \
        f"{dt} {ta}{tb} {m}x{n}x{k} x{batch}: colsum {rel_col:.2e} rowsum {rel_row:.2e} sampled {rel_s:.2e} (tol {tol})"
    return rel_col, rel_row, rel_s


def _chunked_bmm(u, B, absolute=False, chunk=1 << 14):
    """(batch,1,k) @ (batch,k,n) in fp64 without materialising an fp64 copy of B."""
    k = B.shape[1]
    out = torch.zeros(B.shape[0], 1, B.shape[2], device=B.device, dtype=torch.float64)
    for s in range(0, k, chunk):
        blk = B[:, s:s + chunk, :].to(torch.float64)
        out += u[:, :, s:s + chunk] @ (blk.abs() if absolute else blk)
    return out


def _chunked_bmm_left(A, v, absolute=False, chunk=1 << 14):
    """(batch,m,k) @ (batch,k,1) in fp64 without materialising an fp64 copy of A."""
    k = A.shape[2]
    out = torch.zeros(A.shape[0], A.shape[1], 1, device=A.device, dtype=torch.float64)
    for s in range(0, k, chunk):
        blk = A[:, :, s:s + chunk].to(torch.float64)
        out += (blk.abs() if absolute else blk) @ v[:, s:s + chunk, :]
    return out


def test_cfg2_dgemm_8192_all_transposes(handle):
    for (ta, tb), (al, be) in itertools.product([("n", "n"), ("n", "t"), ("t", "n"), ("t", "t")], [(1.0, 0.0), (1.5, 0.5)]):
        _check(handle, torch.float64, ta, tb, 8192, 8192, 8192, al, be)
    assert handle.last_kernel == "dmma"


def test_cfg3_sgemm_16384(handle):
    _check(handle, torch.float32, "n", "n", 16384, 16384, 16384, 1.0, 0.0)
    assert handle.last_kernel == "tcgen05" and handle.last_presplit == 3   # tf32 + 2 x bf16 pre-split
    # one M-block shard of the 8-GPU partition (rows [0, 2048) with the ORIGINAL leading dimensions is what a rank runs;
    # here the compact equivalent) and a transposed, beta != 0 variant
    _check(handle, torch.float32, "t", "n", 2048, 16384, 16384, 1.5, 0.5)


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_cfg4_strided_batched_4096x256(handle, dt):
    _check(handle, dt, "n", "n", 256, 256, 256, 1.0, 0.0, batch=4096)
    assert handle.last_kernel == "tcgen05"
    _check(handle, dt, "t", "t", 256, 256, 256, 1.5, 0.5, batch=512)


def test_cfg5_tall_skinny_split_k(handle):
    _check(handle, torch.float32, "n", "n", 512, 512, 1 << 20, 1.0, 0.0)
    assert handle.last_kernel == "tcgen05" and handle.last_split_k > 1
    _check(handle, torch.float32, "t", "n", 512, 512, 1 << 18, 1.5, 0.5)


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_trsm_symm_large_on_device_residuals(handle, dt):
    """_trsm / _symm at sizes where the recursive substitution has several levels and every GEMM runs on the tensor
    cores: residual |op(A) X - alpha B| against |op(A)||X| + |alpha||B| in fp64 on the device (the unused triangle
    holds NaN), and the _symm product against the fp64 product of the mirrored matrix."""
    tol, f64 = TOL[dt], torch.float64
    gen = torch.Generator(device="cuda").manual_seed(3)
    for side, uplo, trans, diag, m, n in [("l", "l", "n", "n", 4096 + 40, 1536), ("r", "u", "t", "u", 1000, 4096 + 72),
                                          ("l", "u", "n", "n", 3000, 520), ("r", "l", "n", "n", 520, 3000)]:
        kk = m if side == "l" else n
        low = torch.tril(torch.rand(kk, kk, device="cuda", dtype=f64, generator=gen) * 2 - 1, -1) / kk ** 0.5
        t = low + torch.diag(torch.rand(kk, device="cuda", dtype=f64, generator=gen) * 9 + 1)      # lower, well conditioned
        tri = t if uplo == "l" else t.T                                                              # logical A
        stored = torch.where(torch.tril(torch.ones_like(tri, dtype=torch.bool)) if uplo == "l"
                             else torch.triu(torch.ones_like(tri, dtype=torch.bool)), tri, torch.nan)
        a = stored.T.contiguous().view(-1).to(dt)                                                    # column-major
        b0 = _rand(m * n, dt, gen)
        b = b0.clone()
        blas._trsm(handle, side, uplo, trans, diag, m, n, 2.0, a, kk, b, m)
        handle.wait()
        A = torch.nan_to_num(a.view(kk, kk).T.to(f64), nan=0.0)
        if diag == "u":
            A = A - torch.diag(torch.diag(A)) + torch.eye(kk, device="cuda", dtype=f64)
        opA = A.T if trans == "t" else A
        X, R = b.view(n, m).T.to(f64), 2.0 * b0.view(n, m).T.to(f64)
        if side == "l":
            res, den = opA @ X - R, opA.abs() @ X.abs() + R.abs()
        else:
            res, den = X @ opA - R, X.abs() @ opA.abs() + R.abs()
        rel = float((res.abs() / den).max())
        assert torch.isfinite(X).all() and rel <= tol, f"trsm {dt} {side}{uplo}{trans}{diag} {m}x{n}: residual {rel:.2e}"
    for side, uplo, m, n in [("l", "u", 4096, 2048 + 8), ("r", "l", 1536, 4096)]:
        kk = m if side == "l" else n
        a = _rand(kk * kk, dt, gen)
        bmat, c0 = _rand(m * n, dt, gen), _rand(m * n, dt, gen)
        c = c0.clone()
        blas._symm(handle, side, uplo, m, n, 1.5, a, kk, bmat, m, 0.5, c, m)
        handle.wait()
        A = a.view(kk, kk).T.to(f64)
        tri = torch.tril(A) if uplo == "l" else torch.triu(A)
        S = tri + tri.T - torch.diag(torch.diag(A))
        Bm, C0, C = bmat.view(n, m).T.to(f64), c0.view(n, m).T.to(f64), c.view(n, m).T.to(f64)
        want = 1.5 * (S @ Bm if side == "l" else Bm @ S) + 0.5 * C0
        bound = 1.5 * (S.abs() @ Bm.abs() if side == "l" else Bm.abs() @ S.abs()) + 0.5 * C0.abs()
        rel = float(((C - want).abs() / bound).max())
        assert rel <= tol, f"symm {dt} {side}{uplo} {m}x{n}: {rel:.2e}"


@pytest.mark.parametrize("dt", [torch.complex64, torch.complex128])
def test_complex_gemm_large_on_device(handle, dt):
    tol = 1e-5 if dt == torch.complex64 else 1e-12
    rdt = torch.float32 if dt == torch.complex64 else torch.float64
    gen = torch.Generator(device="cuda").manual_seed(5)
    for ta, tb, m, n, k in [("n", "n", 2048, 2048, 2048), ("t", "n", 1000, 1536, 2200), ("n", "t", 1536, 1000, 2200)]:
        def crand(cnt):
            return torch.complex(_rand(cnt, rdt, gen), _rand(cnt, rdt, gen))
        a, b, c0 = crand(m * k), crand(k * n), crand(m * n)
        c = c0.clone()
        al, be = 1.5 + 1.0j, 0.5 - 2.0j
        blas._gemm(handle, ta, tb, m, n, k, al, a, m if ta == "n" else k, b, k if tb == "n" else n, be, c, m)
        handle.wait()
        z = torch.complex128
        A = (a.view(k, m).T if ta == "n" else a.view(m, k)).to(z)
        B = (b.view(n, k).T if tb == "n" else b.view(k, n)).to(z)
        C0, C = c0.view(n, m).T.to(z), c.view(n, m).T.to(z)
        want = al * (A @ B) + be * C0
        l1 = lambda x: x.real.abs() + x.imag.abs()   # noqa: E731
        bound = (abs(al.real) + abs(al.imag)) * (l1(A) @ l1(B)) + (abs(be.real) + abs(be.imag)) * l1(C0)
        rel = float(((C - want).abs() / bound).max())
        assert rel <= tol, f"cgemm {dt} {ta}{tb} {m}x{n}x{k}: {rel:.2e}"
