"""GPU tests of the host-buffer entry point pbx_gemm_host (the end-to-end path of bench.py): the
pipelined panel / batch-chunk schedule must give the same result as one device-resident call, must
write only the M x N window of the caller's C (ld padding preserved) and must honour beta.

Mirrors what a reference caller does around one GEMM (samples/gemm.cpp:50-63: copy_to_device,
_gemm, copy_to_host, wait)."""
from __future__ import annotations

import itertools

import numpy as np
import pytest
import torch

from portblas_b200 import blas

pytestmark = pytest.mark.gpu

TRANS = [("n", "n"), ("n", "t"), ("t", "n"), ("t", "t")]
TOL = {torch.float32: 1e-5, torch.float64: 1e-12, torch.float16: 2e-3, torch.bfloat16: 1.6e-2}


def _check(handle, tdt, ta, tb, m, n, k, alpha, beta, ldc_pad, batch=1, bcast_a=False):
    rng = np.random.default_rng(1234 + m + 3 * n + 7 * k)
    a_rows, a_cols = (k, m) if ta == "t" else (m, k)
    b_rows, b_cols = (n, k) if tb == "t" else (k, n)
    lda, ldb, ldc = a_rows, b_rows, m + ldc_pad
    sa, sb, sc = (0 if bcast_a else lda * a_cols), ldb * b_cols, ldc * n
    na = lda * a_cols * (1 if bcast_a else batch)
    a64 = rng.uniform(-2, 5, na)
    b64 = rng.uniform(-2, 5, sb * batch)
    c64 = rng.uniform(-2, 5, sc * batch)
    a_h = torch.from_numpy(a64).to(tdt).pin_memory()
    b_h = torch.from_numpy(b64).to(tdt).pin_memory()
    c_h = torch.from_numpy(c64).to(tdt).pin_memory()
    c_before = c_h.clone()
    blas.gemm_host(handle, ta, tb, m, n, k, alpha, a_h, lda, b_h, ldb, beta, c_h, ldc,
                   stridea=sa if batch > 1 else 0, strideb=sb if batch > 1 else 0, stridec=sc if batch > 1 else 0,
                   batch_size=batch)
    af = a_h.to(torch.float64).numpy()
    bf = b_h.to(torch.float64).numpy()
    cf = c_before.to(torch.float64).numpy()
    got = c_h.to(torch.float64).numpy()
    for bi in range(batch):
        A = af[(0 if bcast_a else bi * sa):][:lda * a_cols].reshape(a_cols, lda).T
        B = bf[bi * sb:][:ldb * b_cols].reshape(b_cols, ldb).T
        C0 = cf[bi * sc:][:ldc * n].reshape(n, ldc).T
        G = got[bi * sc:][:ldc * n].reshape(n, ldc).T
        opA = A.T if ta == "t" else A
        opB = B.T if tb == "t" else B
        want = alpha * (opA @ opB) + (beta * C0[:m] if beta != 0 else 0.0)
        bound = abs(alpha) * (np.abs(opA) @ np.abs(opB)) + abs(beta) * np.abs(C0[:m])
        err = np.abs(G[:m] - want)
        assert np.all(err <= TOL[tdt] * bound + 1e-300), f"max rel {np.max(err / bound):.3e}"
        assert np.array_equal(G[m:], C0[m:]), "ld padding of the host C buffer was overwritten"


@pytest.mark.parametrize("tdt", [torch.float32, torch.float64, torch.bfloat16])
def test_host_pipelined_panels(handle, tdt):
    """batch == 1, large enough for the column-panel pipeline (several panels, ragged last panel)."""
    for (ta, tb), beta, pad in itertools.product(TRANS, [0.0, 0.5], [0, 8]):
        _check(handle, tdt, ta, tb, 1536, 2304 + 40, 1024, 1.5, beta, pad)


@pytest.mark.parametrize("tdt", [torch.float16, torch.float32])
def test_host_pipelined_batches(handle, tdt):
    for (ta, tb), beta, bcast in itertools.product([("n", "n"), ("t", "n")], [0.0, 0.5], [False, True]):
        _check(handle, tdt, ta, tb, 384, 256, 320, 1.0, beta, 8, batch=44, bcast_a=bcast)


def test_host_small_simple_path(handle):
    for (ta, tb), beta in itertools.product(TRANS, [0.0, 0.5]):
        _check(handle, torch.float32, ta, tb, 65, 33, 47, 1.5, beta, 3)
        _check(handle, torch.float64, ta, tb, 65, 33, 47, 1.5, beta, 3, batch=3)


@pytest.mark.parametrize("tdt", [torch.float32, torch.float64])
def test_host_k_streamed_first_block(handle, tdt):
    """Deep single GEMM with fp32 / fp64 C: the first column block is accumulated over K panels (beta applied
    once, then beta' = 1), later blocks run with A resident; ragged n and k, all transposes, ld padding."""
    for (ta, tb), beta, pad in itertools.product(TRANS, [0.0, 0.5], [0, 4]):
        _check(handle, tdt, ta, tb, 520, 2048 + 300, 2048 + 136, 1.5, beta, pad)
    _check(handle, tdt, "n", "n", 300, 4096 + 24, 9000, -0.5, 2.0, 4)
