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


def test_host_path_validates_before_it_copies(handle):
    """pbx_gemm_host sizes its host <-> device copies from the strides, so pbx_gemm's validation (same order, same texts:
    gemm_interface.hpp:144-165) runs BEFORE any buffer size is derived; batch == 0 is the same no-op as in pbx_gemm."""
    m = n = k = 16
    a = torch.ones(4 * m * k).pin_memory()
    b = torch.ones(4 * k * n).pin_memory()
    c = torch.full((4 * m * n,), 3.0).pin_memory()
    kw = dict(stridea=m * k, strideb=k * n, stridec=m * n, batch_size=4)
    for bad, text in ((dict(kw, stridec=-m * n), "invalid _stridec"), (dict(kw, stridec=m * n - 1), "invalid _stridec"),
                      (dict(kw, stridea=-1), "invalid _stridea"), (dict(kw, strideb=-1), "invalid _strideb")):
        with pytest.raises(ValueError, match=text):
            blas.gemm_host(handle, "n", "n", m, n, k, 1.0, a, m, b, k, 0.0, c, m, **bad)
    with pytest.raises(ValueError, match="invalid _TransA"):
        blas.gemm_host(handle, "x", "n", m, n, k, 1.0, a, m, b, k, 0.0, c, m, **kw)
    with pytest.raises(ValueError, match="invalid _TransB"):
        blas.gemm_host(handle, "n", "y", m, n, k, 1.0, a, m, b, k, 0.0, c, m, **kw)
    blas.gemm_host(handle, "n", "n", m, n, k, 1.0, a, m, b, k, 0.0, c, m, **dict(kw, batch_size=0))   # no-op
    assert bool((c == 3.0).all()), "nothing may be written for batch == 0 or after a rejected call"


def test_calls_leave_the_current_device_alone(handle):
    """Every entry point works on the handle's device and restores the caller's current device (a single-process
    multi-GPU torch program must not be switched to another GPU by a library call).  With one GPU the check is that
    the current device is untouched; with two, a handle on device 1 is driven while device 0 is current."""
    from portblas_b200 import SB_Handle
    before = torch.cuda.current_device()
    dev = 1 if torch.cuda.device_count() > 1 else 0
    with torch.cuda.device(dev):
        stream_ptr = torch.cuda.current_stream(dev).cuda_stream
    h = SB_Handle(dev, stream_ptr)
    assert torch.cuda.current_device() == before
    x = torch.ones(64 * 64, device=f"cuda:{dev}")
    y = torch.zeros(64 * 64, device=f"cuda:{dev}")
    blas._gemm(h, "n", "n", 64, 64, 64, 1.0, x, 64, x, 64, 0.0, y, 64)
    h.wait()
    assert torch.cuda.current_device() == before
    assert bool((y == 64.0).all())
    h.close()
    assert torch.cuda.current_device() == before
