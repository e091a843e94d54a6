"""beta == 0 never reads C: every kernel family and every epilogue, with C (ld padding included) pre-filled with NaN / Inf.

Reference behaviour: the production kernels specialise on beta == 0 and start the accumulator from zero instead of
beta*C (src/operations/blas3/gemm_local.hpp:401-409; the comment at gemm_ref.hpp:245-251 states why: uninitialised C
must not leak NaN into the result).  Here the branch is a runtime `beta0` test in each epilogue, so each one is driven:
tcgen05 direct stores (single CTA, CTA pair 256x128 / 256x256), the TMA-store epilogue of 16-bit outputs on and off,
the skinny-M transposed epilogue, split-K partials + reduce, the pre-split and in-kernel-split fp32 paths, DMMA, SIMT,
the interleaved kernels, pbx_gemm_host and pbx_gemm_multicast.  The M x N window must come out finite and equal to the
fp64 product; every element outside it (rows m..ldc-1 of each column) must still hold its NaN.

alpha == 0 and beta == 0 on NaN-filled C: the CUDA path stores exact zeros (BLAS), where the reference's _scal_matrix
computes 0*C and propagates the NaN (src/interface/blas1_interface.hpp:468-510; confirmed by running the reference,
tests/test_oracle_ref.py::test_front_end_rules_are_the_reference_s).  That deviation is pinned here.
"""
from __future__ import annotations

import os

import pytest
import torch

from portblas_b200 import blas

pytestmark = pytest.mark.gpu

SIMT, TCGEN05, DMMA = 1, 2, 3
TD = {"f32": (torch.float32, torch.float32, 2e-5), "f64": (torch.float64, torch.float64, 1e-12),
      "f16": (torch.float16, torch.float16, 2e-3), "bf16": (torch.bfloat16, torch.bfloat16, 1.6e-2),
      "f16f32": (torch.float16, torch.float32, 2e-5), "bf16f32": (torch.bfloat16, torch.float32, 2e-5)}


class _Env:
    """Environment switches around a call; the handle re-reads them (they are not read per call)."""

    def __init__(self, handle, **kv):
        self.handle = handle
        self.kv = {k: str(v) for k, v in kv.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)
        self.handle.reload_env()

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        self.handle.reload_env()


def _operands(dt, ta, tb, m, n, k, batch, dev, seed=7):
    tin, tout, tol = TD[dt]
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    ar, ac = (k, m) if ta != "n" else (m, k)
    br, bc = (n, k) if tb != "n" else (k, n)
    a = (torch.rand(batch, ac, ar, device=dev, generator=g) * 7 - 2).to(tin)   # [b][col][row]: column-major, ld = rows
    b = (torch.rand(batch, bc, br, device=dev, generator=g) * 7 - 2).to(tin)
    a64 = a.double().transpose(1, 2)
    b64 = b.double().transpose(1, 2)
    opa = a64.transpose(1, 2) if ta != "n" else a64
    opb = b64.transpose(1, 2) if tb != "n" else b64
    want = opa @ opb                       # [b][m][n]
    bound = opa.abs() @ opb.abs()
    return a, b, want, bound, tout, tol


def _check(c, m, n, ldc, batch, want, bound, alpha, tol, poison, what):
    cv = c.view(batch, n, ldc).double()
    win = cv[:, :, :m].transpose(1, 2)
    assert torch.isfinite(win).all(), f"{what}: non-finite values inside the M x N window (beta == 0 read C)"
    err = (win - alpha * want).abs()
    assert (err <= tol * abs(alpha) * bound + 1e-300).all(), f"{what}: max err/bound {(err / bound).max().item():.3e}"
    if ldc > m:
        pad = c.view(batch, n, ldc)[:, :, m:]
        ok = torch.isnan(pad).all() if poison != poison else (pad == poison).all()
        assert ok, f"{what}: ld padding was written"


CASES = [
    # (id, dtype, kernel, ta, tb, m, n, k, batch, split_k, env)
    ("tc_f32_cg1", "f32", TCGEN05, "n", "n", 200, 136, 264, 1, 1, dict(PBX_TC_CONFIG="1,128", PBX_TF32_PRESPLIT=0)),
    ("tc_f32_cg2_128", "f32", TCGEN05, "t", "n", 520, 264, 200, 1, 1, dict(PBX_TC_CONFIG="2,128", PBX_TF32_PRESPLIT=0)),
    ("tc_f32_cg2_256", "f32", TCGEN05, "n", "t", 520, 520, 136, 1, 1, dict(PBX_TC_CONFIG="2,256", PBX_TF32_PRESPLIT=0)),
    ("tc_f32_presplit", "f32", TCGEN05, "n", "n", 520, 264, 200, 1, 1, dict(PBX_TC_CONFIG="2,256", PBX_TF32_PRESPLIT=1)),
    ("tc_f32_deepk_chunks", "f32", TCGEN05, "n", "n", 136, 136, 2056, 1, 1, dict(PBX_TF32_PRESPLIT=0)),
    ("tc_f32_tf32x1", "f32", TCGEN05, "n", "n", 264, 136, 200, 1, 1, dict(SB_ENABLE_JOINT_MATRIX=1)),
    ("tc_f32_swap", "f32", TCGEN05, "n", "n", 40, 520, 264, 1, 1, {}),
    ("tc_f32_swap_splitk", "f32", TCGEN05, "t", "t", 24, 264, 2048, 1, 4, {}),
    ("tc_f32_splitk", "f32", TCGEN05, "n", "n", 136, 136, 4096, 1, 5, {}),
    ("tc_f32_batched", "f32", TCGEN05, "n", "n", 136, 72, 136, 5, 1, {}),
    ("tc_bf16_tmastore", "bf16", TCGEN05, "n", "n", 264, 136, 200, 3, 1, {}),
    ("tc_bf16_direct", "bf16", TCGEN05, "n", "n", 264, 136, 200, 3, 1, dict(PBX_TMA_STORE=0)),
    ("tc_f16_tmastore_cg2", "f16", TCGEN05, "t", "n", 520, 520, 136, 1, 1, dict(PBX_TC_CONFIG="2,256")),
    ("tc_f16_swap_tmastore", "f16", TCGEN05, "n", "t", 40, 520, 264, 1, 1, {}),
    ("tc_f16_swap_direct", "f16", TCGEN05, "n", "t", 40, 520, 264, 1, 1, dict(PBX_TMA_STORE=0)),
    ("tc_bf16_splitk", "bf16", TCGEN05, "n", "n", 136, 136, 4096, 1, 4, {}),
    ("tc_f16f32", "f16f32", TCGEN05, "n", "n", 264, 136, 200, 1, 1, {}),
    ("tc_bf16f32_cg2", "bf16f32", TCGEN05, "t", "t", 520, 264, 136, 1, 1, dict(PBX_TC_CONFIG="2,128")),
    ("dmma", "f64", DMMA, "n", "n", 200, 136, 264, 1, 1, {}),
    ("dmma_tt", "f64", DMMA, "t", "t", 137, 75, 99, 3, 1, {}),
    ("dmma_splitk", "f64", DMMA, "n", "t", 136, 136, 4096, 1, 4, {}),
    ("simt_f32", "f32", SIMT, "n", "n", 70, 33, 51, 2, 1, {}),
    ("simt_f32_splitk", "f32", SIMT, "t", "n", 70, 33, 1024, 1, 3, {}),
    ("simt_f64", "f64", SIMT, "n", "t", 70, 33, 51, 1, 1, {}),
    ("simt_f16", "f16", SIMT, "n", "n", 70, 33, 51, 1, 1, {}),
    ("simt_bf16f32", "bf16f32", SIMT, "t", "t", 70, 33, 51, 1, 1, {}),
    ("auto_f32_oddld", "f32", 0, "n", "n", 131, 77, 200, 1, 1, {}),   # repack path (lda = 131: not TMA-legal)
]


@pytest.mark.parametrize("poison", [float("nan"), float("inf")], ids=["nan", "inf"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_beta_zero_never_reads_c(handle, case, poison):
    name, dt, kernel, ta, tb, m, n, k, batch, sk, env = case
    dev = torch.device("cuda", handle.device)
    a, b, want, bound, tout, tol = _operands(dt, ta, tb, m, n, k, batch, dev)
    lda = a.shape[2]
    ldb = b.shape[2]
    ldc = m + 24   # 16-byte-legal padding for every type, so the TMA-store epilogue stays selected
    if env.get("SB_ENABLE_JOINT_MATRIX"):
        tol = 2e-3   # single-tf32 product: 10-bit-mantissa fragments, the reference's joint_matrix precision
    c = torch.full((batch * n * ldc,), poison, device=dev, dtype=tout)
    handle.set_forced_kernel(kernel)
    handle.set_split_k(sk if sk > 1 else 0)
    try:
        with _Env(handle, **env):
            if batch == 1:
                blas._gemm(handle, ta, tb, m, n, k, 1.5, a.view(-1), lda, b.view(-1), ldb, 0.0, c, ldc)
            else:
                blas._gemm_strided_batched(handle, ta, tb, m, n, k, 1.5, a.view(-1), lda, a[0].numel(), b.view(-1), ldb,
                                           b[0].numel(), 0.0, c, ldc, n * ldc, batch)
            handle.wait()
        used, used_sk = handle.last_kernel, handle.last_split_k
    finally:
        handle.set_forced_kernel(0)
        handle.set_split_k(0)
    if kernel:
        assert used == {SIMT: "simt", TCGEN05: "tcgen05", DMMA: "dmma"}[kernel], used
    if sk > 1:
        assert used_sk > 1, f"split-K was requested but not used ({used_sk})"
    _check(c, m, n, ldc, batch, want, bound, 1.5, tol, poison, f"{name} [{used}, split-K {used_sk}]")


@pytest.mark.parametrize("dt", ["f32", "f64", "f16", "bf16f32"])
@pytest.mark.parametrize("batch", [5, 64])
def test_beta_zero_interleaved_never_reads_c(handle, dt, batch):
    dev = torch.device("cuda", handle.device)
    m, n, k = 23, 19, 37
    a, b, want, bound, tout, tol = _operands(dt, "n", "t", m, n, k, batch, dev)
    ldc = m + 3
    # interleaved layout: element (r, c, b) at (c*ld + r)*batch + b (gemm_interleaved.hpp:265-271)
    a_il = a.permute(1, 2, 0).contiguous().view(-1)
    b_il = b.permute(1, 2, 0).contiguous().view(-1)
    c = torch.full((n * ldc * batch,), float("nan"), device=dev, dtype=tout)
    blas._gemm_batched(handle, "n", "t", m, n, k, 1.5, a_il, m, b_il, n, 0.0, c, ldc, batch,
                       blas.gemm_batch_type_t.interleaved)
    handle.wait()
    assert handle.last_kernel == "interleaved"
    got = c.view(n, ldc, batch).permute(2, 0, 1).contiguous().view(-1)
    _check(got, m, n, ldc, batch, want, bound, 1.5, tol, float("nan"), f"interleaved {dt} x{batch}")


@pytest.mark.parametrize("dt,m,n,k,batch", [("f64", 520, 264, 300, 1), ("f32", 2048, 1024, 1024, 1), ("bf16", 264, 136, 200, 6),
                                            ("f32", 72, 40, 56, 1)])
def test_beta_zero_gemm_host_never_reads_c(handle, dt, m, n, k, batch):
    dev = torch.device("cuda", handle.device)
    a, b, want, bound, tout, tol = _operands(dt, "n", "n", m, n, k, batch, dev)
    ldc = m + 8
    a_h = a.cpu().view(-1).pin_memory()
    b_h = b.cpu().view(-1).pin_memory()
    c_h = torch.full((batch * n * ldc,), float("nan"), dtype=tout).pin_memory()
    blas.gemm_host(handle, "n", "n", m, n, k, 1.5, a_h, m, b_h, k, 0.0, c_h, ldc, stridea=m * k if batch > 1 else 0,
                   strideb=k * n if batch > 1 else 0, stridec=n * ldc if batch > 1 else 0, batch_size=batch)
    _check(c_h.to(dev), m, n, ldc, batch, want, bound, 1.5, tol, float("nan"), f"gemm_host {dt}")


@pytest.mark.parametrize("dt,m,n,k", [("f32", 520, 264, 200), ("bf16", 264, 136, 200), ("f16", 264, 136, 200), ("f64", 200, 136, 96),
                                      ("f32", 33, 20, 40)])
def test_beta_zero_multicast_never_reads_c(handle, dt, m, n, k):
    """Two local copies of C (the peers' role on one GPU): both must come out finite and identical."""
    dev = torch.device("cuda", handle.device)
    a, b, want, bound, tout, tol = _operands(dt, "n", "n", m, n, k, 1, dev)
    ldc = m + 8
    cs = [torch.full((n * ldc,), float("nan"), device=dev, dtype=tout) for _ in range(2)]
    blas._gemm_multicast(handle, "n", "n", m, n, k, 1.5, a.view(-1), m, b.view(-1), k, 0.0, [x.data_ptr() for x in cs], ldc,
                         tout)
    handle.wait()
    for i, c in enumerate(cs):
        _check(c, m, n, ldc, 1, want, bound, 1.5, tol, float("nan"), f"multicast {dt} copy {i}")
    assert torch.equal(cs[0].view(n, ldc)[:, :m], cs[1].view(n, ldc)[:, :m])


@pytest.mark.parametrize("dt", ["f32", "f64", "bf16"])
def test_alpha_zero_beta_zero_stores_exact_zeros_on_nan_c(handle, dt):
    """Deviation from the reference, stated: its plain _gemm takes the _scal_matrix branch and computes 0*C = NaN
    (blas1_interface.hpp:468-510); the CUDA path stores exact zeros (the reference's own _scal branch and BLAS do)."""
    dev = torch.device("cuda", handle.device)
    m, n, k = 33, 20, 17
    tin, tout, _ = TD[dt]
    a = torch.ones(m * k, device=dev, dtype=tin)
    b = torch.ones(k * n, device=dev, dtype=tin)
    ldc = m + 5
    c = torch.full((n * ldc,), float("nan"), device=dev, dtype=tout)
    blas._gemm(handle, "n", "n", m, n, k, 0.0, a, m, b, k, 0.0, c, ldc)
    handle.wait()
    cv = c.view(n, ldc)
    assert (cv[:, :m] == 0).all()
    assert torch.isnan(cv[:, m:]).all()
