"""Shared case runner for the GPU parity tests (and tools/gpu_check.py).

Follows the reference's ``verify_gemm`` (test/unittest/blas3/blas3_gemm_common.hpp:96-230):
leading dimensions are ``rows * ld_mul``, buffers hold ``batch`` matrices plus an element
``offset``, inputs are U(-2,5), the expected result is computed per batch entry on the host, and
the WHOLE C buffer (offset prefix and ld padding included) is compared, so writes outside the MxN
window fail.  Expected values come from the CPU oracle (oracle/); the comparison predicate is the
reference's ``almost_equal`` plus a tighter error-bound check against the long-double truth.
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np
import torch

from oracle import oracle
from portblas_b200 import blas

# dtype name -> (torch in, torch out, numpy compute, reference margin kind, truth tolerance)
# truth tolerance: |got - truth| <= tol * (|alpha| * |A||B| + |beta||C|) elementwise
DTYPES = {
    "f32": (torch.float32, torch.float32, np.float32, "float", 1e-5),
    "f64": (torch.float64, torch.float64, np.float64, "double", 1e-12),
    "f16": (torch.float16, torch.float16, np.float32, "half", 2e-3),
    "f16f32": (torch.float16, torch.float32, np.float32, "float", 1e-5),
    "bf16": (torch.bfloat16, torch.bfloat16, np.float32, "half", 1.6e-2),
    "bf16f32": (torch.bfloat16, torch.float32, np.float32, "float", 1e-5),
}


@dataclasses.dataclass
class Case:
    dtype: str = "f32"
    transa: str = "n"
    transb: str = "n"
    m: int = 16
    n: int = 16
    k: int = 16
    alpha: float = 1.5
    beta: float = 0.5
    lda_mul: int = 1
    ldb_mul: int = 1
    ldc_mul: int = 1
    offset: int = 0
    batch: int = 1
    batch_type: int = 0            # 0 strided, 1 interleaved
    api: str = "gemm"              # gemm | batched | strided
    stride_a_mul: int = 1
    stride_b_mul: int = 1
    stride_c_mul: int = 1
    seed: int = 12345
    kernel: int = 0                # forced pbx_kernel_t (0 = auto)
    split_k: int = 0
    env: tuple = ()                # ((name, value), ...) set around the call (PBX_TC_CONFIG, PBX_TF32_RAW_HI ...)
    zero_low_bits: int = 0         # clear the low mantissa bits of the fp32 inputs (reference set_to_zero_last_nbits,
                                   # test/blas_test.hpp:225-244: 13 for tf32 / half fragments)

    def ident(self) -> str:
        return (f"{self.dtype}-{self.api}-{self.transa}{self.transb}-{self.m}x{self.n}x{self.k}-a{self.alpha}"
                f"b{self.beta}-ld{self.lda_mul}{self.ldb_mul}{self.ldc_mul}-off{self.offset}-bs{self.batch}"
                f"t{self.batch_type}-s{self.stride_a_mul}{self.stride_b_mul}{self.stride_c_mul}"
                f"-k{self.kernel}-sk{self.split_k}" + "".join(f"-{name[4:]}={val}" for name, val in self.env)
                + (f"-z{self.zero_low_bits}" if self.zero_low_bits else ""))


@dataclasses.dataclass
class Result:
    ok: bool
    ref_mismatch: int        # elements failing the reference's almost_equal
    bound_violations: int    # elements failing the truth error bound
    max_rel_bound: float     # max |got-truth| / bound-denominator
    kernel: str
    split_k: int
    detail: str = ""


def _storage_round(x: np.ndarray, dtype: str) -> np.ndarray:
    if dtype.startswith("f16"):
        return oracle.round_to(x, "f16")
    if dtype.startswith("bf16"):
        return oracle.round_to(x, "bf16")
    return x


def _out_round(x: np.ndarray, dtype: str) -> np.ndarray:
    if dtype == "f16":
        return oracle.round_to(x, "f16")
    if dtype == "bf16":
        return oracle.round_to(x, "bf16")
    return x


def run_case(handle: blas.SB_Handle, cs: Case) -> Result:
    tin, tout, npdt, margin_kind, tol = DTYPES[cs.dtype]
    ta, tb = cs.transa.lower() != "n", cs.transb.lower() != "n"
    m, n, k, batch = cs.m, cs.n, cs.k, cs.batch
    lda = (k if ta else m) * cs.lda_mul
    ldb = (n if tb else k) * cs.ldb_mul
    ldc = m * cs.ldc_mul
    size_a, size_b, size_c = m * k * cs.lda_mul, k * n * cs.ldb_mul, m * n * cs.ldc_mul
    if cs.api == "strided":
        sa, sb, sc = size_a * cs.stride_a_mul, size_b * cs.stride_b_mul, size_c * cs.stride_c_mul
    else:
        sa, sb, sc = size_a, size_b, size_c
    interleaved = cs.api == "batched" and cs.batch_type == 1
    # buffer sizes: strided layout needs (batch-1)*stride + size (stride may be 0 or > size)
    buf_a = max(size_a * batch, (batch - 1) * sa + size_a) + cs.offset
    buf_b = max(size_b * batch, (batch - 1) * sb + size_b) + cs.offset
    buf_c = max(size_c * batch, (batch - 1) * sc + size_c) + cs.offset

    rng = np.random.default_rng(cs.seed)
    a_h = _storage_round(oracle.random_uniform(rng, buf_a, npdt), cs.dtype)
    b_h = _storage_round(oracle.random_uniform(rng, buf_b, npdt), cs.dtype)
    c_h = _out_round(oracle.random_uniform(rng, buf_c, npdt), cs.dtype)
    if cs.zero_low_bits and npdt == np.float32:
        mask = np.uint32((0xFFFFFFFF << cs.zero_low_bits) & 0xFFFFFFFF)
        for arr in (a_h, b_h, c_h):
            arr.view(np.uint32)[...] &= mask
    off = cs.offset

    # ---- expected (strided layout, per batch entry), as the reference's tests do ----
    osa, osb, osc = (sa, sb, sc) if batch > 1 else (0, 0, 0)
    exp = c_h.copy()
    truth = c_h.copy()
    bound = np.abs(c_h).copy()
    st_exp = oracle.gemm(cs.transa, cs.transb, m, n, k, cs.alpha, a_h[off:], lda, b_h[off:], ldb, cs.beta,
                         exp[off:], ldc, stridea=osa, strideb=osb, stridec=max(osc, 0), batch=batch,
                         mode=oracle.MODE_REF)
    if st_exp == 0:
        oracle.gemm(cs.transa, cs.transb, m, n, k, cs.alpha, a_h[off:], lda, b_h[off:], ldb, cs.beta, truth[off:],
                    ldc, stridea=osa, strideb=osb, stridec=osc, batch=batch, mode=oracle.MODE_TRUTH)
        oracle.gemm(cs.transa, cs.transb, m, n, k, abs(cs.alpha), np.abs(a_h)[off:], lda, np.abs(b_h)[off:], ldb,
                    abs(cs.beta), bound[off:], ldc, stridea=osa, strideb=osb, stridec=osc, batch=batch,
                    mode=oracle.MODE_TRUTH)

    # ---- device run ----
    if interleaved:
        a_rows, a_cols = (k, m) if ta else (m, k)
        b_rows, b_cols = (n, k) if tb else (k, n)
        a_dev_h = np.concatenate([a_h[:off], oracle.interleave(a_h[off:], a_rows, a_cols, lda, batch, size_a)])
        b_dev_h = np.concatenate([b_h[:off], oracle.interleave(b_h[off:], b_rows, b_cols, ldb, batch, size_b)])
        # the reference interleaves whole ld x cols footprints (padding rows included)
        c_il = np.zeros(ldc * n * batch, dtype=npdt)
        c_il.reshape(n, ldc, batch)[...] = np.transpose(c_h[off:off + size_c * batch].reshape(batch, n, ldc), (1, 2, 0))
        c_dev_h = np.concatenate([c_h[:off], c_il])
    else:
        a_dev_h, b_dev_h, c_dev_h = a_h, b_h, c_h
    dev = torch.device("cuda", handle.device)
    a_d = torch.from_numpy(a_dev_h).to(dev).to(tin)
    b_d = torch.from_numpy(b_dev_h).to(dev).to(tin)
    c_d = torch.from_numpy(c_dev_h).to(dev).to(tout)
    torch.cuda.synchronize()

    handle.set_forced_kernel(cs.kernel)
    handle.set_split_k(cs.split_k)
    import os
    saved_env = {name: os.environ.get(name) for name, _ in cs.env}
    for name, val in cs.env:
        os.environ[name] = str(val)
    if cs.env:
        handle.reload_env()   # the PBX_* switches are read into the handle, not per call
    status_text = ""
    try:
        if cs.api == "gemm":
            blas._gemm(handle, cs.transa, cs.transb, m, n, k, cs.alpha, a_d[off:], lda, b_d[off:], ldb, cs.beta,
                       c_d[off:], ldc)
        elif cs.api == "batched":
            blas._gemm_batched(handle, cs.transa, cs.transb, m, n, k, cs.alpha, a_d[off:], lda, b_d[off:], ldb,
                               cs.beta, c_d[off:], ldc, batch, blas.gemm_batch_type_t(cs.batch_type))
        else:
            blas._gemm_strided_batched(handle, cs.transa, cs.transb, m, n, k, cs.alpha, a_d[off:], lda, sa,
                                       b_d[off:], ldb, sb, cs.beta, c_d[off:], ldc, sc, batch)
        handle.wait()
    except ValueError as e:
        status_text = str(e)
    finally:
        handle.set_forced_kernel(0)
        handle.set_split_k(0)
        for name, old in saved_env.items():
            if old is None:
                os.environ.pop(name, None)
            else:
                os.environ[name] = old
        if cs.env:
            handle.reload_env()
    kern, sk = handle.last_kernel, handle.last_split_k
    if st_exp != 0 or status_text:
        ok = oracle.STATUS_TEXT.get(st_exp, "?") == status_text
        return Result(ok, 0, 0, 0.0, kern, sk, f"status oracle='{oracle.STATUS_TEXT.get(st_exp)}' got='{status_text}'")

    got = c_d.to(torch.float64 if npdt == np.float64 else torch.float32).cpu().numpy()
    if interleaved:
        g = got[off:].reshape(n, ldc, batch)
        got = np.concatenate([got[:off], np.ascontiguousarray(np.transpose(g, (2, 0, 1))).reshape(-1)])
        got = np.concatenate([got, c_h[got.size:]])  # strided buffer may be longer than the interleaved one

    exp_r = _out_round(exp, cs.dtype)
    ref_mismatch = oracle.compare(got, exp_r, margin_kind)
    err = np.abs(got.astype(np.float64) - truth.astype(np.float64))
    denom = bound.astype(np.float64)
    # elements outside every MxN window: truth == original C, must be bit-identical
    viol = err > tol * denom + (0.0 if cs.dtype in ("f32", "f64", "f16f32", "bf16f32") else 0.0)
    untouched = (truth == c_h) & (exp == c_h)
    viol = np.where(untouched, got != c_h, viol)
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(denom > 0, err / denom, 0.0)
    nviol = int(viol.sum())
    ok = ref_mismatch == 0 and nviol == 0
    detail = ""
    if not ok:
        bad = np.flatnonzero(viol)[:4]
        detail = "; ".join(f"[{i}] got={got[i]:.9g} truth={truth[i]:.9g}" for i in bad)
    return Result(ok, ref_mismatch, nviol, float(rel.max()) if rel.size else 0.0, kern, sk, detail)
