"""Non-finite operands (Inf / NaN inside A): which kernel gives what.

IEEE fp32 / fp64 arithmetic -- the reference's scalar-FMA kernels (src/operations/blas3/gemm_local.hpp:752-772), this
repository's CUDA-core kernel, the fp64 DMMA kernel and the 16-bit tensor-core kernels (exact products, fp32 accumulate)
-- turn one +Inf in row i of op(A) into +/-Inf across row i of C (the sign follows the B element it meets) and leave
every other row untouched.  The fp32 tensor-core path computes a*b as split products (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo);
the lo half of an Inf is forced to 0, but Inf * (lo half of a tf32-exact b) = Inf * 0 = NaN still reaches the
accumulator, so row i comes out NaN there where IEEE gives Inf.  That deviation is confined to rows / columns that
are non-finite under IEEE as well, and is pinned here: finite rows stay within the fp32 bound, the affected row is
non-finite on every kernel, and exactly +/-Inf on the IEEE-exact ones.  (A NaN operand gives NaN everywhere alike.)
"""
from __future__ import annotations

import pytest
import torch

from portblas_b200 import blas

pytestmark = pytest.mark.gpu
SIMT, TCGEN05, DMMA = 1, 2, 3
CASES = [("f32-simt", torch.float32, torch.float32, SIMT, True, 1e-5), ("f32-tcgen05", torch.float32, torch.float32, TCGEN05, False, 1e-5),
         ("f64-dmma", torch.float64, torch.float64, DMMA, True, 1e-12), ("bf16-tcgen05", torch.bfloat16, torch.float32, TCGEN05, True, 1e-5),
         ("f16-tcgen05", torch.float16, torch.float16, TCGEN05, True, 2e-3)]


@pytest.mark.parametrize("bad", [float("inf"), float("nan")], ids=["inf", "nan"])
@pytest.mark.parametrize("name,tin,tout,kernel,ieee_exact,tol", CASES, ids=[c[0] for c in CASES])
def test_one_nonfinite_element_of_a(handle, name, tin, tout, kernel, ieee_exact, tol, bad):
    m, n, k = 264, 136, 200
    i0, k0 = 77, 31
    g = torch.Generator(device="cuda").manual_seed(3)
    a = (torch.rand(k, m, device="cuda", generator=g) * 7 - 2).to(tin)      # [col][row], column-major, lda = m
    b = (torch.rand(n, k, device="cuda", generator=g) * 7 - 2).to(tin)
    b[b == 0] = 1                                                          # keep Inf * b away from Inf * 0
    a_ok = a.clone()
    a[k0, i0] = bad
    c = torch.zeros(n * m, device="cuda", dtype=tout)
    handle.set_forced_kernel(kernel)
    try:
        blas._gemm(handle, "n", "n", m, n, k, 1.0, a.view(-1), m, b.view(-1), k, 0.0, c, m)
        handle.wait()
    finally:
        handle.set_forced_kernel(0)
    got = c.view(n, m).T.double()
    want = a_ok.double().T @ b.double().T
    bound = a_ok.double().T.abs() @ b.double().T.abs()
    rows = torch.ones(m, dtype=torch.bool, device="cuda")
    rows[i0] = False
    assert torch.isfinite(got[rows]).all(), f"{name}: a finite row was contaminated"
    assert ((got[rows] - want[rows]).abs() <= tol * bound[rows]).all(), name
    assert (~torch.isfinite(got[i0])).all(), f"{name}: the row of the non-finite element must be non-finite"
    if bad != bad:
        assert torch.isnan(got[i0]).all()
    elif ieee_exact:
        sign = torch.sign(b.double().T[k0])          # +Inf * b[k0, j]
        assert torch.equal(got[i0], sign * float("inf")), f"{name}: IEEE gives +/-Inf along the row"
