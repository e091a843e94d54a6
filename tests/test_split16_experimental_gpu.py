"""EXPERIMENTAL fp32 path, opt-in: tf32 + 2 x bf16 (PBX_F32_SPLIT16=1; gemm_tcgen05.cu PRE == 3, split16_kernel).

A_hi*B_hi stays one tf32 MMA on the raw tiles; the two cross terms run as kind::f16 MMAs on bf16 copies made by a
pre-pass (bf16(a) and bf16(a - trunc_tf32(a))) -- two tf32-MMA times per k-step instead of three, and the split itself
stays two orders of magnitude inside the 1e-5 budget (tools/split_emulation.py, profiles/r01/split_emulation.txt).

The kernel variant was written after round 1's GPU budget was spent and has NEVER run on a GPU: a wrong byte count on an
mbarrier would hang the device, so these tests run only when PBX_RUN_EXPERIMENTAL=1 is set (tools/gpu_split16.sh wraps
them in a timeout).  The default path does not change: the SASS of every pre-existing kernel is byte-identical with and
without the variant compiled in."""
import itertools
import os

import pytest

from gemm_case import Case, run_case

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("PBX_RUN_EXPERIMENTAL") != "1",
                                 reason="experimental kernel variant: opt in with PBX_RUN_EXPERIMENTAL=1 (under a timeout)")]
TRANS = [("n", "n"), ("n", "t"), ("t", "n"), ("t", "t")]
S16 = (("PBX_F32_SPLIT16", "1"), ("PBX_TF32_PRESPLIT", "1"))
TCGEN05 = 2   # PBX_KERNEL_TCGEN05 (include/pbx_gemm.h)


def test_smallest_case_first(handle):
    """One single-CTA tile, one k-block per operand layout: if a descriptor or byte count is wrong it shows here."""
    for ta, tb in TRANS:
        cs = Case(dtype="f32", transa=ta, transb=tb, m=128, n=128, k=32, alpha=1.0, beta=0.0, kernel=TCGEN05,
                  env=S16 + (("PBX_TC_CONFIG", "1,128"),))
        r = run_case(handle, cs)
        assert r.ok and handle.last_presplit == 3, (cs.ident(), r)


def test_every_tile_configuration_and_layout(handle):
    cases = []
    for cfg in ("1,128", "2,128", "2,256"):
        env = S16 + (("PBX_TC_CONFIG", cfg),)
        for (ta, tb), be in itertools.product(TRANS, [0.0, 0.5]):
            cases.append(Case(dtype="f32", transa=ta, transb=tb, m=392, n=520, k=1096, alpha=1.5, beta=be, kernel=TCGEN05,
                              env=env))
        cases.append(Case(dtype="f32", m=300, n=260, k=4104, alpha=1.0, beta=0.5, kernel=TCGEN05, split_k=3, env=env))
        cases.append(Case(dtype="f32", api="strided", transa="t", m=264, n=392, k=200, alpha=1.0, beta=0.0, batch=5,
                          stride_a_mul=0, kernel=TCGEN05, env=env))
        cases.append(Case(dtype="f32", api="strided", transb="t", m=264, n=136, k=328, alpha=-1.0, beta=1.0, batch=4,
                          stride_b_mul=2, stride_c_mul=2, kernel=TCGEN05, env=env))
    for ta, tb in TRANS:   # skinny-M swap, odd-ld repack, ld multipliers
        cases.append(Case(dtype="f32", transa=ta, transb=tb, m=40, n=1000, k=520, alpha=1.5, beta=0.5, kernel=TCGEN05, env=S16))
        cases.append(Case(dtype="f32", transa=ta, transb=tb, m=263, n=131, k=517, alpha=1.5, beta=0.5, offset=1,
                          kernel=TCGEN05, env=S16))
        cases.append(Case(dtype="f32", transa=ta, transb=tb, m=200, n=136, k=264, lda_mul=2, ldb_mul=3, ldc_mul=4,
                          kernel=TCGEN05, env=S16))
    cases.append(Case(dtype="f32", m=512, n=512, k=65536, alpha=1.0, beta=0.0, env=S16))
    for cs in cases:
        r = run_case(handle, cs)
        assert r.ok, (cs.ident(), r)
